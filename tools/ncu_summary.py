"""Selected metrics of every kernel in an .ncu-rep (ncu --set full) -> markdown."""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid")]
print(f"# {title}\n")
print("`ncu --set full --clock-control none` (one capture; per-launch values, cold caches).\n")
print("| " + " | ".join(c[1] for c in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    cells = []
    for key, _ in cols:
        if key not in idx:
            cells.append("n/a")
            continue
        v, u = r[idx[key]], units[idx[key]]
        if key == "Kernel Name":
            v = "`" + v.split("(")[0].replace("void ", "")[:40] + "`"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}"
            except ValueError:
                pass
            if u and u not in ("%",):
                v += " " + u
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
