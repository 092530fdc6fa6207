"""Selected metrics of every kernel in an .ncu-rep (ncu --set full) -> markdown table; optionally a JSON with the
DRAM traffic / tensor-pipe figures bench.py attaches to its roofline object.

usage: python tools/ncu_summary.py REPORT.ncu-rep "title" [--names a,b,c,...] [--json out.json] [--commit SHA]
  --names: label of each captured launch, in capture order (e.g. fwd_layer0,fwd_layer1,...)"""
import csv
import json
import subprocess
import sys

args = sys.argv[1:]
rep, title = args[0], (args[1] if len(args) > 1 and not args[1].startswith("--") else args[0])
names = args[args.index("--names") + 1].split(",") if "--names" in args else None
jpath = args[args.index("--json") + 1] if "--json" in args else None
commit = args[args.index("--commit") + 1] if "--commit" in args else None      # the code the capture ran
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: tensor-core reads %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: LSU %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid")]


def num(r, key):
    v, u = r[idx[key]], units[idx[key]]
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    return x * scale


print(f"# {title}\n")
print("`ncu --set full --clock-control none` (one capture; per-launch values, cold caches).\n")
print("| " + ("launch | " if names else "") + " | ".join(c[1] for c in cols) + " |")
print("|" + "---|" * (len(cols) + (1 if names else 0)))
kern = {}
for n, r in enumerate(data):
    cells = [names[n]] if names and n < len(names) else ([""] if names else [])
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
    if names and n < len(names):
        kern[names[n]] = {
            "dram_bytes_per_launch": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
            "tensor_pipe_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
            "smem_pipe_tensor_pct": float(r[idx["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]].replace(",", "")),
            "smem_pipe_lsu_pct": float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]].replace(",", "")),
            "ncu_time_us": num(r, "gpu__time_duration.sum"),
        }
if jpath:
    with open(jpath, "w") as f:
        json.dump({"source": f"{rep} (ncu --set full, B=24576, C1 dims, T=5): {title}", "commit": commit, "kernels": kern}, f, indent=1)
