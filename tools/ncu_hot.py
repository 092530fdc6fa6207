"""Top SASS instructions by warp-stall samples for one kernel of an .ncu-rep (source page)."""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:120])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
i_src, i_n = hdr.index("Source"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[1:]):
    try:
        n = int(r[i_n])
    except (ValueError, IndexError):
        continue
    data.append((n, k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, k, r in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i]) for i in stalls), reverse=True)[:2]
    print(f"{100 * n / tot:5.1f}%  #{k:5d}  {r[i_src].strip()[:70]:70s}  {st}")
