"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> markdown table."""
import collections
import csv
import sys


def main(path, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        a = agg[row["Kernel Name"]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {title}\n")
    print(f"{n} launches, {tot / 1000:.2f} ms summed device time (ncu per-launch times are cold-cache and serialised: "
          f"compare SHARES).\n")
    print("| kernel | launches | total us | us/launch | share |")
    print("|---|---:|---:|---:|---:|")
    ours = 0.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "snn::" in k:
            ours += v[1]
        if v[1] / tot < 0.004:
            continue
        name = k.split("(")[0].replace("void ", "")[:90]
        print(f"| `{name}` | {v[0]} | {v[1]:.0f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
    print(f"\nKernels of this repo (`snn::*`): {100 * ours / tot:.1f}% of the summed device time; the rest are PyTorch "
          f"library kernels (critic sgemm, PPO-head elementwise ops, Adam, clip, gathers).")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "launch list")
