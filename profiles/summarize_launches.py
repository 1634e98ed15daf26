"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table (markdown)."""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"])
    except (ValueError, KeyError):
        continue
    us = v / 1000 if row["Metric Unit"].startswith("ns") else v
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("cv2::", "")
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.0005:
        continue
    print(f"| `{k[:70]}` | {v[0]} | {v[1] / 1000:.2f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
print(f"\ntotal {tot / 1000:.2f} ms over {sum(v[0] for v in agg.values())} launches")
