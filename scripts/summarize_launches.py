"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = name.replace("void tsfmx::<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(row["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms (gpu__time_duration.sum; cold-cache, serialised)")
print("| ms | share | launches | avg us | kernel |")
print("|---:|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"| {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k[:90]}` |")
