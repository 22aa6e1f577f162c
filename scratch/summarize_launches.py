"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python scratch/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/....md"""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
rows = [r for r in rows if r[0] >= skip]
tot = defaultdict(float); cnt = defaultdict(int)
def short(n):
    n = re.sub(r"\(.*$", "", n)
    return n[:90]
for _, name, us, _, _ in rows:
    tot[short(name)] += us; cnt[short(name)] += 1
total = sum(tot.values())
print(f"launches {len(rows)}, total {total:.1f} us\n")
print("| total us | launches | avg us | share | kernel |\n|---:|---:|---:|---:|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {v:.1f} | {cnt[k]} | {v / cnt[k]:.1f} | {100 * v / total:.1f}% | `{k}` |")
