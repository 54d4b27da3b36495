"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_shares.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0]
    if not name.startswith(("void mmkdev", "mmkdev", "k_")):
        name = "(torch: synthetic genome generation) " + name[:60]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("kernel,launches,total_ms,share_pct")
for k, v in tot.most_common():
    print(f'"{k}",{cnt[k]},{v:.3f},{100 * v / total:.1f}')
