"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total device time, share.
    python bench/launch_summary.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else v  # -> us
    name = r[kn].split("(")[0]
    t[name][0] += v
    t[name][1] += 1
tot = sum(v[0] for v in t.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for k, v in sorted(t.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {v[1]:8d} {v[0]:12.1f} {v[0] / tot * 100:6.1f}%")
