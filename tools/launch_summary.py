"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel.  usage: launch_summary.py in.csv [out.csv]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
K, V = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= V:
        continue
    name = re.sub(r'^void ', '', r[K])
    name = re.sub(r'\(.*', '', name).replace('ctl::<unnamed>::', 'ctl::').replace('at::native::', 'at::')
    try:
        v = float(r[V].replace(',', ''))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v / 1000.0
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print("launches %d, total %.1f us" % (n, tot))
lines = [("kernel", "launches", "total_us", "share_pct", "avg_us")]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append((k[:120], v[0], round(v[1], 1), round(100 * v[1] / tot, 2), round(v[1] / v[0], 1)))
for l in lines[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print("%-90s %6s %10s %7s %8s" % l)
if len(sys.argv) > 2 and sys.argv[2] != '-':
    with open(sys.argv[2], "w", newline="") as f:
        csv.writer(f).writerows(lines)
