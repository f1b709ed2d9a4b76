"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_summary.py csv [out]"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.reader(lines)
h = next(r)
ik, im, iv = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
iu = h.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
n = 0
for row in r:
    if len(row) <= iv or row[im] != 'gpu__time_duration.sum':
        continue
    v = float(row[iv].replace(',', ''))
    u = row[iu]
    v_us = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    name = row[ik].split('(')[0][:90]
    agg[name][0] += 1
    agg[name][1] += v_us
    total += v_us
    n += 1
out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
print('%d launches, %.1f ms total device time (cold-cache, serialised under ncu: compare shares)' % (n, total / 1e3), file=out)
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%6.2f%%  %9.1f us  x%-5d avg %8.1f us  %s' % (100 * t / total, t, c, t / c, name), file=out)
