"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel summary CSV (launches, total, mean, share of the sea:: kernels)."""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')) if r]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or not r[vi]:
        continue
    name = re.sub(r'\(.*$', '', r[ki]).strip()
    us = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for k, v in agg.items() if 'sea::' in k)
print('kernel,launches,total_us,mean_us,share_pct')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if 'sea::' in k:
        print(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / v[0]:.1f},{100 * v[1] / tot:.1f}')
