"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [div]"""
import csv, collections, sys
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'nsecond': 1e-3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki][:100], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'total {tot / div:.1f} us')
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{a[0] / div:6.1f} launches {a[1] / div:10.1f} us {100 * a[1] / tot:5.1f}%  {k}')
