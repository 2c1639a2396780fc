"""Per-(kernel, grid) table of an ncu launch list (second half = the profiled step): count, total/avg us, avg DRAM MB, TB/s."""
import csv, re, sys
from collections import defaultdict, OrderedDict
path = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(path) if not l.startswith('==')]
L = defaultdict(dict)
for r in csv.DictReader(lines):
    i = int(r['ID'])
    L[i]['name'] = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
    L[i]['grid'] = r['Grid Size']
    L[i][r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
ids = sorted(L)
half = len(ids) // 2
agg = OrderedDict()
for i in ids[half:]:
    d = L[i]
    k = (d['name'][:44], d['grid'])
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d['gpu__time_duration.sum'] / 1e3
    a[2] += (d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6
tot = sum(a[1] for a in agg.values())
totb = sum(a[2] for a in agg.values())
print("step: %d launches, %.1f us serialised, %.1f MB DRAM" % (sum(a[0] for a in agg.values()), tot, totb))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if flt and flt not in k[0]:
        continue
    print(f"{k[0]:44s} {k[1]:16s} x{a[0]:3d} tot {a[1]:7.1f} us avg {a[1]/a[0]:7.1f} us  avg {a[2]/a[0]:7.1f} MB  {a[2]/a[1]*1e-6*1e6/1e6:5.2f} TB/s")
