"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    k = r[ki].split('(')[0]
    agg[k][0] += 1; agg[k][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches, {tot/1e6:.2f} ms of kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {v[0]:6d} launches {v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f} %")
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for r in data[a:b]: print(r[ki].split('(')[0][:30], r[gi], r[vi])
