"""Per-iteration kernel times of the wavefront loop from an ncu launch list: python tools/iter_breakdown.py launches.csv"""
import csv, sys
import numpy as np
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
its = []; cur = None
for r in data:
    k = r[ki].split('(')[0].replace('void ', ''); v = float(r[vi].replace(',', '')) / 1e3
    if k == 'k_wf_reset':
        cur = {}; its.append(cur)
    if cur is not None and k.startswith('k_wf'): cur[k] = cur.get(k, 0) + v
print(len(its), 'iterations')
sel = [i for i in (2, 5, 10, 20, 30, 40, 50, 60, 80, 100, 150, 200, 250, 300, 400, 500, 600, 700) if i < len(its)]
for i in sel:
    d = its[i]; print(f"{i:4d} " + ' '.join(f"{k[5:]}={v:.0f}" for k, v in d.items()) + f"  sum={sum(d.values()):.0f} us")
tot = np.array([sum(d.values()) for d in its])
print('cumulative ms:', [(i, round(tot[:i].sum() / 1e3, 1)) for i in (25, 50, 100, 150, 200, 300, 400, 500, 600, 800, 1000) if i <= len(its)], 'total', round(tot.sum() / 1e3, 1))
