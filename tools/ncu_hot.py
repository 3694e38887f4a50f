#!/usr/bin/env python
"""Hot regions of an ncu --page source --csv dump: runs of SASS with similar execution counts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; data = rows[2:]
ia, isrc, iex, ith, ismp = (h.index(x) for x in ('Address', 'Source', 'Instructions Executed', 'Avg. Threads Executed', '# Samples'))
tot = sum(float(r[iex]) for r in data if r[iex]); tots = sum(float(r[ismp] or 0) for r in data)
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
runs = []; cur = None
for r in data:
    e = float(r[iex] or 0)
    if cur and e > 0 and abs(e - cur['e']) / max(cur['e'], 1) < 0.3:
        cur['n'] += 1; cur['sum'] += e; cur['thr'] += float(r[ith] or 0); cur['end'] = r[ia]; cur['smp'] += float(r[ismp] or 0)
    else:
        if cur: runs.append(cur)
        cur = dict(start=r[ia], end=r[ia], e=e, n=1, sum=e, thr=float(r[ith] or 0), smp=float(r[ismp] or 0), first=r[isrc][:50])
runs.append(cur)
print('total warp instructions %.3g, sass lines %d' % (tot, len(data)))
for r in runs:
    if r['sum'] / tot > thresh:
        print(r['start'][-5:], r['end'][-5:], f"n={r['n']:4d} exec/inst={r['e']/1e6:7.1f}M share={r['sum']/tot*100:5.1f}% thr={r['thr']/r['n']:5.1f} smp={r['smp']/tots*100:5.1f}%", r['first'])
