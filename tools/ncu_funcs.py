#!/usr/bin/env python
"""Per-function aggregation of the ncu source page (cuda,sass CSV) for the csrc/*.cuh files: samples, warp instructions, avg active threads."""
import csv, sys, re, os
csvp = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mcmc_eq_b200', 'csrc')
def funcs_of(fn):
    src = open(os.path.join(root, fn)).read().split('\n')
    out = []
    for i, l in enumerate(src):
        m = re.match(r'^(?:template <[^>]*>\s*)?(?:EIK_HD_NOINLINE|EIK_HD|__global__|static|inline)[^;(]*?([A-Za-z_0-9]+)\(', l)
        if m: out.append((i + 1, m.group(1)))
    return out
tables = {}
rows = list(csv.reader(open(csvp)))
cur = None; hdr = None; agg = {}; ts = ti = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    elif len(r) > 8 and r[0] == 'Line No': hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try: smp, ins, thr = float(r[6] or 0), float(r[7] or 0), float(r[8] or 0)
        except ValueError: continue
        l = int(r[0]); name = cur
        if cur and os.path.exists(os.path.join(root, cur)):
            if cur not in tables: tables[cur] = funcs_of(cur)
            f = '?'
            for ln, nm in tables[cur]:
                if ln <= l: f = nm
            name = f'{cur}:{f}'
        a = agg.setdefault(name, [0, 0, 0]); a[0] += smp; a[1] += ins; a[2] += thr
        ts += smp; ti += ins
print(f'samples {ts:.0f}  warp instructions {ti:.3g}')
for n, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{s/ts*100:5.1f}% samples {i/ti*100:5.1f}% instr ({i:.3g})  thr/inst {t/max(i,1):4.1f}  {n}")
