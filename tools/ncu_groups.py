#!/usr/bin/env python
"""Group the per-line profile into the kernel's phases (line ranges of eik_fast.cuh / eik_core.cuh)."""
import csv, sys, re
src = open('mcmc_eq_b200/csrc/eik_fast.cuh').read().split('\n')
def line_of(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    return 10**9
marks = [('sqrt_pos/node_update helpers', line_of('EIK_HD float sqrt_pos')), ('general fast_sweep (box phase, refined grid)', line_of('template <bool ROW>\nEIK_HD bool fast_sweep') if False else line_of('EIK_HD bool fast_sweep')),
         ('march_sweep (literal walk)', line_of('EIK_HD void march_sweep(')), ('march_sweep2 (two-pass march)', line_of('EIK_HD bool march_sweep2(')),
         ('perimeter/slow_line glue', line_of('EIK_HD void load_perimeter')), ('run_grid (round loop, outputs)', line_of('EIK_HD int run_grid(')),
         ('solve_warp (init, copy-back, copy-out)', line_of('EIK_HD int solve_warp(')), ('end', 10**9)]
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None; agg = {}
ts = ti = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    elif len(r) > 8 and r[0] == 'Line No': hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try: smp, ins, thr = float(r[6] or 0), float(r[7] or 0), float(r[8] or 0)
        except ValueError: continue
        l = int(r[0])
        if cur == 'eik_fast.cuh':
            name = 'eik_fast.cuh: header'
            for (n, a), (_, b) in zip(marks, marks[1:]):
                if a <= l < b: name = n
        else: name = cur + ' (generic core / other)'
        a = agg.setdefault(name, [0, 0, 0]); a[0] += smp; a[1] += ins; a[2] += thr
        ts += smp; ti += ins
for n, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{s/ts*100:5.1f}% of samples  {i/ti*100:5.1f}% of warp instructions  avg threads {t/max(i,1):4.1f}  {n}")
