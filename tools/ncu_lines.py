#!/usr/bin/env python
"""Per-source-line samples/instructions from `ncu --page source --csv --print-source cuda,sass` output."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    elif len(r) > 8 and r[0] == 'Line No': hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try: agg.append((cur, int(r[0]), r[1].strip(), float(r[6] or 0), float(r[7] or 0), float(r[10] or 0)))
        except ValueError: pass
ts = sum(a[3] for a in agg); ti = sum(a[4] for a in agg)
print('samples %d warp-instructions %.3g' % (ts, ti))
for f, l, src, smp, ins, thr in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{smp/ts*100:5.1f}% inst={ins/ti*100:5.1f}% {f}:{l:<4d} {src[:95]}")
