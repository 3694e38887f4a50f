#!/bin/bash
# usage: tools/ncu_kernels.sh <tag>  -- basic per-launch metrics of the eikonal kernels of one bench run (csv in gpurun_out/)
tag=$1
envs=$2   # optional: environment assignments for the bench, e.g. MCMCEQ_EIKONAL_PIPE=1
timeout 3000 tools/gpurun_retry.sh 600 "ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:eik_ -c 12 --csv --log-file gpurun_out/kern_$tag.csv env $envs python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/kern_$tag.log 2>&1" 2>&1 | tail -3
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/kern_$tag.csv')) if len(r)>10 and r[0].isdigit()]
out={}
for r in rows:
    out.setdefault((int(r[0]), r[4].split('(')[0][-28:]), {})[r[12]] = r[14]
for (i,k),m in sorted(out.items()):
    if i>=8: print(i,k,{a.split('.')[0].replace('smsp__','').replace('launch__',''):b for a,b in m.items()})
PY
