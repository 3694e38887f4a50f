#!/bin/bash
# usage: tools/prof_pipe.sh <tag>   -- launch list + full capture of eik_pipe_kernel under bench.py (config 3), each after the
# same command has exited 0 without ncu (run on the GPU box through gpurun); summaries with tools/ncu_summary.py / ncu_funcs.py
tag=$1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --parity-chains 2"
$B > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eik_pipe -s 4 -c 1 -o gpurun_out/prof_${tag}pipe -f $B > gpurun_out/${tag}_ncu_pipe.log 2>&1
tail -n 1 gpurun_out/${tag}_plain.log | cut -c1-200
