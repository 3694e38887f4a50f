#!/bin/bash
# Round-2 profile captures, each after the same command has exited 0 without ncu (run on the GPU box through gpurun):
#   r2 launch list + full capture of eik_pipe_kernel under bench.py (config 3), full capture of misfit_kernel at the
#   config-4 shard shape, full capture of eik_fine_kernel on the 565 x 2001 plane.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --parity-chains 2"
$B > gpurun_out/r2_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eik_pipe -s 4 -c 1 -o gpurun_out/prof_r2pipe -f $B > gpurun_out/r2_ncu_pipe.log 2>&1
python tools/misfit_probe.py > gpurun_out/r2_misfit_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:misfit_kernel -s 8 -c 1 -o gpurun_out/prof_r2misfit -f python tools/misfit_probe.py > gpurun_out/r2_ncu_misfit.log 2>&1
python tools/fine_probe.py 4 > gpurun_out/r2_fine_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:eik_fine -s 1 -c 1 -o gpurun_out/prof_r2fine -f python tools/fine_probe.py 4 > gpurun_out/r2_ncu_fine.log 2>&1
tail -1 gpurun_out/r2_plain.log | cut -c1-200; cat gpurun_out/r2_misfit_plain.log gpurun_out/r2_fine_plain.log
