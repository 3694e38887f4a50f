#!/bin/bash
# small launches of the fused eikonal kernel: solves per warp-task fixed at 32 against chosen per launch (run on the GPU box)
MIX=QQQQQQQQQQQVRRRRRRRPBDMN
for ch in 10 64 256; do
  for lanes in 32 0; do
    if [ $lanes == 0 ]; then unset MCMCEQ_EIKONAL_LANES; else export MCMCEQ_EIKONAL_LANES=$lanes; fi
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chains $ch --proposals $MIX --iters-per-step 48 > gpurun_out/lanes_tmp.log 2>&1
    python tools/show_bench.py gpurun_out/lanes_tmp.log chains=$ch lanes=$lanes mix
  done
done
for lanes in 32 0; do
  if [ $lanes == 0 ]; then unset MCMCEQ_EIKONAL_LANES; else export MCMCEQ_EIKONAL_LANES=$lanes; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 10 --proposals P > gpurun_out/lanes_tmp.log 2>&1
  python tools/show_bench.py gpurun_out/lanes_tmp.log chains=10 lanes=$lanes P_full
done
