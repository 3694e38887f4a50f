#!/bin/bash
# usage: tools/ab_variants.sh <tag> <variant> ...   -- on the GPU box: for every variant library built by tools/ab_build.py
# ("main" = the in-tree build): pipelined vs fused kernel bit for bit (tools/ab_pipe.py), then the bench at 1024 and 8192 chains
tag=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset MCMCEQ_LIB; else export MCMCEQ_LIB=mcmc_eq_b200/libmcmceq_b200_$v.so; fi
  echo "== $v" >> gpurun_out/${tag}.log
  python tools/ab_pipe.py 2>&1 | tail -n 3 >> gpurun_out/${tag}.log
  python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python tools/show_bench.py /dev/stdin >> gpurun_out/${tag}.log
  python bench.py --steps 6 --warmup 2 --chains 8192 --no-extras --no-cpu-baseline 2>/dev/null | python tools/show_bench.py /dev/stdin >> gpurun_out/${tag}.log
done
cat gpurun_out/${tag}.log
