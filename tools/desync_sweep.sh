#!/bin/bash
# lock-step against desynchronised stepping for a few proposal strings (run on the GPU box): prints one line per run
run() {  # string iters chains
  for ds in 0 1; do
    MCMCEQ_DESYNC=$ds python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chains $3 --proposals $1 --iters-per-step $2 > gpurun_out/ds_tmp.log 2>&1
    python - "$1" "$2" "$3" "$ds" <<'PY'
import json, sys
s, it, ch, ds = sys.argv[1:5]
try:
    d = json.loads(open('gpurun_out/ds_tmp.log').read().strip().splitlines()[-1])
    print(s[:8], 'iters', it, 'chains', ch, 'desync', ds, 'value', round(d['value']), 'ms/iter', round(d['ms_per_step'] / int(it), 3), 'launches', d['gpu_launches'])
except Exception as e:
    print('FAILED', s, it, ch, ds, e, open('gpurun_out/ds_tmp.log').read()[-300:])
PY
  done
}
MIX=QQQQQQQQQQQVRRRRRRRPBDMN
run QN 48 1024
run QN 48 10
run P 8 1024
run $MIX 480 1024
run $MIX 480 4096
