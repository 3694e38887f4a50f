#!/bin/bash
# usage: tools/prof_cycle.sh <tag> [n_lines] [--tests] [--launches]
#   plain bench, then ncu full capture of the eikonal kernel under bench.py, then line-level summary; --tests runs pytest -m gpu first,
#   --launches adds the ncu launch list (gpu__time_duration.sum of every kernel) of the same command
tag=$1
pre=""; post=""
for a in "$@"; do
  if [ "$a" == "--tests" ]; then pre="python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log;"; fi
  if [ "$a" == "--launches" ]; then post="; ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1"; fi
done
timeout 4000 tools/gpurun_retry.sh 1500 "$pre python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eik_pipe -s 4 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1 $post; ls gpurun_out | tail -1" 2>&1 | tail -6
tail -1 gpurun_out/plain_$tag.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms/step', d['ms_per_step'], 'eik ms', d['roofline']['avg_launch_ms'], 'clocks', d['clocks'])"
python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep 2>/dev/null | head -12
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/src_${tag}_cs.csv
python tools/ncu_funcs.py /tmp/src_${tag}_cs.csv 18
python tools/ncu_lines.py /tmp/src_${tag}_cs.csv ${2:-25}
