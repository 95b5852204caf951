#!/bin/bash
# usage: scripts_tune.sh "ENV1=a ENV2=b" ...   -> one bench line (ms_per_step, frac) per setting
mkdir -p gpurun_out
for cfg in "$@"; do
  out=$(env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms=%.3f frac=%.3f'%(d['ms_per_step'], d['roofline']['frac']))")
  echo "$cfg -> $out" | tee -a gpurun_out/tune.log
done
