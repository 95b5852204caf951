#!/bin/bash
# kernel version 2 (VIPRS_B200_FAST=2) against version 1 with the round's final build
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02p_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02p_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c2 ln small c1; do
  run v1 $wl VIPRS_B200_FAST=1
  run v2 $wl VIPRS_B200_FAST=2
  run v2rot0 $wl VIPRS_B200_FAST=2 VIPRS_B200_SMSP_ROT=0
done
