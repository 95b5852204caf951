#!/bin/bash
# tiled sweep: far part of the forward products on a second side stream vs all of it on the caller's stream
# (experiment of commit history only: measured 13.33 vs 13.54 ms on c5, no gain, code removed -- VIPRS_B200_NO_FWD_SPLIT no longer exists)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q -x ) 2>&1 | tail -n 2
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 30 > gpurun_out/r02v_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02v_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
run split c5 X=1
run nosplit c5 VIPRS_B200_NO_FWD_SPLIT=1
run split512 c5 VIPRS_B200_TILE_ROWS=512
run split2048 c5 VIPRS_B200_TILE_ROWS=2048
run split_graph c5 X=1
