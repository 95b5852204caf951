#!/bin/bash
# int16 / float LD (more than two tiles per thread): dead tiles skipped per panel (-DVB_LIVE_ALL) vs the general form for every tile
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( VIPRS_B200_LIB=$PWD/viprs_b200/_C_live/libviprs_b200.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02u_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02u_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c1 c4; do
  run base $wl X=1
  run live $wl VIPRS_B200_LIB=$PWD/viprs_b200/_C_live/libviprs_b200.so
done
