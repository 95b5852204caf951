#!/bin/bash
# warp-id order of the roles of sweep_fast_kernel (issue arbitration favours some warp ids): A/C/chain/producer permutations
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02y_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02y_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for v in C C_v1 C_v2 C_v3 C_v4; do
  ( VIPRS_B200_LIB=$PWD/viprs_b200/_$v/libviprs_b200.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -1
  for rot in 0 1; do
    run ${v}_rot$rot c2 VIPRS_B200_SMSP_ROT=$rot VIPRS_B200_LIB=$PWD/viprs_b200/_$v/libviprs_b200.so
  done
  run ${v}_rot1 c4 VIPRS_B200_SMSP_ROT=1 VIPRS_B200_LIB=$PWD/viprs_b200/_$v/libviprs_b200.so
  run ${v}_rot1 ln VIPRS_B200_SMSP_ROT=1 VIPRS_B200_LIB=$PWD/viprs_b200/_$v/libviprs_b200.so
done
