#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "version_2" ) 2>&1 | tail -3
run() { name=$1; shift
  for W in c2 small; do
    env "$@" timeout 300 python bench.py --workload $W --no-extras --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/whatif_${name}_$W.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/whatif_${name}_$W.json'));print('$name $W kernel %.4f ms'%(d['roofline']['kernel_ms']))"
  done
}
run v2base VIPRS_B200_FAST=2
run v2skipA VIPRS_B200_FAST=2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipA/libviprs_b200.so
run v2skipC VIPRS_B200_FAST=2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipC/libviprs_b200.so
run v2skipAC VIPRS_B200_FAST=2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipAC/libviprs_b200.so
