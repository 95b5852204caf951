#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02f_pytest.log 2>&1; tail -5 gpurun_out/r02f_pytest.log
run() { # name env...
  name=$1; shift
  for W in c2 small; do
    env "$@" timeout 300 python bench.py --workload $W --no-extras --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/whatif_${name}_$W.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/whatif_${name}_$W.json'));print('$name $W kernel %.4f ms'%(d['roofline']['kernel_ms']))"
  done
}
run base X=1
run skipA VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipA/libviprs_b200.so
run skipC VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipC/libviprs_b200.so
run skipAC VIPRS_B200_LIB=$PWD/viprs_b200/_C_skipAC/libviprs_b200.so
run stage24k VIPRS_B200_STAGE_BYTES=23552
run stage16k VIPRS_B200_STAGE_BYTES=16384
