#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-e2e --no-extras --steps 5 > gpurun_out/r02w_c3_$name.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02w_c3_$name.json'));print('c3 $name sweep %.3f ms pipe %.3f'%(d['roofline']['kernel_ms'], d['roofline']['fp32_pipe_frac']))"
}
run skew0 VIPRS_B200_LIB=$PWD/viprs_b200/_C_skew0/libviprs_b200.so
run skew1400 X=1
run skew2800 VIPRS_B200_LIB=$PWD/viprs_b200/_C_skew2/libviprs_b200.so
