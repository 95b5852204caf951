#!/bin/bash
# row chunks of the host-state round trip after the faster update_q_factor pass; c5 back on the row-at-a-time form
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -n 2
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02z_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02z_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms e2e %.4f ms'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['e2e']['ms_per_step']))"
}
run ch3 c2 VIPRS_B200_CHUNKS=3
run ch4 c2 VIPRS_B200_CHUNKS=4
run ch6 c2 VIPRS_B200_CHUNKS=6
run ch8 c2 VIPRS_B200_CHUNKS=8
run ch8 c4 VIPRS_B200_CHUNKS=8
run rg1 c5 X=1
