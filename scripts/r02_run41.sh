#!/bin/bash
# row_dot_kernel with 4-row groups per warp (x read once per block vector): tests, e2e timeline, c2 / c4 / c5 lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02z_pytest.log 2>&1; tail -n 3 gpurun_out/r02z_pytest.log
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -4
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02z_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02z_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms e2e %.4f ms'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['e2e']['ms_per_step']))"
}
run rg4 c2 X=1
run rg4 c4 X=1
run rg4 c5 X=1
