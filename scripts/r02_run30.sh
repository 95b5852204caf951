#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "tiled or c5 or banded or incoming or param_0 or non_block" ) 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --steps 8 $EXTRA > gpurun_out/r02t_c5_$name.json 2>gpurun_out/r02t_c5_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r02t_c5_$name.json'));print('c5 $name sweep %.3f ms step %.3f ms'%(d['roofline']['kernel_ms'], d['ms_per_step']))" || tail -3 gpurun_out/r02t_c5_$name.err
}
run noside VIPRS_B200_NO_SIDE_STREAM=1
run side X=1
EXTRA=--graph run sidegraph X=1
