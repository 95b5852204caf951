#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x -k "grid or c3 or Grid" ) > gpurun_out/r02o_pytest.log 2>&1; tail -5 gpurun_out/r02o_pytest.log
timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-e2e --no-extras --steps 5 > gpurun_out/r02o_bench_c3.json 2> gpurun_out/r02o_bench_c3.err; tail -3 gpurun_out/r02o_bench_c3.err
python -c "
import json
d=json.load(open('gpurun_out/r02o_bench_c3.json'))
print('c3 value %.4g ms/step %.4f kernel %.4f pipe %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['fp32_pipe_frac']))"
