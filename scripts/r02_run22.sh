#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "incremental or resident" ) > gpurun_out/r02k_pytest.log 2>&1; tail -15 gpurun_out/r02k_pytest.log
timeout 600 python bench.py --no-extras --steps 100 > gpurun_out/r02k_bench_c2.json 2> gpurun_out/r02k_bench_c2.err; tail -3 gpurun_out/r02k_bench_c2.err
python -c "
import json
d=json.load(open('gpurun_out/r02k_bench_c2.json'))
print('c2 value %.4g ms/step %.4f kernel %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms']), d['e2e'], d['cpu_baseline'])"
VIPRS_B200_NO_INCREMENTAL=1 timeout 600 python bench.py --no-extras --steps 100 --no-cpu-baseline > gpurun_out/r02k_bench_c2_noincr.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02k_bench_c2_noincr.json'))
print('no-incremental e2e', d['e2e']['ms_per_step'])"
