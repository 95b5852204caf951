#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02d_pytest.log 2>&1; tail -12 gpurun_out/r02d_pytest.log
for F in 1 0; do
VIPRS_B200_FUSED_SUMS=$F timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 200 > gpurun_out/r02d_bench_c2_fused$F.json 2> gpurun_out/r02d_bench_c2.err; tail -3 gpurun_out/r02d_bench_c2.err
python -c "
import json
d=json.load(open('gpurun_out/r02d_bench_c2_fused$F.json'))
print('fused=$F c2 value %.4g ms/step %.4f kernel %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
done
