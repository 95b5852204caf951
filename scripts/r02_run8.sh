#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_model_gpu.py -m gpu -q -k "device or model or history" ) > gpurun_out/r02c_pytest.log 2>&1; tail -25 gpurun_out/r02c_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err; tail -3 gpurun_out/r02c_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_c2.json'))
print('c2 value %.4g ms/step %.4f kernel %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))
for k,v in d.get('workloads',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload','kernel')})
PY
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extras --graph > gpurun_out/r02c_bench_c2_graph.json 2> gpurun_out/r02c_bench_c2_graph.err; tail -3 gpurun_out/r02c_bench_c2_graph.err
python -c "
import json
d=json.load(open('gpurun_out/r02c_bench_c2_graph.json'))
print('c2 graph value %.4g ms/step %.4f kernel %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms']))"
