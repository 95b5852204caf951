#!/bin/bash
# multi-GPU: strong scaling of the fixed genome (N given as $1)
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r02c_bench_c2_n$N.json 2> gpurun_out/r02c_bench_c2_n$N.err; tail -5 gpurun_out/r02c_bench_c2_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02c_bench_c2_n$N.json'))
print('N=$N c2 value %.4g ms/step %.4f kernel %.4f frac %.3f e2e %s'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'], d.get('e2e',{}).get('ms_per_step')))
for k,v in d.get('workloads',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload','kernel')})
PY
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3; fi
