#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/r02c_bench_c2_n$N.json 2> gpurun_out/r02c_bench_c2_n$N.err; grep -v Warn gpurun_out/r02c_bench_c2_n$N.err | tail -3
timeout 900 $TR bench.py --gpus $N --workload c5 --no-cpu-baseline --no-e2e > gpurun_out/r02c_bench_c5_n$N.json 2> gpurun_out/r02c_bench_c5_n$N.err; grep -v Warn gpurun_out/r02c_bench_c5_n$N.err | tail -3
python - <<PY
import json
for w in ("c2","c5"):
    txt=open('gpurun_out/r02c_bench_%s_n$N.json'%w).read()
    line=[l for l in txt.splitlines() if l.startswith('{')][-1]
    d=json.loads(line)
    print('N=$N',w,'value %.4g ms/step %.4f kernel %.4f frac %.3f e2e %s'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'], (d.get('e2e') or {}).get('ms_per_step')))
    for k,v in d.get('workloads',{}).items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload','kernel')})
PY
