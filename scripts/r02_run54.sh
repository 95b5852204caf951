#!/bin/bash
# host-state round trip: one cudaMemcpyBatchAsync per chunk for the uploads (default) vs eight cudaMemcpyAsync (VIPRS_B200_NO_BATCH_COPY=1)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "resident" ) 2>&1 | tail -n 1
for v in "X=1" "VIPRS_B200_NO_BATCH_COPY=1" "X=1" "VIPRS_B200_NO_BATCH_COPY=1"; do
  env $v timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02m_c2.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02m_c2.json'));print('c2 $v e2e %.4f ms'%(d['e2e']['ms_per_step']))"
done
for v in "X=1" "VIPRS_B200_NO_BATCH_COPY=1"; do
  env $v timeout 300 python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 30 > gpurun_out/r02m_c4.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02m_c4.json'));print('c4 $v e2e %.4f ms'%(d['e2e']['ms_per_step']))"
done
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -n 4
