#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for n in 1 2 3 4 6 8; do
  VIPRS_B200_CHUNKS=$n timeout 300 python bench.py --no-extras --steps 50 --no-cpu-baseline > gpurun_out/r02m_chunks_$n.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02m_chunks_$n.json'));print('chunks $n e2e %.4f ms'%d['e2e']['ms_per_step'])"
done
