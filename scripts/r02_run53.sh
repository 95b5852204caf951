#!/bin/bash
# row chunks of the host-state round trip with the final code of the round
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for ch in 4 5 6 4 6; do
  VIPRS_B200_CHUNKS=$ch timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02n_c2_ch$ch.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02n_c2_ch$ch.json'));print('c2 chunks $ch e2e %.4f ms'%(d['e2e']['ms_per_step']))"
done
VIPRS_B200_CHUNKS=6 timeout 300 python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 30 > gpurun_out/r02n_c4_ch6.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02n_c4_ch6.json'));print('c4 chunks 6 e2e %.4f ms'%(d['e2e']['ms_per_step']))"
