#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_floor.json
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02a_pytest.log 2>&1; tail -6 gpurun_out/r02a_pytest.log
python scripts/dbg_tiled.py 2>&1 | head -8
for L in 3 4; do
  VIPRS_B200_LIMBS=$L timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r02a_c2_limbs$L.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02a_c2_limbs$L.json'));print('limbs $L ms/step %.4f kernel %.4f e2e %.3f ms'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']))"
done
