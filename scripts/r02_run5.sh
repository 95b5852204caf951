#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02b_pytest.log 2>&1; tail -8 gpurun_out/r02b_pytest.log
for V in 1 2; do
  VIPRS_B200_FAST=$V timeout 300 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02b_c2_fast$V.json 2>gpurun_out/r02b_c2_fast$V.err
  python -c "import json;d=json.load(open('gpurun_out/r02b_c2_fast$V.json'));print('FAST=$V c2 ms/step %.4f kernel %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))" || tail -3 gpurun_out/r02b_c2_fast$V.err
  for W in c4 c1 ln small; do
    VIPRS_B200_FAST=$V timeout 300 python bench.py --workload $W --no-extras --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/r02b_${W}_fast$V.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02b_${W}_fast$V.json'));print('FAST=$V $W ms/step %.4f kernel %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
  done
done
