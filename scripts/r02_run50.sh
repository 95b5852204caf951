#!/bin/bash
# chain -> C hand-off per group of four rows vs per panel (experiment of the commit history only: measured 1.093 vs
# 0.986 ms on c2 -- slower, code removed; -DVB_NO_FINE_C no longer exists)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py tests/test_full_size_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r02q_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02q_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f e2e %s'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'] if d.get('e2e') else None))"
}
for wl in c2 c4 ln small c1; do
  run fine $wl X=1
  run panel $wl VIPRS_B200_LIB=$PWD/viprs_b200/_C_nofine/libviprs_b200.so
done
