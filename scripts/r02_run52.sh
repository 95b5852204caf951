#!/bin/bash
# counter polls (wait_progress / wait_ge) with the spin bound checked once per 256 polls vs on every poll: no effect
# (c2 0.9858 vs 0.9857 ms), not kept
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02o_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02o_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c2 c4 small c1 c5; do
  run new $wl X=1
  run prev $wl VIPRS_B200_LIB=$PWD/viprs_b200/_C_prev/libviprs_b200.so
done
