#!/bin/bash
# A role: software-pipelined row-group loop for regular panels vs the plain loop (experiment of the commit history only:
# measured 1.097 vs 0.987 ms on c2 -- slower, code removed; -DVB_NO_PIPELINED_A no longer exists)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02t_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02t_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c2 ln small; do
  run pipe $wl X=1
  run plain $wl VIPRS_B200_LIB=$PWD/viprs_b200/_C_nopipe/libviprs_b200.so
done
