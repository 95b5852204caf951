#!/bin/bash
# what-if: N never-executed instructions in the middle of the A row-group loop (-DVB_WHATIF_PAD=N): sensitivity of the
# sweep to the size of a role's loop body (instruction supply)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02s_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02s_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
run base c2 X=1
run pad64 c2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_pad64/libviprs_b200.so
run pad256 c2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_pad256/libviprs_b200.so
run base small X=1
run pad256 small VIPRS_B200_LIB=$PWD/viprs_b200/_C_pad256/libviprs_b200.so
