#!/bin/bash
# placement of the bulk warps' tile shares on the SM sub-partitions (VIPRS_B200_SMSP_ROT 0 / 1 / 2), chain warps of the
# two co-resident CTAs on different sub-partitions (-DVB_CHAIN_SPLIT=148), phase trace of the Latin-square placement
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -2
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02x_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02x_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c2 c4 ln; do
  run rot0 $wl VIPRS_B200_SMSP_ROT=0
  run rot1 $wl VIPRS_B200_SMSP_ROT=1
  run rot2 $wl VIPRS_B200_SMSP_ROT=2
  run rot2split $wl VIPRS_B200_SMSP_ROT=2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_split/libviprs_b200.so
done
run rot0 c1 VIPRS_B200_SMSP_ROT=0
run rot2 c1 VIPRS_B200_SMSP_ROT=2
VIPRS_B200_LIB=$PWD/viprs_b200/_C_trace/libviprs_b200.so VIPRS_B200_TRACE=gpurun_out/trace.bin \
    timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
python scripts/trace_report.py gpurun_out/trace.bin 100 103 > gpurun_out/r02x_c2_rot2_trace.txt 2>&1
rm -f gpurun_out/trace.bin
