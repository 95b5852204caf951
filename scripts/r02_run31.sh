#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --steps 8 > gpurun_out/r02u_c5_$name.json 2>gpurun_out/r02u_c5_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r02u_c5_$name.json'));print('c5 $name sweep %.3f ms step %.3f ms'%(d['roofline']['kernel_ms'], d['ms_per_step']))" || tail -3 gpurun_out/r02u_c5_$name.err
}
run base X=1
run f2 VIPRS_B200_LIB=$PWD/viprs_b200/_C_f2/libviprs_b200.so
run f4 VIPRS_B200_LIB=$PWD/viprs_b200/_C_f4/libviprs_b200.so
run f2noside VIPRS_B200_LIB=$PWD/viprs_b200/_C_f2/libviprs_b200.so VIPRS_B200_NO_SIDE_STREAM=1
