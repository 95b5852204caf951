#!/bin/bash
# int16 / float LD (more than two tiles per thread): dead tiles skipped per panel (default since; was -DVB_LIVE_ALL, now -DVB_NO_LIVE_ALL turns it off) vs the general form for every tile
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( VIPRS_B200_LIB=$PWD/viprs_b200/_C_live/libviprs_b200.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02u_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02u_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
}
for wl in c1 c4; do
  run base $wl X=1
  run live $wl VIPRS_B200_LIB=$PWD/viprs_b200/_C_live/libviprs_b200.so
done
# host-state round trip: outputs other than q downloaded next to the update_q_factor pass
( timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "resident" ) 2>&1 | tail -n 1
for i in 1 2; do
timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02u_c2_e2e$i.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02u_c2_e2e$i.json'));print('c2 e2e %.4f ms'%(d['e2e']['ms_per_step']))"
done
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -n 4
