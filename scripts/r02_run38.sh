#!/bin/bash
# host-state round trip with kernel-driven ("zero copy") transfers of page-locked state arrays vs DMA copies
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "resident or incremental" ) 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --no-extras --no-cpu-baseline --steps 50 > gpurun_out/r02z_${wl}_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02z_${wl}_${name}.json'));print('$wl $name sweep %.4f ms step %.4f ms e2e %.4f ms'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['e2e']['ms_per_step']))"
}
run dma c2 VIPRS_B200_NO_ZERO_COPY=1
run zc4 c2 X=1
run zc2 c2 VIPRS_B200_CHUNKS=2
run zc6 c2 VIPRS_B200_CHUNKS=6
run zc8 c2 VIPRS_B200_CHUNKS=8
run dma c4 VIPRS_B200_NO_ZERO_COPY=1
run zc4 c4 X=1
run zc8 c4 VIPRS_B200_CHUNKS=8
