#!/bin/bash
# blocked decomposition on the C2 shape: LD blocks of 4096 swept as 2 / 4 diagonal tiles + streaming rectangle products
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( VIPRS_B200_TILE_LIMIT=255 VIPRS_B200_TILE_ROWS=256 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x ) 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/r02j_tiles_${name}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02j_tiles_${name}.json'));print('$name sweep %.4f ms step %.4f ms'%(d['roofline']['kernel_ms'], d['ms_per_step']))"
}
run whole X=1
run two VIPRS_B200_TILE_LIMIT=4095 VIPRS_B200_TILE_ROWS=2048
run four VIPRS_B200_TILE_LIMIT=4095 VIPRS_B200_TILE_ROWS=1024
run eight VIPRS_B200_TILE_LIMIT=4095 VIPRS_B200_TILE_ROWS=512
