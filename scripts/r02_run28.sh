#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "tiled or c5 or banded or incoming or param_0" ) 2>&1 | tail -3
for t in 2048 1024 512; do
  VIPRS_B200_TILE_ROWS=$t timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/r02r_c5_tile$t.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02r_c5_tile$t.json'));print('c5 tile $t sweep %.3f ms step %.3f ms'%(d['roofline']['kernel_ms'], d['ms_per_step']))"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 40 --csv --log-file gpurun_out/r02_c5_launches.csv \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
