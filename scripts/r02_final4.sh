#!/bin/bash
# default bench line + e2e timeline + resident tests with the final code of the round
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02i_bench_c2.json 2> gpurun_out/r02i_bench_c2.err; cut -c1-160 gpurun_out/r02i_bench_c2.json
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -n 4 > gpurun_out/r02i_c2_e2e_timeline.txt
( timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -n 1
