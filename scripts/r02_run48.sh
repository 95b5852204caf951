#!/bin/bash
# ncu full capture of the sums / prepare kernels at G = 256 (c3): which pipe bounds them
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'sums_kernel|prepare_kernel' -s 2 -c 2 -o gpurun_out/r02_c3_em_full -f \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
ls -la gpurun_out/r02_c3_em_full.ncu-rep
