#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'grid_sweep' -s 2 -c 1 -o gpurun_out/r02h_c3_grid_full -f \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2> gpurun_out/r02h_ncu.err
tail -3 gpurun_out/r02h_ncu.err; ls -la gpurun_out/r02h_c3_grid_full.ncu-rep
