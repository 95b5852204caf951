#!/bin/bash
# round 2, GPU call 1: all GPU tests (incl. tiled / q-in / BASELINE-parameter parity), new bench lines, grid kernel source profile
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_floor.json
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02a_pytest.log 2>&1; tail -15 gpurun_out/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; cut -c1-600 gpurun_out/r02a_bench_c2.json; tail -3 gpurun_out/r02a_bench_c2.err
timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err; cut -c1-400 gpurun_out/r02a_bench_c5.json; tail -3 gpurun_out/r02a_bench_c5.err
timeout 600 python bench.py --workload ln --no-cpu-baseline > gpurun_out/r02a_bench_ln.json 2> gpurun_out/r02a_bench_ln.err; cut -c1-400 gpurun_out/r02a_bench_ln.json; tail -3 gpurun_out/r02a_bench_ln.err
timeout 700 ncu --set full --import-source on --clock-control none -k regex:'grid_sweep' -s 4 -c 1 -o gpurun_out/r02a_c3_grid_full -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r02a_ncu_grid.err
ls -la gpurun_out | tail -12
