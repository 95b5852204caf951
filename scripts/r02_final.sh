#!/bin/bash
# final single-GPU evidence of the round: tests, smoke, default bench (all legs), reference arm, ln / c5 lines, the c2 launch
# list and full capture of the sweep kernel, micro-benchmarks quoted in DESIGN.md
TAG=r02f
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; cut -c1-160 gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_c2.json 2>/dev/null
for w in ln c5; do timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/${TAG}_c2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
timeout 800 ncu --set full --import-source on --clock-control none -k regex:'sweep_fast' -s 3 -c 1 -o gpurun_out/${TAG}_c2_fast_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -n 4 > gpurun_out/${TAG}_c2_e2e_timeline.txt
./scratch/idp_bench > gpurun_out/${TAG}_idp_bench.txt 2>&1
./scratch/h2d_bench > gpurun_out/${TAG}_h2d_bench.txt 2>&1
./scratch/grid_loop_bench > gpurun_out/${TAG}_grid_loop_bench.txt 2>&1
python scripts/pcie_probe.py > gpurun_out/${TAG}_pcie_probe.txt 2>&1
ls -la gpurun_out/${TAG}_*
