#!/bin/bash
# One gpurun call that reproduces the round's single-GPU evidence on a B200 box (run from the repo root of the snapshot):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round_check.sh r02'
# GPU tests, smoke, the default bench (c2 + c3/c4/c1 sub-measurements) with every leg, the reference arm, the other
# workloads, the ncu launch lists of the same commands and one ncu --set full capture of the c2 sweep and the c3 grid sweep.
TAG=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; cut -c1-200 gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_c2.json 2>/dev/null
for w in c3 c4 c1 ln c5; do timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/${TAG}_c2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
for w in c3 c4; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/${TAG}_${w}_launches.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
done
timeout 800 ncu --set full --import-source on --clock-control none -k regex:'sweep_fast' -s 3 -c 1 -o gpurun_out/${TAG}_c2_fast_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'grid_sweep' -s 2 -c 1 -o gpurun_out/${TAG}_c3_grid_full -f \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
