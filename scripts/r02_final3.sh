#!/bin/bash
# launch lists of the c2 / c3 EM step with the final code (sums_kernel changed after r02_final2.sh), final default bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for w in c2 c3; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/r02h_${w}_launches.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
done
timeout 900 python bench.py > gpurun_out/r02h_bench_c2.json 2> gpurun_out/r02h_bench_c2.err; cut -c1-160 gpurun_out/r02h_bench_c2.json
timeout 900 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02h_bench_c3.json 2>/dev/null
( timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -n 1
