#!/bin/bash
# last verification of the round with the final code: GPU tests, smoke, default bench, reference arm, c4 / c1 / ln lines
TAG=r02g
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; cut -c1-160 gpurun_out/${TAG}_bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_c2.json 2>/dev/null
for w in c4 c1 ln c3; do timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -n 4 > gpurun_out/${TAG}_c2_e2e_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/${TAG}_c4_launches.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
ls gpurun_out/${TAG}_* | wc -l
