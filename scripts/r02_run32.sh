#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02v_pytest.log 2>&1; tail -3 gpurun_out/r02v_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02_bench_c5.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02_bench_c5.json'));print('c5 sweep %.3f ms step %.3f ms frac %.3f e2e %s'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']))"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 48 --csv --log-file gpurun_out/r02_c5_launches.csv \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
