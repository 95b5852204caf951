#!/bin/bash
# sums_kernel with the next row's operands requested before the current row's arithmetic
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_model_gpu.py tests/test_round2_gpu.py tests/test_grid_gpu.py -m gpu -q -x ) 2>&1 | tail -n 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sums_kernel|prepare_kernel' -c 8 --csv --log-file gpurun_out/r02r_c3_em.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
grep -o 'sums_kernel.*\|prepare_kernel.*' gpurun_out/r02r_c3_em.csv | awk -F'","' '{print substr($1,1,16), $NF}' | tail -n 4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sums_kernel|prepare_kernel' -c 8 --csv --log-file gpurun_out/r02r_c2_em.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
grep -o 'sums_kernel.*\|prepare_kernel.*' gpurun_out/r02r_c2_em.csv | awk -F'","' '{print substr($1,1,16), $NF}' | tail -n 4
for wl in c2 c3; do
timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02r_$wl.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r02r_$wl.json'));print('$wl sweep %.4f ms step %.4f ms'%(d['roofline']['kernel_ms'], d['ms_per_step']))"
done
