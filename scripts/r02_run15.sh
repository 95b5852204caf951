#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -k "grid or c3" ) > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log
timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-e2e --no-extras --steps 5 > gpurun_out/r02e_bench_c3.json 2> gpurun_out/r02e_bench_c3.err; tail -3 gpurun_out/r02e_bench_c3.err
python -c "
import json
d=json.load(open('gpurun_out/r02e_bench_c3.json'))
print('c3 value %.4g ms/step %.4f kernel %.4f pipe %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['fp32_pipe_frac']))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv --log-file gpurun_out/r02e_c3_launches.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02e_c3_launches.csv')) if len(r)>=15 and r[0].isdigit()]
for r in rows[-24:]:
    if r[12]=='gpu__time_duration.sum': print(r[4][:50], r[14])
PY
