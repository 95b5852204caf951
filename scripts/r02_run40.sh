#!/bin/bash
# per-kernel durations inside the host-state round trip (serialised under ncu: each chunk's kernels alone on the GPU)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
VIPRS_B200_NO_ZERO_COPY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'sweep_fast|row_dot|host_copy' -c 120 --csv --log-file gpurun_out/r02z_e2e_launches.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02z_e2e_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-24:]:
    print(r[4][:70], r[8], r[-1])
PY
