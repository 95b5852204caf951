#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=r02
for w in c2 c3 c4; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'sweep|prepare_kernel|sums_kernel|em_update|row_dot|forward_axpy' -c 24 --csv --log-file gpurun_out/${TAG}_${w}_launches.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
done
# chain-warp placement experiment (second CTA of an SM swaps its chain / producer warps)
for v in base split; do
  if [ $v = split ]; then export VIPRS_B200_LIB=$PWD/viprs_b200/_C_split/libviprs_b200.so; fi
  timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02p_chain_$v.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02p_chain_$v.json'));print('$v sweep %.4f ms'%d['roofline']['kernel_ms'])"
done
