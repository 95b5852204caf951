#!/bin/bash
# 2-GPU strong scaling of the fixed genome (c2 headline + c3 / c4 / c1 sub-measurements + multi_gpu_parity), and the 2-GPU test
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err
tail -2 gpurun_out/r02_bench_c2_n$N.err; cut -c1-300 gpurun_out/r02_bench_c2_n$N.json
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q ) 2>&1 | tail -2; fi
