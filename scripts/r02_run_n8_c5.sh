#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --workload c5 --no-e2e > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err
tail -1 gpurun_out/r02_bench_c5_n$N.json | cut -c1-300
