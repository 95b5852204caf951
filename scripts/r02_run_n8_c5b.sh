#!/bin/bash
# BASELINE configs[4] on 8 GPUs with the final code of the round (586 float64 blocks of 10,240 SNPs, one all-reduce per iteration)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 8 --workload c5 --no-e2e > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
tail -n 2 gpurun_out/r02_bench_c5_n8.err; tail -n 1 gpurun_out/r02_bench_c5_n8.json | cut -c1-260
