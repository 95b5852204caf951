cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r01b_bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo rc=$?; cut -c1-700 gpurun_out/r01b_bench_c2_n2.json; tail -3 gpurun_out/bench_c2_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
