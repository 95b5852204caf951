cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 900 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"; cat gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep|prepare|sums|backward_dot' -c 60 --csv --log-file gpurun_out/launches_r1_c2_em.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c2.json 2>&1; cat gpurun_out/bench_ref_c2.json
