cd $GRAFT_REPO_ROOT
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r01b_bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/r01b_bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01b_bench_reference_c2.json 2>/dev/null; cut -c1-400 gpurun_out/r01b_bench_reference_c2.json
for w in c4 c1; do
timeout 600 python bench.py --workload $w > gpurun_out/r01b_bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/r01b_bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01b_c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -c sweep gpurun_out/r01b_c2_launches.csv
