cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
for w in c2 c4 c3; do
timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('fp32_pipe_frac'), d['e2e']['value'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'sweep' -s 3 -c 2 --csv --log-file gpurun_out/launches_r1_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep' -s 3 -c 2 --csv --log-file gpurun_out/launches_r1_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -h "sweep" gpurun_out/launches_r1_c3.csv gpurun_out/launches_r1_c4.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-160
