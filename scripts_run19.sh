cd $GRAFT_REPO_ROOT
( timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
for w in c2 c4; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
export VIPRS_B200_LIB=$GRAFT_REPO_ROOT/viprs_b200/_C_trace/libviprs_b200.so
VIPRS_B200_TRACE=gpurun_out/trace.bin timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-20
python scripts_trace.py gpurun_out/trace.bin 100 103 > gpurun_out/trace_report.txt 2>&1
cat gpurun_out/trace_report.txt
rm -f gpurun_out/trace.bin
