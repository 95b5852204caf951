cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for w in c2 c4 c1; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
done
VIPRS_B200_LIB=$GRAFT_REPO_ROOT/viprs_b200/_C_nobal/libviprs_b200.so timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_nobal.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_nobal.json')); print('c2 nobal', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
