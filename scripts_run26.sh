cd $GRAFT_REPO_ROOT
( timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -2
for w in c2 c4 c1; do
timeout 300 python bench.py --workload $w --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('$w', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
