cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'])"
timeout 900 python bench.py --workload c3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['fp32_pipe_frac'])"
python - <<'PY'
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
wl = bench.WORKLOADS["small"]
m, M, nb = bench.build_model(wl, 0, 1)
m.initialize({"pi": 0.01, "sigma_epsilon": 0.8})
for _ in range(5): m.e_step(); m.m_step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): m.e_step(); m.m_step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:4000])
PY
