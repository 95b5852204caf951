cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for cfg in "X=1" "VIPRS_B200_STAGE_BYTES=32768" "VIPRS_B200_STAGE_BYTES=28672" "X=2"; do
env $cfg timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
VIPRS_B200_STAGE_BYTES=32768 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
