cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in _C _C_u4 _C; do
VIPRS_B200_LIB=$PWD/viprs_b200/$v/libviprs_b200.so timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
