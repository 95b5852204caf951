cd $GRAFT_REPO_ROOT
VIPRS_B200_BUILD_TRACE=1 python -c "
from viprs_b200 import build; build.build(force=True)" 2>&1 | tail -1
VIPRS_B200_TRACE=gpurun_out/trace.bin timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-200
python scripts_trace.py gpurun_out/trace.bin 100 104 > gpurun_out/trace_report.txt 2>&1
cat gpurun_out/trace_report.txt
rm -f gpurun_out/trace.bin
