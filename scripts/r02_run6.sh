#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02b_pytest.log 2>&1; tail -8 gpurun_out/r02b_pytest.log
for V in 1 2; do
  for W in c2 small; do
    VIPRS_B200_FAST=$V timeout 300 python bench.py --workload $W --no-extras --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/r02b_${W}_fast$V.json 2>/dev/null
    python -c "import json;d=json.load(open('gpurun_out/r02b_${W}_fast$V.json'));print('FAST=$V $W ms/step %.4f kernel %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))"
    VIPRS_B200_FAST=$V VIPRS_B200_LIB=$PWD/viprs_b200/_C_trace/libviprs_b200.so VIPRS_B200_TRACE=gpurun_out/trace.bin \
        timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
    python scripts/trace_report.py gpurun_out/trace.bin 100 103 > gpurun_out/r02b_${W}_fast${V}_trace.txt 2>&1
    rm -f gpurun_out/trace.bin
  done
done
