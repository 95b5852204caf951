#!/bin/bash
# fine-grained trace of one A row group (-DVB_TRACE2): where do its ~900 cycles go?
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
VIPRS_B200_LIB=$PWD/viprs_b200/_C_trace/libviprs_b200.so VIPRS_B200_TRACE=gpurun_out/trace.bin \
    timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
python scripts/trace_report.py gpurun_out/trace.bin 100 102 > gpurun_out/r02x_c2_trace2.txt 2>&1
rm -f gpurun_out/trace.bin
tail -n 8 gpurun_out/r02x_c2_trace2.txt
