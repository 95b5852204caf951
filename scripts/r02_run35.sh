#!/bin/bash
# phase trace of a lone CTA on a small LD block (c1: 22 blocks of 400..1200 SNPs)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for wl in c1 ln; do
VIPRS_B200_LIB=$PWD/viprs_b200/_C_trace/libviprs_b200.so VIPRS_B200_TRACE=gpurun_out/trace.bin \
    timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
python scripts/trace_report.py gpurun_out/trace.bin 20 24 > gpurun_out/r02x_${wl}_trace.txt 2>&1
rm -f gpurun_out/trace.bin
done
cat gpurun_out/r02x_c1_trace.txt
