cd $GRAFT_REPO_ROOT
export VIPRS_B200_LIB=$GRAFT_REPO_ROOT/viprs_b200/_C_trace/libviprs_b200.so
VIPRS_B200_TRACE=gpurun_out/trace.bin timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-20
python scripts_trace.py gpurun_out/trace.bin 100 103 > gpurun_out/trace_report.txt 2>&1
cat gpurun_out/trace_report.txt | cut -c1-220
rm -f gpurun_out/trace.bin
