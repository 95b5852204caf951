cd $GRAFT_REPO_ROOT
for v in _C_oldtrace _C_trace; do
export VIPRS_B200_LIB=$GRAFT_REPO_ROOT/viprs_b200/$v/libviprs_b200.so
VIPRS_B200_TRACE=gpurun_out/trace.bin timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | cut -c1-20
python scripts_trace.py gpurun_out/trace.bin 100 102 > gpurun_out/trace_report$v.txt 2>&1
grep -v "^  [AC][4-7]:" gpurun_out/trace_report$v.txt | cut -c1-200
rm -f gpurun_out/trace.bin
done
