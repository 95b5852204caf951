#!/bin/bash
# One gpurun call that reproduces the round's evidence on a B200 box (run from the repo root of the snapshot):
#   gpurun --timeout 1800 -- 'bash scripts/gpu_round_check.sh'
# GPU tests, smoke, the default bench (c2) with every leg, the reference arm, the other workloads, the ncu launch
# list of the same command, one ncu --set full capture of the sweep, and the phase timeline of CTA 0 (needs the
# trace build: VIPRS_B200_OUT=_C_trace VIPRS_B200_BUILD_TRACE=1 python -m viprs_b200.build, done here on the CPU box).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_c2.json 2>/dev/null
for w in c4 c1 c3; do timeout 900 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'prepare_kernel|sweep|sums_kernel' -c 15 --csv --log-file gpurun_out/c2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 800 ncu --set full --import-source on --clock-control none -k regex:'sweep_fast' -s 3 -c 1 -o gpurun_out/c2_fast_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
if [ -f viprs_b200/_C_trace/libviprs_b200.so ]; then
    VIPRS_B200_LIB=$PWD/viprs_b200/_C_trace/libviprs_b200.so VIPRS_B200_TRACE=gpurun_out/trace.bin \
        timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    python scripts/trace_report.py gpurun_out/trace.bin 100 103 > gpurun_out/c2_phase_trace.txt 2>&1
    rm -f gpurun_out/trace.bin
fi
