cd $GRAFT_REPO_ROOT
for v in _C_old _C; do
export VIPRS_B200_LIB=$GRAFT_REPO_ROOT/viprs_b200/$v/libviprs_b200.so
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__issue_active.max.pct_of_peak_sustained_active,smsp__issue_active.min.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:'sweep_fast' -s 3 -c 1 --csv --log-file gpurun_out/ncu_inst$v.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_inst$v.csv')) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    print('$v', r[h.index('Metric Name')], r[h.index('Metric Value')])
PY
done
