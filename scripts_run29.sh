cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'prepare_kernel|sweep|sums_kernel' -c 15 --csv --log-file gpurun_out/r01b_c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -c sweep gpurun_out/r01b_c2_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep' -s 3 -c 2 --csv --log-file gpurun_out/r01b_c4_launches.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 800 ncu --set full --import-source on --clock-control none -k regex:'sweep_fast' -s 3 -c 1 -o gpurun_out/r01b_c2_fast_full -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out/ | head -30
