cd $GRAFT_REPO_ROOT
rm -f gpurun_out/tune.log
bash scripts_tune.sh "VIPRS_B200_L2_AHEAD=0" "VIPRS_B200_L2_AHEAD=2" "VIPRS_B200_L2_AHEAD=4" "VIPRS_B200_L2_AHEAD=8" "VIPRS_B200_L2_AHEAD=16" "VIPRS_B200_L2_AHEAD=8 VIPRS_B200_STAGE_BYTES=16384" "VIPRS_B200_L2_AHEAD=4 VIPRS_B200_STAGE_BYTES=16384" "VIPRS_B200_LIMBS=4"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep|pack_rows|backward_dot' -c 40 --csv --log-file gpurun_out/launches_r1_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
VIPRS_B200_L2_AHEAD=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep' -s 3 -c 3 --csv --log-file gpurun_out/launches_l2ahead0.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
