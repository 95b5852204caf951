cd $GRAFT_REPO_ROOT
timeout 800 ncu --set full --import-source on --clock-control none -k regex:'sweep_fast' -s 3 -c 1 -o gpurun_out/c2_fast_full -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out/
