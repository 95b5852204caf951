cd $GRAFT_REPO_ROOT
python scratch/grid_perf.py 37 256 3
python scratch/grid_perf.py 37 64 3
python scratch/grid_perf.py 148 8 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 1 -c 1 -o gpurun_out/prof_grid python scratch/grid_perf.py 37 32 2 > gpurun_out/ncu_grid.log 2>&1
