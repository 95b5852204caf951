cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_grid_gpu.py -m gpu -x -q 2>&1 | tail -5
python scratch/grid_perf.py 37 256 3
python scratch/grid_perf.py 148 8 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_sweep -s 1 -c 1 -o gpurun_out/prof_grid2 python scratch/grid_perf.py 37 32 2 > gpurun_out/ncu_grid.log 2>&1
