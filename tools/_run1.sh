python tools/run_calibration_round.py 1000 520 fast r02 2>&1 | tail -2
python tools/run_screening_grid.py 500 520 260 fast r02 2>&1 | tail -7
python -m pytest tests/test_gpu_calibration.py tests/test_gpu_netsim_stats.py -x -q -m gpu 2>&1 | tail -2
