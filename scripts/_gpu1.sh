cd /root/repo
timeout 900 python -m pytest tests/test_brick_gpu.py tests/test_ljengine_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02e_pytest_brick.log
tail -30 gpurun_out/r02e_pytest_brick.log
UB200_LJ_WIDEN=0 scripts/_bin/lj_col_ab > gpurun_out/r02e_lj_ab_nowiden.json; cat gpurun_out/r02e_lj_ab_nowiden.json
