cd /root/repo
timeout 1200 python -m pytest tests/test_dropin_gpu.py tests/test_brick_gpu.py tests/test_lj_gpu.py tests/test_verlet_gpu.py -q -m gpu 2>&1 | tail -12
