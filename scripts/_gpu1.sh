cd /root/repo
timeout 900 python -m pytest tests/test_poisson_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -30
