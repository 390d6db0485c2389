cd /root/repo
timeout 600 oracle/_ref/dropin_poisson 2000 200000 2>&1 | tail -1 | tee gpurun_out/r02z_poisson.json
timeout 900 python -m pytest tests/test_poisson_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -3
