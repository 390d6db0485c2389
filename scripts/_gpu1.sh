cd /root/repo
timeout 600 oracle/_ref/dropin_poisson 2000 200000 2>&1 | tail -1 | tee gpurun_out/r03f_poisson.json
timeout 900 python -m pytest tests/test_poisson_gpu.py "tests/test_dropin_gpu.py::test_poisson_dropin_matches_reference" -q -x 2>&1 | grep -v "^\[W" | tail -3
