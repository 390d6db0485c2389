cd /root/repo
timeout 600 oracle/_ref/dropin_poisson 2000 200000 2>&1 | tail -1 | tee gpurun_out/r03e_poisson.json
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^\[W" | tail -5
