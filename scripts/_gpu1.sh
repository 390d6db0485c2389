cd /root/repo
timeout 30 python -m pytest tests/test_pse_dist_gpu.py -q -rs 2>&1 | grep -v "^\[W" | grep -E "^E  |passed|failed|FAILED|Error|^\.|SKIP" > gpurun_out/r04d_pse_dist_pytest.log
cat gpurun_out/r04d_pse_dist_pytest.log
