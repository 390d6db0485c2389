cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^\[W" | tail -8 > gpurun_out/r02z_pytest_gpu.log; tail -6 gpurun_out/r02z_pytest_gpu.log
