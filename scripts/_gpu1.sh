cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02j_pytest_gpu.log
tail -25 gpurun_out/r02j_pytest_gpu.log
