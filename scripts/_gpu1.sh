cd /root/repo
timeout 900 python -m pytest tests/test_fcm_gpu.py tests/test_poisson_gpu.py tests/test_pse_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -4
timeout 600 oracle/_ref/dropin_poisson 2000 0 2>&1 | tail -1
timeout 300 scripts/_bin/fft_vs_cufft 10 2>&1 | tail -4
