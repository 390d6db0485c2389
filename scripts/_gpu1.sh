cd /root/repo
oracle/_ref/gen_golden_windows gpurun_out six; ls -la gpurun_out/windows_six_f64.bin
timeout 900 python -m pytest tests/test_fcm_gpu.py tests/test_oracle_fcm.py -q -x 2>&1 | grep -v "^\[W" | tail -8
