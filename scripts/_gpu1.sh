cd /root/repo
timeout 900 python -m pytest tests/test_nvt_gpu.py tests/test_oracle_nvt.py "tests/test_dropin_gpu.py::test_langevin_verlet_dropin_matches_reference" -q -x 2>&1 | grep -v "^\[W" | tail -8
ls gpurun_out/*.npz
