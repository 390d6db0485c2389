cd /root/repo
timeout 900 python -m pytest tests/test_vlist_gpu.py tests/test_verlet_gpu.py tests/test_nvt_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -12
timeout 600 python scripts/vlist_time.py > gpurun_out/r03l_vlist_time.json 2> gpurun_out/r03l_vlist_time.err; tail -3 gpurun_out/r03l_vlist_time.err; cat gpurun_out/r03l_vlist_time.json
UB200_VERLET_TRAVERSAL=gather timeout 600 python scripts/vlist_time.py 2>/dev/null | cut -c1-420
