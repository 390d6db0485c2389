cd /root/repo
timeout 900 python -m pytest tests/test_celllist_gpu.py tests/test_abi.py -q -x 2>&1 | grep -v "^\[W" | tail -5
