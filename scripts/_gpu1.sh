cd /root/repo
timeout 300 python scripts/verlet_fill_time.py | tee gpurun_out/r02o_verlet_fill.json
timeout 300 ncu --set full --clock-control none -k regex:verletFill -s 3 -c 1 -f -o gpurun_out/r02o_verlet python scripts/verlet_fill_time.py > gpurun_out/r02o_verlet_ncu.log 2>&1
ncu -i gpurun_out/r02o_verlet.ncu-rep --page raw --csv > gpurun_out/r02o_verlet_raw.csv
timeout 300 python -m pytest tests/test_verlet_gpu.py -q -m gpu 2>&1 | tail -3
