cd /root/repo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r03h_launches.csv python bench.py --steps 4 --warmup 3 --equil 3 --fcm-steps 4 --no-cpu-baseline > gpurun_out/r03h_ncu_bench.log 2>&1; tail -1 gpurun_out/r03h_ncu_bench.log | cut -c1-200
