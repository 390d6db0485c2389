cd /root/repo
scripts/_bin/lj_col_ab > gpurun_out/r02d_lj_ab.json; cat gpurun_out/r02d_lj_ab.json
timeout 600 python bench.py --steps 50 --warmup 5 --no-fcm --no-extra --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
cat gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
