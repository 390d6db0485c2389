cd /root/repo
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2986$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r03d_bench$N.json 2> gpurun_out/r03d_bench$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f'/root/repo/gpurun_out/r03d_bench{n}.json'))
print(n, 'GPUs', {k: d[k] for k in ('value','ms_per_step','value_back_to_back','gpu_launches','scaling')}, 'e2e', d['e2e']['value'], d['bricks'])
print('  fcm', {k: d['fcm'][k] for k in ('value','ms_per_step','value_back_to_back','barrier_timeouts')})
print('  dpd', {k: d['dpd'].get(k) for k in ('value','ms_per_step','kT_from_velocities','rank_grid','error_flags','error')})
PY
tail -2 gpurun_out/r03d_bench$N.err | cut -c1-200
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | grep -v "^\[W" | tail -3 | tee gpurun_out/r03d_pytest_${N}gpu.log
