cd /root/repo
timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_brick_gpu.py tests/test_domain_gpu.py -q -m gpu 2>&1 | grep -v "^\[W" | tail -6 | tee gpurun_out/r03b_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r03b_bench2.json 2> gpurun_out/r03b_bench2.err
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/r03b_bench2.json'))
print({k: d[k] for k in ('value','ms_per_step','value_back_to_back','gpu_launches')}, 'e2e', d['e2e']['value'], d['bricks'])
print('  fcm', {k: d['fcm'][k] for k in ('value','ms_per_step','value_back_to_back','barrier_timeouts')})
print('  dpd', {k: d['dpd'].get(k) for k in ('value','ms_per_step','kT_from_velocities','rank_grid','error_flags','error')})
PY
tail -2 gpurun_out/r03b_bench2.err | cut -c1-300
