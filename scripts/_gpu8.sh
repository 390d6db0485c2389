cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus 8 --steps 100 --warmup 10 --no-extra > gpurun_out/r02m_bench8.json 2> gpurun_out/r02m_bench8.err
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/r02m_bench8.json'))
print({k: d[k] for k in ('value','ms_per_step','value_back_to_back','gpu_launches')}, d['e2e']['value'], d['bricks'])
print('fcm', {k: d['fcm'][k] for k in ('value','ms_per_step','value_back_to_back','barrier_timeouts','gpu_launches')})
PY
tail -3 gpurun_out/r02m_bench8.err
UB200_DIST_GRAPH=0 UB200_BRICK_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29842 bench.py --gpus 8 --steps 100 --warmup 10 --no-extra > gpurun_out/r02m_bench8_nograph.json 2> gpurun_out/r02m_bench8_nograph.err
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/r02m_bench8_nograph.json'))
print('nograph', {k: d[k] for k in ('value','ms_per_step','value_back_to_back')}, 'fcm', {k: d['fcm'][k] for k in ('value','ms_per_step','value_back_to_back')})
PY
