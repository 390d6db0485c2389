cd /root/repo
N=$1
UB200_BENCH_DUMP=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2989$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r03o_bench$N.json 2> gpurun_out/r03o_bench$N.err
grep "per-step" gpurun_out/r03o_bench$N.err | cut -c1-150
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f'/root/repo/gpurun_out/r03o_bench{n}.json'))
print(n, 'GPUs', {k: d[k] for k in ('value','ms_per_step','value_back_to_back')})
print('  fcm', {k: d['fcm'][k] for k in ('value','ms_per_step','value_back_to_back','barrier_timeouts')})
print('  pse_far', {k: d['pse_far'].get(k) for k in ('value','ms_per_step','single_gpu_ms','error')})
print('  dpd', {k: d['dpd'].get(k) for k in ('value','ms_per_step','kT_from_velocities','error')})
PY
