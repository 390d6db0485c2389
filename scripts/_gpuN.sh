cd /root/repo
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2990$N bench.py --gpus $N --steps 30 --warmup 5 --no-extra --no-fcm > gpurun_out/r03p_bench$N.json 2> gpurun_out/r03p_bench$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f'/root/repo/gpurun_out/r03p_bench{n}.json'))
print(n, 'GPUs', {k: d[k] for k in ('value','ms_per_step','value_back_to_back')}, 'e2e', d['e2e']['value'])
PY
tail -1 gpurun_out/r03p_bench$N.err | cut -c1-200
