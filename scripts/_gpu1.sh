cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^\[W" | tail -6 > gpurun_out/r02x_pytest_gpu.log; tail -4 gpurun_out/r02x_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02x_bench_ours.json 2> gpurun_out/r02x_ours.err; tail -2 gpurun_out/r02x_ours.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open('/root/repo/gpurun_out/r02x_bench_ours.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'])
for k, v in d.get('extra', {}).items():
    print(k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'kT', 'kT_from_velocities', 'rebuilds_per_step', 'error')})
PY
