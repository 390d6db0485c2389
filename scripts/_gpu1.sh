cd /root/repo
T=r03k
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^\[W" | tail -6 > gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_ref.err; tail -1 gpurun_out/${T}_ref.err | cut -c1-200
timeout 900 python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_ours.err; tail -1 gpurun_out/${T}_ours.err | cut -c1-200
python - <<'PY'
import json
T='r03k'
o = json.load(open(f'/root/repo/gpurun_out/{T}_bench_ours.json')); r = json.load(open(f'/root/repo/gpurun_out/{T}_bench_reference.json'))
print('LJ', o['value'], r['value'], o['value']/r['value'], 'e2e', o['e2e']['value'], 'resident', o['e2e_resident']['value'])
for k in ('fcm','verlet','pse','bd','langevin','dpd','poisson'):
    a, b = o.get(k, {}), r.get(k, {})
    print(k, a.get('value'), b.get('value'), (a.get('value') or 0)/(b.get('value') or 1e30), a.get('error'), b.get('error'))
PY
