cd /root/repo
for kb in 52 26; do
UB200_IBM_SPREAD_KB=$kb timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench_fcm, bench_extra
from bench import measured_peaks
r = bench_fcm.run(torch.device('cuda:0'), measured_peaks()[0], steps=50)
p = bench_extra.pse(torch.device('cuda:0'), steps=6, warmup=2)
print(os.environ['UB200_IBM_SPREAD_KB'], 'fcm', round(r['ms_per_step'], 4), round(r['value'], 1), 'pse', round(p['ms_per_step'], 3), 'far', round(p['far_field_T0_ms'], 3))
PY
done
UB200_IBM_SPREAD_KB=26 timeout 600 python -m pytest tests/test_fcm_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -2
