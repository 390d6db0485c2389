cd /root/repo
for g in sorted warp; do
UB200_IBM_GATHER=$g timeout 600 python - <<'PY'
import os, json, torch, sys
sys.path.insert(0, '/root/repo')
import bench_extra as b
r = b.pse(torch.device('cuda:0'), steps=10, warmup=3)
print(os.environ.get('UB200_IBM_GATHER'), {k: r[k] for k in ('value', 'ms_per_step', 'far_field_T0_ms', 'near_field_T0_ms')})
PY
done
timeout 900 python -m pytest tests/test_pse_gpu.py tests/test_fcm_gpu.py -q -x 2>&1 | grep -v "^\[W" | tail -3
