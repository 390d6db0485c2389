cd /root/repo
timeout 900 python -m pytest tests/test_pse_gpu.py "tests/test_dropin_gpu.py::test_verlet_bd_pse_dropin_matches_reference" -q -x 2>&1 | grep -v "^\[W" | tail -5
timeout 600 python - <<'PY'
import torch, sys
sys.path.insert(0, '/root/repo')
import bench_extra as b
r = b.pse(torch.device('cuda:0'), steps=10, warmup=3)
print({k: r[k] for k in ('value', 'ms_per_step', 'far_field_T0_ms', 'near_field_T0_ms', 'lanczos_iterations')})
PY
