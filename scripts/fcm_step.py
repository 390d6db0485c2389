"""Run a few FCM steps of BASELINE config 3 (used under ncu)."""
import sys; sys.path.insert(0, '.')
import torch
import bench_fcm as fcm_bench
r = fcm_bench.run(torch.device('cuda:0'), 6550.7, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 5, warmup=3)
print({k: r[k] for k in ('value', 'steps_per_s', 'ms_per_step')})
