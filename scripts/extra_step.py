"""Run the secondary bench legs with few steps (used under ncu): python scripts/extra_step.py [verlet] [pse] [bd]"""
import sys; sys.path.insert(0, '.')
import json
import torch
import bench_extra as extra_bench
from uammd_b200 import synthetic as syn
dev = torch.device('cuda:0')
which = sys.argv[1:] or ['verlet', 'pse', 'bd']
if 'verlet' in which:
    N = 1_000_000; Lb = syn.lj_box_length(N)
    print(json.dumps(extra_bench.verlet(dev, N, Lb, syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0, seed=7), 2.5, 0.005, steps=20, warmup=5, equil=40)))
if 'pse' in which:
    print(json.dumps(extra_bench.pse(dev, steps=3, warmup=2)))
if 'bd' in which:
    print(json.dumps(extra_bench.bd_ideal(dev, steps=20)))
