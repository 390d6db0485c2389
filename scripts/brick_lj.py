"""One rank of the brick-decomposed LJ MD under torchrun (one process per GPU): timing and, with --check, comparison with
the single-GPU engine on rank 0. Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uammd_b200 import synthetic as syn  # noqa: E402
from uammd_b200.brickmd import BrickLJMD  # noqa: E402
from uammd_b200.md import Box, LJ, LJMD  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=63)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 4 * args.cells ** 3
Lb = syn.lj_box_length(N, 0.8)
pos = syn.fcc_lattice(N, Lb)
pos[:, :3] += np.random.default_rng(1).normal(0, 0.05, (N, 3)).astype(np.float32)
vel = syn.maxwell_velocities(N, 1.0, seed=7)
pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
box = Box(Lb)
md = BrickLJMD(box, pot, 0.005, N, rank, world)
md.connect()
md.setGlobalState(torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev))
md.run(0)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
md.run(args.steps)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
no, nl, err = md.counts()
p, v, f = md.gatherGlobalState()
if rank == 0:
    out = {"N": N, "world": world, "rankGrid": md.rankGrid, "ms_per_step": float(ms), "owned_rank0": no, "local_rank0": nl, "err": err}
    if os.environ.get("UB200_BRICK_PROFILE") == "1":
        out["phases_ms_rank0"] = md.profile()
    if args.check:
        os.environ["UB200_LJ_WIDEN"] = "0"  # eight lanes per particle throughout, like the bricks (see tests/test_brick_gpu.py)
        one = LJMD(box, pot, 0.005)
        ps, vs, fs = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
        one.run(ps, vs, fs, args.steps)
        torch.cuda.synchronize()
        out["bit_identical"] = bool(np.array_equal(p, ps.cpu().numpy()) and np.array_equal(v, vs.cpu().numpy()))
        out["max_pos_diff"] = float(np.abs(p - ps.cpu().numpy()).max())
    print(json.dumps(out))
dist.destroy_process_group()
sys.exit(0 if (not args.check or rank != 0 or out["bit_identical"]) and err == 0 else 1)
