"""Brick-decomposed DPD fluid with the ghost-cell halo exchange over NCCL (BASELINE config 4 shape), launched with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P scripts/brick_dpd.py [N] [steps]
Times `steps` steps (CUDA events, barrier on both sides, max over ranks, L2 flushed before every step) and then checks the
final state bit for bit against a single-GPU run of the same trajectory on rank 0. Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    warm = 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = dist.new_group(ranks=[0])  # the single-GPU check below must not enter collectives the other ranks never join
    from uammd_b200 import synthetic as syn
    from uammd_b200.domain import make_dpd
    from uammd_b200.md import Box, DPD
    from uammd_b200.multigpu import DistributedDPDMD
    L = (N / 3.0) ** (1.0 / 3.0)
    pos, vel = syn.uniform_cloud(N, L, seed=21), syn.maxwell_velocities(N, 1.0, seed=22)
    mk = lambda: DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99)
    md = make_dpd(Box(L), mk(), 0.01, N)
    md.setGlobalState(torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev))
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        md.forwardTime()
    md.stats = {"migrated": 0, "ghosts": 0}
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    dist.barrier(); torch.cuda.synchronize()
    for a, b in evs:
        scrub.fill_(1)
        a.record()
        md.forwardTime()
        b.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counts = torch.tensor([md.nOwned, md.pos.shape[0] - md.nOwned, md.stats["migrated"]], device=dev, dtype=torch.int64)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    gp, gv = md.gatherGlobalState()
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"timing_only": True, "ms_per_step": float(ms.item()), "n_gpus": world}), flush=True)
        p, v, f = torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), torch.zeros(N, 4, device=dev)
        single = DistributedDPDMD(Box(L), mk(), 0.01, N, group=solo)
        assert single.world == 1
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(warm + steps):
            if k == warm:
                s0.record()
            single.forwardTime(p, v, f)
        s1.record()
        torch.cuda.synchronize()
        same = bool(torch.equal(gp.view(torch.int32), p.view(torch.int32)) and torch.equal(gv.view(torch.int32), v.view(torch.int32)))
        c = torch.stack(allc).cpu().numpy()
        print(json.dumps({"metric": "DPD MD steps/s @%d particles, brick decomposition + ghost-cell halo exchange" % N,
                          "value": 1000.0 / float(ms.item()), "unit": "steps/s", "ms_per_step": float(ms.item()), "n_gpus": world,
                          "rank_grid": list(md.dec.rankGrid), "owned_per_rank": c[:, 0].tolist(), "ghosts_per_rank": c[:, 1].tolist(),
                          "migrated_per_step": float(c[:, 2].sum()) / steps, "bit_identical_to_single_gpu": same,
                          "single_gpu_ms_per_step_back_to_back": s0.elapsed_time(s1) / steps,
                          "halo_bytes_per_rank_per_step": int(c[:, 1].mean()) * 32,
                          "l2": "flushed before every step (256 MiB write)"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
