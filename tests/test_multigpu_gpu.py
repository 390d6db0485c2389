"""GPU (needs >= 2 devices, skipped otherwise): the NCCL particle decomposition reproduces the single-GPU fused
engine bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, steps, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, LJ
    from uammd_b200.multigpu import DistributedLJMD
    dev = torch.device("cuda", rank)
    Lb = syn.lj_box_length(N)
    pos = torch.from_numpy(syn.fcc_lattice(N, Lb)).to(dev)
    vel = syn.maxwell_velocities(N, 1.0)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    md = DistributedLJMD(Box(Lb), pot, 0.005, N, engine="cuda")
    vb = torch.from_numpy(vel[md.dec.lo:md.dec.hi].copy()).to(dev)
    force = torch.zeros(N, 4, device=dev)
    md.run(pos, vb, force, steps)
    md._gather(pos)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out, pos.cpu().numpy())
    dist.destroy_process_group()


def test_nccl_decomposition_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, LJ, LJMD
    N, steps = 4 * 20 ** 3, 25
    out = str(tmp_path / "pos.npy")
    mp.spawn(_worker, args=(2, 29533, N, steps, out), nprocs=2, join=True)
    got = np.load(out)
    Lb = syn.lj_box_length(N)
    dev = torch.device("cuda:0")
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    p = torch.from_numpy(syn.fcc_lattice(N, Lb)).to(dev)
    v = torch.from_numpy(syn.maxwell_velocities(N, 1.0)).to(dev)
    f = torch.zeros(N, 4, device=dev)
    LJMD(Box(Lb), pot, 0.005).run(p, v, f, steps)
    assert np.array_equal(got.view(np.uint32), p.cpu().numpy().view(np.uint32))
