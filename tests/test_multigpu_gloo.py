"""CPU: the multi-GPU particle decomposition (uammd_b200/multigpu.py) exercised with world_size 2 over gloo, with the
checker engines of tests/_checker_engines.py (oracle arithmetic) injected in place of the CUDA engine: two ranks must
reproduce the single-process trajectory bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "tests")) if p not in sys.path]  # also in the spawned workers


def _worker(rank, world, port, N, steps, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, LJ
    from uammd_b200.multigpu import DistributedLJMD
    Lb = syn.lj_box_length(N)
    pos = torch.from_numpy(syn.fcc_lattice(N, Lb))
    vel = torch.from_numpy(syn.maxwell_velocities(N, 1.0))
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    from _checker_engines import OracleLJEngine
    md = DistributedLJMD(Box(Lb), pot, 0.004, N, engine=OracleLJEngine(Box(Lb), pot, 0.004))
    force = torch.zeros(N, 4)
    vb = vel[md.dec.lo:md.dec.hi].clone()
    md.run(pos, vb, force, steps)
    md.run(pos, vb, force, 2)
    # final positions of the other ranks' blocks are one gather behind the owners: gather once more to compare
    md._gather(pos)
    allv = [torch.zeros_like(vb) for _ in range(world)]
    dist.all_gather(allv, vb)
    if rank == 0:
        np.save(out, np.concatenate([pos.numpy().ravel(), torch.cat(allv).numpy().ravel()]))
    dist.destroy_process_group()


def test_block_decomposition_rules():
    sys.path.insert(0, ROOT)
    from uammd_b200.multigpu import BlockDecomposition
    d = BlockDecomposition(1000, 4, 2)
    assert (d.lo, d.hi, d.block) == (500, 750, 250)
    try:
        BlockDecomposition(1001, 4, 0)
        assert False
    except ValueError:
        pass


def test_two_ranks_reproduce_single_process_trajectory(tmp_path, orc):
    from uammd_b200 import synthetic as syn
    N, steps = 4 * 6 ** 3, 4
    out = str(tmp_path / "dist.npy")
    mp.spawn(_worker, args=(2, 29517, N, steps, out), nprocs=2, join=True)
    got = np.load(out)
    Lb = syn.lj_box_length(N)
    ref = orc.MDOracle((Lb,) * 3, 2.5, syn.lj_params(), 0.004, syn.fcc_lattice(N, Lb), syn.maxwell_velocities(N, 1.0))
    ref.step(steps + 2)
    assert np.array_equal(got[:4 * N].view(np.uint32), ref.pos.ravel().view(np.uint32))
    assert np.array_equal(got[4 * N:].view(np.uint32), ref.vel.ravel().view(np.uint32))


def _blob_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uammd_b200.multigpu import exchange_blobs
    blob = bytes([rank * 16 + k for k in range(8)]) + bytes(56)   # 64 bytes like a cudaIpcMemHandle_t
    allb = exchange_blobs(blob)
    if rank == 1:
        open(out, "wb").write(allb)
    dist.destroy_process_group()


def test_slab_ranges_and_blob_exchange(tmp_path):
    sys.path.insert(0, ROOT)
    from uammd_b200.multigpu import slab_ranges
    assert slab_ranges(128, 8) == [(16 * r, 16 * r + 16) for r in range(8)]
    assert slab_ranges(64, 2) == [(0, 32), (32, 64)]
    try:
        slab_ranges(130, 8)
        assert False
    except ValueError:
        pass
    out = str(tmp_path / "blobs.bin")
    mp.spawn(_blob_worker, args=(2, 29519, out), nprocs=2, join=True)
    b = open(out, "rb").read()
    assert len(b) == 128 and b[:8] == bytes(range(8)) and b[64:72] == bytes(range(16, 24))


def _dpd_setup(N):
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, DPD
    L = (N / 3.0) ** (1.0 / 3.0)
    pos = syn.uniform_cloud(N, L, seed=21)
    vel = syn.maxwell_velocities(N, 1.0, seed=22)
    return Box(L), DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99), pos, vel


def _dpd_worker(rank, world, port, N, steps, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uammd_b200.multigpu import DistributedDPDMD
    box, pot, pos, vel = _dpd_setup(N)
    p, v, f = torch.from_numpy(pos), torch.from_numpy(vel), torch.zeros(N, 4)
    from _checker_engines import OracleDPDEngine
    md = DistributedDPDMD(box, pot, 0.01, N, engine=OracleDPDEngine(box, pot, 0.01, N))
    for _ in range(steps):
        md.forwardTime(p, v, f)
    md.gatherState(p, v)
    if rank == 0:
        np.save(out, np.concatenate([p.numpy().ravel(), v.numpy().ravel()]))
    dist.destroy_process_group()


def test_dpd_two_ranks_reproduce_single_process(tmp_path, orc):
    """Particle-decomposed DPD (positions + velocities all-gathered, noise keyed on global pair indices): two gloo ranks
    give the single-process trajectory bit for bit."""
    sys.path.insert(0, ROOT)
    from uammd_b200.multigpu import DistributedDPDMD
    N, steps = 3000, 3
    out = str(tmp_path / "dpd.npy")
    mp.spawn(_dpd_worker, args=(2, 29521, N, steps, out), nprocs=2, join=True)
    got = np.load(out)
    box, pot, pos, vel = _dpd_setup(N)
    p, v, f = torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy()), torch.zeros(N, 4)
    from _checker_engines import OracleDPDEngine
    md = DistributedDPDMD(box, pot, 0.01, N, engine=OracleDPDEngine(box, pot, 0.01, N))
    for _ in range(steps):
        md.forwardTime(p, v, f)
    assert np.array_equal(got[:4 * N].view(np.uint32), p.numpy().ravel().view(np.uint32))
    assert np.array_equal(got[4 * N:].view(np.uint32), v.numpy().ravel().view(np.uint32))
    assert not np.array_equal(p.numpy(), pos)
