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


def _fcm_worker(rank, world, port, N, n, T, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from uammd_b200 import synthetic as syn
    from uammd_b200.fcm import Peskin3
    from uammd_b200.multigpu import DistributedFCM
    dev = torch.device("cuda", rank)
    L = float(n)
    pos = np.zeros((N, 4)); pos[:, :3] = syn.uniform_cloud(N, L, seed=11)[:, :3].astype(np.float64)
    force = np.zeros((N, 4)); force[:, :3] = syn.gaussian_forces(N, seed=12)
    fcm = DistributedFCM(L, (n, n, n), Peskin3(L / n), 1.0, N, seed=1234)
    dp, df = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    res = []
    for _ in range(3):   # several calls: buffer reuse across the barriers, advancing noise counter
        res.append(fcm.computeHydrodynamicDisplacements(dp, df, temperature=T, prefactor=1.0).cpu().numpy())
    assert fcm.errorFlag() == 0
    np.save(out + f".{rank}.npy", np.stack(res))
    fcm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [0.0, 1.0])
def test_slab_fcm_matches_single_gpu(tmp_path, T):
    """z-slab FCM over all visible GPUs (2, 4 or 8): bit-identical to the single-GPU pipeline on every rank."""
    world = torch.cuda.device_count()
    world = 8 if world >= 8 else (4 if world >= 4 else world)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from uammd_b200 import synthetic as syn
    from uammd_b200.fcm import FCM_impl, Peskin3
    N, n = 60000, 64
    out = str(tmp_path / "mf")
    mp.spawn(_fcm_worker, args=(world, 29541 + int(T), N, n, T, out), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    L = float(n)
    pos = np.zeros((N, 4)); pos[:, :3] = syn.uniform_cloud(N, L, seed=11)[:, :3].astype(np.float64)
    force = np.zeros((N, 4)); force[:, :3] = syn.gaussian_forces(N, seed=12)
    single = FCM_impl(L, (n, n, n), Peskin3(L / n), 1.0, seed=1234)
    dp, df = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    want = np.stack([single.computeHydrodynamicDisplacements(dp, df, temperature=T, prefactor=1.0).cpu().numpy() for _ in range(3)])
    for r in range(world):
        got = np.load(out + f".{r}.npy")
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), \
            f"rank {r}: max |d| = {np.abs(got - want).max()} (rel {np.abs(got - want).max() / np.abs(want).max():.2e})"


def _dpd_gpu_worker(rank, world, port, N, steps, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, DPD
    from uammd_b200.multigpu import DistributedDPDMD
    dev = torch.device("cuda", rank)
    L = (N / 3.0) ** (1.0 / 3.0)
    p = torch.from_numpy(syn.uniform_cloud(N, L, seed=21)).to(dev)
    v = torch.from_numpy(syn.maxwell_velocities(N, 1.0, seed=22)).to(dev)
    f = torch.zeros(N, 4, device=dev)
    md = DistributedDPDMD(Box(L), DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99), 0.01, N)
    for _ in range(steps):
        md.forwardTime(p, v, f)
    md.gatherState(p, v)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(out, np.concatenate([p.cpu().numpy().ravel(), v.cpu().numpy().ravel()]))
    dist.destroy_process_group()


def test_dpd_decomposition_matches_single_gpu(tmp_path):
    """BASELINE config 4 shape (rho = 3, rc = 1, A = 25, gamma = 4.5) at N = 240 000: 2 ranks vs the single-GPU
    VerletNVE + PairForcesDPD classes, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from uammd_b200 import synthetic as syn
    from uammd_b200.md import Box, DPD, PairForcesDPD, VerletNVE
    N, steps = 240_000, 5
    out = str(tmp_path / "dpd.npy")
    mp.spawn(_dpd_gpu_worker, args=(2, 29547, N, steps, out), nprocs=2, join=True)
    got = np.load(out)
    dev = torch.device("cuda:0")
    L = (N / 3.0) ** (1.0 / 3.0)
    p = torch.from_numpy(syn.uniform_cloud(N, L, seed=21)).to(dev)
    v = torch.from_numpy(syn.maxwell_velocities(N, 1.0, seed=22)).to(dev)
    pf = PairForcesDPD(DPD(cutOff=1.0, dt=0.01, gamma=4.5, temperature=1.0, A=25.0, seed=99), Box(L))

    class _It:
        def sum(self, pos, force=None):
            pf.sum(pos, v, force)
    integ = VerletNVE(p, v, 0.01)
    integ.addInteractor(_It())
    for _ in range(steps):
        integ.forwardTime()
    torch.cuda.synchronize()
    assert np.array_equal(got[:4 * N].view(np.uint32), p.cpu().numpy().ravel().view(np.uint32))
    assert np.array_equal(got[4 * N:].view(np.uint32), v.cpu().numpy().ravel().view(np.uint32))
