"""GPU: BDHI::PSE near field and Lanczos noise over ranks (ub200_pse_dist_*).

The reference is single GPU (PSE/NearField.cuh:236-282, misc/LanczosAlgorithm/LanczosAlgorithm.cu:202-228), so the oracle of
the decomposition is our own single-GPU path, which tests/test_pse_gpu.py pins against the compiled reference.

  * one GPU is enough for the first tests: VIRTUAL ranks = one handle, one stream and one host thread per rank on the
    same device, wired with ub200_pse_dist_attach_local (the peer stores and the one-warp reduce / barrier kernels are the
    ones the NVLink path runs; only the CUDA IPC import is replaced). Up to three virtual ranks: a B200 did not co-schedule
    the kernels of a fourth stream next to three spinning barriers (gpurun_out/r04c_pse_dist_pytest.log);
  * with >= 2 GPUs: one process per GPU over CUDA IPC, the whole of PSE::computeHydrodynamicDisplacements
    (near rows + far slabs, BASELINE config 3) against the single-GPU call."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

from uammd_b200 import bd, pse
from uammd_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _cloud(N, L, dtype):
    pos = np.zeros((N, 4), dtype); pos[:, :3] = syn.uniform_cloud(N, L, seed=51)[:, :3]
    force = np.zeros((N, 4), dtype); force[:, :3] = syn.gaussian_forces(N, seed=52).astype(dtype)
    return pos, force


def _virtual_ranks(cuda, world, p, f, par, T, pref, seed2, calls=1, wantIt=None):
    """runs Mdot and the Lanczos noise on `world` virtual ranks; returns per rank (near, noise, iterations, error flag)"""
    from uammd_b200.multigpu import DistributedPSENearField
    N = p.shape[0]
    ranks = [DistributedPSENearField(p, par, sys=bd.System(7), rank=r, world=world) for r in range(world)]
    DistributedPSENearField.attachLocal(ranks)
    streams = [torch.cuda.Stream(device=cuda) for _ in range(world)]
    near = [torch.zeros(N, 3, dtype=p.dtype, device=cuda) for _ in range(world)]
    noise = [torch.zeros(N, 3, dtype=p.dtype, device=cuda) for _ in range(world)]
    its, errs, fails = [0] * world, [0] * world, []
    torch.cuda.synchronize()
    for r in range(world):            # list builds: no communication; done before the lockstep phase so that no
        ranks[r].prepare(streams[r])  # device-wide synchronisation (allocation growth) can meet a spinning barrier
    torch.cuda.synchronize()

    def work(r):
        try:
            torch.cuda.set_device(cuda)
            for _ in range(calls):
                near[r].zero_(); noise[r].zero_()
                streams[r].wait_stream(torch.cuda.current_stream())
                ranks[r].Mdot(f, near[r], streams[r])
                its[r] = ranks[r].noiseAdd(noise[r], T, pref, seed2, streams[r])
                streams[r].synchronize()
            errs[r] = ranks[r].errorFlag(streams[r])
        except Exception as e:  # noqa: BLE001 - reported by the main thread
            try:
                flag = ranks[r].errorFlag(streams[r])
            except Exception:  # noqa: BLE001
                flag = -1
            fails.append((r, repr(e), f"timed-out barrier (0 = none): {flag}"))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "a virtual rank hangs"
    if fails and all("timed-out barrier (0 = none): 0" not in f[2] for f in fails):
        # Virtual ranks need the barrier kernels of ALL ranks resident at once on the one device; CUDA does not promise that
        # kernels of different streams run concurrently (measured on B200: three ranks do, the fourth rank's first kernel
        # only starts once a spinning barrier of another rank has given up). That is a limit of the one-device stand-in,
        # not of the one-process-per-GPU path: no verdict rather than a false one.
        pytest.skip(f"virtual ranks were not co-scheduled on this device: {fails}")
    assert not fails, (fails, f"single-GPU iterations {wantIt}")
    torch.cuda.synchronize()
    return [(near[r].cpu().numpy(), noise[r].cpu().numpy(), its[r], errs[r]) for r in range(world)]


def _single(cuda, p, f, par, T, pref, seed2):
    from uammd_b200._lib import check
    from uammd_b200.md import _ptr, _stream_ptr
    import ctypes as C
    m = pse.PSE(p, par, sys=bd.System(7), force=f)
    N = p.shape[0]
    near = torch.zeros(N, 3, dtype=p.dtype, device=cuda)
    noise = torch.zeros_like(near)
    m.computeMFNearField(near)
    it = C.c_int(0)
    check(m.lib.ub200_pse_near_noise_add(m._h, _ptr(p), N, float(T), float(pref), seed2, _ptr(noise), C.byref(it), _stream_ptr()))
    torch.cuda.synchronize()
    return near.cpu().numpy(), noise.cpu().numpy(), it.value


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,shear,dtype", [
    (1, 0.0, np.float64), (2, 0.0, np.float64), (3, 0.15, np.float64),
    # single precision: written when the round's GPU minutes were spent - its first run is the driver's; not allowed to
    # stop the suite (-x) before it has passed once
    pytest.param(2, 0.15, np.float32, marks=pytest.mark.xfail(strict=False, reason="first execution pending (round 2 GPU budget spent)")),
])
def test_virtual_ranks_match_single_gpu(cuda, world, shear, dtype):
    N, L, T, pref, seed2 = 6000, 32.0, 0.9, 1.3, 0xC0FFEE
    tol = 1e-6 if dtype == np.float64 else 1e-4
    pos, force = _cloud(N, L, dtype)
    p, f = torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda)
    par = pse.Parameters(L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=tol, psi=0.6, temperature=T, dt=0.01, shearStrain=shear)
    wantNear, wantNoise, wantIt = _single(cuda, p, f, par, T, pref, seed2)
    got = _virtual_ranks(cuda, world, p, f, par, T, pref, seed2, calls=2, wantIt=wantIt)   # the second call reuses arena, slots and epochs
    eps = 1e-12 if dtype == np.float64 else 2e-5
    for r, (near, noise, it, err) in enumerate(got):
        assert err == 0, f"rank {r}: a peer barrier timed out"
        # the rows are the same sums in the same order: only the (particle-order vs row-order) placement differs
        assert _rel(near, wantNear) < eps, f"rank {r} near"
        # Lanczos: same Saru streams, dot products summed in another order; an iteration more or less moves the result
        # by at most the requested tolerance
        assert it >= 2 and abs(it - wantIt) <= 1
        assert _rel(noise, wantNoise) < (eps * 100 if it == wantIt else 20 * tol), f"rank {r} noise, {it} vs {wantIt} iterations"
        # every rank holds the same bits
        assert np.array_equal(near.view(np.uint8), got[0][0].view(np.uint8))
        assert np.array_equal(noise.view(np.uint8), got[0][1].view(np.uint8))


@pytest.mark.xfail(strict=False, reason="first execution pending (round 2 GPU budget spent)")
def test_dist_calls_need_setup(cuda):
    """products before create / attach / prepare are refused, as are more particles than the arena was sized for"""
    from uammd_b200._lib import UB200Error
    from uammd_b200.multigpu import DistributedPSENearField
    pos, force = _cloud(500, 32.0, np.float64)
    p, f = torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda)
    par = pse.Parameters(32.0, tolerance=1e-4, psi=0.6, temperature=1.0)
    a = DistributedPSENearField(p, par, sys=bd.System(1), rank=0, world=2)
    out = torch.zeros(500, 3, dtype=p.dtype, device=cuda)
    with pytest.raises(UB200Error):
        a.Mdot(f, out)                      # not attached, not prepared
    with pytest.raises(UB200Error):
        a.noiseAdd(out, 1.0, 1.0, 3)
    with pytest.raises(UB200Error):
        a.check(a.lib.ub200_pse_dist_create(a.pse._h, 0, 2, 500))   # twice
    other = pse.PSE(p, par, sys=bd.System(1))
    with pytest.raises(UB200Error):
        a.check(a.lib.ub200_pse_dist_create(other._h, 9, 2, 500))   # rank outside the world
    with pytest.raises(UB200Error):
        a.check(a.lib.ub200_pse_dist_create(other._h, 0, 9, 500))   # more ranks than one NVSwitch domain holds


def _pse_worker(rank, world, port, N, L, T, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from uammd_b200.multigpu import DistributedPSE
    dev = torch.device("cuda", rank)
    pos, force = _cloud(N, L, np.float32)
    p, f = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    par = pse.Parameters(L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=1e-3, psi=0.593, temperature=T, dt=1.0)
    m = DistributedPSE(p, par, sys=bd.System(11))
    MF = torch.zeros(N, 3, device=dev)
    for _ in range(2):
        m.sys = bd.System(11); m.sys.rng().next32(); m.sys.rng().next32()   # the generator state after the constructor draws
        it = m.computeHydrodynamicDisplacements(f, MF, T, 1.0)
    torch.cuda.synchronize()
    assert m.errorFlag() == 0
    np.save(out + f".{rank}.npy", MF.cpu().numpy())
    if rank == 0:
        np.save(out + ".it.npy", np.array([it]))
    m.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.xfail(strict=False, reason="first execution pending (round 2 GPU budget spent before a 2-GPU box could be used)")
@pytest.mark.parametrize("T", [0.0, 1.0])
def test_nccl_launched_pse_matches_single_gpu(tmp_path, T):
    """BASELINE config 3 end to end over the GPUs of the box: near rows + Lanczos over peer memory, far field on FFT slabs."""
    world = torch.cuda.device_count()
    world = 8 if world >= 8 else (4 if world >= 4 else world)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    N, L = 40000, 64.0
    out = str(tmp_path / "pse")
    mp.spawn(_pse_worker, args=(world, 29571 + int(T), N, L, T, out), nprocs=world, join=True)
    dev = torch.device("cuda:0")
    pos, force = _cloud(N, L, np.float32)
    p, f = torch.from_numpy(pos).to(dev), torch.from_numpy(force).to(dev)
    par = pse.Parameters(L, viscosity=1.0, hydrodynamicRadius=1.0, tolerance=1e-3, psi=0.593, temperature=T, dt=1.0)
    m = pse.PSE(p, par, sys=bd.System(11), force=f)
    want = torch.zeros(N, 3, device=dev)
    m.computeHydrodynamicDisplacements(f, want, T, 1.0)
    torch.cuda.synchronize()
    want = want.cpu().numpy()
    first = np.load(out + ".0.npy")
    for r in range(world):
        got = np.load(out + f".{r}.npy")
        assert np.array_equal(got.view(np.uint32), first.view(np.uint32)), f"rank {r} differs from rank 0"
        assert _rel(got, want) < (2e-5 if T == 0 else 2e-2), f"rank {r}: rel {_rel(got, want):.2e}"
