"""GPU: brick domain decomposition with the device-side halo exchange (ub200_brick_*, ub200_halo_exchange_*).

Oracle: the single-GPU fused engine (LJMD), itself checked against the fp64 oracle and the compiled reference. The
bricks must reproduce its trajectory BIT FOR BIT for every rank grid. Virtual ranks (several handles in one process,
driven phase by phase on one stream) exercise every kernel of the path on one GPU; the NCCL-launched test below does the
same over real peer mappings when at least two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.brickmd import BrickDPDMD, BrickLJMD, assemble
from uammd_b200.md import Box, DPD, LJ, LJMD, PairForcesDPD, VerletNVE

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _system(n, jitter=0.05, seed=1):
    N = 4 * n ** 3
    Lb = syn.lj_box_length(N, 0.8)
    pos = syn.fcc_lattice(N, Lb)
    pos[:, :3] += np.random.default_rng(seed).normal(0, jitter, (N, 3)).astype(np.float32)
    vel = syn.maxwell_velocities(N, 1.0, seed=7)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    return N, Lb, pos, vel, pot


def _single(cuda, N, Lb, pos, vel, pot, steps, dt=0.005):
    """Single-GPU engine with eight lanes per particle in every pass (UB200_LJ_WIDEN=0), the summation order the bricks
    use: by default the last pass of a column spreads one or two left-over particles over 16 / 32 lanes, which makes the
    last bits of a force depend on what else shares the particle's column."""
    old = os.environ.get("UB200_LJ_WIDEN")
    os.environ["UB200_LJ_WIDEN"] = "0"
    try:
        md = LJMD(Box(Lb), pot, dt)
        p, v, f = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), torch.zeros(N, 4, device=cuda)
        md.run(p, v, f, steps)
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["UB200_LJ_WIDEN"]
        else:
            os.environ["UB200_LJ_WIDEN"] = old
    return p.cpu().numpy(), v.cpu().numpy(), f.cpu().numpy()


@pytest.mark.parametrize("rankGrid", [(1, 1, 1), (2, 1, 1), (1, 2, 2), (2, 2, 2), (1, 1, 3), (4, 2, 1)])
def test_virtual_ranks_bit_identical_to_single_gpu(cuda, rankGrid):
    N, Lb, pos, vel, pot = _system(14)
    steps = 25
    world = int(np.prod(rankGrid))
    ranks = [BrickLJMD(Box(Lb), pot, 0.005, N, r, world, rankGrid) for r in range(world)]
    BrickLJMD.connectLocal(ranks)
    dp, dv = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda)
    for r in ranks:
        r.setGlobalState(dp, dv)
    BrickLJMD.runLocal(ranks, steps)
    parts = [r.owned() for r in ranks]
    for r in ranks:
        no, nl, err = r.counts()
        assert err == 0 and 0 < no <= nl
    p, v, f = assemble(parts, N)
    ps, vs, fs = _single(cuda, N, Lb, pos, vel, pot, steps)
    assert np.array_equal(p, ps) and np.array_equal(v, vs), "trajectory differs from the single-GPU engine"
    assert np.array_equal(f[:, :3], fs[:, :3])
    if world > 1:
        assert sum(r.counts()[1] - r.counts()[0] for r in ranks) > 0  # ghosts were exchanged


@pytest.mark.parametrize("graph", ["1", "0"])
def test_one_rank_run_entry_point_with_and_without_graph(cuda, graph):
    """ub200_brick_lj_nve_run_f32 (the entry point multi-process runs use): a captured CUDA graph of four steps replayed on
    an internal stream, closing kicks fused into the next step's push; 26 steps = 6 graphs + 2 plain steps. One rank is a
    1 x 1 x 1 brick exchanging with itself: every kernel of the path runs, and the result must be the single-GPU bits."""
    N, Lb, pos, vel, pot = _system(14)
    steps = 26
    old = os.environ.get("UB200_BRICK_GRAPH")
    os.environ["UB200_BRICK_GRAPH"] = graph
    try:
        md = BrickLJMD(Box(Lb), pot, 0.005, N, 0, 1, (1, 1, 1))
    finally:
        if old is None:
            del os.environ["UB200_BRICK_GRAPH"]
        else:
            os.environ["UB200_BRICK_GRAPH"] = old
    md.setGlobalState(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda))
    md.run(10)
    md.run(steps - 10)
    p, v, f = assemble([md.owned()], N)
    assert md.counts()[2] == 0
    ps, vs, fs = _single(cuda, N, Lb, pos, vel, pot, steps)
    assert np.array_equal(p, ps) and np.array_equal(v, vs) and np.array_equal(f[:, :3], fs[:, :3])


def test_particles_migrate_between_bricks(cuda):
    """A drifting gas: after enough steps a good share of the particles has changed owner; nothing is lost or duplicated
    and the trajectory still equals the single-GPU one."""
    N, Lb, pos, vel, pot = _system(10, jitter=0.02)
    vel[:, 0] += 3.0  # common drift along x: every particle crosses brick boundaries in turn
    steps, world, rankGrid = 120, 4, (2, 2, 1)
    ranks = [BrickLJMD(Box(Lb), pot, 0.005, N, r, world, rankGrid) for r in range(world)]
    BrickLJMD.connectLocal(ranks)
    dp, dv = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda)
    for r in ranks:
        r.setGlobalState(dp, dv)
    first = [set(r.owned()[2].cpu().numpy().tolist()) for r in ranks]
    BrickLJMD.runLocal(ranks, steps)
    parts = [r.owned() for r in ranks]
    moved = sum(len(first[k] - set(parts[k][2].cpu().numpy().tolist())) for k in range(world))
    assert moved > N // 20
    p, v, _ = assemble(parts, N)
    ps, vs, _ = _single(cuda, N, Lb, pos, vel, pot, steps)
    assert np.array_equal(p, ps) and np.array_equal(v, vs)


@pytest.mark.parametrize("rankGrid", [(1, 1, 1), (2, 1, 1), (2, 2, 2)])
def test_dpd_bricks_bit_identical_to_single_gpu(cuda, rankGrid):
    """BASELINE config 4 scaled down (DPD fluid at rho = 3, rc = 1, A = 25, gamma = 4.5, kT = 1, dt = 0.01): ghosts carry
    velocities, the pair noise is keyed on global ids, particles migrate during the run. Oracle: VerletNVE +
    PairForcesDPD on one GPU (itself checked against the reference's compiled transverser)."""
    N, steps, dt = 24000, 15, 0.01
    L = (N / 3.0) ** (1.0 / 3.0)
    pos, vel = syn.uniform_cloud(N, L, seed=21), syn.maxwell_velocities(N, 1.0, seed=22)
    box = Box(L)
    mk = lambda: DPD(cutOff=1.0, dt=dt, gamma=4.5, temperature=1.0, A=25.0, seed=4321)
    world = int(np.prod(rankGrid))
    ranks = [BrickDPDMD(box, mk(), dt, N, r, world, rankGrid) for r in range(world)]
    BrickLJMD.connectLocal(ranks)
    dp, dv = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda)
    for r in ranks:
        r.setGlobalState(dp, dv)
    BrickLJMD.runLocal(ranks, steps)
    parts = [r.owned() for r in ranks]
    assert all(r.counts()[2] == 0 for r in ranks)
    p, v, f = assemble(parts, N)

    class _It:  # Interactor over the single-GPU DPD path: VerletNVE hands it positions, it needs the velocities too
        def __init__(self, vel):
            self.pf, self.vel = PairForcesDPD(mk(), box), vel

        def sum(self, pos, force=None, **kw):
            self.pf.sum(pos, self.vel, force)

    ps, vs = dp.clone(), dv.clone()
    nve = VerletNVE(ps, vs, dt)
    nve.addInteractor(_It(vs))
    for _ in range(steps):
        nve.forwardTime()
    torch.cuda.synchronize()
    assert np.array_equal(p, ps.cpu().numpy()) and np.array_equal(v, vs.cpu().numpy())
    assert np.array_equal(f[:, :3], nve.force.cpu().numpy()[:, :3])


def test_unsupported_decompositions_are_reported(cuda):
    from uammd_b200._lib import UB200Error
    _, Lb, _, _, pot = _system(6)
    with pytest.raises(UB200Error):
        BrickLJMD(Box(Lb), pot, 0.005, 864, 0, 8, (8, 1, 1))   # bricks thinner than two half cells / window wider than the grid
    with pytest.raises(UB200Error):
        BrickLJMD(Box(Lb), pot, 0.005, 864, 0, 16, (4, 2, 2))  # more than 8 ranks


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_nccl_launched_bricks_match_single_gpu():
    n = min(torch.cuda.device_count(), 8)
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "scripts", "brick_lj.py"),
                          "--cells", "20", "--steps", "30", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit_identical\": true" in out.stdout
