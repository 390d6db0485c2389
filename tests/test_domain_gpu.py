"""GPU: brick domain decomposition with ghost-cell halo exchange (uammd_b200/domain.py, csrc/domain.cu) and the NBody
fallback, through the C ABI.

One GPU is enough: the virtual ranks of tests/_checker_engines.lockstep run the real step generators with the CUDA
engines and do the all-to-all by hand, so every kernel of the decomposed path (classification, cell list over
[owned | ghosts], owner-restricted traversal, DPD noise keyed on global ids, half kicks) is exercised. The bar is the
single-GPU trajectory BIT FOR BIT. tests/test_multigpu_gpu.py runs the same over NCCL when more GPUs are visible."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "tests")) if p not in sys.path]

from uammd_b200 import synthetic as syn  # noqa: E402
from uammd_b200.md import Box, CellList, DPD, LJ, PairForces, VerletNVE  # noqa: E402

pytestmark = pytest.mark.gpu


def _lj(rc=2.5, shift=False):
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=rc, shift=shift)
    return pot


@pytest.mark.parametrize("rankGrid", [(2, 2, 2), (3, 1, 2)])
def test_brick_classify_matches_oracle(orc, cuda, rankGrid):
    """Integer work: cell, owner and ghost mask bit exact against the C restatement (uniform cloud incl. positions
    outside the primary box, which the reference folds in Grid::getCell)."""
    from uammd_b200.domain import CudaLJBrickEngine
    N = 200000
    Lb = syn.lj_box_length(N, 0.8)
    pos = syn.uniform_cloud(N, Lb, seed=3)
    pos[::7, :3] += np.float32(Lb)          # unfolded coordinates
    pos[::11, :3] -= np.float32(2 * Lb)
    box = Box(Lb)
    eng = CudaLJBrickEngine(box, _lj(), 0.005)
    dpos = torch.from_numpy(pos).to(cuda)
    cell = torch.empty(N, dtype=torch.int32, device=cuda)
    owner = torch.empty(N, dtype=torch.int32, device=cuda)
    mask = torch.empty(N, dtype=torch.int32, device=cuda)
    eng.check(eng.lib.ub200_brick_classify_f32(eng._ptr(dpos), N, eng.f3(box.boxSize), eng.i3([1, 1, 1]), eng.i3(eng.cellDim),
                                               eng.i3(rankGrid), eng._ptr(cell), eng._ptr(owner), eng._ptr(mask), eng._stream()))
    torch.cuda.synchronize()
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, 2.5))
    assert tuple(g.cellDim) == tuple(eng.cellDim)
    ocell, oowner, omask = orc.brick_classify(g, pos, rankGrid)
    assert np.array_equal(cell.cpu().numpy(), ocell)
    assert np.array_equal(owner.cpu().numpy(), oowner)
    assert np.array_equal(mask.cpu().numpy().view(np.uint32), omask)
    o2, m2 = eng.classify(dpos, rankGrid)
    assert torch.equal(o2, owner) and torch.equal(m2, mask)


def test_brick_classify_rejects_bad_rank_grids(cuda):
    from uammd_b200._lib import UB200Error
    from uammd_b200.domain import CudaLJBrickEngine
    box = Box(10.0)
    eng = CudaLJBrickEngine(box, _lj(), 0.005)      # 4 cells per dimension
    pos = torch.zeros(8, 4, device=cuda)
    with pytest.raises(UB200Error):
        eng.classify(pos, (5, 1, 1))                # a brick without a cell
    with pytest.raises(UB200Error):
        eng.classify(pos, (4, 4, 4))                # 64 ranks: more than the 32 mask bits


@pytest.mark.parametrize("rankGrid", [(2, 1, 1), (2, 2, 2)])
def test_lj_bricks_bit_identical_to_single_gpu(cuda, rankGrid):
    from _checker_engines import gather_lockstep, lockstep
    from uammd_b200.domain import CudaLJBrickEngine, DomainDecomposedMD
    N, steps, dt = 4 * 16 ** 3, 6, 0.004
    Lb = syn.lj_box_length(N)
    pos = syn.fcc_lattice(N, Lb)
    pos[:, :3] += np.random.default_rng(5).normal(0, 0.05, (N, 3)).astype(np.float32)
    vel = syn.maxwell_velocities(N, 1.5)
    box, pot = Box(Lb), _lj()
    world = rankGrid[0] * rankGrid[1] * rankGrid[2]
    ranks = [DomainDecomposedMD(CudaLJBrickEngine(box, pot, dt), N, r, world, rankGrid) for r in range(world)]
    for r in ranks:
        r.setGlobalState(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda))
    assert sum(r.nOwned for r in ranks) == N
    for _ in range(steps):
        lockstep(ranks)
    torch.cuda.synchronize()
    gp, gv = gather_lockstep(ranks, N)
    p, v = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda)
    nve = VerletNVE(p, v, dt)
    nve.addInteractor(PairForces(pot, box, nl=CellList()))  # the cell traversal these (legacy) bricks use: same bits
    for _ in range(steps):
        nve.forwardTime()
    torch.cuda.synchronize()
    assert np.array_equal(gp.view(np.uint32), p.cpu().numpy().view(np.uint32))
    assert np.array_equal(gv.view(np.uint32), v.cpu().numpy().view(np.uint32))
    assert all(0 < r.stats["ghosts"] for r in ranks) and all(r.pos.shape[0] < N for r in ranks)


def test_dpd_bricks_bit_identical_to_single_gpu(cuda):
    """DPD at rho = 3 (BASELINE config 4 shape, scaled down): ghosts carry velocities, the noise is keyed on global ids,
    particles migrate between bricks during the run."""
    from _checker_engines import gather_lockstep, lockstep
    from uammd_b200.domain import CudaDPDBrickEngine, DomainDecomposedMD
    from uammd_b200.multigpu import DistributedDPDMD
    N, steps, dt = 24000, 12, 0.01
    L = (N / 3.0) ** (1.0 / 3.0)
    pos, vel = syn.uniform_cloud(N, L, seed=21), syn.maxwell_velocities(N, 1.0, seed=22)
    box = Box(L)
    mk = lambda: DPD(cutOff=1.0, dt=dt, gamma=4.5, temperature=1.0, A=25.0, seed=99)
    rankGrid = (2, 2, 2)
    ranks = [DomainDecomposedMD(CudaDPDBrickEngine(box, mk(), dt), N, r, 8, rankGrid) for r in range(8)]
    for r in ranks:
        r.setGlobalState(torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda))
    for _ in range(steps):
        lockstep(ranks)
    torch.cuda.synchronize()
    gp, gv = gather_lockstep(ranks, N)
    p, v, f = torch.from_numpy(pos).to(cuda), torch.from_numpy(vel).to(cuda), torch.zeros(N, 4, device=cuda)
    single = DistributedDPDMD(box, mk(), dt, N)
    for _ in range(steps):
        single.forwardTime(p, v, f)
    torch.cuda.synchronize()
    assert np.array_equal(gp.view(np.uint32), p.cpu().numpy().view(np.uint32))
    assert np.array_equal(gv.view(np.uint32), v.cpu().numpy().view(np.uint32))
    assert sum(r.stats["migrated"] for r in ranks) > 0, "the test must exercise migration"


@pytest.mark.parametrize("N,shift", [(300, False), (257, True), (1, False)])
def test_nbody_fallback_matches_oracle(orc, cuda, N, shift):
    """Box <= 3 cut-offs in every dimension: PairForces::sumTransverser takes NBody (PairForces.cu:49-53). Oracle: the fp64
    restatement over a one-cell grid (all pairs, per-pair minimum image), fp32 tolerance model of tests/test_lj_gpu.py."""
    L = 7.0
    pot = _lj(shift=shift)
    n = int(np.ceil(N ** (1 / 3)))
    rng = np.random.default_rng(N)
    ijk = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)[:N]
    pos = np.zeros((N, 4), np.float32)
    pos[:, :3] = ((ijk + 0.5) * (L / n) - 0.5 * L + 0.2 * (L / n) * rng.uniform(-0.5, 0.5, (N, 3))).astype(np.float32)
    box = Box(L)
    dpos = torch.from_numpy(pos).to(cuda)
    force, e, v = torch.zeros(N, 4, device=cuda), torch.zeros(N, device=cuda), torch.zeros(N, device=cuda)
    force[:, 0] = 1.0                                   # sum() accumulates like Transverser::set
    PairForces(pot, box).sum(dpos, force=force, energy=e, virial=v)
    torch.cuda.synchronize()
    g = orc.make_grid_f(box.boxSize, (1, 1, 1))
    cl = orc.celllist_build(g, pos)
    f64, e64, v64, sc = orc.lj_f64(g, cl, pot.table(), 1, N)
    F = force.cpu().numpy()
    F[:, 0] -= 1.0
    tol = sc.force_tol(box.boxSize, 2.5) + 1e-6 * np.abs(f64).max(initial=0.0)
    assert np.all(np.abs(F[:, :3] - f64).max(axis=1) <= tol)
    assert np.all(np.abs(e.cpu().numpy() - e64) <= 2.5 * tol + 1e-6 * np.abs(e64) + 1e-6)
    assert np.all(np.abs(v.cpu().numpy() - v64) <= 5.0 * tol + 1e-6 * np.abs(v64) + 1e-5)
    if N > 1:
        assert np.abs(f64).max() > 1.0                  # the system does interact
    # group index list: only the listed particles interact and are written
    idx = torch.arange(0, N, 2, dtype=torch.int32, device=cuda)
    f2 = torch.zeros(N, 4, device=cuda)
    PairForces(pot, box).sumNBody(dpos, force=f2, globalIndex=idx)
    torch.cuda.synchronize()
    sub = pos[::2].copy()
    cl2 = orc.celllist_build(g, sub)
    f64s, _, _, sc2 = orc.lj_f64(g, cl2, pot.table(), 1, sub.shape[0])
    F2 = f2.cpu().numpy()
    assert np.all(F2[1::2] == 0)
    assert np.all(np.abs(F2[::2, :3] - f64s).max(axis=1) <= sc2.force_tol(box.boxSize, 2.5) + 1e-6 * np.abs(f64s).max(initial=0.0))
