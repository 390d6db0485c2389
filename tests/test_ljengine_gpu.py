"""GPU: the LJ engine (ub200_ljengine_*: private half-cell list + TMA-staged column traversal, lj_column.cu) against the
cell traversal over the reference-layout list (pair_lj.cu, itself checked against the fp64 oracle and the compiled
reference in test_lj_gpu.py / test_ref_parity_gpu.py). Both evaluate the same pairs with the same pair arithmetic, so
they differ by fp32 summation order only; the bound used here is 1e-5 of the per-particle sum of |f_ij| (+ the pairs
inside the rounding band of the cut-off, which either side may count: oracle.LJScale)."""
import os

import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList, LJ, LJEngine, PairForces

pytestmark = pytest.mark.gpu


def _lj(rc=2.5, **kw):
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=rc, **kw)
    return pot


def _both(cuda, pos, L, pot, periodic=(1, 1, 1), stage=None):
    N = pos.shape[0]
    box = Box(L); box.setPeriodicity(*periodic)
    dpos = torch.from_numpy(pos).to(cuda)
    out = []
    old = os.environ.get("UB200_LJ_STAGE")
    if stage:
        os.environ["UB200_LJ_STAGE"] = stage
    try:
        for nl in (None, CellList()):
            pf = PairForces(pot, box, nl=nl)
            f = torch.zeros(N, 4, device=cuda); e = torch.zeros(N, device=cuda); v = torch.zeros(N, device=cuda)
            pf.sum(dpos, force=f, energy=e, virial=v)
            f2 = torch.zeros(N, 4, device=cuda)
            pf.sum(dpos, force=f2)
            torch.cuda.synchronize()
            # the force-only instantiation folds sigma / epsilon into two constants: same pairs, last-bit roundings differ
            scale = f[:, :3].abs().max() + 1e-30
            assert (f - f2).abs().max() <= 2e-5 * scale, "force-only and force+energy+virial instantiations differ"
            out.append((f.cpu().numpy(), e.cpu().numpy(), v.cpu().numpy(), pf))
    finally:
        if stage:
            if old is None:
                del os.environ["UB200_LJ_STAGE"]
            else:
                os.environ["UB200_LJ_STAGE"] = old
    return out


def _compare(orc, a, b, pos, L, pot, periodic=(1, 1, 1)):
    box = Box(L); box.setPeriodicity(*periodic)
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, pot.getCutOff()), periodic)
    cl = orc.celllist_build(g, pos)
    _, _, _, sc = orc.lj_f64(g, cl, pot.table(), pot.ntypes, pos.shape[0])
    tol = sc.force_tol(box.boxSize, pot.getCutOff())
    df = np.abs(a[0][:, :3] - b[0][:, :3]).max(axis=1)
    assert (df / tol).max() < 1.0, f"column vs cell traversal: {(df / tol).max():.3e} x fp32 tolerance"
    assert np.all(np.abs(a[1] - b[1]) <= pot.getCutOff() * tol + 1e-6 * np.abs(b[1]) + 1e-6)
    assert np.all(np.abs(a[2] - b[2]) <= 2 * pot.getCutOff() * tol + 1e-6 * np.abs(b[2]) + 1e-5)
    return df, sc


@pytest.mark.parametrize("stage", ["tma", "ldg"])
@pytest.mark.parametrize("N,L", [(30000, 33.5), (4000, (14.0, 29.0, 13.0)), (300, 12.6)])
def test_column_matches_cell_traversal(orc, cuda, N, L, stage):
    L = (L,) * 3 if np.isscalar(L) else L
    pos = syn.uniform_cloud(N, L, seed=5)
    pos[::5, :3] += np.float32(L) * np.array([2, -1, 1], np.float32)  # particles outside the primary box
    pot = _lj()
    col, cell = _both(cuda, pos, L, pot, stage=stage)
    assert col[3]._engine.lastPath() == "column" and col[3]._engine.errorFlag() == 0
    df, sc = _compare(orc, col, cell, pos, L, pot)
    # summation order only: far below the fp32 model on a cloud without close contacts dominating
    assert np.median(df / np.maximum(sc.abssum, 1e-30)) < 1e-6


def test_liquid_1e6_flat_tolerance(orc, cuda):
    """BASELINE.md 3.4 on the jittered FCC liquid at the headline size: rel-Linf <= 1e-5 of max|F| between the column
    engine and the cell traversal, and <= 1e-6 |F|inf... is asserted against the fp64 oracle in test_ref_parity_gpu."""
    N = 4 * 63 ** 3
    Lb = syn.lj_box_length(N, 0.8)
    pos = syn.fcc_lattice(N, Lb)
    pos[:, :3] += np.random.default_rng(3).normal(0, 0.06, (N, 3)).astype(np.float32)
    pot = _lj()
    box = Box(Lb)
    dpos = torch.from_numpy(pos).to(cuda)
    fc = torch.zeros(N, 4, device=cuda); fr = torch.zeros(N, 4, device=cuda)
    pfc, pfr = PairForces(pot, box), PairForces(pot, box, nl=CellList())
    pfc.sum(dpos, force=fc); pfr.sum(dpos, force=fr)
    torch.cuda.synchronize()
    assert pfc._engine.lastPath() == "column" and pfc._engine.grid() == (86, 86, 86)
    fc, fr = fc.cpu().numpy()[:, :3], fr.cpu().numpy()[:, :3]
    assert np.abs(fc - fr).max() <= 1e-5 * np.abs(fr).max()
    assert np.abs(fc.astype(np.float64).sum(0)).max() < 1e-3 * np.abs(fr).max()  # Newton's third law over the box


def test_non_periodic_and_thin_boxes(orc, cuda):
    pot = _lj()
    for L, per, N in [((30.0, 30.0, 30.0), (1, 0, 1), 20000), ((30.0, 7.0, 26.0), (1, 1, 1), 4000),
                      ((26.0, 26.0, 26.0), (0, 0, 0), 12000), ((40.0, 40.0, 0.0), (1, 1, 0), 1200)]:
        pos = syn.uniform_cloud(N, tuple(l if l > 0 else 1.0 for l in L), seed=N)
        if L[2] == 0.0:
            pos[:, 2] = 0.0
        col, cell = _both(cuda, pos, L, pot, periodic=per)
        assert col[3]._engine.lastPath() == "column", (L, per)
        _compare(orc, col, cell, pos, L, pot, per)


def test_small_periodic_box_falls_back(orc, cuda):
    """A periodic dimension under five half cells: the engine takes the reference-layout traversal; boxes <= 3 cut-offs
    take the all-pairs path (PairForces.cu:49-53)."""
    pot = _lj()
    pos = syn.uniform_cloud(700, (30.0, 5.5, 30.0), seed=3)
    col, cell = _both(cuda, pos, (30.0, 5.5, 30.0), pot)
    assert col[3]._engine.lastPath() == "cell"
    assert np.array_equal(col[0], cell[0])
    pos = syn.uniform_cloud(200, (7.0, 7.0, 7.0), seed=4)
    col, cell = _both(cuda, pos, (7.0, 7.0, 7.0), pot)
    assert col[3]._engine.lastPath() == "nbody"
    assert np.allclose(col[0], cell[0], rtol=1e-4, atol=1e-3 * np.abs(cell[0]).max())


def test_dense_columns_take_direct_path_and_multitype(orc, cuda):
    L = (30.0, 30.0, 30.0)
    pos = syn.uniform_cloud(6000, L, seed=10, ntypes=3)
    rng = np.random.default_rng(0)
    k = np.arange(2744)
    lat = np.stack([k % 14, (k // 14) % 14, k // 196], -1) * 0.17 + 1.0   # > 416 candidates around these columns
    pos[:2744, :3] = (lat + rng.random((2744, 3)) * 0.02).astype(np.float32)
    pot = LJ()
    for a in range(3):
        for b in range(a, 3):
            pot.setPotParameters(a, b, cutOff=2.0 + 0.25 * (a + b), sigma=0.9 + 0.1 * a + 0.05 * b,
                                 epsilon=1.0 + 0.5 * a * b, shift=(a == b))
    col, cell = _both(cuda, pos, L, pot)
    assert col[3]._engine.lastPath() == "column"
    _compare(orc, col, cell, pos, L, pot)


def test_owner_restriction_and_write_mode(cuda):
    N, L = 20000, (30.0,) * 3
    pos = torch.from_numpy(syn.uniform_cloud(N, L, seed=12)).to(cuda)
    pot, box = _lj(), Box(L)
    eng = LJEngine()
    full = torch.zeros(N, 4, device=cuda)
    eng.sum(pos, box, pot.table(), 1, force=full)
    part = torch.full((N, 4), 7.0, device=cuda)
    eng.sum(pos, box, pot.table(), 1, force=part, accumulate=False, owner=(5000, 12000))
    torch.cuda.synchronize()
    assert torch.equal(part[5000:12000, :3], full[5000:12000, :3]) and torch.all(part[5000:12000, 3] == 0)
    assert torch.all(part[:5000] == 7.0) and torch.all(part[12000:] == 7.0)
    # group indirection: a subset of the particles, outputs scattered to their global slots
    idx = torch.arange(0, N, 3, device=cuda, dtype=torch.int32)
    sub = torch.zeros(N, 4, device=cuda)
    eng.sum(pos, box, pot.table(), 1, force=sub, groupIndex=idx, globalIndex=idx)
    ref = torch.zeros(idx.numel(), 4, device=cuda)
    eng.sum(pos[idx.long()].contiguous(), box, pot.table(), 1, force=ref)
    torch.cuda.synchronize()
    assert torch.equal(sub[idx.long()], ref)
    mask = torch.ones(N, dtype=torch.bool, device=cuda); mask[idx.long()] = False
    assert torch.all(sub[mask] == 0)
