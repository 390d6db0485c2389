"""Cases of tests/test_lj_f64_gpu.py, run in a CHILD process (python tests/_lj_f64_cases.py prints one JSON object
{case: "ok" | error text}): ub200_lj_sum_f64, PairForces<LJ, CellList> in double precision (the path of a -DDOUBLE_PRECISION
reference build), against the oracle's fp64 pass on the same (single-precision-exact) coordinates. The pair arithmetic is
the same in both, only the order of the sums differs, so the bound is flat: 1e-12 of the largest force / energy / virial."""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT,) if p not in sys.path]

import numpy as np
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, PairForcesLJ64



def _check(orc, cuda, pos32, L, pot, periodic=(1, 1, 1), rel=1e-12):
    N = pos32.shape[0]
    box = Box(L); box.setPeriodicity(*periodic)
    dpos = torch.from_numpy(pos32.astype(np.float64)).to(cuda)
    force = torch.zeros(N, 4, dtype=torch.float64, device=cuda)
    e = torch.zeros(N, dtype=torch.float64, device=cuda)
    v = torch.zeros(N, dtype=torch.float64, device=cuda)
    PairForcesLJ64(box, pot).sum(dpos, force=force, energy=e, virial=v)
    torch.cuda.synchronize()
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, pot.getCutOff()), periodic)
    cl = orc.celllist_build(g, pos32)
    f64, e64, v64, _ = orc.lj_f64(g, cl, pot.table(), pot.ntypes, N)
    F = force.cpu().numpy()
    assert np.all(F[:, 3] == 0)
    assert np.abs(F[:, :3] - f64).max() <= rel * np.abs(f64).max()
    assert np.abs(e.cpu().numpy() - e64).max() <= rel * np.abs(e64).max()
    assert np.abs(v.cpu().numpy() - v64).max() <= rel * np.abs(v64).max()
    return F


def _lj(rc=2.5, shift=False):
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=rc, sigma=1.0, epsilon=1.0, shift=shift)
    return pot


def test_two_particle_kat(cuda):
    pos = np.zeros((2, 4)); pos[0, 0] = -0.5; pos[1, 0] = 0.5
    force = torch.zeros(2, 4, dtype=torch.float64, device=cuda)
    PairForcesLJ64(Box(20.0), _lj()).sum(torch.from_numpy(pos).to(cuda), force=force)
    assert np.allclose(force.cpu().numpy()[:, 0], [-24.0, 24.0], rtol=1e-14)


def test_jittered_fcc_matches_the_fp64_oracle(orc, cuda, N, shift):
    L = syn.lj_box_length(N)
    pos = syn.fcc_lattice(N, L)
    rng = np.random.default_rng(3)
    pos[:, :3] += rng.uniform(-0.08, 0.08, (N, 3)).astype(np.float32)
    _check(orc, cuda, pos, L, _lj(shift=shift))


def test_two_types_and_a_non_periodic_dimension(orc, cuda):
    N = 4 * 8 ** 3
    L = syn.lj_box_length(N)
    pos = syn.fcc_lattice(N, L)
    pos[:, :3] += np.random.default_rng(4).uniform(-0.05, 0.05, (N, 3)).astype(np.float32)
    pos[:, :3] *= 0.98          # stays inside the box in the open dimension
    pos[::3, 3] = 1
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=2.5, sigma=1.0, epsilon=1.0)
    pot.setPotParameters(0, 1, cutOff=2.0, sigma=0.875, epsilon=0.5, shift=True)
    pot.setPotParameters(1, 1, cutOff=1.5, sigma=0.75, epsilon=2.0)
    _check(orc, cuda, pos, L, pot, periodic=(1, 1, 0))


def test_small_box_collapses_to_one_cell(orc, cuda):
    """L <= 3 rc in every dimension: one cell, all pairs with the per-pair minimum image (CellList.cuh:117-122)"""
    N, L = 200, 7.0
    pos = np.zeros((N, 4), np.float32)
    pos[:, :3] = syn.uniform_cloud(N, L, seed=9)[:, :3]
    # keep clear of overlaps: forces of 1e12 would make the flat bound meaningless
    keep = [0]
    for i in range(1, N):
        d = pos[keep, :3] - pos[i, :3]
        d -= L * np.round(d / L)
        if (np.einsum("ij,ij->i", d, d) > 0.8 ** 2).all():
            keep.append(i)
    _check(orc, cuda, np.ascontiguousarray(pos[keep]), L, _lj())


def test_accumulates_like_transverser_set(cuda):
    pos = np.zeros((2, 4)); pos[0, 0] = -0.5; pos[1, 0] = 0.5
    force = torch.ones(2, 4, dtype=torch.float64, device=cuda)
    e = torch.full((2,), 3.0, dtype=torch.float64, device=cuda)
    pf = PairForcesLJ64(Box(20.0), _lj())
    pf.sum(torch.from_numpy(pos).to(cuda), force=force, energy=e)
    pf.sum(torch.from_numpy(pos).to(cuda), force=force, energy=e)
    F = force.cpu().numpy()
    assert np.allclose(F[:, 0], [1 - 48.0, 1 + 48.0], rtol=1e-14) and np.all(F[:, 3] == 1.0)
    # energy per particle of a pair at r = sigma: 0.5 * 4 eps (1 - 1) = 0 (unshifted)
    assert np.allclose(e.cpu().numpy(), 3.0, atol=1e-13)


CASES = {
    "two_particle_kat": lambda orc, cuda: test_two_particle_kat(cuda),
    "jittered_fcc_2048": lambda orc, cuda: test_jittered_fcc_matches_the_fp64_oracle(orc, cuda, 4 * 8 ** 3, False),
    "jittered_fcc_6912_shifted": lambda orc, cuda: test_jittered_fcc_matches_the_fp64_oracle(orc, cuda, 4 * 12 ** 3, True),
    "two_types_open_dimension": test_two_types_and_a_non_periodic_dimension,
    "one_cell_box": test_small_box_collapses_to_one_cell,
    "accumulates": lambda orc, cuda: test_accumulates_like_transverser_set(cuda),
}


def main():
    from oracle import oracle as orc
    orc.lib()
    cuda = torch.device("cuda:0")
    out = {}
    for name, fn in CASES.items():
        try:
            fn(orc, cuda)
            torch.cuda.synchronize()
            out[name] = "ok"
        except BaseException:  # noqa: BLE001 - every case reports, a broken context shows up in the later ones
            out[name] = traceback.format_exc()[-600:]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
