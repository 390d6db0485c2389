"""GPU: LJ pair traversal through the C ABI vs the oracle.

Tolerance (fp32), per particle: |F_gpu - F_fp64| <= 2^-23 (L + 8 rc) sum_j |df_ij/dr| + 2e-6 sum_j |f_ij|
(oracle.LJScale.force_tol): separations are differences of fp32 positions of box scale, pairs across the
periodic boundary carry one more rounding of size ulp(L), and the LJ force is r^-13 steep, so the error of a
pair term is its r-derivative times that separation error. The reference's own fp32 arithmetic obeys the same
bound and no tighter one; tests/test_ref_parity_gpu.py checks the compiled reference under the same metric.
"""
import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, CellList, LJ, PairForces

pytestmark = pytest.mark.gpu
RC_MAX = 3.1


def _lj(rc=2.5, sigma=1.0, eps=1.0, shift=False):
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=rc, sigma=sigma, epsilon=eps, shift=shift)
    return pot


def _run(orc, cuda, pos, L, pot, periodic=(1, 1, 1), energy=True, virial=True):
    N = pos.shape[0]
    box = Box(L); box.setPeriodicity(*periodic)
    pf = PairForces(pot, box)
    dpos = torch.from_numpy(pos).to(cuda)
    force = torch.zeros(N, 4, device=cuda)
    e = torch.zeros(N, device=cuda) if energy else None
    v = torch.zeros(N, device=cuda) if virial else None
    pf.sum(dpos, force=force, energy=e, virial=v)
    torch.cuda.synchronize()
    g = orc.make_grid_f(box.boxSize, orc.neighbour_celldim(box.boxSize, pot.getCutOff()), periodic)
    cl = orc.celllist_build(g, pos)
    f64, e64, v64, sc = orc.lj_f64(g, cl, pot.table(), pot.ntypes, N)
    F = force.cpu().numpy()
    assert np.all(F[:, 3] == 0)
    tol = sc.force_tol(box.boxSize, pot.getCutOff())
    err = np.abs(F[:, :3] - f64).max(axis=1) / tol
    assert err.max() < 1.0, f"force error {err.max():.3e} x tolerance"
    a = sc.abssum
    if energy:   # e_ij ~ f_ij r / 6..12: the force scale times rc bounds the energy error
        ee = np.abs(e.cpu().numpy() - e64)
        assert np.all(ee <= pot.getCutOff() * tol + 1e-6 * np.abs(e64) + 1e-6)
    if virial:
        vv = np.abs(v.cpu().numpy() - v64)
        assert np.all(vv <= 2 * pot.getCutOff() * tol + 1e-6 * np.abs(v64) + 1e-5)
    return F, f64, a


def test_two_particle_kat(cuda):
    pot = _lj()
    pos = np.zeros((2, 4), np.float32); pos[0, 0] = -0.5; pos[1, 0] = 0.5
    force = torch.zeros(2, 4, device=cuda)
    PairForces(pot, Box(20.0)).sum(torch.from_numpy(pos).to(cuda), force=force)
    assert np.allclose(force.cpu().numpy()[:, 0], [-24.0, 24.0], rtol=1e-6)


@pytest.mark.parametrize("N", [100, 5000, 100000])
def test_liquid_density_cloud(orc, cuda, N):
    Lb = max(syn.lj_box_length(N, 0.8), 10.5)
    _run(orc, cuda, syn.uniform_cloud(N, Lb, seed=N + 1), (Lb,) * 3, _lj())


def test_fcc_lattice_forces_cancel(orc, cuda):
    N = 4 * 20 ** 3
    Lb = syn.lj_box_length(N, 0.8)
    F, f64, a = _run(orc, cuda, syn.fcc_lattice(N, Lb), (Lb,) * 3, _lj())
    assert np.abs(F[:, :3]).max() < 1e-3 * a.max()  # perfect lattice: net force ~ 0


def test_accumulates_into_existing_forces(orc, cuda):
    N = 3000
    L = (14.0,) * 3
    pos = syn.uniform_cloud(N, L, seed=2)
    pot = _lj()
    pf = PairForces(pot, Box(L))
    dpos = torch.from_numpy(pos).to(cuda)
    f1 = torch.zeros(N, 4, device=cuda)
    pf.sum(dpos, force=f1)
    f2 = torch.full((N, 4), 1.5, device=cuda)
    pf.sum(dpos, force=f2)
    assert torch.allclose(f2[:, :3], f1[:, :3] + 1.5, rtol=1e-5, atol=1e-3)
    assert torch.all(f2[:, 3] == 1.5)  # Transverser::set adds make_real4(F, 0)


def test_shifted_potential_and_scaled_parameters(orc, cuda):
    N = 20000
    Lb = 40.0
    _run(orc, cuda, syn.uniform_cloud(N, Lb, seed=4), (Lb,) * 3, _lj(rc=3.1, sigma=1.15, eps=0.7, shift=True))


def test_multiple_types(orc, cuda):
    N = 20000
    Lb = 32.0
    pos = syn.uniform_cloud(N, Lb, seed=6, ntypes=3)
    pot = LJ()
    for a in range(3):
        for b in range(a, 3):
            pot.setPotParameters(a, b, cutOff=2.0 + 0.25 * (a + b), sigma=0.9 + 0.1 * a + 0.05 * b,
                                 epsilon=1.0 + 0.5 * a * b, shift=(a == b))
    _run(orc, cuda, pos, (Lb,) * 3, pot)


def test_collapsed_dimension_uses_pair_minimum_image(orc, cuda):
    L = (30.0, 7.0, 26.0)  # y collapses to one (periodic) cell: per-pair MIC path
    _run(orc, cuda, syn.uniform_cloud(15000, L, seed=7), L, _lj())


def test_non_periodic_dimension(orc, cuda):
    L = (30.0, 30.0, 30.0)
    _run(orc, cuda, syn.uniform_cloud(20000, L, seed=8), L, _lj(), periodic=(1, 0, 1))


def test_particles_outside_primary_box(orc, cuda):
    L = (24.0, 24.0, 24.0)
    pos = syn.uniform_cloud(12000, L, seed=9)
    pos[::4, :3] += np.float32(24.0) * np.array([1, -2, 3], np.float32)
    _run(orc, cuda, pos, L, _lj())


def test_dense_cluster_takes_direct_path(orc, cuda):
    # > 1024 candidates around one cell: exercises the un-staged (global memory) traversal
    L = (30.0, 30.0, 30.0)
    pos = syn.uniform_cloud(6000, L, seed=10)
    rng = np.random.default_rng(0)
    k = np.arange(2744)
    lat = np.stack([k % 14, (k // 14) % 14, k // 196], -1) * 0.17 + 1.0
    pos[:2744, :3] = (lat + rng.random((2744, 3)) * 0.02).astype(np.float32)
    _run(orc, cuda, pos, L, _lj())


def test_coincident_particles_give_zero_like_reference(orc, cuda):
    L = (20.0, 20.0, 20.0)
    pos = syn.uniform_cloud(2000, L, seed=11)
    pos[1] = pos[0]  # r2 == 0 -> pair skipped (RadialPotential.cuh:111-113)
    F, f64, a = _run(orc, cuda, pos, L, _lj())
    assert np.all(np.isfinite(F))


def test_full_size_1e6_newton_third_law(orc, cuda):
    """BASELINE config 2 (uniform cloud): parity vs fp64 oracle + sum of forces vanishes."""
    N = 1_000_000
    Lb = syn.lj_box_length(N)
    F, f64, a = _run(orc, cuda, syn.uniform_cloud(N, Lb, seed=2024), (Lb,) * 3, _lj(), energy=False, virial=False)
    tot = F[:, :3].astype(np.float64).sum(0)
    assert np.abs(tot).max() < 1e-6 * a.sum()
