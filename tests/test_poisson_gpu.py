"""GPU: spectral Ewald Poisson solver (uammd_b200/poisson.py -> ub200_poisson_*) against
  * the reference tests' analytic known answers (test/Potentials/Poisson/TriplyPeriodic/test_poisson.cu:13-23,192-222: field of
    two Gaussian charges, 1e-3 relative),
  * the numpy restatement of the reference algorithm (oracle/oracle_poisson.py) on random charge clouds,
  * its own invariants: the result does not depend on the Ewald split (to the tolerance), forces are q E, energies q phi.
Tolerances are written next to each assertion."""
import math

import numpy as np
import pytest
import torch

from uammd_b200._lib import UB200Error
from uammd_b200.poisson import Parameters, Poisson

pytestmark = pytest.mark.gpu


def _theoretical_field(r, gw):     # test_poisson.cu:13-18
    return -math.exp(-r * r / (4.0 * gw * gw)) / (4 * math.pi * math.sqrt(math.pi) * gw * r) - \
        math.erf(r / (2.0 * gw)) / (4 * math.pi * r * r)


def _three_charges(L, r, seed, dtype, cuda):
    ori = (np.random.default_rng(seed).random(3) - 0.5) * L
    pos = np.zeros((3, 4))
    pos[0, :3] = ori + [-0.5 * r, 0, 0]
    pos[1, :3] = ori + [0.5 * r, 0, 0]
    pos[2, :3] = ori + [0.5 * r, 0, 0]
    q = np.array([1.0, -0.5, -0.5])
    return torch.from_numpy(pos.astype(dtype)).to(cuda), torch.from_numpy(q.astype(dtype)).to(cuda)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_two_charges_field_known_answer(cuda, seed):
    """PoissonTest.SingleSimulationTest (test_poisson.cu:192-222), double precision like the reference's test build."""
    L, r, tol, gw, split = 100.0, 2.0, 1e-7, 0.001, 0.2
    pos, q = _three_charges(L, r, seed, np.float64, cuda)
    p = Poisson(pos, q, Parameters(L, epsilon=1.0, tolerance=tol, gw=gw, split=split))
    force = torch.zeros(3, 4, dtype=torch.float64, device=cuda)
    p.sum(force=force)
    f = force.cpu().numpy()[0]
    want = _theoretical_field(r, gw)
    assert abs(f[1]) < 1e-10 and abs(f[2]) < 1e-10 and f[0] > 0
    assert abs(1.0 - abs(f[0] / want)) < 1e-3
    fp = p.computeFieldPotentialAtParticles().cpu().numpy()[0]
    assert abs(fp[1]) < 1e-10 and abs(fp[2]) < 1e-10 and fp[0] > 0
    assert abs(1.0 - abs(fp[0] / want)) < 1e-3
    assert abs(f[0] - fp[0]) < 1e-12 * abs(want)           # q_0 = 1: force and field are the same numbers


def _cloud(N, L, seed):
    rng = np.random.default_rng(seed)
    pos = np.zeros((N, 4))
    pos[:, :3] = (rng.random((N, 3)) - 0.5) * L
    q = rng.choice([-1.0, 1.0], N)
    q[-1] -= q.sum()                                        # neutral
    return pos, q


@pytest.mark.parametrize("split", [-1.0, 0.5, 0.9])
def test_random_cloud_matches_the_restatement(cuda, split):
    """Field and potential at 300 random charges against the numpy restatement (exact Green's functions, numpy FFT); without
    splitting (far field only, gw resolved on the grid) and with two splits. 1e-9 of the largest value far-field only (same
    arithmetic, different FFT), 2e-5 with the near field (the product interpolates the reference's 4096+ point tables)."""
    from oracle.oracle_poisson import PoissonOracle
    N, L, gw, tol = 300, 40.0, 0.5, 1e-6
    pos, q = _cloud(N, L, 5)
    orc = PoissonOracle(L, 1.0, tol, gw, split)
    want = orc.field_potential(pos, q)
    dpos, dq = torch.from_numpy(pos).to(cuda), torch.from_numpy(q).to(cuda)
    p = Poisson(dpos, dq, Parameters(L, epsilon=1.0, tolerance=tol, gw=gw, split=split))
    inf = p.info()
    assert tuple(inf.cells) == tuple(orc.cells) and inf.support == orc.support
    assert abs(inf.nearFieldCutOff - orc.nearCut) < 1e-9 * max(1.0, orc.nearCut)
    got = p.computeFieldPotentialAtParticles().cpu().numpy()
    scale = np.abs(want).max(axis=0)
    lim = 1e-9 if split <= 0 else 2e-5
    assert (np.abs(got - want).max(axis=0) < lim * scale).all(), np.abs(got - want).max(axis=0) / scale
    # Poisson::sum: force += q E, energy += q phi, accumulating
    force = torch.ones(N, 4, dtype=torch.float64, device=cuda)
    energy = torch.full((N,), 2.0, dtype=torch.float64, device=cuda)
    p.sum(force=force, energy=energy)
    f, e = force.cpu().numpy(), energy.cpu().numpy()
    assert np.abs(f[:, :3] - 1.0 - q[:, None] * got[:, :3]).max() < 1e-11 * scale[:3].max()
    assert np.abs(f[:, 3] - 1.0).max() == 0.0
    assert np.abs(e - 2.0 - q * got[:, 3]).max() < 1e-11 * scale[3]
    with pytest.raises(UB200Error):
        p.sum(force=force, virial=energy)


def test_result_does_not_depend_on_the_split(cuda):
    """The reference's acceptance criterion for the Ewald mode (SpectralEwaldPoisson.cuh:41-44): two different splits agree
    to the tolerance."""
    N, L, gw, tol = 500, 48.0, 0.5, 1e-5
    pos, q = _cloud(N, L, 9)
    dpos, dq = torch.from_numpy(pos).to(cuda), torch.from_numpy(q).to(cuda)
    res = [Poisson(dpos, dq, Parameters(L, epsilon=1.3, tolerance=tol, gw=gw, split=s)).computeFieldPotentialAtParticles().cpu().numpy()
           for s in (0.45, 0.6, 0.9)]
    # the tolerance bounds the absolute error of ONE pair of unit charges; the truncation errors of the tens of charges within
    # the cut-off of a particle add up: 50 tolerances for this cloud (the mean difference is ~1 tolerance)
    for other in res[1:]:
        assert np.abs(other - res[0]).max() < 50 * tol and np.abs(other - res[0]).mean() < 3 * tol


def test_single_precision_and_parameter_errors(cuda):
    N, L, gw, tol = 400, 40.0, 0.5, 1e-4
    pos, q = _cloud(N, L, 11)
    d64 = Poisson(torch.from_numpy(pos).to(cuda), torch.from_numpy(q).to(cuda),
                  Parameters(L, epsilon=1.0, tolerance=tol, gw=gw, split=0.6)).computeFieldPotentialAtParticles().cpu().numpy()
    p32 = Poisson(torch.from_numpy(pos.astype(np.float32)).to(cuda), torch.from_numpy(q.astype(np.float32)).to(cuda),
                  Parameters(L, epsilon=1.0, tolerance=tol, gw=gw, split=0.6))
    d32 = p32.computeFieldPotentialAtParticles().cpu().numpy()
    scale = np.abs(d64).max(axis=0)
    assert (np.abs(d32 - d64).max(axis=0) < 5e-4 * scale).all()
    with pytest.raises(UB200Error):      # near field cut-off beyond half the box ("increase splitting parameter")
        Poisson(torch.from_numpy(pos).to(cuda), torch.from_numpy(q).to(cuda), Parameters(8.0, epsilon=1.0, tolerance=1e-8, gw=0.5, split=0.05))
