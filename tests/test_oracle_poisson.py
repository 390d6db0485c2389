"""CPU: the numpy restatement of the reference's spectral Ewald Poisson solver (oracle/oracle_poisson.py) against the
reference tests' analytic known answers (test/Potentials/Poisson/TriplyPeriodic/test_poisson.cu:13-23,192-222) and its
own split invariance (SpectralEwaldPoisson.cuh:41-44)."""
import math

import numpy as np

from oracle.oracle_poisson import PoissonOracle, greens_function, greens_function_field


def test_two_charges_field_and_potential_known_answer():
    L, r, tol, gw, split = 100.0, 2.0, 1e-7, 0.001, 0.2
    o = PoissonOracle(L, 1.0, tol, gw, split)
    assert tuple(o.cells) == (66, 66, 66) and o.support == 18
    for ori in ([3.1, -7.7, 12.3], [49.2, -49.9, 0.4]):
        pos = np.zeros((3, 3))
        pos[0] = np.array(ori) + [-0.5 * r, 0, 0]
        pos[1] = pos[2] = np.array(ori) + [0.5 * r, 0, 0]
        fp = o.field_potential(pos, np.array([1.0, -0.5, -0.5]))
        want = -math.exp(-r * r / (4.0 * gw * gw)) / (4 * math.pi * math.sqrt(math.pi) * gw * r) - \
            math.erf(r / (2.0 * gw)) / (4 * math.pi * r * r)
        assert abs(fp[0, 1]) < 1e-10 and abs(fp[0, 2]) < 1e-10 and fp[0, 0] > 0
        assert abs(1.0 - abs(fp[0, 0] / want)) < 1e-3            # the reference's own bound


def test_green_functions_are_consistent():
    # greensFunctionField really is the radial derivative of greensFunction (central differences), on both branches of the piecewise definitions
    gw, split, eps = 0.4, 0.7, 1.3
    for r in (0.01, 0.1, 0.5, 1.7, 3.0):
        d = 1e-5 * max(r, 0.1)
        num = (greens_function(np.array([(r + d) ** 2]), gw, split, eps) - greens_function(np.array([(r - d) ** 2]), gw, split, eps)) / (2 * d)
        # greensFunctionField is -dG/dr: the Transversers put the minus sign of F = -grad U in front of it
        assert abs(num[0] + greens_function_field(np.array([r]), gw, split, eps)[0]) < 1e-6 * max(1.0, abs(num[0]))


def test_split_invariance():
    rng = np.random.default_rng(3)
    N, L, gw, tol = 60, 40.0, 0.5, 1e-5
    pos = (rng.random((N, 3)) - 0.5) * L
    q = rng.choice([-1.0, 1.0], N); q[-1] -= q.sum()
    res = [PoissonOracle(L, 1.0, tol, gw, s).field_potential(pos, q) for s in (0.5, 0.9)]
    # the tolerance bounds the ABSOLUTE error of the potential of unit charges (epsilon = 1): two splits agree to a few of them
    assert np.abs(res[1] - res[0]).max() < 5 * tol
