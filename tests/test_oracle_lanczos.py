"""CPU: the Lanczos square-root oracle (oracle/oracle_lanczos.py, restating misc/LanczosAlgorithm/LanczosAlgorithm.cu) on the
near-field mobility of the PSE oracle, and the row-block iteration the rank decomposition runs (ub200_pse_dist_*)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT,) if p not in sys.path]

from oracle import oracle_lanczos as OL  # noqa: E402


def _near_matrix(orc, N=60, L=16.0, psi=0.8, rh=1.0, seed=3, shear=0.0):
    """dense near-field mobility from the oracle's mat-vec on unit vectors (symmetric positive definite by construction of
    the positively split Ewald sum)"""
    rng = np.random.default_rng(seed)
    pos = np.zeros((N, 4)); pos[:, :3] = rng.uniform(-L / 2, L / 2, (N, 3))
    par = orc.pse_params(L, 1.0, rh, 1e-6, psi)
    table = orc.pse_near_table(par, rh, psi)
    M = np.zeros((3 * N, 3 * N))
    for k in range(3 * N):
        e = np.zeros((N, 3)); e.flat[k] = 1.0
        M[:, k] = orc.pse_near_mdot(par, rh, psi, pos, e, shear=shear, table=table).ravel()
    return M


def _dense_sqrt(M):
    lam, Q = np.linalg.eigh(0.5 * (M + M.T))
    assert lam.min() > 0
    return (Q * np.sqrt(lam)) @ Q.T


@pytest.mark.parametrize("tol", [1e-3, 1e-7])
def test_lanczos_matches_the_dense_square_root(orc, tol):
    M = _near_matrix(orc)
    assert np.abs(M - M.T).max() < 1e-12 * np.abs(M).max()
    z = np.random.default_rng(5).standard_normal(M.shape[0])
    Bz, it = OL.lanczos_sqrt(lambda v: M @ v, z, tol)
    want = _dense_sqrt(M) @ z
    assert 2 <= it < 60
    # the stopping rule bounds the CHANGE of the estimate; the distance to the limit is of the same order
    assert np.linalg.norm(Bz - want) / np.linalg.norm(want) < 20 * tol


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_blocks_reproduce_the_whole_vector_iteration(orc, world):
    """what the ranks of ub200_pse_dist_* compute: same iteration count, result equal up to the summation order of the dots"""
    M = _near_matrix(orc, shear=0.1)
    z = np.random.default_rng(6).standard_normal(M.shape[0])
    one, it1 = OL.lanczos_sqrt(lambda v: M @ v, z, 1e-6)
    rows, itr = OL.lanczos_sqrt_rows(lambda v: M @ v, z, 1e-6, world)
    assert itr == it1
    assert np.linalg.norm(rows - one) / np.linalg.norm(one) < 1e-12


def test_convergence_check_cadence_adapts_like_the_reference():
    """registerRequiredStepsForConverge (LanczosAlgorithm.cu:253-262): more iterations than expected -> check one step later
    next time; fewer -> check up to two steps earlier, never before step 1"""
    s = OL.Solver()
    assert s.check_convergence_steps == 3
    s._register(9); assert s.check_convergence_steps == 4
    s._register(5); assert s.check_convergence_steps == 2
    s._register(2); assert s.check_convergence_steps == 1
    s._register(2); assert s.check_convergence_steps == 1


def test_zero_residual_takes_the_unit_vector_branch():
    """an eigenvector as input: w - h v vanishes, hsup is clamped to 0 and the next basis vector is e1 (nextIteration :146-155)"""
    M = np.diag([4.0, 9.0, 16.0])
    z = np.array([0.0, 2.0, 0.0])
    Bz, it = OL.lanczos_sqrt(lambda v: M @ v, z, 1e-10)
    assert np.allclose(Bz, [0.0, 6.0, 0.0], atol=1e-12) and it >= 1
