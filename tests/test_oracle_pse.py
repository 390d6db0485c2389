"""CPU: the PSE oracle (oracle/oracle_pse.c) pinned by the reference test's known answer - near + far self mobility
equals Hasimoto's periodic self mobility (test/BDHI/PSE/pse_test.cu:28-41,64-117, restated at tolerances a CPU grid
affords) - and by internal consistency: table vs closed form, F/G continuity at r = 2a, far+near independence of psi."""
import math

import numpy as np
import pytest


def _self_mobility(rh, vis, L):  # pse_test.cu:28-41 / PSE/initialization.cu:31-49
    a = rh / L
    a3 = a ** 3
    c, b = 2.83729747948061947666591710460773907, 0.19457
    a6pref = 16.0 * math.pi ** 2 / 45.0 + 630.0 * b * b
    return 1.0 / (6.0 * math.pi * vis * rh) * (1.0 - c * a + (4.0 / 3.0) * math.pi * a3 - a6pref * a3 * a3)


def _mobility(orc, L, vis, rh, tol, psi, pos, force):
    par = orc.pse_params(L, vis, rh, tol, psi)
    far = orc.pse_far_mdot(par, vis, rh, psi, pos, force[:, :3])
    near = orc.pse_near_mdot(par, rh, psi, pos, force, table=orc.pse_near_table(par, rh, psi))
    return far + near, par


@pytest.mark.parametrize("psi", [0.4, 0.7])
def test_self_mobility_known_answer(orc, psi):
    rh, vis = 1.012312, 1.12321
    L = 32 * rh
    tol = 1e-5
    rng = np.random.default_rng(1234)
    for _ in range(3):
        pos = np.zeros((1, 4)); pos[0, :3] = (rng.random(3) - 0.5) * L
        for d in range(3):
            f = np.zeros((1, 4)); f[0, d] = 1.0
            u, par = _mobility(orc, L, vis, rh, tol, psi, pos, f)
            want = np.zeros(3); want[d] = _self_mobility(rh, vis, L)
            assert np.abs(u[0] - want).max() < 10 * tol, (psi, par["cells"], u, want)


def test_pair_mobility_is_independent_of_the_split(orc):
    rh, vis, L, tol = 1.0, 1.0, 24.0, 1e-6
    rng = np.random.default_rng(5)
    N = 12
    pos = np.zeros((N, 4)); pos[:, :3] = (rng.random((N, 3)) - 0.5) * L
    f = np.zeros((N, 4)); f[:, :3] = rng.normal(size=(N, 3))
    u1, _ = _mobility(orc, L, vis, rh, tol, 0.5, pos, f)
    u2, _ = _mobility(orc, L, vis, rh, tol, 0.8, pos, f)
    assert np.abs(u1 - u2).max() < 20 * tol * np.abs(u1).max()


def test_near_table_and_closed_form(orc):
    rh, psi, vis = 1.0, 0.6, 1.0
    par = orc.pse_params(32.0, vis, rh, 1e-4, psi)
    t = orc.pse_near_table(par, rh, psi)
    assert t.shape == (par["nTable"], 2) and par["nTable"] >= 1 << 14
    assert t[0, 1] == 0 and abs(t[0, 0] * par["normalization"] - orc.rpy_near_fg(0.0, rh, psi, par["rcut"])[0]) < 1e-15
    # continuity of F and G across the overlap boundary r = 2a
    lo, hi = orc.rpy_near_fg(2 * rh - 1e-9, rh, psi, par["rcut"]), orc.rpy_near_fg(2 * rh + 1e-9, rh, psi, par["rcut"])
    assert abs(lo[0] - hi[0]) < 1e-8 and abs(lo[1] - hi[1]) < 1e-8
    # tabulated mat-vec vs closed form: linear interpolation error ~ (dr)^2 F''
    rng = np.random.default_rng(2)
    N = 40
    pos = np.zeros((N, 4)); pos[:, :3] = (rng.random((N, 3)) - 0.5) * 8.0
    v = rng.normal(size=(N, 3))
    a = orc.pse_near_mdot(par, rh, psi, pos, v, table=t)
    b = orc.pse_near_mdot(par, rh, psi, pos, v, table=None)
    assert np.abs(a - b).max() < 1e-7 * np.abs(b).max()
    # the near-field matrix is symmetric: u.(M v) == v.(M u)
    u = rng.normal(size=(N, 3))
    assert abs((u * orc.pse_near_mdot(par, rh, psi, pos, v, table=t)).sum() - (v * orc.pse_near_mdot(par, rh, psi, pos, u, table=t)).sum()) < 1e-12


def test_sheared_minimum_image(orc):
    # a lattice translation of the sheared cell leaves the near-field product unchanged
    rh, psi, vis, L, g = 1.0, 0.6, 1.0, 20.0, 0.3
    par = orc.pse_params(L, vis, rh, 1e-4, psi)
    rng = np.random.default_rng(3)
    N = 30
    pos = np.zeros((N, 4)); pos[:, :3] = (rng.random((N, 3)) - 0.5) * L
    v = rng.normal(size=(N, 3))
    a = orc.pse_near_mdot(par, rh, psi, pos, v, shear=g)
    pos2 = pos.copy(); pos2[::2, 1] += L; pos2[1::3, 2] -= L; pos2[::5, 0] += L
    b = orc.pse_near_mdot(par, rh, psi, pos2, v, shear=g)
    assert np.abs(a - b).max() < 1e-12


def test_fft_wise_sizes(orc):
    assert orc.next_fft_wise_size(255) == 256 and orc.next_fft_wise_size(355) == 360 and orc.next_fft_wise_size(65) == 66
    assert orc.next_fft_wise_size(129) == 132 and orc.next_fft_wise_size(2) == 2 and orc.next_fft_wise_size(23) == 24
    assert orc.pse_params(256.0, 1.0, 1.0, 1e-3, 0.593)["cells"] == (256, 256, 256)   # SURVEY 8(d) C4 recipe
    assert orc.pse_params(256.0, 1.0, 1.0, 1e-3, 0.593)["support"] == 7
