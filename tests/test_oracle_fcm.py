"""CPU: path-2 oracle pinned against golden vectors from the reference's own host code (tests/golden/
gen_golden_f64.cu), against the reference tests' analytic known answers (Hasimoto self mobility,
test/BDHI/FCM/fcm_test.cu:85-144; Peskin spread/gather vs manual loops, test/misc/ibm/test_ibm_regular.cu)
and against numpy.fft."""
import numpy as np
import pytest

from uammd_b200 import synthetic as syn


def test_peskin_kernels_match_reference(orc, golden_dir):
    raw = np.fromfile(f"{golden_dir}/peskin_f64.bin", dtype=np.float64).reshape(-1, 3)
    k3, k4 = orc.peskin3(0.7), orc.peskin4(0.7)
    for r, p3, p4 in raw:
        assert abs(orc.lib().orc_ibm_phi(k3, r) - p3) <= 1e-15 * max(1.0, abs(p3))
        assert abs(orc.lib().orc_ibm_phi(k4, r) - p4) <= 1e-15 * max(1.0, abs(p4))


def test_barnett_magland_window_matches_reference(orc, golden_dir):
    """IBM_kernels::BarnettMagland (misc/IBM_kernels.cuh:83-113) evaluated by the reference's own host code
    (tests/golden/gen_golden_windows.cu): the oracle and the host-side mirror of the product bindings."""
    from uammd_b200.fcm import BarnettMagland
    raw = np.fromfile(f"{golden_dir}/windows_bm_f64.bin", dtype=np.float64).reshape(3, 3 + 129)
    for row in raw:
        alpha, beta, phi0 = row[:3]
        k = orc.barnett_magland(alpha, beta, 6)
        mine = BarnettMagland(alpha, beta, 6)
        assert abs(k.prefactor - phi0) <= 1e-13 * phi0 and abs(1.0 / mine.norm - phi0) <= 1e-13 * phi0
        for i, v in enumerate(row[3:]):
            r = (-1.05 + 2.1 * i / 128.0) * alpha
            assert abs(orc.lib().orc_ibm_phi(k, r) - v) <= 1e-13 * max(1.0, abs(v))
            assert abs(mine.phi(r) - v) <= 1e-13 * max(1.0, abs(v))
        # unit integral (what the norm is for)
        x = np.linspace(-alpha, alpha, 20001)
        assert abs(np.trapezoid([orc.lib().orc_ibm_phi(k, float(t)) for t in x], x) - 1.0) < 1e-6


def test_six_point_window(orc, golden_dir):
    """GaussianFlexible::sixPoint (misc/IBM_kernels.cuh:163-237). Its constructor fills a device-side table, so the golden
    vector comes from a GPU box (windows_six_f64.bin, same generator); without it the defining moment conditions of Bao, Kaye
    and Peskin pin the closed form: sum_j phi(r - j) = 1, sum_j (r - j) phi = 0, sum_j (r - j)^2 phi = K, sum_j (r - j)^3 phi
    = 0, even/odd sums 1/2 each, at every r."""
    import os
    K = 59.0 / 60.0 - np.sqrt(29.0) / 20.0
    k = orc.six_point(1.0)
    phi = lambda r: orc.lib().orc_ibm_phi(k, float(r))
    for r in np.linspace(0.0, 1.0, 41):
        j = np.arange(-4, 5)
        w = np.array([phi(r - jj) for jj in j])
        d = r - j
        assert abs(w.sum() - 1.0) < 1e-13 and abs((d * w).sum()) < 1e-13
        assert abs((d * d * w).sum() - K) < 1e-13 and abs((d ** 3 * w).sum()) < 1e-12
        assert abs(w[j % 2 == 0].sum() - 0.5) < 1e-13
    path = f"{golden_dir}/windows_six_f64.bin"
    if os.path.exists(path):
        raw = np.fromfile(path, dtype=np.float64)
        h, rows = raw[0], raw[1:].reshape(-1, 2)
        kh = orc.six_point(h)
        for r, v in rows:
            assert abs(orc.lib().orc_ibm_phi(kh, r) - v) <= 1e-14 * max(1.0, abs(v))


def test_gaussian_kernel_parameters_match_reference(orc, golden_dir):
    from uammd_b200.fcm import Gaussian
    raw = np.fromfile(f"{golden_dir}/gaussian_f64.bin", dtype=np.float64).reshape(6, 70)
    for row in raw:
        h, tol, support, rmax, a = row[:5]
        k, ka = orc.gaussian_fcm(h, tol)
        assert k.support == int(support) and abs(k.rmax - rmax) < 1e-14 and abs(ka - a) < 1e-14
        mine = Gaussian(h, tol)  # the host-side mirror used by the product bindings
        assert mine.support == int(support) and abs(mine.rmax - rmax) < 1e-14 and abs(mine.a - a) < 1e-14
        for i, v in enumerate(row[5:]):
            r = rmax * 1.05 * i / 64.0
            assert abs(orc.lib().orc_ibm_phi(k, r) - v) <= 1e-14 * max(1.0, abs(v))


def test_getcell_f64_matches_reference(orc, golden_dir):
    import ctypes as C
    raw = np.fromfile(f"{golden_dir}/getcell_f64.bin", dtype=np.uint8)
    off = 0
    for _ in range(2):
        L = raw[off:off + 24].view(np.float64); off += 24
        cd = raw[off:off + 12].view(np.int32); off += 12
        rec = raw[off:off + 2048 * 36].reshape(2048, 36); off += 2048 * 36
        pts = rec[:, :24].copy().view(np.float64)
        cells = rec[:, 24:].copy().view(np.int32)
        g = orc.make_grid_d(tuple(L), tuple(int(x) for x in cd))
        c = (C.c_int * 3)()
        bad = 0
        for p, ref in zip(pts, cells):
            orc.lib().orc_get_cell_d(C.byref(g), np.ascontiguousarray(p).ctypes.data_as(C.c_void_p), c)
            bad += tuple(c) != tuple(ref)
        assert bad <= 1


def _manual_spread(pos, val, L, n, h):
    """test/misc/ibm/test_ibm_regular.cu:89-111 manual_spread: plain triple loop over the 27 nearest cells."""
    grid = np.zeros((n, n, n, 3))
    k3 = lambda r: (1 + np.sqrt(1 - 3 * (r / h) ** 2)) / (3 * h) if abs(r) < 0.5 * h else \
        ((5 - 3 * abs(r) / h - np.sqrt(1 - 3 * (1 - abs(r) / h) ** 2)) / (6 * h) if abs(r) < 1.5 * h else 0.0)
    for p, v in zip(pos, val):
        c = np.floor((p[:3] + 0.5 * L) / h).astype(int)
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    cj = c + np.array([dx, dy, dz])
                    centre = (cj + 0.5) * h - 0.5 * L
                    r = p[:3] - centre
                    w = k3(r[0]) * k3(r[1]) * k3(r[2])
                    cw = cj % n
                    grid[cw[2], cw[1], cw[0]] += v * w
    return grid


def test_peskin_spread_and_gather_equal_manual_loops(orc):
    n, L = 16, 16.0
    h = L / n
    rng = np.random.default_rng(123)
    pos = np.zeros((40, 4)); pos[:, :3] = (rng.random((40, 3)) - 0.5) * L
    val = rng.normal(size=(40, 3))
    g = orc.make_grid_d((L,) * 3, (n,) * 3)
    sp = orc.ibm_spread(g, orc.peskin3(h), pos, val, n)
    man = _manual_spread(pos, val, L, n, h)
    assert np.abs(sp - man).max() < 1e-10                       # test_ibm_regular.cu:113-136 tolerance
    field = rng.normal(size=(n, n, n, 3))
    ga = orc.ibm_gather(g, orc.peskin3(h), pos, field, n)
    ones = np.eye(3)
    for i in range(5):                                           # J = S^T dV : interpolate == manual spread . field
        s1 = _manual_spread(pos[i:i + 1], np.ones((1, 3)), L, n, h)
        assert np.abs(ga[i] - (s1 * field).sum((0, 1, 2)) * h ** 3).max() < 1e-10


def test_spread_conserves_total_and_is_adjoint_to_gather(orc):
    L, cells = (6.4, 3.2, 0.7 * 7), (64, 32, 7)                 # non cubic cells like test_ibm_regular.cu:156-214
    g = orc.make_grid_d(L, cells)
    h = min(L[d] / cells[d] for d in range(3))
    rng = np.random.default_rng(5)
    N = 300
    pos = np.zeros((N, 4)); pos[:, :3] = (rng.random((N, 3)) - 0.5) * np.array(L)
    val = rng.normal(size=(N, 3))
    kern = orc.peskin3(h)
    sp = orc.ibm_spread(g, kern, pos, val, cells[0])
    field = rng.normal(size=sp.shape)
    ga = orc.ibm_gather(g, kern, pos, field, cells[0])
    dV = np.prod([L[d] / cells[d] for d in range(3)])
    assert abs((sp * field).sum() * dV - (ga * val).sum()) < 1e-10 * abs((ga * val).sum())


def test_dft_matches_numpy(orc):
    nx, ny, nz = 12, 10, 6
    rng = np.random.default_rng(2)
    grid = rng.normal(size=(nz, ny, 2 * (nx // 2 + 1), 3))
    ghat = orc.dft3_r2c(grid, nx)
    ref = np.fft.rfftn(grid[:, :, :nx, :], axes=(0, 1, 2))
    assert np.abs(ghat - ref).max() < 1e-11
    back = orc.dft3_c2r(ghat, nx, grid.shape[2])
    assert np.abs(back[:, :, :nx, :] - grid[:, :, :nx, :] * nx * ny * nz).max() < 1e-9


def _hasimoto(a, eta, L):
    from uammd_b200.fcm import hasimotoSelfMobility
    return hasimotoSelfMobility(a, eta, L)


def test_fcm_self_mobility_peskin3_coarse(orc):
    # fcm_test.cu:19-22 expects only a few digits from the Peskin 3pt kernel (its effective hydrodynamic radius
    # is h only approximately and varies with the position inside the cell); the strict pin is the Gaussian KAT
    n, L, eta = 64, 64.0, 1.12321
    pos = np.zeros((1, 4)); pos[0, :3] = [3.3, -7.21, 11.17]
    u = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(L / n), eta, pos, np.array([[1.0, 1.0, 1.0]]))
    m0 = _hasimoto(L / n, eta, L)
    assert np.abs(u[0] / m0 - 1).max() < 3e-2


@pytest.mark.timeout(600)
def test_fcm_self_mobility_gaussian_reference_kat(orc):
    """test/BDHI/FCM/fcm_test.cu:85-144 restated: Gaussian kernel, tolerance 1e-8, eta = 1.12321, a = 1.012312,
    L = 96 h ceil(a/h) -> 288^3 grid; self mobility equals the Hasimoto expression to 1e-8."""
    tol, a, eta = 1e-8, 1.012312, 1.12321
    from uammd_b200.fcm import Gaussian
    h = Gaussian.adviseGridSize(a, tol)
    L = 96 * h * np.ceil(a / h)
    n = int(L / h + 1e-9)
    assert n == 288
    kern, ka = orc.gaussian_fcm(h, tol)
    pos = np.zeros((1, 4)); pos[0, :3] = np.array([0.2137, -0.3871, 0.0713]) * L
    # cubic symmetry: M = M0 * identity, so F = (1,1,1) probes the three rows at once
    u = orc.fcm_mdot((L,) * 3, (n,) * 3, kern, eta, pos, np.array([[1.0, 1.0, 1.0]]))
    m0 = _hasimoto(a, eta, L)
    assert np.abs(u[0] - m0).max() < 3 * tol


def test_noise_is_hermitian_and_projected(orc):
    n, L = 8, 8.0
    g = orc.make_grid_d((L,) * 3, (n,) * 3)
    ghat = orc.fcm_add_noise(g, 1.0, 0.37, 42, 1, np.zeros((n, n, n // 2 + 1, 3), np.complex128))
    assert np.abs(ghat[0, 0, 0]).max() == 0                      # k = 0 gets no noise
    for kx in (0, n // 2):                                       # planes stored twice: v(-k) = conj v(k)
        pl = ghat[:, :, kx, :]
        conj = np.conj(pl[(-np.arange(n)) % n][:, (-np.arange(n)) % n])
        if kx == 0:
            assert np.abs(pl - conj).max() < 1e-12
    back = np.fft.irfftn(ghat, s=(n, n, n), axes=(0, 1, 2))
    # divergence free: k . v(k) = 0 away from Nyquist components
    k = 2 * np.pi * np.fft.fftfreq(n, d=L / n)
    kz, ky, kx = np.meshgrid(k, k, 2 * np.pi * np.fft.rfftfreq(n, d=L / n), indexing="ij")
    div = kx * ghat[..., 0] + ky * ghat[..., 1] + kz * ghat[..., 2]
    inner = np.ones_like(div, bool)
    inner[n // 2, :, :] = inner[:, n // 2, :] = inner[:, :, n // 2] = False
    assert np.abs(div[inner]).max() < 1e-12 and np.isfinite(back).all()


def test_rotational_self_mobility_known_answer(orc):
    """Rotational FCM (torque path, FCM_impl.cuh:306-358,583-649): with the torque kernel the reference derives from the
    hydrodynamic radius (width = a / (6 sqrt(pi))^(1/3), BDHI_FCM.cuh:69-80) a unit torque spins the particle with
    1/(8 pi eta a^3), the rotational mobility of a sphere, up to periodic corrections O((a/L)^3); a pure torque moves nothing."""
    import math
    eta, tol, n, L = 1.3, 1e-6, 64, 40.0
    h = L / n
    kern, a = orc.gaussian_fcm(h, tol)
    kt = orc.gaussian_torque(a / (6 * math.sqrt(math.pi)) ** (1 / 3.0), h, tol)
    want = 1.0 / (8 * math.pi * eta * a ** 3)
    for d, p in ((0, (0.3, -0.2, 0.1)), (2, (-11.1, 7.7, 19.9))):
        pos = np.zeros((1, 4)); pos[0, :3] = p
        tor = np.zeros((1, 3)); tor[0, d] = 1.0
        lin, ang = orc.fcm_mdot((L,) * 3, (n,) * 3, kern, eta, pos, None, torque3=tor, kernTorque=kt)
        assert abs(ang[0, d] - want) < 5e-4 * want
        assert np.abs(np.delete(ang[0], d)).max() < 1e-12 and np.abs(lin).max() < 1e-4 * want
