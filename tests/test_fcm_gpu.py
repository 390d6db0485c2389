"""GPU: path 2 through the C ABI - hand-written 3-D FFT vs numpy.fft, IBM spread/gather vs the oracle, the FCM
pipeline vs the oracle (fp64 rel-L2 <= 1e-12 as BASELINE.md asks), the reference's own FCM KAT (Gaussian self
mobility to 1e-8, test/BDHI/FCM/fcm_test.cu:85-144), Brownian noise vs the oracle restatement, and parity with
the compiled reference (oracle/_ref/ref_fcm)."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.fcm import BarnettMagland, SixPoint, FCM_impl, FFT3D, Gaussian, IBM, Peskin3, Peskin4, hasimotoSelfMobility

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FCM = os.path.join(ROOT, "oracle", "_ref", "ref_fcm")


# ---------------- FFT ----------------
@pytest.mark.parametrize("shape", [(16, 16, 16), (128, 128, 128), (24, 20, 18), (30, 14, 6), (9, 15, 7), (64, 32, 4),
                                   (96, 96, 96), (22, 44, 66), (256, 64, 512), (360, 8, 8), (250, 50, 25)])
def test_fft3d_matches_numpy_f64(cuda, shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx * 7 + ny)
    real = rng.normal(size=(nz, ny, nx, 3))
    grid = torch.zeros(nz, ny, 2 * (nx // 2 + 1), 3, dtype=torch.float64, device=cuda)
    grid[:, :, :nx, :] = torch.from_numpy(real).to(cuda)
    fft = FFT3D(nx, ny, nz, torch.float64)
    spec = fft.forward(grid).cpu().numpy()
    ref = np.fft.rfftn(real, axes=(0, 1, 2))
    scale = np.abs(ref).max()
    assert np.abs(spec - ref).max() < 1e-13 * scale * np.log2(nx * ny * nz)
    back = fft.inverse(grid).cpu().numpy()[:, :, :nx, :]
    assert np.abs(back - real * nx * ny * nz).max() < 1e-12 * nx * ny * nz


def test_fft3d_f32(cuda):
    nx, ny, nz = 64, 48, 40
    rng = np.random.default_rng(1)
    real = rng.normal(size=(nz, ny, nx, 3)).astype(np.float32)
    grid = torch.zeros(nz, ny, 2 * (nx // 2 + 1), 3, dtype=torch.float32, device=cuda)
    grid[:, :, :nx, :] = torch.from_numpy(real).to(cuda)
    spec = FFT3D(nx, ny, nz, torch.float32).forward(grid).cpu().numpy()
    ref = np.fft.rfftn(real.astype(np.float64), axes=(0, 1, 2))
    assert np.abs(spec - ref).max() < 2e-6 * np.abs(ref).max() * np.log2(nx * ny * nz)


def test_fft3d_rejects_unsupported_size(cuda):
    from uammd_b200 import UB200Error
    with pytest.raises(UB200Error):
        FFT3D(26, 16, 16)  # factor 13 (2, 3, 5, 7 and 11 are supported: nextFFTWiseSize3D emits all of them)


# ---------------- IBM ----------------
def _cloud(N, L, seed, dtype=np.float64):
    pos = np.zeros((N, 4), dtype)
    pos[:, :3] = syn.uniform_cloud(N, L, seed=seed)[:, :3].astype(dtype)
    return pos


@pytest.mark.parametrize("kname,cells,L", [("p3", (32, 32, 32), (32.0,) * 3), ("p4", (24, 20, 16), (12.0, 10.0, 8.0)),
                                           ("p4", (40, 36, 32), (20.0, 18.0, 16.0)), ("p3", (19, 21, 23), (9.5, 10.5, 11.5)),
                                           ("p3", (64, 32, 7), (6.4, 3.2, 4.9)), ("gauss", (32, 32, 32), (32.0,) * 3),
                                           ("six", (32, 30, 28), (16.0, 15.0, 14.0)), ("bm", (32, 32, 32), (32.0,) * 3),
                                           ("bm5", (36, 32, 40), (18.0, 16.0, 20.0))])
def test_ibm_spread_gather_match_oracle(orc, cuda, kname, cells, L):
    h = min(L[d] / cells[d] for d in range(3))
    if kname == "six":    # GaussianFlexible::sixPoint: even support, warp-per-particle path
        kern, ok = SixPoint(h), orc.six_point(h)
    elif kname == "bm":   # Barnett-Magland, w = 6 points (alpha = w h / 2, beta = 1.8 w)
        kern, ok = BarnettMagland(3.0 * h, 1.8 * 6, 6), orc.barnett_magland(3.0 * h, 1.8 * 6, 6)
    elif kname == "bm5":  # w = 5: takes the node-centred (atomic-free) spread
        kern, ok = BarnettMagland(2.5 * h, 1.8 * 5, 5), orc.barnett_magland(2.5 * h, 1.8 * 5, 5)
    elif kname == "p3":
        kern, ok = Peskin3(h), orc.peskin3(h)
    elif kname == "p4":
        kern, ok = Peskin4(h), orc.peskin4(h)
    else:
        kern = Gaussian(h, 1e-5)
        ok, _ = orc.gaussian_fcm(h, 1e-5)
    N = 5000
    pos = _cloud(N, L, 3)
    pos[::9, :3] *= 2.3  # outside the primary box
    val = syn.gaussian_forces(N, seed=4)
    nxPad = 2 * (cells[0] // 2 + 1)
    g = orc.make_grid_d(L, cells)
    ref = orc.ibm_spread(g, ok, pos, val, nxPad)
    ibm = IBM(kern, L, cells, nxPad)
    dpos, dval = torch.from_numpy(pos).to(cuda), torch.from_numpy(val).to(cuda)
    grid = torch.full((cells[2], cells[1], nxPad, 3), 7.0, dtype=torch.float64, device=cuda)
    ibm.spread(dpos, dval, grid, overwrite=True)            # writes every node, no zero fill needed
    sp = grid.cpu().numpy()
    scale = np.abs(ref).max()
    assert np.abs(sp[:, :, :cells[0]] - ref[:, :, :cells[0]]).max() < 1e-13 * scale
    grid2 = torch.ones_like(grid)
    ibm.spread(dpos, dval, grid2)                            # reference semantics: accumulates
    assert np.abs(grid2.cpu().numpy()[:, :, :cells[0]] - 1.0 - ref[:, :, :cells[0]]).max() < 1e-12 * scale
    field = torch.from_numpy(np.random.default_rng(5).normal(size=ref.shape)).to(cuda)
    out = torch.full((N, 3), 2.0, dtype=torch.float64, device=cuda)
    ibm.gather(dpos, field, out)                             # accumulates
    gref = orc.ibm_gather(g, ok, pos, field.cpu().numpy(), nxPad)
    assert np.abs(out.cpu().numpy() - 2.0 - gref).max() < 1e-12 * max(1.0, np.abs(gref).max())


def test_ibm_non_periodic_corner_clips_like_reference(orc, cuda):
    # test_ibm_regular.cu:16-64: a particle in the corner of a non periodic box touches 8 cells instead of 27
    L, cells = (16.0,) * 3, (16,) * 3
    pos = np.zeros((1, 4)); pos[0, :3] = [-7.9, -7.9, -7.9]
    val = np.ones((1, 3))
    ibm = IBM(Peskin3(1.0), L, cells, 16, periodic=(0, 0, 0))
    grid = torch.zeros(16, 16, 16, 3, dtype=torch.float64, device=cuda)
    ibm.spread(torch.from_numpy(pos).to(cuda), torch.from_numpy(val).to(cuda), grid)
    assert int((grid[..., 0] != 0).sum()) == 8
    ref = orc.ibm_spread(orc.make_grid_d(L, cells, (0, 0, 0)), orc.peskin3(1.0), pos, val, 16)
    assert np.abs(grid.cpu().numpy() - ref).max() < 1e-14


# ---------------- FCM ----------------
def _fcm_inputs(N, L, seed=11):
    pos = _cloud(N, (L,) * 3, seed)
    force = np.zeros((N, 4)); force[:, :3] = syn.gaussian_forces(N, seed=seed + 1)
    return pos, force


@pytest.mark.parametrize("kname,n,N", [("p3", 32, 3000), ("p4", 48, 3000), ("gauss", 36, 500), ("p3", 64, 20000)])
def test_fcm_mdot_matches_oracle(orc, cuda, kname, n, N):
    L, eta = float(n), 1.3
    h = L / n
    kern, ok = {"p3": (Peskin3(h), orc.peskin3(h)), "p4": (Peskin4(h), orc.peskin4(h)),
                "gauss": (Gaussian(h, 1e-4), orc.gaussian_fcm(h, 1e-4)[0])}[kname]
    pos, force = _fcm_inputs(N, L)
    fcm = FCM_impl(L, (n,) * 3, kern, eta, seed=1)
    out = fcm.computeHydrodynamicDisplacements(torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda))
    ref = orc.fcm_mdot((L,) * 3, (n,) * 3, ok, eta, pos, force[:, :3])
    err = np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-12, err


def test_fcm_f32(orc, cuda):
    n, N, L = 32, 3000, 32.0
    pos, force = _fcm_inputs(N, L)
    fcm = FCM_impl(L, (n,) * 3, Peskin3(1.0), 1.0, seed=1, dtype=torch.float32)
    out = fcm.computeHydrodynamicDisplacements(torch.from_numpy(pos.astype(np.float32)).to(cuda),
                                               torch.from_numpy(force.astype(np.float32)).to(cuda))
    ref = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(1.0), 1.0, pos.astype(np.float32).astype(np.float64), force[:, :3])
    assert np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref) < 2e-5


def test_fcm_config3_full_size(orc, cuda):
    """BASELINE config 3: N = 5e5, 128^3, Peskin 3pt, fp64, T = 0: rel-L2 <= 1e-12 vs the oracle."""
    N, n, L = 500_000, 128, 128.0
    pos, force = _fcm_inputs(N, L)
    fcm = FCM_impl(L, (n,) * 3, Peskin3(1.0), 1.0, seed=1)
    out = fcm.computeHydrodynamicDisplacements(torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda))
    ref = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(1.0), 1.0, pos, force[:, :3])
    err = np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-12, err
    # size independent property: linearity M(a f1 + b f2) = a M f1 + b M f2
    f2 = force.copy(); f2[:, :3] = syn.gaussian_forces(N, seed=99)
    dpos = torch.from_numpy(pos).to(cuda)
    o2 = fcm.computeHydrodynamicDisplacements(dpos, torch.from_numpy(f2).to(cuda))
    o3 = fcm.computeHydrodynamicDisplacements(dpos, torch.from_numpy(2.0 * force - 0.5 * f2).to(cuda))
    assert float((o3 - (2.0 * out - 0.5 * o2)).norm() / o3.norm()) < 1e-12
    # symmetric positive operator: f . M f > 0
    assert float((torch.from_numpy(force[:, :3]).to(cuda) * out).sum()) > 0


def test_fcm_self_mobility_reference_kat(cuda):
    """test/BDHI/FCM/fcm_test.cu:85-144: Gaussian kernel at tolerance 1e-8 on the 288^3 grid the test derives;
    20 Saru-free random positions x 3 directions, each within 1e-8 of the Hasimoto self mobility."""
    tol, a, eta = 1e-8, 1.012312, 1.12321
    h = Gaussian.adviseGridSize(a, tol)
    L = 96 * h * np.ceil(a / h)
    n = int(L / h + 1e-9)
    fcm = FCM_impl(L, (n,) * 3, Gaussian(h, tol), eta, hydrodynamicRadius=a, seed=1)
    assert abs(fcm.getHydrodynamicRadius() - a) < 1e-12
    m0 = hasimotoSelfMobility(a, eta, L)
    rng = np.random.default_rng(1234)
    for _ in range(20):
        pos = np.zeros((1, 4)); pos[0, :3] = (rng.random(3) - 0.5) * L
        dpos = torch.from_numpy(pos).to(cuda)
        for d in range(3):
            f = np.zeros((1, 4)); f[0, d] = 1.0
            u = fcm.computeHydrodynamicDisplacements(dpos, torch.from_numpy(f).to(cuda)).cpu().numpy()[0]
            expect = np.zeros(3); expect[d] = m0
            assert np.abs(u - expect).max() < tol, (u, expect)


def test_fcm_noise_matches_oracle_restatement(orc, cuda):
    n, N, L, eta, T, dt = 16, 200, 16.0, 0.9, 1.7, 0.01
    pos, force = _fcm_inputs(N, L)
    fcm = FCM_impl(L, (n,) * 3, Peskin3(1.0), eta, seed=4242)
    dpos, dforce = torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda)
    for call in (1, 2):  # the call counter is Saru's third seed (FCM_impl.cuh:517,523)
        out = fcm.computeHydrodynamicDisplacements(dpos, dforce, temperature=T, prefactor=1 / np.sqrt(dt)).cpu().numpy()
        ref = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(1.0), eta, pos, force[:, :3], T, 1 / np.sqrt(dt), 4242, call)
        det = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(1.0), eta, pos, force[:, :3])
        # device vs host logf/sinf/cosf differ in the last float ulp
        assert np.abs(out - ref).max() < 5e-6 * np.abs(ref - det).max()
    only = fcm.computeHydrodynamicDisplacements(dpos, None, temperature=T, prefactor=1 / np.sqrt(dt)).cpu().numpy()
    ref = orc.fcm_mdot((L,) * 3, (n,) * 3, orc.peskin3(1.0), eta, pos, None, T, 1 / np.sqrt(dt), 4242, 3)
    assert np.abs(only - ref).max() < 5e-6 * np.abs(ref).max()


def test_fcm_fluctuation_dissipation(cuda):
    """pse_test.cu:121-159 style: <dx^2> = 2 T M0 for the Brownian displacements of an isolated particle
    (prefactor 1, noise only), 1000 draws, 2% tolerance at this sample size."""
    n, L, eta, T = 32, 32.0, 1.0, 0.5
    fcm = FCM_impl(L, (n,) * 3, Gaussian(1.0, 1e-4), eta, seed=777)
    a = fcm.getHydrodynamicRadius()
    m0 = hasimotoSelfMobility(a, eta, L)
    pos = np.zeros((64, 4)); pos[:, :3] = (np.random.default_rng(3).random((64, 3)) - 0.5) * L
    dpos = torch.from_numpy(pos).to(cuda)
    acc = torch.zeros(3, dtype=torch.float64, device=cuda)
    ndraw = 400
    for _ in range(ndraw):
        dx = fcm.computeHydrodynamicDisplacements(dpos, None, temperature=T, prefactor=1.0)
        acc += (dx * dx).mean(0)
    var = (acc / ndraw).cpu().numpy()
    assert np.abs(var / (2 * T * m0) - 1).max() < 0.05, var / (2 * T * m0)


# ---------------- compiled reference ----------------
@pytest.mark.parametrize("kname,N,n", [("peskin3", 20000, 64), ("peskin3", 500_000, 128), ("gaussian", 2000, 48)])
def test_reference_parity(cuda, tmp_path, kname, N, n):
    if not os.path.exists(REF_FCM):
        pytest.skip("oracle/_ref/ref_fcm not built")
    L, eta, tol = float(n), 1.0, 1e-5
    pos, force = _fcm_inputs(N, L)
    pos.tofile(tmp_path / "pos.bin"); force.tofile(tmp_path / "force.bin")
    subprocess.run([REF_FCM, "mdot", kname, str(N), str(L), str(n), str(eta), str(tol), "0", "0", "1",
                    str(tmp_path / "pos.bin"), str(tmp_path / "force.bin"), str(tmp_path / "out.bin")],
                   check=True, capture_output=True, timeout=600)
    ref = np.fromfile(tmp_path / "out.bin", np.float64).reshape(N, 3)
    kern = Peskin3(1.0) if kname == "peskin3" else Gaussian(1.0, tol)
    fcm = FCM_impl(L, (n,) * 3, kern, eta, seed=1)
    out = fcm.computeHydrodynamicDisplacements(torch.from_numpy(pos).to(cuda), torch.from_numpy(force).to(cuda))
    err = np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"[FCM parity {kname} N={N} n={n}] rel-L2 vs compiled reference {err:.3e}")
    assert err < 1e-12, err


# ---------------- rotational FCM (torques) ----------------
def _torque_inputs(N, L, seed):
    pos = _cloud(N, (L,) * 3, seed)
    force = np.zeros((N, 4)); force[:, :3] = syn.gaussian_forces(N, seed=seed + 1)
    torque = np.zeros((N, 4)); torque[:, :3] = syn.gaussian_forces(N, seed=seed + 2)
    return pos, force, torque


@pytest.mark.parametrize("with_force,T", [(True, 0.0), (False, 0.0), (True, 0.8)])
def test_fcm_torques_match_oracle(orc, cuda, with_force, T):
    from uammd_b200.fcm import GaussianTorque
    N, n, L, tol, eta = 3000, 32, 32.0, 1e-5, 1.3
    h = L / n
    pos, force, torque = _torque_inputs(N, L, 41)
    kern = Gaussian(h, tol)
    a = kern.fixHydrodynamicRadius(0, h)
    kt = GaussianTorque.forHydrodynamicRadius(a, h, tol)
    fcm = FCM_impl(L, (n, n, n), kern, eta, seed=77, kernelTorque=kt)
    dp, df, dt_ = (torch.from_numpy(x).to(cuda) for x in (pos, force, torque))
    lin, ang = fcm.computeHydrodynamicDisplacements(dp, df if with_force else None, temperature=T, prefactor=1.0, torque=dt_)
    torch.cuda.synchronize()
    ok, _ = orc.gaussian_fcm(h, tol)
    okt = orc.gaussian_torque(kt.width, h, tol)
    assert okt.support == kt.support
    rlin, rang = orc.fcm_mdot((L,) * 3, (n, n, n), ok, eta, pos, force[:, :3] if with_force else None, temperature=T,
                              prefactor=1.0, seed=77, seed2=1, torque3=torque[:, :3], kernTorque=okt)
    rel = lambda a_, b_: np.linalg.norm(a_ - b_) / np.linalg.norm(b_)
    tol_rel = 1e-11 if T == 0 else 1e-5   # host vs device libm in the float Box-Muller of the noise
    assert rel(lin.cpu().numpy(), rlin) < tol_rel and rel(ang.cpu().numpy(), rang) < tol_rel


def test_fcm_torques_match_reference(cuda, tmp_path):
    from uammd_b200.fcm import GaussianTorque
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_fcm")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_fcm not built")
    N, n, L, tol, eta = 20000, 64, 64.0, 1e-6, 1.1
    h = L / n
    pos, force, torque = _torque_inputs(N, L, 51)
    files = {k: str(tmp_path / f"{k}.bin") for k in ("p", "f", "t", "lin", "ang")}
    pos.tofile(files["p"]); force.tofile(files["f"]); torque.tofile(files["t"])
    out = subprocess.run([ref, "mdott", "gaussian", str(N), repr(L), str(n), repr(eta), repr(tol), "0.0", "0.0", "5", files["p"],
                          files["f"], files["t"], files["lin"], files["ang"]], check=True, capture_output=True, text=True,
                         timeout=600).stdout
    info = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    kern = Gaussian(h, tol)
    kt = GaussianTorque.forHydrodynamicRadius(info["a"], h, tol)
    assert kt.support == info["supportTorque"] and abs(kt.width - info["widthTorque"]) < 1e-14
    fcm = FCM_impl(L, (n, n, n), kern, eta, seed=5, kernelTorque=kt)
    dp, df, dt_ = (torch.from_numpy(x).to(cuda) for x in (pos, force, torque))
    lin, ang = fcm.computeHydrodynamicDisplacements(dp, df, torque=dt_)
    torch.cuda.synchronize()
    rlin = np.fromfile(files["lin"], np.float64).reshape(N, 3)
    rang = np.fromfile(files["ang"], np.float64).reshape(N, 3)
    rel = lambda a_, b_: np.linalg.norm(a_ - b_) / np.linalg.norm(b_)
    print(f"[fcm torques vs reference] linear {rel(lin.cpu().numpy(), rlin):.2e} angular {rel(ang.cpu().numpy(), rang):.2e}")
    assert rel(lin.cpu().numpy(), rlin) < 1e-11 and rel(ang.cpu().numpy(), rang) < 1e-11


def test_spread_dense_cloud_multiple_staging_chunks(orc, cuda):
    """1.5 particles per cell: the row-brick spread stages its region in several chunks (capacity 512 records)."""
    cells, L, N = (32, 32, 32), (32.0,) * 3, 50000
    h = 1.0
    pos = _cloud(N, L, 13)
    val = syn.gaussian_forces(N, seed=14)
    g = orc.make_grid_d(L, cells)
    ref = orc.ibm_spread(g, orc.peskin3(h), pos, val, 34)
    ibm = IBM(Peskin3(h), L, cells, 34)
    grid = torch.full((32, 32, 34, 3), -3.0, dtype=torch.float64, device=cuda)
    ibm.spread(torch.from_numpy(pos).to(cuda), torch.from_numpy(val).to(cuda), grid, overwrite=True)
    sp = grid.cpu().numpy()
    assert np.abs(sp[:, :, :32] - ref[:, :, :32]).max() < 1e-12 * np.abs(ref).max()
    assert np.all(sp[:, :, 32:] == 0)
    # a clustered cloud: every particle in one corner brick, the rest of the grid must come out exactly zero
    pos2 = pos.copy(); pos2[:, :3] = pos2[:, :3] * 0.1 - 14.0
    ref2 = orc.ibm_spread(g, orc.peskin3(h), pos2, val, 34)
    ibm.spread(torch.from_numpy(pos2).to(cuda), torch.from_numpy(val).to(cuda), grid, overwrite=True)
    sp2 = grid.cpu().numpy()
    assert np.abs(sp2[:, :, :32] - ref2[:, :, :32]).max() < 1e-11 * np.abs(ref2).max()
