"""CPU: the oracle restatement against golden vectors produced by the reference's own host code
(tests/golden/gen_golden*.cu, run in the build container) and against known answers."""
import numpy as np

from uammd_b200 import synthetic as syn


def test_saru_streams_match_reference(orc, golden_dir):
    raw = np.fromfile(f"{golden_dir}/saru.bin", dtype=np.uint32).reshape(64, 15)
    for row in raw:
        s1, s2, s3 = (int(x) for x in row[:3])
        assert np.array_equal(orc.saru3_u32(s1, s2, s3, 8), row[3:11])
        fl = row[11:15].view(np.float32)
        u = orc.saru3_u32(s1, s2, s3, 2)
        f = ((u >> 1).astype(np.int32)).astype(np.float32) * np.float32(1.0 / 2147483648.0)
        assert np.array_equal(f, fl[:2])
        g = orc.saru3_gf(s1, s2, s3, 0.5, 2.0)
        # host libm sinf/cosf/logf of the generator build vs ours: same glibc -> exact
        assert np.allclose(g, fl[2:], rtol=1e-6, atol=1e-6)


def test_morton_hash_matches_reference(orc, golden_dir):
    raw = np.fromfile(f"{golden_dir}/morton.bin", dtype=np.uint32).reshape(-1, 4)
    for cx, cy, cz, h in raw[:1024]:
        assert orc.lib().orc_morton_hash(int(cx), int(cy), int(cz)) == int(h)


def test_getcell_f32_matches_reference(orc, golden_dir):
    raw = np.fromfile(f"{golden_dir}/getcell_f32.bin", dtype=np.uint8)
    off = 0
    for _ in range(3):
        L = raw[off:off + 12].view(np.float32); off += 12
        cd = raw[off:off + 12].view(np.int32); off += 12
        rec = raw[off:off + 4096 * 24].reshape(4096, 24); off += 4096 * 24
        pts = rec[:, :12].copy().view(np.float32)
        cells = rec[:, 12:].copy().view(np.int32)
        g = orc.make_grid_f(tuple(float(x) for x in L), tuple(int(x) for x in cd))
        pos4 = np.zeros((4096, 4), np.float32); pos4[:, :3] = pts
        mine = orc.get_cells(g, pos4)
        # the generator is a host build without FMA contraction; the oracle mirrors the device (FMA) arithmetic.
        # They may differ only for points within one rounding of a cell face.
        bad = np.any(mine != cells, axis=1)
        assert bad.sum() <= 2, f"{bad.sum()} cell mismatches"
        assert mine.min() >= 0 and np.all(mine < cd[None, :])


def test_two_particle_lj_kat(orc):
    # examples/uammd_as_a_library/wrapper.py:32 : two particles at r = sigma feel F = -/+ 24 eps/sigma
    L = (20.0, 20.0, 20.0)
    g = orc.make_grid_f(L, orc.neighbour_celldim(L, 2.5))
    pos = np.zeros((2, 4), np.float32); pos[0, 0] = -0.5; pos[1, 0] = 0.5
    cl = orc.celllist_build(g, pos)
    f, e, v = orc.lj_f32(g, cl, syn.lj_params(), 1, 2, energy=True, virial=True)
    assert np.allclose(f[:, 0], [-24.0, 24.0]) and np.all(f[:, 1:] == 0)
    assert np.allclose(e, 0.0, atol=1e-6)  # U(sigma) = 0
    assert np.allclose(v, -24.0)           # F . r12 per particle


def test_neighbour_grid_rule(orc):
    # CellList.cuh:100-126
    assert orc.neighbour_celldim((107.7217,) * 3, 2.5) == (43, 43, 43)
    assert orc.neighbour_celldim((10.0, 7.4, 100.0), 2.5) == (4, 1, 40)


def _brute_force_lj(pos, L, par):
    x = pos[:, :3].astype(np.float64)
    d = x[None, :, :] - x[:, None, :]
    d -= np.floor(d / L + 0.5) * L
    r2 = (d * d).sum(-1)
    np.fill_diagonal(r2, np.inf)
    inr = r2 < par[0]
    r2s = np.where(inr, r2, 1.0)
    invr2 = par[1] / r2s
    invr6 = invr2 ** 3
    fm = np.where(inr, par[2] * (-48.0 * invr6 + 24.0) * invr6 * invr2, 0.0)
    return (fm[:, :, None] * d).sum(1)


def test_cell_traversal_equals_brute_force(orc):
    N = 1500
    Lb = syn.lj_box_length(N, 0.5)
    L = np.array([Lb, Lb, Lb])
    pos = syn.uniform_cloud(N, Lb, seed=3)
    par = syn.lj_params()
    g = orc.make_grid_f(tuple(L), orc.neighbour_celldim(tuple(L), 2.5))
    cl = orc.celllist_build(g, pos)
    f64, _, _, sc = orc.lj_f64(g, cl, par, 1, N)
    ref = _brute_force_lj(pos, L.astype(np.float32).astype(np.float64), par.astype(np.float64))
    assert np.max(np.abs(f64 - ref) / np.maximum(sc.abssum, 1e-30)[:, None]) < 1e-10
    f32, _, _ = orc.lj_f32(g, cl, par, 1, N)
    assert np.max(np.abs(f32[:, :3] - f64) / sc.force_tol(L, 2.5)[:, None]) < 1.0


def test_celllist_invariants(orc):
    N = 5000
    L = (30.0, 22.0, 41.0)
    pos = syn.uniform_cloud(N, L, seed=8)
    pos[::7, :3] *= 3.0  # some particles outside the primary box
    g = orc.make_grid_f(L, (9, 7, 12))
    cl = orc.celllist_build(g, pos)
    assert cl["error"] == 0
    assert np.array_equal(np.sort(cl["index"]), np.arange(N))
    assert np.array_equal(cl["sortPos"], pos[cl["index"]])
    cells = orc.get_cells(g, cl["sortPos"])
    keys = np.array([orc.lib().orc_morton_hash(int(c[0]), int(c[1]), int(c[2])) for c in cells], dtype=np.int64)
    assert np.all(np.diff(keys) >= 0)
    same = np.diff(keys) == 0
    assert np.all(np.diff(cl["index"])[same] > 0)  # stable: ties keep ascending original index
    lin = cells[:, 0] + 9 * (cells[:, 1] + 7 * cells[:, 2])
    for c in np.unique(lin)[:200]:
        w = np.nonzero(lin == c)[0]
        assert cl["cellStart"][c] == w[0] and cl["cellEnd"][c] == w[-1] + 1
    empty = np.setdiff1d(np.arange(9 * 7 * 12), lin)
    assert np.all(cl["cellStart"][empty] == -1)


def test_dpd_pairwise_momentum_conservation(orc):
    # Frij = Frji by seeding Saru with (min,max) (DPD.cuh:126-129): total force vanishes
    N = 3000
    L = (10.0, 10.0, 10.0)
    pos = syn.uniform_cloud(N, L, seed=21)
    vel = syn.maxwell_velocities(N, 1.0, seed=22)
    g = orc.make_grid_f(L, orc.neighbour_celldim(L, 1.0))
    cl = orc.celllist_build(g, pos)
    sigma = np.sqrt(2.0 * 1.0) / np.sqrt(0.01)
    f32, f64 = orc.dpd_f32(g, cl, vel, 25.0, 4.5, sigma, 1.0, 1234, 7, N)
    assert np.abs(f64.sum(0)).max() < 1e-6 * np.abs(f64).sum()
    assert np.abs(f64).max() > 1.0
    f32b, _ = orc.dpd_f32(g, cl, vel, 25.0, 4.5, sigma, 1.0, 1234, 8, N)
    assert not np.allclose(f32, f32b)  # step changes the noise


def test_nve_energy_conservation_cpu(orc):
    N = 2048
    Lb = syn.lj_box_length(N, 0.8)
    pos = syn.fcc_lattice(N, Lb)
    vel = syn.maxwell_velocities(N, 0.5)
    par = syn.lj_params()
    md = orc.MDOracle((Lb,) * 3, 2.5, par, 0.004, pos, vel)
    # the truncated force is the derivative of the SHIFTED potential (continuous at rc): measure that energy
    par_shift = syn.lj_params(shift=True)

    def energy():
        cl = orc.celllist_build(md.grid, md.pos)
        _, e, _, _ = orc.lj_f64(md.grid, cl, par_shift, 1, N)
        return e.sum() + 0.5 * (md.vel.astype(np.float64) ** 2).sum()

    e0 = energy()
    md.step(50)
    e1 = energy()
    assert abs(e1 - e0) / N < 2e-3
