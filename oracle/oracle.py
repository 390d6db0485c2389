"""TEST INFRASTRUCTURE - ctypes binding of the C restatement oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. Nothing under uammd_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

KERNEL_PESKIN3, KERNEL_PESKIN4, KERNEL_GAUSSIAN, KERNEL_BARNETT_MAGLAND, KERNEL_SIXPOINT = 0, 1, 2, 3, 4


class GridF(C.Structure):
    _fields_ = [("L", C.c_float * 3), ("minusInvL", C.c_float * 3), ("cellDim", C.c_int * 3),
                ("cellSize", C.c_float * 3), ("invCellSize", C.c_float * 3)]


class GridD(C.Structure):
    _fields_ = [("L", C.c_double * 3), ("minusInvL", C.c_double * 3), ("cellDim", C.c_int * 3),
                ("cellSize", C.c_double * 3), ("invCellSize", C.c_double * 3)]


class IBMKernel(C.Structure):
    _fields_ = [("kind", C.c_int), ("support", C.c_int), ("h", C.c_double), ("prefactor", C.c_double),
                ("tau", C.c_double), ("rmax", C.c_double)]


def build():
    """Compile the C oracle (gcc). Safe to call repeatedly."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_morton_hash.restype = C.c_uint32
        _lib.orc_ibm_phi.restype = C.c_double
        _lib.orc_ibm_phi.argtypes = [C.POINTER(IBMKernel), C.c_double]
        _lib.orc_md_scratch_new.restype = C.c_void_p
        _lib.orc_md_scratch_new.argtypes = [C.c_int, C.c_int]
        _lib.orc_md_scratch_free.argtypes = [C.c_void_p]
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def make_grid_f(L, cellDim, periodic=(1, 1, 1)):
    g = GridF()
    lib().orc_grid_init_f(C.byref(g), (C.c_float * 3)(*L), (C.c_int * 3)(*[int(p) for p in periodic]),
                          (C.c_int * 3)(*cellDim))
    return g


def make_grid_d(L, cellDim, periodic=(1, 1, 1)):
    g = GridD()
    lib().orc_grid_init_d(C.byref(g), (C.c_double * 3)(*L), (C.c_int * 3)(*[int(p) for p in periodic]),
                          (C.c_int * 3)(*cellDim))
    return g


def neighbour_celldim(L, rc):
    cd = (C.c_int * 3)()
    lib().orc_neighbour_celldim_f((C.c_float * 3)(*L), C.c_float(rc), cd)
    return tuple(cd)


def get_cells(grid, pos4):
    pos4 = np.ascontiguousarray(pos4, dtype=np.float32)
    out = np.zeros((pos4.shape[0], 3), dtype=np.int32)
    c = (C.c_int * 3)()
    l = lib()
    for i in range(pos4.shape[0]):
        l.orc_get_cell_f(C.byref(grid), _p(pos4[i]), c)
        out[i] = c[:]
    return out


def celllist_build(grid, pos4):
    pos4 = np.ascontiguousarray(pos4, dtype=np.float32)
    N = pos4.shape[0]
    ncells = grid.cellDim[0] * grid.cellDim[1] * grid.cellDim[2]
    sortPos = np.empty((N, 4), np.float32)
    index = np.empty(N, np.int32)
    cs = np.empty(ncells, np.int32)
    ce = np.empty(ncells, np.int32)
    err = lib().orc_celllist_build_f(C.byref(grid), _p(pos4), N, _p(sortPos), _p(index), _p(cs), _p(ce))
    return dict(sortPos=sortPos, index=index, cellStart=cs, cellEnd=ce, error=err)


def lj_f32(grid, cl, params, ntypes, N, energy=False, virial=False):
    force = np.zeros((N, 4), np.float32)
    e = np.zeros(N, np.float32) if energy else None
    v = np.zeros(N, np.float32) if virial else None
    params = np.ascontiguousarray(params, np.float32)
    lib().orc_lj_f32(C.byref(grid), _p(cl["sortPos"]), _p(cl["index"]), _p(cl["cellStart"]), _p(cl["cellEnd"]),
                     N, _p(params), ntypes, _p(force), _p(e), _p(v))
    return force, e, v


def _sep_uncertainty(L, rc):
    """fp32 uncertainty of a pair separation: see LJScale."""
    return 2.0 ** -23 * (float(np.max(L)) + 8.0 * rc)


def lj_f64(grid, cl, params, ntypes, N):
    force = np.zeros((N, 3), np.float64)
    e = np.zeros(N, np.float64)
    v = np.zeros(N, np.float64)
    a = np.zeros(N, np.float64)
    s = np.zeros(N, np.float64)
    g = np.zeros(N, np.float64)
    params = np.ascontiguousarray(params, np.float32)
    rcmax = float(np.sqrt(params.reshape(-1, 4)[:, 0].max()))
    band = 2.0 * rcmax * 2.0 * _sep_uncertainty(tuple(grid.L), rcmax)  # |d r2| = 2 r |dr|, both particles
    lib().orc_lj_f64(C.byref(grid), _p(cl["sortPos"]), _p(cl["index"]), _p(cl["cellStart"]), _p(cl["cellEnd"]),
                     N, _p(params), ntypes, _p(force), _p(e), _p(v), _p(a), _p(s), C.c_double(band), _p(g))
    return force, e, v, LJScale(a, s, g)


class LJScale:
    """Per particle error scales from the fp64 pass: abssum = sum_j |f_ij|, sens = sum_j |df_ij/dr|.

    force_tol(L, rc): the fp32 error model used by the parity tests. A separation is known to
    ~2^-23 * max(|x|) <= 2^-23 * L (positions are fp32 numbers of box scale and pairs across the periodic
    boundary pick up one more rounding of that size), r2 and the force polynomial add a few relative
    roundings (8 * 2^-23 * rc in separation terms), and the sum itself 2e-6 relative to sum |f_ij|.
    The unshifted LJ force is discontinuous at the cut-off (|f(rc)| ~ 0.04 eps/sigma): edge = sum of |f_ij|
    over the pairs whose r2 lies within that separation uncertainty of rc2, which either side may count."""

    def __init__(self, abssum, sens, edge=0.0):
        self.abssum, self.sens, self.edge = abssum, sens, edge

    def force_tol(self, L, rc):
        return _sep_uncertainty(L, rc) * self.sens + 2e-6 * self.abssum + 1.01 * self.edge + 1e-30

    def energy_tol(self, L, rc, escale):
        """|dE_i| <= 1/2 sum_j |dE_ij/dr| dsep = dsep/2 * abssum; a pair inside the cut-off band adds or drops
        |e(rc)|/2 ~ 0.21 |f(rc)|; summation roundings relative to the largest per-particle energy."""
        return 0.5 * _sep_uncertainty(L, rc) * self.abssum + 0.25 * self.edge + 4e-6 * escale + 1e-30

    def virial_tol(self, L, rc, vscale):
        """V_ij = f_ij . r_ij: |dV| <= rc |df| + |f| dsep."""
        return rc * (_sep_uncertainty(L, rc) * self.sens + 1.01 * self.edge) + _sep_uncertainty(L, rc) * self.abssum \
            + 4e-6 * vscale + 1e-30

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.abssum, dtype=dtype)


def dpd_f32(grid, cl, vel3, A, gamma, sigma, rcut, seed, step, N):
    force = np.zeros((N, 4), np.float32)
    force64 = np.zeros((N, 3), np.float64)
    vel3 = np.ascontiguousarray(vel3, np.float32)
    lib().orc_dpd_f32(C.byref(grid), _p(cl["sortPos"]), _p(cl["index"]), _p(cl["cellStart"]), _p(cl["cellEnd"]),
                      N, _p(vel3), C.c_float(A), C.c_float(gamma), C.c_float(sigma), C.c_float(rcut),
                      C.c_uint32(seed), C.c_uint32(step), 0, _p(force), _p(force64))
    return force, force64


def dpd_ids_f32(grid, cl, vel3, A, gamma, sigma, rcut, seed, step, N, noiseId, idStride):
    """DPD forces on local arrays with the noise keyed on noiseId[array index] (brick decomposition)."""
    force = np.zeros((N, 4), np.float32)
    vel3 = np.ascontiguousarray(vel3, np.float32)
    noiseId = np.ascontiguousarray(noiseId, np.int32)
    lib().orc_dpd_ids_f32(C.byref(grid), _p(cl["sortPos"]), _p(cl["index"]), _p(cl["cellStart"]), _p(cl["cellEnd"]),
                          N, _p(vel3), C.c_float(A), C.c_float(gamma), C.c_float(sigma), C.c_float(rcut),
                          C.c_uint32(seed), C.c_uint32(step), 0, _p(force), None, _p(noiseId), int(idStride))
    return force


def brick_classify(grid, pos4, rankGrid):
    pos4 = np.ascontiguousarray(pos4, dtype=np.float32)
    N = pos4.shape[0]
    cell, owner, mask = np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.uint32)
    lib().orc_brick_classify_f(C.byref(grid), _p(pos4), N, (C.c_int * 3)(*[int(p) for p in rankGrid]), _p(cell), _p(owner),
                               _p(mask))
    return cell, owner, mask


def saru3_u32(s1, s2, s3, n):
    out = np.zeros(n, np.uint32)
    lib().orc_saru3_u32(C.c_uint32(s1), C.c_uint32(s2), C.c_uint32(s3), n, _p(out))
    return out


def saru3_gf(s1, s2, s3, mean, std):
    out = (C.c_float * 2)()
    lib().orc_saru3_gf(C.c_uint32(s1), C.c_uint32(s2), C.c_uint32(s3), C.c_float(mean), C.c_float(std), out)
    return out[0], out[1]


def nve_half(pos4, vel3, force4, dt, mass, step):
    lib().orc_nve_half_f32(_p(pos4), _p(vel3), _p(force4), pos4.shape[0], C.c_float(dt), C.c_float(mass), step)


def nvt_gj_half(pos4, vel3, force4, dt, friction, noiseAmplitude, stepNum, seed, step, defaultMass=1.0, mass=None,
                is2D=False):
    """In place VerletNVT::GronbechJensen half step (float32 arrays)."""
    lib().orc_nvt_gj_half_f32(_p(pos4), _p(vel3), _p(force4), _p(mass), C.c_float(defaultMass), pos4.shape[0],
                              C.c_float(dt), C.c_float(friction), int(is2D), C.c_float(noiseAmplitude),
                              C.c_uint32(stepNum), C.c_uint32(seed), int(step))


def nvt_basic_half(pos4, vel3, force4, dt, friction, noiseAmplitude, stepNum, seed, step, defaultMass=1.0, mass=None,
                   is2D=False, Ngroup=None):
    """In place VerletNVT::Basic half step (float32 arrays). Ngroup: size of the whole group when the arrays hold a prefix."""
    lib().orc_nvt_basic_half_f32(_p(pos4), _p(vel3), _p(force4), _p(mass), C.c_float(defaultMass), pos4.shape[0],
                                 int(Ngroup if Ngroup is not None else pos4.shape[0]), C.c_float(dt), C.c_float(friction), int(is2D), C.c_float(noiseAmplitude),
                                 C.c_uint32(stepNum), C.c_uint32(seed), int(step))


def nvt_initial_velocities(N, vamp, seed, is2D=False):
    vel = np.zeros((N, 3), np.float32)
    lib().orc_nvt_initial_velocities_f32(_p(vel), N, C.c_float(vamp), int(is2D), C.c_uint32(seed))
    return vel


class MDOracle:
    """CPU (OpenMP) VerletNVE + PairForces<LJ,CellList>; used as correctness oracle and cpu_baseline."""

    def __init__(self, L, rc, params, dt, pos4, vel3):
        self.L = (C.c_float * 3)(*L)
        self.rc, self.dt = float(rc), float(dt)
        self.params = np.ascontiguousarray(params, np.float32)
        self.pos = np.ascontiguousarray(pos4, np.float32).copy()
        self.vel = np.ascontiguousarray(vel3, np.float32).copy()
        self.N = self.pos.shape[0]
        cd = neighbour_celldim(L, rc)
        self.grid = make_grid_f(L, cd)
        self.ncells = cd[0] * cd[1] * cd[2]
        self.scratch = lib().orc_md_scratch_new(self.N, self.ncells)
        self.force = np.zeros((self.N, 4), np.float32)
        cl = celllist_build(self.grid, self.pos)
        self.force, _, _ = lj_f32(self.grid, cl, self.params, 1, self.N)

    def step(self, n=1):
        for _ in range(n):
            rc = lib().orc_md_step_f32(self.L, C.c_float(self.rc), _p(self.params), C.c_float(self.dt), self.N,
                                       _p(self.pos), _p(self.vel), _p(self.force), C.c_void_p(self.scratch))
            if rc:
                raise RuntimeError(f"oracle md step failed: {rc}")

    def __del__(self):
        try:
            lib().orc_md_scratch_free(C.c_void_p(self.scratch))
        except Exception:
            pass


# ---------------- path 2 ----------------
def peskin3(h):
    return IBMKernel(KERNEL_PESKIN3, 3, h, 0.0, 0.0, 0.0)


def peskin4(h):
    return IBMKernel(KERNEL_PESKIN4, 4, h, 0.0, 0.0, 0.0)


def barnett_magland(alpha, beta, support):
    """IBM_kernels::BarnettMagland(alpha, beta) (misc/IBM_kernels.cuh:99-113) with the support its wrapper declares."""
    _lib.orc_ibm_bm_norm.restype = C.c_double
    _lib.orc_ibm_bm_norm.argtypes = [C.c_double, C.c_double]
    norm = _lib.orc_ibm_bm_norm(alpha, beta)
    return IBMKernel(KERNEL_BARNETT_MAGLAND, support, 2.0 * alpha / support, 1.0 / norm, beta, alpha)


def six_point(h):
    return IBMKernel(KERNEL_SIXPOINT, 6, h, 0.0, 0.0, 0.0)


def gaussian_fcm(h, tolerance):
    """FCM_ns::Kernels::Gaussian constructor (Integrator/BDHI/FCM/FCM_kernels.cuh:22-46)."""
    amin, amax = 0.55, 1.65
    x = -np.log10(3 * tolerance) / 10.0
    ups = min(amin + x * (amax - amin), amax)
    width = h * ups
    pref = (2.0 * np.pi * width * width) ** -0.5
    tau = -0.5 / (width * width)
    dr = 0.5 * h
    r = dr
    while pref * np.exp(tau * r * r) > tolerance:
        r += dr
    support = max(3, int(2 * r / h + 0.5))
    k = IBMKernel(KERNEL_GAUSSIAN, support, h, pref, tau, support * h)
    k_a = h * ups * np.sqrt(np.pi)
    return k, k_a


def ibm_spread(grid, kern, pos4, val3, nxPad):
    pos4 = np.ascontiguousarray(pos4, np.float64)
    val3 = np.ascontiguousarray(val3, np.float64)
    n = grid.cellDim
    out = np.zeros((n[2], n[1], nxPad, 3), np.float64)
    lib().orc_ibm_spread_d(C.byref(grid), C.byref(kern), _p(pos4), _p(val3), pos4.shape[0], nxPad, _p(out))
    return out


def ibm_gather(grid, kern, pos4, grid3, nxPad):
    pos4 = np.ascontiguousarray(pos4, np.float64)
    grid3 = np.ascontiguousarray(grid3, np.float64)
    out = np.zeros((pos4.shape[0], 3), np.float64)
    lib().orc_ibm_gather_d(C.byref(grid), C.byref(kern), _p(pos4), pos4.shape[0], nxPad, _p(grid3), _p(out))
    return out


def fcm_force2vel(grid, viscosity, ghat):
    """ghat: complex128 [nz, ny, nx/2+1, 3]; returns the scaled copy."""
    g = np.ascontiguousarray(ghat, np.complex128).copy()
    lib().orc_fcm_force2vel_d(C.byref(grid), C.c_double(viscosity), _p(g))
    return g


def fcm_add_noise(grid, viscosity, noisePrefactor, seed1, seed2, ghat):
    g = np.ascontiguousarray(ghat, np.complex128).copy()
    lib().orc_fcm_add_noise_d(C.byref(grid), C.c_double(viscosity), C.c_double(noisePrefactor), C.c_uint32(seed1),
                              C.c_uint32(seed2), _p(g))
    return g


def dft3_r2c(grid3, nx):
    nz, ny, nxPad, _ = grid3.shape
    out = np.zeros((nz, ny, nx // 2 + 1, 3), np.complex128)
    g = np.ascontiguousarray(grid3, np.float64)
    lib().orc_dft3_r2c_d(nx, ny, nz, nxPad, _p(g), _p(out))
    return out


def dft3_c2r(ghat, nx, nxPad):
    nz, ny, _, _ = ghat.shape
    out = np.zeros((nz, ny, nxPad, 3), np.float64)
    g = np.ascontiguousarray(ghat, np.complex128)
    lib().orc_dft3_c2r_d(nx, ny, nz, nxPad, _p(g), _p(out))
    return out


def gaussian_torque(width, h, tolerance):
    """FCM_ns::Kernels::GaussianTorque (Integrator/BDHI/FCM/FCM_kernels.cuh:60-80)."""
    pref = (2.0 * np.pi * width * width) ** -0.5
    tau = -0.5 / (width * width)
    dr = 0.5 * h
    r = dr
    while pref * np.exp(tau * r * r) > tolerance:
        r += dr
    support = max(3, int(2 * r / h + 0.5))
    return IBMKernel(KERNEL_GAUSSIAN, support, h, pref, tau, support * h)


def _half_curl_fourier(L, cells, vhat):
    """1/2 i dk x v in Fourier space with the unpaired (Nyquist) components of dk zeroed: addTorqueCurl /
    computeVelocityCurlFourier (FCM_impl.cuh:306-325,593-615), getGradientFourier (FCM/utils.cuh:41-51)."""
    nx, ny, nz = cells

    def dk(n, Ld, m):
        i = np.arange(m)
        f = i - n * (i >= (n // 2 + 1))
        k = 2.0 * np.pi / Ld * f
        k[f == n - f] = 0.0
        return k
    kx = dk(nx, L[0], nx // 2 + 1)[None, None, :]
    ky = dk(ny, L[1], ny)[None, :, None]
    kz = dk(nz, L[2], nz)[:, None, None]
    vx, vy, vz = vhat[..., 0], vhat[..., 1], vhat[..., 2]
    out = np.empty_like(vhat)
    out[..., 0] = 0.5j * (ky * vz - kz * vy)
    out[..., 1] = 0.5j * (kz * vx - kx * vz)
    out[..., 2] = 0.5j * (kx * vy - ky * vx)
    return out


def fcm_mdot(L, cells, kern, viscosity, pos4, force3, temperature=0.0, prefactor=0.0, seed=0, seed2=1, torque3=None,
             kernTorque=None):
    """FCM pipeline FCM_impl::computeHydrodynamicDisplacements (Integrator/BDHI/FCM/FCM_impl.cuh:652-693)
    with numpy.fft standing in for cuFFT. force3 may be None (noise only); seed2 = number of noisy calls so far.
    With torque3 (and kernTorque) returns the pair (linear, angular) like the reference."""
    g = make_grid_d(L, cells)
    nx, ny, nz = cells
    nxPad = 2 * (nx // 2 + 1)
    if force3 is not None:
        sp = ibm_spread(g, kern, pos4, force3, nxPad)
        ghat = np.fft.rfftn(sp[:, :, :nx, :], axes=(0, 1, 2))
    else:
        ghat = np.zeros((nz, ny, nx // 2 + 1, 3), np.complex128)
    if torque3 is not None:
        spt = ibm_spread(g, kernTorque, pos4, torque3, nxPad)
        ghat = ghat + _half_curl_fourier(L, cells, np.fft.rfftn(spt[:, :, :nx, :], axes=(0, 1, 2)))
    if force3 is not None or torque3 is not None:
        ghat = fcm_force2vel(g, viscosity, ghat)
    if temperature > 0:
        dV = g.cellSize[0] * g.cellSize[1] * g.cellSize[2]
        noisePrefactor = prefactor * np.sqrt((1.0 / (float(nx) * ny * nz)) * 2 * temperature / dV)
        ghat = fcm_add_noise(g, viscosity, noisePrefactor, seed, seed2, ghat)

    vel = np.zeros((nz, ny, nxPad, 3))
    vel[:, :, :nx, :] = np.fft.irfftn(ghat, s=(nz, ny, nx), axes=(0, 1, 2)) * (nx * ny * nz)
    linear = ibm_gather(g, kern, pos4, vel, nxPad)
    if torque3 is None:
        return linear
    ang = np.zeros((nz, ny, nxPad, 3))
    ang[:, :, :nx, :] = np.fft.irfftn(_half_curl_fourier(L, cells, ghat), s=(nz, ny, nx), axes=(0, 1, 2)) * (nx * ny * nz)
    return linear, ibm_gather(g, kernTorque, pos4, ang, nxPad)


def bd_euler_maruyama_f64(pos4, force4, selfMobility, dt, temperature, step, seed, K9=None, radius=None, is2D=False):
    """In place BD::EulerMaruyama update (BrownianDynamics.cu:117-145), fp64."""
    assert pos4.dtype == np.float64 and pos4.flags.c_contiguous
    k = np.ascontiguousarray(K9, np.float64) if K9 is not None else None
    lib().orc_bd_euler_maruyama_f64(_p(pos4), _p(force4), _p(k), C.c_double(selfMobility), _p(radius), C.c_double(dt),
                                    int(is2D), C.c_double(temperature), pos4.shape[0], C.c_uint32(step), C.c_uint32(seed))


# ---------------- BDHI::PSE ----------------
def next_fft_wise_size(n):
    """nextFFTWiseSize3D (utils/Grid.cuh:142-213), one dimension."""
    best = None
    for m in range(4):
        for l in range(5):
            for k in range(6):
                base = 11 ** m * 7 ** l * 5 ** k
                p3 = 1
                while base * p3 <= (1 << 40):
                    p2 = 2
                    while base * p3 * p2 < n:
                        p2 *= 2
                    v = base * p3 * p2
                    if best is None or v < best:
                        best = v
                    p3 *= 3
    return best


def pse_params(L, viscosity, rh, tolerance, psi, cells=None):
    """Derived PSE parameters in double precision: NearField::initializeDeterministicPart (PSE/NearField.cuh:65-102),
    FarField::initializeGrid / initializeKernel (PSE/FarField.cuh:605-654)."""
    import math
    L = (L, L, L) if np.isscalar(L) else tuple(L)
    rcut = math.sqrt(-math.log(tolerance)) / psi
    nTable = int(min(1 << 22, max(1 << 14, int(rcut / (rh * tolerance) + 0.5))))
    kcut = 2 * psi * math.sqrt(-math.log(tolerance))
    hgrid = 2 * math.pi / kcut
    if cells is None:
        cells = tuple(next_fft_wise_size(int(2 * L[d] / hgrid) + 1) for d in range(3))
    Cc, m = 0.976, 1.0
    while math.erfc(m / math.sqrt(2)) > 0.1 * tolerance:
        m += 0.01
    while True:
        support = int((m / Cc) ** 2 / math.pi + 0.5) + 1
        if support % 2 == 1:
            break
        m += tolerance
    P = support // 2
    if support > min(cells):
        support = min(cells)
        if support % 2 == 0:
            support -= 1
        P = support // 2
        m = Cc * math.sqrt(math.pi * support)
    h = min(L[d] / cells[d] for d in range(3))
    w = (2 * P + 1) * h / 2.0
    eta = (2.0 * psi * w / m) ** 2
    width = math.sqrt(eta) / (2.0 * psi)
    kern = IBMKernel(KERNEL_GAUSSIAN, 2 * P + 1, h, (1.0 / (width ** 3 * (2.0 * math.pi) ** 1.5)) ** (1.0 / 3.0),
                     -0.5 / (width * width), float("inf"))
    return dict(L=L, cells=tuple(cells), support=2 * P + 1, eta=eta, rcut=rcut, nTable=nTable, kernel=kern,
                normalization=6 * math.pi * rh * viscosity)


def pse_far_mdot(par, viscosity, rh, psi, pos4, force3, shear=0.0, temperature=0.0, prefactor=0.0, seed=0, seed2=0):
    """FarField::computeHydrodynamicDisplacements (PSE/FarField.cuh:535-553) with numpy.fft standing in for cuFFT."""
    L, cells, kern = par["L"], par["cells"], par["kernel"]
    g = make_grid_d(L, cells)
    nx, ny, nz = cells
    nxPad = 2 * (nx // 2 + 1)
    if force3 is not None:
        sp = ibm_spread(g, kern, pos4, force3, nxPad)
        ghat = np.ascontiguousarray(np.fft.rfftn(sp[:, :, :nx, :], axes=(0, 1, 2)))
        lib().orc_pse_force2vel_d(C.byref(g), C.c_double(shear), C.c_double(rh), C.c_double(viscosity), C.c_double(psi),
                                  C.c_double(par["eta"]), _p(ghat))
    else:
        ghat = np.zeros((nz, ny, nx // 2 + 1, 3), np.complex128)
    if temperature > 0:
        dV = g.cellSize[0] * g.cellSize[1] * g.cellSize[2]
        lib().orc_pse_add_noise_d(C.byref(g), C.c_double(shear), C.c_double(rh), C.c_double(viscosity), C.c_double(psi),
                                  C.c_double(par["eta"]), C.c_double(prefactor * np.sqrt(2 * temperature / dV)),
                                  C.c_uint32(seed), C.c_uint32(seed2), _p(ghat))
    vel = np.zeros((nz, ny, nxPad, 3))
    vel[:, :, :nx, :] = np.fft.irfftn(ghat, s=(nz, ny, nx), axes=(0, 1, 2)) * (nx * ny * nz)
    return ibm_gather(g, kern, pos4, vel, nxPad)


def rpy_near_fg(r, rh, psi, rcut):
    out = (C.c_double * 2)()
    lib().orc_rpy_near_fg(C.c_double(r), C.c_double(rh), C.c_double(psi), C.c_double(rcut), out)
    return out[0], out[1]


def pse_near_table(par, rh, psi):
    t = np.zeros((par["nTable"], 2))
    lib().orc_pse_near_table_d(par["nTable"], C.c_double(rh), C.c_double(psi), C.c_double(par["normalization"]),
                               C.c_double(par["rcut"]), _p(t))
    return t


def pse_near_mdot(par, rh, psi, pos4, v, shear=0.0, table=None):
    """Direct O(N^2) near-field mat-vec (NearField.cuh:131-182); table=None evaluates F, G in closed form."""
    pos4 = np.ascontiguousarray(pos4, np.float64)
    v = np.ascontiguousarray(v, np.float64)
    out = np.zeros((pos4.shape[0], 3))
    lib().orc_pse_near_mdot_d(pos4.shape[0], _p(pos4), _p(v), v.shape[1], (C.c_double * 3)(*par["L"]), C.c_double(shear),
                              C.c_double(rh), C.c_double(psi), C.c_double(par["normalization"]), C.c_double(par["rcut"]),
                              _p(table), par["nTable"], 1 if table is not None else 0, _p(out))
    return out
