"""Seeded synthetic particle clouds for tests and bench (numpy only)."""
import numpy as np


def lj_box_length(N, rho=0.8):
    return float(np.float32((N / rho) ** (1.0 / 3.0)))


def uniform_cloud(N, L, seed=2024, ntypes=1):
    """Uniform random positions in [-L/2, L/2)^3 as real4 (x, y, z, type)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    L = np.broadcast_to(np.asarray(L, dtype=np.float64), (3,))
    pos = np.zeros((N, 4), dtype=np.float32)
    pos[:, :3] = ((rng.random((N, 3)) - 0.5) * L).astype(np.float32)
    if ntypes > 1:
        pos[:, 3] = rng.integers(0, ntypes, N).astype(np.float32)
    return pos


def fcc_lattice(N, L):
    """First N sites of the smallest FCC lattice with >= N sites that fills a cubic box of side L."""
    n = int(np.ceil((N / 4.0) ** (1.0 / 3.0) - 1e-9))
    a = L / n
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    ii = np.arange(n)
    cells = np.stack(np.meshgrid(ii, ii, ii, indexing="ij"), -1).reshape(-1, 1, 3)
    sites = ((cells + basis[None]) * a + 0.25 * a - 0.5 * L).reshape(-1, 3)
    pos = np.zeros((N, 4), dtype=np.float32)
    pos[:, :3] = sites[:N].astype(np.float32)
    return pos


def maxwell_velocities(N, T=1.0, seed=7):
    rng = np.random.Generator(np.random.MT19937(seed))
    v = rng.normal(0.0, np.sqrt(T), (N, 3))
    v -= v.mean(axis=0, keepdims=True)
    return v.astype(np.float32)


def gaussian_forces(N, seed=12, dtype=np.float64):
    rng = np.random.Generator(np.random.MT19937(seed))
    return rng.normal(0.0, 1.0, (N, 3)).astype(dtype)


def lj_params(sigma=1.0, epsilon=1.0, rc=2.5, shift=False):
    """One row of LJFunctor::PairParameters {cutOff2, sigma2, epsDivSigma2, shift} in fp32."""
    f = np.float32
    s2 = f(f(sigma) * f(sigma))
    c2 = f(f(rc) * f(rc))
    sh = f(0)
    if shift:
        i2 = f(s2 / c2)
        i6 = f(f(i2 * i2) * i2)
        sh = f(f(f(f(epsilon) * f(4)) * i6) * f(i6 - f(1)))
    return np.array([c2, s2, f(f(epsilon) / s2), sh], dtype=np.float32)


class Xorshift128plus:
    """The reference's host generator (utils/utils.h:38-115), restated: used to reproduce the README example's
    initial positions (sys->rng().uniform3) and the Saru seed an integrator draws at construction (next32)."""
    M64 = (1 << 64) - 1

    def __init__(self, s0=None):
        if s0 is None:
            self.s = [12679825035178159220, 15438657923749336752]
        else:
            self.setSeed(s0)

    def setSeed(self, s0):
        self.s = [s0 & self.M64, ((s0 + 15438657923749336752) & self.M64) % self.M64]

    def next(self):
        x, y = self.s
        self.s[0] = y
        x ^= (x << 23) & self.M64
        x ^= x >> 17
        x ^= y ^ (y >> 26)
        self.s[1] = x
        return (x + y) & self.M64

    def next32(self):
        return self.next() % 0xFFFFFFFF

    def uniform(self, lo, hi):
        return lo + (self.next() / float(self.M64)) * (hi - lo)

    def uniform3(self, lo, hi):
        return (self.uniform(lo, hi), self.uniform(lo, hi), self.uniform(lo, hi))
