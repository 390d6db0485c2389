"""TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH. CPU (numpy, float64) restatement of the reference's spectral Ewald
Poisson solver (/root/reference/src/Interactor/SpectralEwaldPoisson.cu); file:line citations are relative to that file.
Only tests/ may import this module. Pinned by the reference's own analytic known answers (test/Potentials/Poisson/
TriplyPeriodic/test_poisson.cu:13-23: field and potential of two Gaussian charges) and, on a GPU box, by the compiled
reference (oracle/_ref/dropin_poisson)."""
import math

import numpy as np
from scipy.special import erf


def greens_function(r2, gw, split, epsilon):
    """Poisson_ns::greensFunction :16-39 (argument r^2)."""
    r2 = np.asarray(r2, np.float64)
    out = np.empty_like(r2)
    big = r2 > gw ** 4
    r = np.sqrt(r2[big])
    out[big] = 1.0 / (4.0 * math.pi * epsilon * r) * (erf(r / (2 * gw)) - erf(r / math.sqrt(4 * gw * gw + 1 / (split * split))))
    pi32 = math.pi ** 1.5
    gw2, invsp2 = gw * gw, 1.0 / (split * split)
    selfterm = 1.0 / (4 * pi32 * gw) - 1.0 / (2 * pi32 * math.sqrt(4 * gw2 + invsp2))
    r2term = 1.0 / (6.0 * pi32 * (4.0 * gw2 + invsp2) ** 1.5) - 1.0 / (48.0 * pi32 * gw2 * gw)
    r4term = 1.0 / (640.0 * pi32 * gw2 * gw2 * gw) - 1.0 / (20.0 * pi32 * (4 * gw2 + invsp2) ** 2.5)
    s = r2[~big]
    out[~big] = 1.0 / epsilon * (selfterm + s * r2term + s * s * r4term)
    return out


def greens_function_field(r, gw, split, epsilon):
    """Poisson_ns::greensFunctionField :41-63 (argument r)."""
    r = np.asarray(r, np.float64)
    r2, gw2 = r * r, gw * gw
    newgw = math.sqrt(gw2 + 1 / (4.0 * split * split))
    newgw2 = newgw * newgw
    out = np.zeros_like(r)
    big = r2 > gw ** 4
    rb, r2b = r[big], r2[big]
    invrterm = np.exp(-0.25 * r2b / newgw2) / math.sqrt(math.pi * newgw2) - np.exp(-0.25 * r2b / gw2) / math.sqrt(math.pi * gw2)
    invr2term = erf(0.5 * rb / newgw) - erf(0.5 * rb / gw)
    out[big] = 1 / (4 * math.pi) * (invrterm / rb - invr2term / r2b)
    small = (~big) & (r2 > 0)
    pi32 = math.pi ** 1.5
    rterm = 1 / (24 * pi32) * (1.0 / (gw2 * gw) - 1 / (newgw2 * newgw))
    r3term = 1 / (160 * pi32) * (1.0 / (newgw2 * newgw2 * newgw) - 1.0 / (gw2 * gw2 * gw))
    out[small] = r[small] * rterm + r2[small] * r[small] * r3term
    return out / epsilon


def next_fft_wise_size(n):
    """nextFFTWiseSize3D utils/Grid.cuh:142-213, one dimension."""
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


class PoissonOracle:
    """Poisson::Poisson :74-156 (parameter resolution in double) + farField :337-366 + the near-field Transversers :212-335
    with the EXACT Green's functions (the product interpolates the reference's tables: agreement to the table's accuracy)."""

    def __init__(self, L, epsilon, tolerance, gw, split=-1.0, upsampling=-1.0):
        self.L = np.broadcast_to(np.asarray(L, np.float64), (3,)).copy()
        self.epsilon, self.tolerance, self.gw, self.split = epsilon, tolerance, gw, split
        self.width = gw if split <= 0 else math.sqrt(gw * gw + 1.0 / (4.0 * split * split))
        h = 1.0 / upsampling if upsampling > 0 else (1.3 - min(-math.log10(tolerance) / 10.0, 0.9)) * self.width
        h = min(h, self.L[0] / 32.0)
        self.cells = np.array([next_fft_wise_size(int(self.L[d] / h)) for d in range(3)])
        self.h = self.L[0] / self.cells[0]
        w = self.width
        self.prefactor = (2 * math.pi * w * w) ** -0.5         # cbrt(pow(2 pi w^2, -1.5)), SpectralEwaldPoisson.cuh:69
        self.tau = -1.0 / (2.0 * w * w)
        rmax = math.sqrt(math.log(tolerance * math.sqrt(2 * math.pi * w * w)) / self.tau)
        self.support = max(3, int(2 * rmax / self.h + 0.5))
        assert self.support <= self.cells[0] // 2 - 1, "Kernel support is too large"
        self.support = min(self.support, self.cells[0] // 2 - 2)
        self.nearCut = 0.0
        if split > 0:
            E, r = 1.0, self.width
            root = math.sqrt(4 * gw * gw + 1 / (split * split))
            while abs(E) > tolerance:       # the search runs on the far branch of greensFunction (r^2 > gw^4)
                r += 0.001 * gw
                E = (1.0 / (4.0 * math.pi * epsilon * r) * (math.erf(r / (2 * gw)) - math.erf(r / root))) if r * r > gw ** 4 \
                    else float(greens_function(np.array([r * r]), gw, split, epsilon)[0])
            self.nearCut = r

    # IBM_ns::detail::computeSupportShift / fillSharedWeights misc/IBM.cu:11-66 (cell-centred distances)
    def _stencil(self, p):
        n, P = self.cells, self.support // 2
        idx, wts = [], []
        for d in range(3):
            hd = self.L[d] / n[d]
            x = p[d] - self.L[d] * math.floor(p[d] / self.L[d] + 0.5)          # fold into [-L/2, L/2)
            c = int((x + 0.5 * self.L[d]) / hd)
            c = 0 if c == n[d] else c
            if self.support % 2 == 0:
                # even supports start half a cell towards the particle (misc/IBM.cu:17-27)
                first = c - P + (1 if (x + 0.5 * self.L[d]) - c * hd >= 0.5 * hd else 0)
            else:
                first = c - P
            cells = first + np.arange(self.support)
            dist = x + 0.5 * self.L[d] - hd * (cells + 0.5)
            idx.append(np.mod(cells, n[d]))
            wts.append(self.prefactor * np.exp(self.tau * dist * dist))
        return idx, wts

    def far(self, pos, charge):
        """(E, phi) of the far field at the particles."""
        n = self.cells
        rho = np.zeros((n[2], n[1], n[0]))
        st = [self._stencil(p) for p in pos]
        for (idx, w), q in zip(st, charge):
            rho[np.ix_(idx[2], idx[1], idx[0])] += q * w[2][:, None, None] * w[1][None, :, None] * w[0][None, None, :]
        rk = np.fft.fftn(rho)
        k = [2 * math.pi * np.fft.fftfreq(n[d], d=1.0 / n[d]) / self.L[d] for d in range(3)]
        kz, ky, kx = np.meshgrid(k[2], k[1], k[0], indexing="ij")
        k2 = kx * kx + ky * ky + kz * kz
        k2[0, 0, 0] = 1.0
        B = 1.0 / (k2 * self.epsilon)
        B[0, 0, 0] = 0.0
        # isNyquist :428-444 (the reference zeroes the unpaired modes of even grids)
        def ny(i, m):
            return (m % 2 == 0) & (i == m - i)
        iz, iy, ix = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
        xn, yn, zn = ny(ix, n[0]), ny(iy, n[1]), ny(iz, n[2])
        x0, y0, z0 = ix == 0, iy == 0, iz == 0
        nyq = (xn & y0 & z0) | (xn & yn & z0) | (x0 & yn & z0) | (xn & y0 & zn) | (x0 & y0 & zn) | (x0 & yn & zn) | (xn & yn & zn)
        B[nyq] = 0.0
        phi = np.fft.ifftn(rk * B).real
        E = [np.fft.ifftn(-1j * kk * rk * B).real for kk in (kx, ky, kz)]
        dV = np.prod(self.L / n)
        out = np.zeros((len(pos), 4))
        for i, (idx, w) in enumerate(st):
            W = w[2][:, None, None] * w[1][None, :, None] * w[0][None, None, :] * dV
            sel = np.ix_(idx[2], idx[1], idx[0])
            out[i] = [np.sum(E[0][sel] * W), np.sum(E[1][sel] * W), np.sum(E[2][sel] * W), np.sum(phi[sel] * W)]
        return out

    def near(self, pos, charge):
        """(E, phi) of the near field at the particles (all pairs, minimum image, exact Green's functions)."""
        out = np.zeros((len(pos), 4))
        if self.split <= 0:
            return out
        d = pos[None, :, :3] - pos[:, None, :3]                                  # rij = pj - pi
        d -= self.L * np.floor(d / self.L + 0.5)
        r2 = (d * d).sum(-1)
        inside = r2 < self.nearCut ** 2
        G = np.where(inside, greens_function(r2, self.gw, self.split, self.epsilon), 0.0)
        r = np.sqrt(r2)
        Gf = np.where(inside & (r2 > 0), greens_function_field(r, self.gw, self.split, self.epsilon), 0.0)
        out[:, 3] = (G * charge[None, :]).sum(1)
        with np.errstate(invalid="ignore", divide="ignore"):
            unit = np.where(r2[..., None] > 0, d / r[..., None], 0.0)
        out[:, :3] = (-(charge[None, :] * Gf)[..., None] * unit).sum(1)
        return out

    def field_potential(self, pos, charge):
        return self.far(pos, charge) + self.near(pos, charge)
