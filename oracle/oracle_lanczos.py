"""TEST INFRASTRUCTURE (CPU oracle, numpy) - never imported by the product path.

Restatement of the reference's Lanczos square root, lanczos::Solver::run with detail::KrylovSubspace
(/root/reference/src/misc/LanczosAlgorithm/LanczosAlgorithm.cu): setFirstBasisVector :110-121, nextIteration :123-157,
computeSquareRoot :65-91 + computeCurrentResultEstimation :160-172, run :202-228, computeError :230-251,
registerRequiredStepsForConverge :253-262. `dot(v)` is the MatrixDot (here: any callable returning M v).

Pinned by properties rather than golden vectors (the reference's Lanczos needs a GPU, cuBLAS and LAPACKE): on a small SPD
matrix the result must equal the dense sqrtm(M) z to the requested tolerance (tests/test_oracle_lanczos.py); the GPU
implementation itself is pinned against the compiled reference in tests/test_pse_gpu.py.

`lanczos_sqrt_rows` is the same iteration as ub200_pse_dist_* runs it (uammd_b200/csrc/pse.cu, distNoise / lanczosSqrt with
dist = true): every vector is cut into contiguous row blocks, one per rank; a product needs the complete basis vector
(published to every rank), a scalar is the sum of the ranks' partial dot products taken IN RANK ORDER, so that every rank
holds the same bits and takes the same decisions. It must reproduce `lanczos_sqrt` up to the summation order of the dots."""
import numpy as np


def _sqrt_h_e1(hdiag, hsup, m):
    """H^1/2 e1 for the m x m tridiagonal H (computeSquareRoot :65-91: eigen-decomposition, sqrt of the eigenvalues)"""
    H = np.diag(np.asarray(hdiag[:m], float))
    for i in range(m - 1):
        H[i, i + 1] = H[i + 1, i] = hsup[i]
    lam, P = np.linalg.eigh(H)
    return P @ (np.sqrt(np.maximum(lam, 0.0)) * P[0, :])


class Solver:
    """lanczos::Solver: keeps check_convergence_steps between runs like the reference object does."""

    def __init__(self):
        self.check_convergence_steps = 3
        self.iterationHardLimit = 200
        self.lastRunRequiredSteps = 0

    def _register(self, steps_needed):
        self.lastRunRequiredSteps = steps_needed
        if steps_needed - 2 > self.check_convergence_steps:
            self.check_convergence_steps += 1
        else:
            self.check_convergence_steps = max(1, self.check_convergence_steps - 2)

    def run(self, dot, z, tolerance, blocks=None):
        """returns (Bz, iterations). blocks = None: the reference's iteration on whole vectors; blocks = [(lo, hi), ...]:
        the row-block iteration of the rank decomposition (vectors held as one slice per rank)."""
        z = np.asarray(z, float)
        n = z.shape[0]
        blocks = [(0, n)] if blocks is None else list(blocks)

        def vdot(a, b):
            s = 0.0
            for lo, hi in blocks:                  # partial sums of the ranks, added in rank order
                s += float(np.dot(a[lo:hi], b[lo:hi]))
            return s

        def rows_dot(v):
            # every rank evaluates its rows of M v from the complete, published v
            full = dot(v)
            w = np.empty(n)
            for lo, hi in blocks:
                w[lo:hi] = full[lo:hi]
            return w

        normz = np.sqrt(vdot(z, z))
        V = [z / normz]
        hdiag, hsup = [], []
        oldBz = np.zeros(n)
        Bz = np.zeros(n)
        check = min(self.check_convergence_steps, self.iterationHardLimit - 2)
        for i in range(self.iterationHardLimit):
            w = rows_dot(V[i])
            if i > 0:
                w = w - hsup[i - 1] * V[i - 1]
            hdiag.append(vdot(w, V[i]))
            w = w - hdiag[i] * V[i]
            hs = np.sqrt(vdot(w, w))
            if hs < 1e-3 * hdiag[i] / normz:
                hs = 0.0
            hsup.append(hs)
            if hs > 0.0:
                V.append(w / hs)
            else:
                e1 = np.zeros(n); e1[0] = 1.0
                V.append(e1)
            if i >= check:
                m = i + 1
                c = _sqrt_h_e1(hdiag, hsup, m)
                Bz = normz * (np.stack(V[:m], axis=1) @ c)
                if i > 0:
                    prev = np.sqrt(vdot(oldBz, oldBz))
                    d = oldBz - Bz
                    with np.errstate(divide="ignore", invalid="ignore"):   # the first check divides by ||0||, like the reference
                        err = abs(np.sqrt(vdot(d, d)) / prev)
                    if np.isnan(err):
                        raise RuntimeError("[Lanczos] Unknown error (found NaN in result guess)")
                    if err <= tolerance:
                        self._register(i)
                        return Bz, i
                oldBz = Bz.copy()
        raise RuntimeError("[Lanczos] Could not converge")


def lanczos_sqrt(dot, z, tolerance):
    return Solver().run(dot, z, tolerance)


def lanczos_sqrt_rows(dot, z, tolerance, world):
    n = np.asarray(z).shape[0]
    rows = n // 3
    blocks = [(3 * ((r * rows) // world), 3 * (((r + 1) * rows) // world)) for r in range(world)]
    return Solver().run(dot, z, tolerance, blocks=blocks)
