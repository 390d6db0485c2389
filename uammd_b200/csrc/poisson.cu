// Spectral Ewald Poisson solver (triply periodic electrostatics of Gaussian charges), sm_100a.
//
// Replaces Poisson::sum / Poisson::computeFieldPotentialAtParticles (Interactor/SpectralEwaldPoisson.cuh:110-135,
// SpectralEwaldPoisson.cu:57-580):
//   far field : spread the charges with a Gaussian of width sqrt(gw^2 + 1/(4 split^2)) -> FFT -> (E, phi)(k) =
//               (-i k, 1) rho(k) / (eps k^2 N) -> inverse FFT -> interpolate: F_i += q_i E(x_i), U_i += q_i phi(x_i)
//   near field: pair sum within the cut-off of the tabulated real-space corrections G(r^2) and G'(r) (split > 0)
// built from the pieces paths 2 and 1 already have: the IBM stencil arithmetic (ibm.cuh), the hand-written 3-D FFT with the
// component count as a template parameter (one scalar transform forward, one four-component transform back: the
// reference's cufftPlan3d + batch-4 cufftPlanMany, SpectralEwaldPoisson.cu:158-210) and the reference-layout cell list.
#include "fft3d.cuh"
#include "ibm_state.cuh"
#include "pair_common.cuh"
#include <cmath>
#include <limits>
#include <vector>

namespace ub200 {

int nextFFTWiseSizeOf(int n); // pse.cu

// Poisson_ns::greensFunction (SpectralEwaldPoisson.cu:16-39): real-space correction of the potential, argument r^2
static double poissonG(double r2, double gw, double split, double epsilon) {
  double G = 0;
  if (r2 > gw * gw * gw * gw) {
    const double r = sqrt(r2);
    G = (1.0 / (4.0 * M_PI * epsilon * r) * (erf(r / (2 * gw)) - erf(r / sqrt(4 * gw * gw + 1 / (split * split)))));
  } else {
    const double pi32 = pow(M_PI, 1.5);
    const double gw2 = gw * gw;
    const double invsp2 = 1.0 / (split * split);
    const double selfterm = 1.0 / (4 * pi32 * gw) - 1.0 / (2 * pi32 * sqrt(4 * gw2 + invsp2));
    const double r2term = 1.0 / (6.0 * pi32 * pow(4.0 * gw2 + invsp2, 1.5)) - 1.0 / (48.0 * pi32 * gw2 * gw);
    const double r4term = 1.0 / (640.0 * pi32 * gw2 * gw2 * gw) - 1.0 / (20.0 * pi32 * pow(4 * gw2 + invsp2, 2.5));
    G = 1.0 / epsilon * (selfterm + r2 * r2term + r2 * r2 * r4term);
  }
  return G;
}
// Poisson_ns::greensFunctionField (SpectralEwaldPoisson.cu:41-63): its radial derivative, argument r
static double poissonGField(double r, double gw, double split, double epsilon) {
  const double r2 = r * r;
  const double gw2 = gw * gw;
  const double newgw = sqrt(gw2 + 1 / (4.0 * split * split));
  const double newgw2 = newgw * newgw;
  double fmod = 0;
  if (r2 > gw * gw * gw * gw) {
    const double invrterm = exp(-0.25 * r2 / newgw2) / sqrt(M_PI * newgw2) - exp(-0.25 * r2 / gw2) / sqrt(M_PI * gw2);
    const double invr2term = erf(0.5 * r / newgw) - erf(0.5 * r / gw);
    fmod += 1 / (4 * M_PI) * (invrterm / r - invr2term / r2);
  } else if (r2 > 0) {
    const double pi32 = pow(M_PI, 1.5);
    const double rterm = 1 / (24 * pi32) * (1.0 / (gw2 * gw) - 1 / (newgw2 * newgw));
    const double r3term = 1 / (160 * pi32) * (1.0 / (newgw2 * newgw2 * newgw) - 1.0 / (gw2 * gw2 * gw));
    fmod += r * rterm + r2 * r * r3term;
  }
  return fmod / epsilon;
}

// TabulatedFunction<real, LinearInterpolation>::operator() (misc/TabulatedFunction.cuh:62-74,148-158)
template <class T> struct ScalarTable {
  const T *table;
  int Ntable;
  T rmin, rmax, interval, dr;
};
template <class T> __device__ __forceinline__ T tableValue(const ScalarTable<T> &tb, T rs) {
  const T r = (rs - tb.rmin) * tb.interval;
  if (rs >= tb.rmax) return T(0);
  if (r <= T(0)) return __ldg(tb.table);
  const int i = (int)(r * tb.Ntable);
  const T r0 = i * tb.dr;
  const T v0 = __ldg(tb.table + i), v1 = __ldg(tb.table + i + 1);
  const T t = (r - r0) * (T)tb.Ntable;
  return fma(t, v1, fma(-t, v0, v0));
}

// IBM<Kernel>::spread of the scalar charges (particles2GridD, misc/IBM.cu:83-147): one warp per particle, atomics into the
// zero-filled charge grid [nz][ny][nxPad]
template <class T4, class T>
__global__ void __launch_bounds__(128)
poissonSpreadCharges(const T4 *__restrict__ pos, const T *__restrict__ charge, int N, GridT<T> g, IbmKernel<T> k, int nxPad,
                     T *__restrict__ gridQ, const int *__restrict__ order) {
  // order: optional cell-sorted permutation (the near field's cell list): neighbouring warps then touch neighbouring nodes
  __shared__ T wsh[4][3 * kMaxSupport];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * 4 + warp;
  if (slot >= N) return;
  const int i = order ? order[slot] : slot;
  const T4 p = pos[i];
  const T pr[3] = {p.x, p.y, p.z};
  int o[3];
  T *w = wsh[warp];
  const int S = k.support;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int c = cellOfT(g, d, pr[d]);
    o[d] = supportOrigin(g, k, d, pr[d], c);
    if (lane < S) w[d * kMaxSupport + lane] = supportWeight(g, k, d, pr[d], o[d], lane);
  }
  __syncwarp();
  const T q = charge[i];
  const int total = S * S * S;
  for (int t = lane; t < total; t += 32) {
    const int ii = t % S, jj = (t / S) % S, kk = t / (S * S);
    const int cx = wrapCell(g, 0, o[0] + ii), cy = wrapCell(g, 1, o[1] + jj), cz = wrapCell(g, 2, o[2] + kk);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    const T v = q * w[ii] * w[kMaxSupport + jj] * w[2 * kMaxSupport + kk];
    if (v != T(0)) atomicAdd(gridQ + ((size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g.n[1] * cz)), v);
  }
}

// IBM<Kernel>::gather of the real4 grid (Ex, Ey, Ez, phi) (grid2ParticlesDTPP, misc/IBM.cu:168-235; quadrature weight = cell
// volume) into UnZip2Real4 (SpectralEwaldPoisson.cu:535-559): force_i += q_i (E, 0), energy_i += q_i phi - or, for
// computeFieldPotentialAtParticles, fieldPotential_i += (E, phi)
template <class T4, class T>
__global__ void __launch_bounds__(128)
poissonGatherField(const T4 *__restrict__ pos, const T *__restrict__ charge, int N, GridT<T> g, IbmKernel<T> k, int nxPad,
                   const T4 *__restrict__ gridF, T4 *__restrict__ force, T *__restrict__ energy, T4 *__restrict__ fieldPotential,
                   const int *__restrict__ order) {
  __shared__ T wsh[4][3 * kMaxSupport];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * 4 + warp;
  if (slot >= N) return;
  const int i = order ? order[slot] : slot;
  const T4 p = pos[i];
  const T pr[3] = {p.x, p.y, p.z};
  int o[3];
  T *w = wsh[warp];
  const int S = k.support;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int c = cellOfT(g, d, pr[d]);
    o[d] = supportOrigin(g, k, d, pr[d], c);
    if (lane < S) w[d * kMaxSupport + lane] = supportWeight(g, k, d, pr[d], o[d], lane);
  }
  __syncwarp();
  T ax = T(0), ay = T(0), az = T(0), aw = T(0);
  const int total = S * S * S;
  for (int t = lane; t < total; t += 32) {
    const int ii = t % S, jj = (t / S) % S, kk = t / (S * S);
    const int cx = wrapCell(g, 0, o[0] + ii), cy = wrapCell(g, 1, o[1] + jj), cz = wrapCell(g, 2, o[2] + kk);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    const T4 v = gridF[(size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g.n[1] * cz)];
    const T wgt = w[ii] * w[kMaxSupport + jj] * w[2 * kMaxSupport + kk];
    ax += g.cellVolume * (v.x * wgt);
    ay += g.cellVolume * (v.y * wgt);
    az += g.cellVolume * (v.z * wgt);
    aw += g.cellVolume * (v.w * wgt);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, off);
    ay += __shfl_xor_sync(0xffffffffu, ay, off);
    az += __shfl_xor_sync(0xffffffffu, az, off);
    aw += __shfl_xor_sync(0xffffffffu, aw, off);
  }
  if (lane != 0) return;
  if (fieldPotential) {
    T4 v = fieldPotential[i];
    v.x += ax; v.y += ay; v.z += az; v.w += aw;
    fieldPotential[i] = v;
    return;
  }
  const T q = charge[i];
  if (force) {
    T4 f = force[i];
    f.x += q * ax; f.y += q * ay; f.z += q * az;
    force[i] = f;
  }
  if (energy) energy[i] += q * aw;
}

// Poisson_ns::chargeFourier2FieldAndPotential (SpectralEwaldPoisson.cu:446-478) with isNyquist (:428-444) and
// cellToWaveNumber (:412-426): rho(k) [nz][ny][nkx] -> (Ex, Ey, Ez, phi)(k) [nz][ny][nkx][4]
template <class T> struct PoissonSpectral {
  int nx, ny, nz, nkx;
  T kfx, kfy, kfz, epsilon, ncells;
};
template <class T>
__global__ void __launch_bounds__(256)
poissonFieldAndPotential(const typename Vec2<T>::type *__restrict__ Q, typename Vec2<T>::type *__restrict__ F, PoissonSpectral<T> s) {
  using C = typename Vec2<T>::type;
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t nk = (size_t)s.nkx * s.ny * s.nz;
  if (id >= nk) return;
  const int cx = (int)(id % s.nkx), cy = (int)((id / s.nkx) % s.ny), cz = (int)(id / ((size_t)s.nkx * s.ny));
  const C zero = mk2<T>(T(0), T(0));
  C ex = zero, ey = zero, ez = zero, ph = zero;
  const bool xn = (cx == s.nx - cx) && (s.nx % 2 == 0), yn = (cy == s.ny - cy) && (s.ny % 2 == 0),
             zn = (cz == s.nz - cz) && (s.nz % 2 == 0);
  const bool nyquist = (xn && cy == 0 && cz == 0) || (xn && yn && cz == 0) || (cx == 0 && yn && cz == 0) ||
                       (xn && cy == 0 && zn) || (cx == 0 && cy == 0 && zn) || (cx == 0 && yn && zn) || (xn && yn && zn);
  if (!(cx == 0 && cy == 0 && cz == 0) && !nyquist) {
    T kx = cx * s.kfx, ky = cy * s.kfy, kz = cz * s.kfz;
    if (cx >= s.nx / 2 + 1) kx -= T(s.nx) * s.kfx;
    if (cy >= s.ny / 2 + 1) ky -= T(s.ny) * s.kfy;
    if (cz >= s.nz / 2 + 1) kz -= T(s.nz) * s.kfz;
    const T k2 = kx * kx + ky * ky + kz * kz;
    const C fk = Q[id];
    const T B = T(1.0) / (k2 * s.epsilon * s.ncells);
    ex = mk2<T>(kx * fk.y * B, -kx * fk.x * B);
    ey = mk2<T>(ky * fk.y * B, -ky * fk.x * B);
    ez = mk2<T>(kz * fk.y * B, -kz * fk.x * B);
    ph = mk2<T>(fk.x * B, fk.y * B);
  }
  F[4 * id] = ex; F[4 * id + 1] = ey; F[4 * id + 2] = ez; F[4 * id + 3] = ph;
}

// positions + charge in the sorted order of the cell list, in the precision of the solver
template <class T4, class T>
__global__ void __launch_bounds__(256)
poissonGatherSorted(const int *__restrict__ groupIndex, const T4 *__restrict__ pos, const T *__restrict__ charge, int N,
                    T4 *__restrict__ sortedPQ) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = groupIndex[k];
  T4 p = pos[i];
  p.w = charge[i];
  sortedPQ[k] = p;
}
__global__ void __launch_bounds__(256) poissonToFloat4(const double4 *__restrict__ in, float4 *__restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = make_float4((float)in[i].x, (float)in[i].y, (float)in[i].z, 0.f);
}

// The three near-field Transversers (SpectralEwaldPoisson.cu:212-335) over the cell list (transverseList,
// NeighbourList/common.cuh:10-34): one warp per home cell, the lanes share the candidates of a home particle.
//   MODE 0: NearFieldForceTransverser            force_i  += sum_j -q_i q_j G'(r) rij / r
//   MODE 1: NearFieldEnergyTransverser           energy_i += sum_j  q_i q_j G(r^2)      (j = i included: the self term)
//   MODE 2: NearFieldFieldPotentialTransverser   (E, phi)_i += sum_j (-q_j G'(r) rij / r, q_j G(r^2))
template <class T> struct PoissonBox { T Lx, Ly, Lz, mx, my, mz; };
template <class T4, class T, int MODE>
__global__ void __launch_bounds__(kPairThreads)
poissonNearTraversal(const T4 *__restrict__ sortedPQ, const int *__restrict__ groupIndex, const uint32_t *__restrict__ binStart,
                     GridF g, int ncells, ScalarTable<T> tabG, ScalarTable<T> tabF, PoissonBox<T> box, T4 *__restrict__ out4,
                     T *__restrict__ out1) {
  constexpr int kStage = 512; // candidate slots of the 27 cells listed flat per warp (denser neighbourhoods walk cell by cell)
  __shared__ int candAll[kPairWarps][kStage];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *cand = candAll[warp];
  const int warpsTotal = gridDim.x * kPairWarps;
  for (int cell = blockIdx.x * kPairWarps + warp; cell < ncells; cell += warpsTotal) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue;
    const bool flat = nc.total <= kStage;
    __syncwarp();
    if (flat)
      for (int t = 0; t < nc.count; t++) cand[nc.off + t] = nc.start + t; // lane = neighbour cell: a few entries each
    __syncwarp();
    for (int h = 0; h < hCount; h++) {
      const T4 pi = sortedPQ[hStart + h];
      T ax = T(0), ay = T(0), az = T(0), aw = T(0);
      auto pair = [&](const T4 pj) {
        // Box::apply_pbc (utils/Box.cuh:51-58)
        T dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        dx += floor(dx * box.mx + T(0.5)) * box.Lx;
        dy += floor(dy * box.my + T(0.5)) * box.Ly;
        dz += floor(dz * box.mz + T(0.5)) * box.Lz;
        const T r2 = dx * dx + dy * dy + dz * dz;
        if (MODE == 1) {
          aw += pi.w * pj.w * tableValue(tabG, r2);
        } else {
          if (MODE == 2) aw += pj.w * tableValue(tabG, r2);
          if (r2 > T(0)) {
            const T r = sqrt(r2);
            const T fmod = (MODE == 0 ? -pi.w * pj.w : -pj.w) * tableValue(tabF, r);
            ax += fmod * dx / r; ay += fmod * dy / r; az += fmod * dz / r;
          }
        }
      };
      if (flat) {
        for (int t = lane; t < nc.total; t += 32) pair(sortedPQ[cand[t]]);
      } else {
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          for (int t = lane; t < cnt; t += 32) pair(sortedPQ[st + t]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
        aw += __shfl_xor_sync(0xffffffffu, aw, o);
      }
      if (lane == 0) {
        const int i = groupIndex[hStart + h];
        if (MODE == 1) {
          out1[i] += aw;
        } else {
          T4 v = out4[i];
          v.x += ax; v.y += ay; v.z += az;
          if (MODE == 2) v.w += aw;
          out4[i] = v;
        }
      }
    }
  }
}

template <class T> struct PoissonState {
  using T4 = typename Real4<T>::type;
  using C = typename Vec2<T>::type;
  Fft3dPlan<T, 1> planQ; // charges
  Fft3dPlan<T, 4> planF; // (Ex, Ey, Ez, phi)
  GridT<T> grid;
  IbmKernel<T> kern;
  DevBuf Q, F, tableG, tableF, posF, sortedPQ;
  ub200_celllist *cl = nullptr;
  double L[3] = {0, 0, 0}, epsilon = 1, split = -1, gw = 1, tolerance = 1e-5, nearCut = 0, farWidth = 0, h = 0;
  int support = 0, nTable = 0, cells[3] = {0, 0, 0};

  // Poisson::Poisson (SpectralEwaldPoisson.cu:74-156)
  int init(const ub200_poisson_params &par) {
    for (int d = 0; d < 3; d++) L[d] = par.L[d];
    epsilon = par.epsilon; split = par.split; gw = par.gw; tolerance = par.tolerance;
    if (!(L[0] > 0) || !(L[1] > 0) || !(L[2] > 0) || !(gw > 0) || !(epsilon > 0) || !(tolerance > 0)) return UB200_ERR_INVALID_ARGUMENT;
    farWidth = gw;
    if (split > 0) farWidth = sqrt(gw * gw + 1.0 / (4.0 * split * split));
    if (par.upsampling > 0) h = 1.0 / par.upsampling;
    else h = (1.3 - std::min((-log10(tolerance)) / 10.0, 0.9)) * farWidth;
    h = std::min(h, (double)(T)L[0] / 32.0);
    for (int d = 0; d < 3; d++) {
      cells[d] = nextFFTWiseSizeOf((int)((T)L[d] / (T)h)); // make_int3(box.boxSize / h): real arithmetic, truncated
      if (cells[d] < 1) return UB200_ERR_GRID_TOO_LARGE;
    }
    h = (double)((T)L[0] / (T)cells[0]); // grid.cellSize.x
    // Poisson_ns::Gaussian(tolerance, width, h) (SpectralEwaldPoisson.cuh:66-82), real arithmetic
    const T width = (T)farWidth, tol = (T)tolerance;
    const T prefactor = (T)cbrt(pow(2 * M_PI * width * width, -1.5));
    const T tau = (T)(-1.0 / (2.0 * width * width));
    const T rmax = (T)sqrt(log(tol * sqrt(2 * M_PI * width * width)) / tau);
    support = std::max(3, int(2 * rmax / (T)h + 0.5));
    if (support > cells[0] / 2 - 1) return UB200_ERR_UNSUPPORTED; // "Kernel support is too large"
    support = std::min(support, cells[0] / 2 - 2);
    if (support > kMaxSupport) return UB200_ERR_UNSUPPORTED;
    if (split > 0) {
      long double E = 1, r = farWidth;
      while (fabsl(E) > tolerance) {
        r += 0.001l * (T)gw;
        E = (T)poissonG((double)(T)(r * r), (T)gw, (T)split, (T)epsilon); // greensFunction takes and returns real
      }
      nearCut = (double)(T)r;
      if (nearCut > (T)L[0] / 2.0) return UB200_ERR_INVALID_ARGUMENT; // "Near field cut off is too large"
    }
    int rc;
    if ((rc = planQ.init(cells[0], cells[1], cells[2])) || (rc = planF.init(cells[0], cells[1], cells[2]))) return rc;
    const int periodic[3] = {1, 1, 1};
    grid = makeGridT<T>(L, periodic, cells);
    kern.kind = kKernelGaussian; // Poisson_ns::Gaussian::phi has no radial cut: every point of the support counts
    kern.support = support;
    kern.invh = (T)(1.0 / h);
    kern.prefactor = prefactor; kern.tau = tau; kern.rmax = std::numeric_limits<T>::max();
    if ((rc = Q.reserve(planQ.gridBytes())) || (rc = F.reserve(planF.gridBytes()))) return rc;
    if (split > 0) {
      nTable = std::max(4096, std::min(1 << 16, int((T)nearCut / ((T)gw * (T)tolerance * 1e3))));
      std::vector<T> tg(nTable), tf(nTable);
      const int Nt = nTable - 1; // TabulatedFunction keeps N - 1 intervals (misc/TabulatedFunction.cuh:113-123)
      const double rmaxF = (double)(T)nearCut, rmaxG = (double)(T)((T)nearCut * (T)nearCut);
      for (int i = 0; i <= Nt; i++) {
        const double xf = (i / (double)Nt) * rmaxF, xg = (i / (double)Nt) * rmaxG;
        tf[i] = (T)poissonGField((double)(T)xf, (T)gw, (T)split, (T)epsilon);
        tg[i] = (T)poissonG((double)(T)xg, (T)gw, (T)split, (T)epsilon);
      }
      if ((rc = tableG.reserve(sizeof(T) * nTable)) || (rc = tableF.reserve(sizeof(T) * nTable))) return rc;
      UB200_CUDA(cudaMemcpy(tableG.p, tg.data(), sizeof(T) * nTable, cudaMemcpyHostToDevice));
      UB200_CUDA(cudaMemcpy(tableF.p, tf.data(), sizeof(T) * nTable, cudaMemcpyHostToDevice));
      if ((rc = ub200_celllist_create(&cl))) return rc;
    }
    return UB200_OK;
  }
  void release() {
    planQ.release(); planF.release();
    DevBuf *b[] = {&Q, &F, &tableG, &tableF, &posF, &sortedPQ};
    for (auto *x : b) x->release();
    if (cl) ub200_celllist_destroy(cl);
    cl = nullptr;
  }
  ScalarTable<T> view(const DevBuf &buf, T rmax) const {
    ScalarTable<T> tb;
    tb.table = buf.as<T>();
    tb.Ntable = nTable - 1;
    tb.rmin = T(0); tb.rmax = rmax;
    tb.interval = (T)(1.0 / (rmax - T(0)));
    tb.dr = (T)(1.0 / (T)(nTable - 1));
    return tb;
  }

  // Poisson::farField (SpectralEwaldPoisson.cu:337-366): zero fill, spread, forward FFT, convolution, inverse FFT,
  // interpolation into (force, energy) or into fieldPotential
  int farField(const void *pos, const void *charge, int N, T4 *force, T *energy, T4 *fieldPotential, const int *order,
               cudaStream_t st) {
    int rc;
    T *q = Q.as<T>(), *f = F.as<T>();
    UB200_CUDA(cudaMemsetAsync(q, 0, planQ.gridBytes(), st));
    const int nb = (N + 3) / 4;
    poissonSpreadCharges<T4, T><<<nb, 128, 0, st>>>((const T4 *)pos, (const T *)charge, N, grid, kern, planQ.nxPad, q, order);
    UB200_LAUNCHED();
    if ((rc = launchPassX<T, true>(planQ, q, st)) || (rc = launchPassY<T, -1>(planQ, q, st)) || (rc = launchPassZ<T, -1>(planQ, q, st)))
      return rc;
    PoissonSpectral<T> s;
    s.nx = planQ.nx; s.ny = planQ.ny; s.nz = planQ.nz; s.nkx = planQ.nkx;
    s.kfx = (T)(T(2.0) * T(M_PI) / (T)L[0]);
    s.kfy = (T)(T(2.0) * T(M_PI) / (T)L[1]);
    s.kfz = (T)(T(2.0) * T(M_PI) / (T)L[2]);
    s.epsilon = (T)epsilon;
    s.ncells = (T)((double)planQ.nx * planQ.ny * planQ.nz);
    const size_t nk = (size_t)planQ.nkx * planQ.ny * planQ.nz;
    poissonFieldAndPotential<T><<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(reinterpret_cast<const C *>(q), reinterpret_cast<C *>(f), s);
    UB200_LAUNCHED();
    if ((rc = launchPassZ<T, +1>(planF, f, st)) || (rc = launchPassY<T, +1>(planF, f, st)) || (rc = launchPassX<T, false>(planF, f, st)))
      return rc;
    poissonGatherField<T4, T><<<nb, 128, 0, st>>>((const T4 *)pos, (const T *)charge, N, grid, kern, planF.nxPad,
                                                    reinterpret_cast<const T4 *>(f), force, energy, fieldPotential, order);
    UB200_LAUNCHED();
    return UB200_OK;
  }

  // cell list at the near-field cut-off and the sorted (position, charge) records
  int nearPrepare(const void *pos, const void *charge, int N, cudaStream_t st) {
    int rc;
    const float Lf[3] = {(float)L[0], (float)L[1], (float)L[2]};
    const float Lmax = std::max({Lf[0], Lf[1], Lf[2]});
    // the list is built from fp32 coordinates: a rounding margin keeps every pair the exact test can accept
    const float rcList = (float)nearCut * (1.0f + 1e-5f) + 16.0f * Lmax * 1.2e-7f;
    const int periodic[3] = {1, 1, 1};
    const void *posf = pos;
    if (sizeof(T) == 8) {
      if ((rc = posF.reserve(sizeof(float4) * (size_t)N))) return rc;
      poissonToFloat4<<<(N + 255) / 256, 256, 0, st>>>((const double4 *)pos, posF.as<float4>(), N);
      UB200_LAUNCHED();
      posf = posF.p;
    }
    int cd[3];
    if ((rc = ub200_neighbour_celldim_f32(Lf, rcList, cd))) return rc;
    if ((rc = ub200_celllist_build_f32(cl, posf, nullptr, N, Lf, periodic, cd, (void *)st))) return rc;
    if ((rc = sortedPQ.reserve(sizeof(T4) * (size_t)N))) return rc;
    poissonGatherSorted<T4, T><<<(N + 255) / 256, 256, 0, st>>>(cl->groupIndex.as<int>(), (const T4 *)pos, (const T *)charge, N,
                                                                sortedPQ.as<T4>());
    UB200_LAUNCHED();
    return UB200_OK;
  }
  template <int MODE> int nearLaunch(T4 *out4, T *out1, cudaStream_t st) {
    PoissonBox<T> box;
    box.Lx = (T)L[0]; box.Ly = (T)L[1]; box.Lz = (T)L[2];
    box.mx = T(-1.0) / box.Lx; box.my = T(-1.0) / box.Ly; box.mz = T(-1.0) / box.Lz;
    const T rc = (T)nearCut;
    const ScalarTable<T> tg = view(tableG, rc * rc), tf = view(tableF, rc);
    const int needed = (cl->ncells + kPairWarps - 1) / kPairWarps;
    const int grid = needed < kNumSMs * 8 ? needed : kNumSMs * 8;
    poissonNearTraversal<T4, T, MODE><<<grid, kPairThreads, 0, st>>>(sortedPQ.as<T4>(), cl->groupIndex.as<int>(),
                                                                    cl->binStart.as<uint32_t>(), cl->grid, cl->ncells, tg, tf, box,
                                                                    out4, out1);
    UB200_LAUNCHED();
    return UB200_OK;
  }

  // Poisson::sum (SpectralEwaldPoisson.cuh:110-122): far field always (it interpolates forces AND energies into whichever
  // array is given), near field per requested computable
  int sum(const void *pos, const void *charge, int N, void *force4, void *energy, bool nearForce, bool nearEnergy,
          cudaStream_t st) {
    int rc;
    nearForce = nearForce && force4;
    nearEnergy = nearEnergy && energy;
    const bool near = split > 0 && (nearForce || nearEnergy);
    // the near field's cell list first: its order also serves the spreading and the interpolation (grid locality)
    if (near && (rc = nearPrepare(pos, charge, N, st))) return rc;
    if ((rc = farField(pos, charge, N, (T4 *)force4, (T *)energy, nullptr, near ? cl->groupIndex.as<int>() : nullptr, st))) return rc;
    if (near) {
      if (nearForce && (rc = nearLaunch<0>((T4 *)force4, nullptr, st))) return rc;
      if (nearEnergy && (rc = nearLaunch<1>(nullptr, (T *)energy, st))) return rc;
    }
    return UB200_OK;
  }
  // Poisson::computeFieldPotentialAtParticles (SpectralEwaldPoisson.cuh:124-135): (Ex, Ey, Ez, phi) ADDED to out4
  int fieldPotential(const void *pos, const void *charge, int N, void *out4, cudaStream_t st) {
    int rc;
    if (split > 0 && (rc = nearPrepare(pos, charge, N, st))) return rc;
    if ((rc = farField(pos, charge, N, nullptr, nullptr, (T4 *)out4, split > 0 ? cl->groupIndex.as<int>() : nullptr, st))) return rc;
    if (split > 0) {
      if ((rc = nearLaunch<2>((T4 *)out4, nullptr, st))) return rc;
    }
    return UB200_OK;
  }
};

} // namespace ub200

using namespace ub200;

struct ub200_poisson {
  int precision;
  PoissonState<float> f;
  PoissonState<double> d;
};

extern "C" {

int ub200_poisson_create(ub200_poisson **out, int precisionBytes, const ub200_poisson_params *par) {
  if (!out || !par || (precisionBytes != 4 && precisionBytes != 8)) return UB200_ERR_INVALID_ARGUMENT;
  ub200_poisson *h = new (std::nothrow) ub200_poisson();
  if (!h) return UB200_ERR_ALLOC;
  h->precision = precisionBytes;
  const int rc = precisionBytes == 4 ? h->f.init(*par) : h->d.init(*par);
  if (rc) { h->f.release(); h->d.release(); delete h; return rc; }
  *out = h;
  return UB200_OK;
}
int ub200_poisson_destroy(ub200_poisson *h) {
  if (!h) return UB200_OK;
  h->f.release(); h->d.release();
  delete h;
  return UB200_OK;
}
int ub200_poisson_info(ub200_poisson *h, ub200_poisson_info_t *info) {
  if (!h || !info) return UB200_ERR_INVALID_ARGUMENT;
#define UB200_PINFO(s)                                                                                                   \
  for (int d = 0; d < 3; d++) info->cells[d] = s.cells[d];                                                               \
  info->support = s.support; info->nTable = s.nTable; info->h = s.h; info->farFieldGaussianWidth = s.farWidth;           \
  info->nearFieldCutOff = s.nearCut;
  if (h->precision == 4) { UB200_PINFO(h->f) } else { UB200_PINFO(h->d) }
#undef UB200_PINFO
  return UB200_OK;
}
int ub200_poisson_sum_ex(ub200_poisson *h, const void *d_pos, const void *d_charge, int N, void *d_force4, void *d_energy,
                         int nearFieldForce, int nearFieldEnergy, void *stream) {
  if (!h || !d_pos || !d_charge || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4
             ? h->f.sum(d_pos, d_charge, N, d_force4, d_energy, nearFieldForce != 0, nearFieldEnergy != 0, (cudaStream_t)stream)
             : h->d.sum(d_pos, d_charge, N, d_force4, d_energy, nearFieldForce != 0, nearFieldEnergy != 0, (cudaStream_t)stream);
}
int ub200_poisson_sum(ub200_poisson *h, const void *d_pos, const void *d_charge, int N, void *d_force4, void *d_energy,
                      void *stream) {
  return ub200_poisson_sum_ex(h, d_pos, d_charge, N, d_force4, d_energy, 1, 1, stream);
}
int ub200_poisson_field_potential(ub200_poisson *h, const void *d_pos, const void *d_charge, int N, void *d_fieldPotential4,
                                  void *stream) {
  if (!h || !d_pos || !d_charge || !d_fieldPotential4 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? h->f.fieldPotential(d_pos, d_charge, N, d_fieldPotential4, (cudaStream_t)stream)
                           : h->d.fieldPotential(d_pos, d_charge, N, d_fieldPotential4, (cudaStream_t)stream);
}
}
