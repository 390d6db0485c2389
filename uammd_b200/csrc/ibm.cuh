// Immersed boundary spreading / interpolation for sm_100a.
//
// Replaces IBM<Kernel>::spread / gather (misc/IBM.cuh:117-184) and their kernels particles2GridD /
// grid2ParticlesDTPP (misc/IBM.cu:83-147,168-235): same support-cell rule (computeSupportShift, IBM.cu:11-31),
// same cell-centred distances and weights (fillSharedWeights, IBM.cu:33-66), same products
// value*phiX*phiY*phiZ (DefaultWeightCompute, IBM.cuh:88-97) and quadrature weight = cell volume (:80-86).
//
// Small supports (Peskin 3/4 point): spreading is turned into an atomic-free, write-once NODE gather:
//   binByCell -> stable order + cell-sorted copies of position/value -> one CTA per 16x8x4 brick of grid nodes
//   stages the particles of the brick's halo region in shared memory with their 1-D weights and every node is
//   summed and written exactly once (no memset, no atomics, deterministic). The reference launches one
//   32-thread block per particle issuing 3*support^3 scalar atomics into a zero-filled grid.
// Large supports (Gaussian, up to 32 points per dimension): one warp per particle, weights in shared memory,
//   atomics for spreading / shuffle reduction for interpolation.
#pragma once
#include "common.cuh"

namespace ub200 {

enum { kKernelPeskin3 = 0, kKernelPeskin4 = 1, kKernelGaussian = 2, kKernelBarnettMagland = 3, kKernelSixPoint = 4 };
constexpr int kMaxSupport = 32;

template <class T> struct IbmKernel {
  int kind, support;
  T invh;                // Peskin, six point
  T prefactor, tau, rmax; // Gaussian: prefactor exp(tau r^2), r < rmax. Barnett-Magland: 1/norm, beta, alpha
};

// GaussianFlexible::sixPoint::phi_impl (misc/IBM_kernels.cuh:168-216): the C3 six-point kernel of Bao, Kaye and Peskin
// (J. Comput. Phys. 316 (2016) 139) for r = |distance| / h in [0, 3); every branch shares one root of a quadratic in the
// fractional part R of r.
template <class T> __host__ __device__ __forceinline__ T ibmSixPoint(T r) {
  if (r >= T(3)) return T(0);
  const T K = T(0.714075092976608); // 59/60 - sqrt(29)/20
  const T R = r - ceil(r) + T(1.0), R2 = R * R, R3 = R2 * R;
  const T alpha = T(28.);
  const T beta = T(9.0 / 4.0) - T(1.5) * (K + R2) + (T(22. / 3) - T(7.0) * K) * R - T(7. / 3.) * R3;
  const T gamma = T(0.25) * (T(0.5) * (T(161.) / T(36) - T(59.) / T(6) * K + T(5) * K * K) * R2 +
                             T(1.) / T(3) * (T(-109.) / T(24) + T(5) * K) * R2 * R2 + T(5.) / T(18) * R3 * R3);
  const T discr = beta * beta - T(4.0) * alpha * gamma;
  const T pre = T(1.) / (T(2) * alpha) * (-beta + sqrt(discr)); // sign(3/2 - K) = +1
  if (r <= T(0)) {
    const T rp1 = r + T(1.0);
    return T(2.) * pre + T(0.25) + T(1. / 6) * (T(4) - T(3) * K) * rp1 - T(1. / 6) * rp1 * rp1 * rp1;
  }
  if (r <= T(1)) return T(2.0) * pre + T(5. / 8) - T(0.25) * (K + r * r);
  if (r <= T(2)) {
    const T rm1 = r + T(-1.0);
    return T(-3.0) * pre + T(0.25) - T(1. / 6.) * (T(4) - T(3) * K) * rm1 + T(1. / 6) * rm1 * rm1 * rm1;
  }
  const T rm2 = r + T(-2.0);
  return pre - T(1. / 16) + T(1. / 8) * (K + rm2 * rm2) - T(1. / 12) * (T(3) * K - T(1)) * rm2 - T(1. / 12) * rm2 * rm2 * rm2;
}

// window functions: misc/IBM_kernels.cuh:118-137 (3 point), :140-157 (4 point), :28-40 + FCM_kernels.cuh:54-56 (Gaussian),
// :83-113 (Barnett-Magland "exponential of a semicircle"), :163-237 (six point)
template <class T> __device__ __forceinline__ T ibmPhi(const IbmKernel<T> &k, T rr) {
  if (k.kind == kKernelPeskin3) {
    const T r = fabs(rr) * k.invh;
    if (r < T(0.5)) return k.invh * T(1 / 3.0) * (T(1.0) + sqrt(T(1.0) + T(-3.0) * r * r));
    if (r < T(1.5)) {
      const T omr = T(1.0) - r;
      return k.invh * T(1 / 6.0) * (T(5.0) - T(3.0) * r - sqrt(T(1.0) + T(-3.0) * omr * omr));
    }
    return T(0);
  } else if (k.kind == kKernelPeskin4) {
    const T r = fabs(rr) * k.invh;
    if (r < T(1.0)) return k.invh * T(0.125) * (T(3.0) - T(2.0) * r + sqrt(T(1.0) + T(4.0) * r * (T(1.0) - r)));
    if (r < T(2.0)) return k.invh * T(0.125) * (T(5.0) - T(2.0) * r - sqrt(T(-7.0) + T(12.0) * r - T(4.0) * r * r));
    return T(0);
  }
  if (k.kind == kKernelBarnettMagland) { // BM(zz, alpha, beta) / norm, IBM_kernels.cuh:83-90,110-112
    const T z = rr / k.rmax;
    const T dz2 = T(1.0) - z * z;
    return dz2 < T(0.0) ? T(0) : exp(k.tau * (sqrt(dz2) - T(1.0))) * k.prefactor;
  }
  if (k.kind == kKernelSixPoint) return ibmSixPoint(fabs(rr) * k.invh) * k.invh;
  return rr >= k.rmax ? T(0) : k.prefactor * exp(k.tau * rr * rr);
}

// Box + Grid in precision T (utils/Box.cuh, utils/Grid.cuh)
template <class T> struct GridT {
  T L[3], m[3], cs[3], ics[3]; // box, minusInvBoxSize (0 = non periodic), cellSize, invCellSize
  int n[3];
  T cellVolume;
  // z window of the cells that are binned / sorted (slab decomposition over GPUs): cells zwin0 .. zwin0+zwinN-1
  // (periodic wrap). The whole grid on one GPU: zwin0 = 0, zwinN = n[2].
  int zwin0, zwinN;
};
// window-local z of global cell plane cz (valid when < zwinN)
template <class T> __host__ __device__ __forceinline__ int windowZ(const GridT<T> &g, int cz) {
  int lz = cz - g.zwin0;
  if (lz < 0) lz += g.n[2];
  return lz;
}

template <class T> inline GridT<T> makeGridT(const double L[3], const int periodic[3], const int cells[3]) {
  GridT<T> g;
  for (int d = 0; d < 3; d++) {
    g.L[d] = (T)L[d];
    g.m[d] = T(-1.0) / g.L[d];
    if (g.L[d] == T(0) || std::isinf((double)g.L[d]) || !periodic[d]) g.m[d] = T(0);
    g.n[d] = cells[d];
    if (d == 2 && g.n[d] == 0) g.n[d] = 1;
    g.cs[d] = g.L[d] / (T)g.n[d];
    g.ics[d] = T(1.0) / g.cs[d];
  }
  if (g.L[2] == T(0)) g.ics[2] = T(0);
  g.cellVolume = g.cs[0] * g.cs[1];
  if (g.n[2] > 1) g.cellVolume *= g.cs[2];
  g.zwin0 = 0;
  g.zwinN = g.n[2];
  return g;
}

template <class T> __device__ __forceinline__ T pbcT(T r, T L, T m) {
  const T off = floor(r * m + T(0.5));
  return m != T(0) ? r + off * L : r;
}
template <class T> __device__ __forceinline__ int cellOfT(const GridT<T> &g, int d, T r) {
  int c = (int)((pbcT(r, g.L[d], g.m[d]) + T(0.5) * g.L[d]) * g.ics[d]);
  return c == g.n[d] ? 0 : c;
}
// Grid::distanceToCellCenter (utils/Grid.cuh:124-131), one coordinate
template <class T> __device__ __forceinline__ T distToCentre(const GridT<T> &g, int d, T r, int cell) {
  return pbcT(r + g.L[d] * T(0.5) - g.cs[d] * ((T)cell + T(0.5)), g.L[d], g.m[d]);
}
// Grid::pbc_cell_coord (utils/Grid.cuh:90-106)
template <class T> __device__ __forceinline__ int wrapCell(const GridT<T> &g, int d, int c) {
  const int nc = g.m[d] != T(0) ? g.n[d] : 0;
  if (c <= -1) c += nc;
  else if (c >= nc) c -= nc;
  return c;
}

// support origin (cell - P) and 1-D weights of one particle in one dimension
template <class T>
__device__ __forceinline__ int supportOrigin(const GridT<T> &g, const IbmKernel<T> &k, int d, T r, int cell) {
  int P = k.support / 2;
  const T dl = fabs(distToCentre(g, d, r, cell - P));
  if (g.cs[d] > T(0) && dl > T(k.support) * g.cs[d] / T(2.0)) P -= 1;
  return cell - P;
}
template <class T>
__device__ __forceinline__ T supportWeight(const GridT<T> &g, const IbmKernel<T> &k, int d, T r, int origin, int i) {
  const int cj = wrapCell(g, d, origin + i);
  return cj >= 0 ? ibmPhi(k, distToCentre(g, d, r, cj)) : T(0);
}

// ---------------- small supports: cell-sorted particles + brick-tiled node-centric spread ----------------
constexpr int kSmallSupport = 7; // largest support served by the sorted (node-centric) kernels

template <class T4>
__global__ void __launch_bounds__(256)
ibmBinByCell(const T4 *__restrict__ pos, int N, GridT<decltype(T4::x)> g, uint32_t *__restrict__ binCount,
             uint2 *__restrict__ codeSlot) {
  using T = decltype(T4::x);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const T4 p = pos[i];
  int cz = cellOfT(g, 2, p.z);
  cz = min(max(cz, 0), g.n[2] - 1);
  const int lz = windowZ(g, cz);
  if (lz >= g.zwinN) { codeSlot[i] = make_uint2(0xffffffffu, 0u); return; } // slab sort: most particles leave here
  int cx = cellOfT(g, 0, p.x), cy = cellOfT(g, 1, p.y);
  cx = min(max(cx, 0), g.n[0] - 1);
  cy = min(max(cy, 0), g.n[1] - 1);
  const uint32_t code = (uint32_t)cx + (uint32_t)g.n[0] * ((uint32_t)cy + (uint32_t)g.n[1] * (uint32_t)lz);
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, code);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  const int rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(binCount + code, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  codeSlot[i] = make_uint2(code, base + rank);
}

// packed per-particle record of the row-brick spread: 3S one-dimensional weights, the spread value, the support origin
template <class T, int S> struct RecGeom {
  static constexpr int W3 = 3 * S;
  static constexpr int REC = ((sizeof(T) == 8 ? W3 + 4 : W3 + 5)) | 1; // odd number of T words per record
  // origin: 3 x 16 bit, biased by 16 (origins range from -S to n)
  // cz: the particle's own z plane (slab ownership of the distributed gather)
  __device__ __forceinline__ static void packOrigin(T *rec, int ox, int oy, int oz, int cz) {
    const uint32_t w0 = (uint32_t)((ox + 16) & 0xffff) | ((uint32_t)((oy + 16) & 0xffff) << 16),
                   w1 = (uint32_t)((oz + 16) & 0xffff) | ((uint32_t)(cz & 0xffff) << 16);
    if (sizeof(T) == 8) {
      const unsigned long long u = ((unsigned long long)w1 << 32) | w0;
      rec[W3 + 3] = (T)__longlong_as_double((long long)u);
    } else {
      rec[W3 + 3] = (T)__uint_as_float(w0);
      rec[W3 + 4] = (T)__uint_as_float(w1);
    }
  }
  __device__ __forceinline__ static void unpackOrigin(const T *rec, int &ox, int &oy, int &oz) {
    int cz;
    unpackOrigin(rec, ox, oy, oz, cz);
  }
  __device__ __forceinline__ static void unpackOrigin(const T *rec, int &ox, int &oy, int &oz, int &cz) {
    uint32_t w0, w1;
    if (sizeof(T) == 8) {
      const unsigned long long u = (unsigned long long)__double_as_longlong((double)rec[W3 + 3]);
      w0 = (uint32_t)u; w1 = (uint32_t)(u >> 32);
    } else {
      w0 = __float_as_uint((float)rec[W3 + 3]); w1 = __float_as_uint((float)rec[W3 + 4]);
    }
    ox = (int)(w0 & 0xffff) - 16; oy = (int)(w0 >> 16) - 16; oz = (int)(w1 & 0xffff) - 16; cz = (int)(w1 >> 16);
  }
};

// stable order inside each cell; cell-sorted copies of the spread value, of the support origin and of the 3*S
// one-dimensional weights (window evaluations happen once per particle, here)
template <class T4, int S>
__global__ void __launch_bounds__(256)
ibmOrderSorted(const int *__restrict__ unstable, const uint2 *__restrict__ codeSlot,
               const uint32_t *__restrict__ binStart, const T4 *__restrict__ pos,
               const decltype(T4::x) *__restrict__ val, int valStride, int N, GridT<decltype(T4::x)> g,
               IbmKernel<decltype(T4::x)> k, int *__restrict__ sortedIndex, decltype(T4::x) *__restrict__ sortedRec) {
  using T = decltype(T4::x);
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  if (g.zwinN != g.n[2] && slot >= (int)binStart[g.n[0] * g.n[1] * g.zwinN]) return; // only the window's particles are sorted
  const int i = unstable[slot];
  const uint32_t code = codeSlot[i].x;
  const int s = (int)binStart[code], e = (int)binStart[code + 1];
  int rank = 0;
  for (int j = s; j < e; j++) rank += (__ldg(unstable + j) < i);
  const int dst = s + rank;
  sortedIndex[dst] = i;
  const T4 p = pos[i];
  T v0 = T(0), v1 = T(0), v2 = T(0);
  if (val) {
    const T *vp = val + (size_t)i * valStride;
    v0 = vp[0]; v1 = vp[1]; v2 = vp[2];
  }
  const T pr[3] = {p.x, p.y, p.z};
  int o[3];
  using R = RecGeom<T, S>;
  T *rec = sortedRec + (size_t)dst * R::REC;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int cell = cellOfT(g, d, pr[d]);
    o[d] = supportOrigin(g, k, d, pr[d], cell);
#pragma unroll
    for (int i = 0; i < S; i++) rec[d * S + i] = supportWeight(g, k, d, pr[d], o[d], i);
  }
  rec[3 * S] = v0; rec[3 * S + 1] = v1; rec[3 * S + 2] = v2;
  R::packOrigin(rec, o[0], o[1], o[2], cellOfT(g, 2, pr[2]));
}

// ---- row-brick spread ----
// A CTA owns a brick of kRbX x kRbY x kRbZ grid nodes; a THREAD owns kRbX nodes in x times two consecutive z planes
// and keeps their accumulators in registers: no atomics, every node is written exactly once (no zero fill of the
// grid, deterministic summation order). The particles that can reach the brick lie in (kRbY+W) x (kRbZ+W) ROWS of
// cells, and because the records are sorted by cell with x fastest, each row is ONE contiguous range of the record
// array (two when the brick touches the periodic x boundary): the set-up is two binStart loads per row and a block
// scan over a few hundred rows, staging is a flat coalesced copy of packed records (odd word stride: conflict-free
// in shared memory), and a thread's candidates for a given z row-plane are one contiguous staged range.
// Particle cells reaching node X in one dimension: X + lo .. X + hi with hi = S/2, lo = -(S-1) + S/2 - (S even)
// (computeSupportShift lowers P by at most one).
constexpr int kRbX = 4, kRbY = 16, kRbZ = 16, kRbThreads = kRbY * kRbZ / 2;

// KB: shared-memory budget of the staged records per pass (a brick with more records takes several passes)
template <class T, int S, int KB = 52> struct RowBrickGeom {
  static constexpr int hi = S / 2, lo = -(S - 1) + S / 2 - ((S & 1) ? 0 : 1);
  static constexpr int W = hi - lo;
  static constexpr int RX = kRbX + W, RY = kRbY + W, RZ = kRbZ + W, nrows = RY * RZ;
  static constexpr int REC = RecGeom<T, S>::REC;
  static constexpr int cap = (KB * 1024) / (REC * (int)sizeof(T)); // staged records per pass
  static constexpr size_t headBytes = (((size_t)nrows * sizeof(int4) + (size_t)(nrows + 1) * sizeof(int) + (size_t)cap * sizeof(int) + (size_t)cap * sizeof(unsigned short)) + 15) / 16 * 16;
  static constexpr size_t smemBytes = headBytes + (size_t)cap * REC * sizeof(T);
};

// d = node - origin folded into [0, n) for periodic dimensions (a single wrap suffices: |d| < n)
__device__ __forceinline__ int foldIndex(int d, int n, bool periodic) {
  if (periodic) { d += d < 0 ? n : 0; d -= d >= n ? n : 0; }
  return d;
}

template <class T, int S, int KB = 52>
__global__ void __launch_bounds__(kRbThreads)
ibmSpreadRows(const T *__restrict__ sortedRec, const uint32_t *__restrict__ binStart, GridT<T> g, int nxPad,
              T *__restrict__ grid3, int zPlane0, int nzLocal) {
  // zPlane0 / nzLocal: the z planes [zPlane0, zPlane0 + nzLocal) this launch writes; grid3 starts at plane zPlane0
  // (the whole grid on one GPU: 0, n[2])
  using G = RowBrickGeom<T, S, KB>;
  constexpr int REC = G::REC;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  int4 *rowSeg = reinterpret_cast<int4 *>(smemRaw);                  // [nrows] {startA, countA, startB, countB}
  int *rowOff = reinterpret_cast<int *>(rowSeg + G::nrows);          // [nrows + 1] staged-order prefix
  int *relOrg = rowOff + G::nrows + 1;                               // [cap] packed brick-relative origin of a staged record
  unsigned short *recRow = reinterpret_cast<unsigned short *>(relOrg + G::cap); // [cap] row of a staged record
  T *recs = reinterpret_cast<T *>(smemRaw + G::headBytes);           // [cap][REC]
  __shared__ int warpTot[kRbThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid % kRbY, tzp = tid / kRbY; // this thread: nodes (bx0.., by0 + ty, bz0 + 2 tzp {, +1})
  const int bx0 = blockIdx.x * kRbX, by0 = blockIdx.y * kRbY, bz0 = zPlane0 + blockIdx.z * kRbZ;
  const bool px = g.m[0] != T(0), py = g.m[1] != T(0), pz = g.m[2] != T(0);

  // ---- rows of the region -> record ranges -> exclusive prefix (block scan) ----
  constexpr int per = (G::nrows + kRbThreads - 1) / kRbThreads;
  int cnt[per];
  int mine = 0;
#pragma unroll
  for (int q = 0; q < per; q++) {
    const int r = tid * per + q;
    cnt[q] = 0;
    if (r < G::nrows) {
      int4 seg = make_int4(0, 0, 0, 0);
      int cy = by0 + G::lo + r % G::RY, cz = bz0 + G::lo + r / G::RY;
      bool ok = true;
      if (cy < 0 || cy >= g.n[1]) { if (py) cy += cy < 0 ? g.n[1] : -g.n[1]; else ok = false; }
      if (cz < 0 || cz >= g.n[2]) { if (pz) cz += cz < 0 ? g.n[2] : -g.n[2]; else ok = false; }
      ok = ok && cy >= 0 && cy < g.n[1] && cz >= 0 && cz < g.n[2] && bx0 < g.n[0]; // bricks of pure padding write zeros
      int lz = 0;
      if (ok) { lz = windowZ(g, cz); ok = lz < g.zwinN; }
      if (ok) {
        const uint32_t *rowBin = binStart + (size_t)g.n[0] * ((size_t)cy + (size_t)g.n[1] * (size_t)lz);
        const int a = bx0 + G::lo, b = bx0 + kRbX - 1 + G::hi; // unwrapped x cells of the region
        const int a0 = max(a, 0), b0 = min(b, g.n[0] - 1);
        if (a0 <= b0) { seg.x = (int)__ldg(rowBin + a0); seg.y = (int)__ldg(rowBin + b0 + 1) - seg.x; }
        if (px && a < 0) { seg.z = (int)__ldg(rowBin + a + g.n[0]); seg.w = (int)__ldg(rowBin + g.n[0]) - seg.z; }
        else if (px && b >= g.n[0]) { seg.z = (int)__ldg(rowBin); seg.w = (int)__ldg(rowBin + min(b - g.n[0], g.n[0] - 1) + 1) - seg.z; }
      }
      rowSeg[r] = seg;
      cnt[q] = seg.y + seg.w;
    }
    mine += cnt[q];
  }
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warpTot[warp] = inc;
  __syncthreads();
  int warpOff = 0, total = 0;
#pragma unroll
  for (int q = 0; q < kRbThreads / 32; q++) {
    if (q < warp) warpOff += warpTot[q];
    total += warpTot[q];
  }
  int run = warpOff + inc - mine;
#pragma unroll
  for (int q = 0; q < per; q++) {
    const int r = tid * per + q;
    if (r < G::nrows) rowOff[r] = run;
    run += cnt[q];
  }
  if (tid == 0) rowOff[G::nrows] = total;
  __syncthreads();

  T acc[kRbX][2][3];
#pragma unroll
  for (int q = 0; q < kRbX; q++)
#pragma unroll
    for (int z = 0; z < 2; z++) acc[q][z][0] = acc[q][z][1] = acc[q][z][2] = T(0);
  const int Y = by0 + ty, Z0 = bz0 + 2 * tzp;

  for (int chunk0 = 0; chunk0 < total; chunk0 += G::cap) {
    // ---- stage: row owners label the staged slots of this chunk, then one thread per record copies it ----
#pragma unroll
    for (int q = 0; q < per; q++) {
      const int r = tid * per + q;
      if (r < G::nrows) {
        const int lo = max(rowOff[r], chunk0), hi2 = min(rowOff[r] + cnt[q], chunk0 + G::cap);
        for (int j = lo; j < hi2; j++) recRow[j - chunk0] = (unsigned short)r;
      }
    }
    __syncthreads();
    const int nstage = min(G::cap, total - chunk0);
    for (int slot = tid; slot < nstage; slot += kRbThreads) {
      const int r = recRow[slot];
      const int4 seg = rowSeg[r];
      const int j = chunk0 + slot - rowOff[r]; // position inside the row: segment A first, then B
      const size_t src = j < seg.y ? (size_t)seg.x + j : (size_t)seg.z + (j - seg.y);
      const T *sp = sortedRec + src * REC;
      T *dp = recs + (size_t)slot * REC;
#pragma unroll
      for (int w = 0; w < 3 * S + 3; w++) dp[w] = __ldg(sp + w);
      // origin relative to the brick, at the periodic image that overlaps it: the node loops then need no folding
      int ox, oy, oz;
      RecGeom<T, S>::unpackOrigin(sp, ox, oy, oz);
      int rx = ox - bx0, ry = oy - by0, rzz = oz - bz0;
      // (a staged record lies within one support of the brick, so exactly one image is in [-(S-1), brick size))
      if (px) { if (rx > kRbX - 1) rx -= g.n[0]; else if (rx < -(S - 1)) rx += g.n[0]; }
      if (py) { if (ry > kRbY - 1) ry -= g.n[1]; else if (ry < -(S - 1)) ry += g.n[1]; }
      if (pz) { if (rzz > kRbZ - 1) rzz -= g.n[2]; else if (rzz < -(S - 1)) rzz += g.n[2]; }
      relOrg[slot] = ((rx + 128) & 0xff) | (((ry + 128) & 0xff) << 8) | (((rzz + 128) & 0xff) << 16);
    }
    __syncthreads();
    // ---- candidates of my nodes: for every z row-plane the rows ty .. ty+W are one contiguous staged range; the
    //      W+2 planes are walked as ONE loop so that a lane idles only when its own candidates are exhausted ----
    int plane = 0, sIdx = 0, sEnd = 0;
    while (true) {
      while (sIdx >= sEnd) {
        if (plane > G::W + 1) break;
        const int r0 = ty + G::RY * (2 * tzp + plane);
        sIdx = max(rowOff[r0] - chunk0, 0);
        sEnd = min(rowOff[r0 + G::W + 1] - chunk0, G::cap);
        plane++;
      }
      if (sIdx >= sEnd) break;
      const int s = sIdx++;
      const int pk = relOrg[s];
      const int iy = ty - (((pk >> 8) & 0xff) - 128);
      const int iz0 = 2 * tzp - (((pk >> 16) & 0xff) - 128);
      const bool v0ok = (unsigned)iz0 < (unsigned)S, v1ok = (unsigned)(iz0 + 1) < (unsigned)S;
      if ((unsigned)iy >= (unsigned)S || !(v0ok || v1ok)) continue;
      const T *rec = recs + (size_t)s * REC;
      const T wy = rec[S + iy];
      const T wz0 = v0ok ? rec[2 * S + iz0] : T(0), wz1 = v1ok ? rec[2 * S + iz0 + 1] : T(0);
      const T wyz0 = wy * wz0, wyz1 = wy * wz1;
      const T f0 = rec[3 * S], f1 = rec[3 * S + 1], f2 = rec[3 * S + 2];
      const int ix0 = -((pk & 0xff) - 128);
#pragma unroll
      for (int lx = 0; lx < kRbX; lx++) {
        const int ix = ix0 + lx;
        const T wx = (unsigned)ix < (unsigned)S ? rec[ix] : T(0);
        const T w0 = wx * wyz0, w1 = wx * wyz1;
        acc[lx][0][0] += f0 * w0; acc[lx][0][1] += f1 * w0; acc[lx][0][2] += f2 * w0;
        acc[lx][1][0] += f0 * w1; acc[lx][1][1] += f1 * w1; acc[lx][1][2] += f2 * w1;
      }
    }
    __syncthreads();
  }
  if (Y < g.n[1]) {
#pragma unroll
    for (int z = 0; z < 2; z++) {
      const int Z = Z0 + z;
      if (Z >= zPlane0 + nzLocal) continue;
#pragma unroll
      for (int lx = 0; lx < kRbX; lx++) {
        const int X = bx0 + lx;
        if (X < nxPad) {
          T *out = grid3 + 3 * ((size_t)X + (size_t)nxPad * ((size_t)Y + (size_t)g.n[1] * (Z - zPlane0)));
          const bool real = X < g.n[0];
          out[0] = real ? acc[lx][z][0] : T(0);
          out[1] = real ? acc[lx][z][1] : T(0);
          out[2] = real ? acc[lx][z][2] : T(0);
        }
      }
    }
  }
}

// Interpolation: one thread per particle slot (cell-sorted, so neighbouring threads read neighbouring nodes),
// support origin and weights precomputed by ibmOrderSorted, S^3 node loads fully unrolled.
template <class T, int S, bool ACCUMULATE>
__global__ void __launch_bounds__(128)
ibmGatherSorted(const T *__restrict__ sortedRec, const int *__restrict__ sortedIndex, int N, GridT<T> g, int nxPad,
                const T *__restrict__ grid3, T *__restrict__ out3) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  const T *wsrc = sortedRec + (size_t)slot * RecGeom<T, S>::REC;
  int o[3];
  RecGeom<T, S>::unpackOrigin(wsrc, o[0], o[1], o[2]);
  T w[3][S];
  int cidx[3][S];
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int i = 0; i < S; i++) {
      const int cj = wrapCell(g, d, o[d] + i);
      cidx[d][i] = (cj >= 0 && cj < g.n[d]) ? cj : -1;
      w[d][i] = wsrc[d * S + i];
    }
  T ax = T(0), ay = T(0), az = T(0);
#pragma unroll
  for (int kk = 0; kk < S; kk++)
#pragma unroll
    for (int jj = 0; jj < S; jj++)
#pragma unroll
      for (int ii = 0; ii < S; ii++) {
        if (cidx[0][ii] < 0 || cidx[1][jj] < 0 || cidx[2][kk] < 0) continue;
        const T *gp = grid3 + 3 * ((size_t)cidx[0][ii] + (size_t)nxPad * ((size_t)cidx[1][jj] + (size_t)g.n[1] * cidx[2][kk]));
        const T wx = w[0][ii], wy = w[1][jj], wz = w[2][kk];
        ax += g.cellVolume * (__ldg(gp) * wx * wy * wz);
        ay += g.cellVolume * (__ldg(gp + 1) * wx * wy * wz);
        az += g.cellVolume * (__ldg(gp + 2) * wx * wy * wz);
      }
  T *op = out3 + 3 * (size_t)sortedIndex[slot];
  if (ACCUMULATE) { op[0] += ax; op[1] += ay; op[2] += az; }
  else { op[0] = ax; op[1] = ay; op[2] = az; }
}

// Slab-decomposed interpolation: this rank owns the z planes [z0, z0 + nzl) of the grid and the particles whose
// cell lies in them; its slab carries `halo` planes of its neighbours on either side (pushed there by the fused z
// pass), so every load is local. The sorted window also holds the neighbours' boundary particles (needed by the
// spread), which are skipped. Results go, in sorted-slot order, into a packed local buffer {index, row}; a separate
// kernel pushes that buffer to every rank with wide coalesced stores and the ranks scatter what they received.
constexpr int kMaxPeers = 8;
template <class T> struct PeerTable { T *p[kMaxPeers]; };

template <class T, int S>
__global__ void __launch_bounds__(128)
ibmGatherSortedSlab(const T *__restrict__ sortedRec, const int *__restrict__ sortedIndex, const uint32_t *__restrict__ binStart,
                    GridT<T> g, int nxPad, const T *__restrict__ slab, int z0, int nzl, int halo, int *__restrict__ packIdx,
                    T *__restrict__ packRows, int *__restrict__ packCount) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int nLocal = (int)binStart[g.n[0] * g.n[1] * g.zwinN];
  if (slot == 0) *packCount = nLocal;
  if (slot >= nLocal) return;
  const T *wsrc = sortedRec + (size_t)slot * RecGeom<T, S>::REC;
  int o[3], cz;
  RecGeom<T, S>::unpackOrigin(wsrc, o[0], o[1], o[2], cz);
  if (cz < z0 || cz >= z0 + nzl) { packIdx[slot] = -1; return; } // a neighbour's particle
  T w[3][S];
  int cidx[2][S];
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int i = 0; i < S; i++) {
      w[d][i] = wsrc[d * S + i];
      if (d < 2) {
        const int cj = wrapCell(g, d, o[d] + i);
        cidx[d][i] = (cj >= 0 && cj < g.n[d]) ? cj : -1;
      }
    }
  // support planes relative to the slab: the origin is at most one support below the particle's own plane
  const int lz0 = (o[2] - cz) + (cz - z0) + halo;
  T ax = T(0), ay = T(0), az = T(0);
#pragma unroll
  for (int kk = 0; kk < S; kk++) {
    const T *plane = slab + 3 * (size_t)nxPad * g.n[1] * (size_t)(lz0 + kk);
#pragma unroll
    for (int jj = 0; jj < S; jj++)
#pragma unroll
      for (int ii = 0; ii < S; ii++) {
        if (cidx[0][ii] < 0 || cidx[1][jj] < 0) continue;
        const T *gp = plane + 3 * ((size_t)cidx[0][ii] + (size_t)nxPad * (size_t)cidx[1][jj]);
        const T wx = w[0][ii], wy = w[1][jj], wz = w[2][kk];
        ax += g.cellVolume * (__ldg(gp) * wx * wy * wz);
        ay += g.cellVolume * (__ldg(gp + 1) * wx * wy * wz);
        az += g.cellVolume * (__ldg(gp + 2) * wx * wy * wz);
      }
  }
  packIdx[slot] = sortedIndex[slot];
  packRows[3 * (size_t)slot] = ax; packRows[3 * (size_t)slot + 1] = ay; packRows[3 * (size_t)slot + 2] = az;
}

// push the packed {count, index[], rows[]} block of this rank into inbox[rank] of every peer (coalesced remote stores)
template <class T>
__global__ void __launch_bounds__(256)
slabPushPacked(const int *__restrict__ packCount, const int *__restrict__ packIdx, const T *__restrict__ packRows,
               PeerTable<int> inboxCount, PeerTable<int> inboxIdx, PeerTable<T> inboxRows, int world) {
  const int n = *packCount;
  const int dest = blockIdx.y;
  if (dest >= world) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) *inboxCount.p[dest] = n;
  int *di = inboxIdx.p[dest];
  T *dr = inboxRows.p[dest];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) di[i] = packIdx[i];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * (size_t)n; i += (size_t)gridDim.x * blockDim.x) dr[i] = packRows[i];
}

// out3[index] = row for everything the ranks pushed into my inboxes (each particle is owned by exactly one rank)
template <class T>
__global__ void __launch_bounds__(256)
slabScatterInbox(PeerTable<int> inboxCount, PeerTable<int> inboxIdx, PeerTable<T> inboxRows, int world, T *__restrict__ out3) {
  const int src = blockIdx.y;
  if (src >= world) return;
  const int n = *inboxCount.p[src];
  const int *ii = inboxIdx.p[src];
  const T *rr = inboxRows.p[src];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int idx = ii[i];
    if (idx < 0) continue;
    out3[3 * (size_t)idx] = rr[3 * (size_t)i]; out3[3 * (size_t)idx + 1] = rr[3 * (size_t)i + 1]; out3[3 * (size_t)idx + 2] = rr[3 * (size_t)i + 2];
  }
}

// ---------------- any support: one warp per particle ----------------
template <class T4, class V, bool SPREAD, bool ACCUMULATE>
__global__ void __launch_bounds__(128)
ibmWarpPerParticle(const T4 *__restrict__ pos, const V *__restrict__ val, int valStride, int N,
                   GridT<decltype(T4::x)> g, IbmKernel<decltype(T4::x)> k, int nxPad,
                   decltype(T4::x) *__restrict__ grid3, decltype(T4::x) *__restrict__ out3, const int *__restrict__ order) {
  // order: optional cell-sorted permutation (the sort of a preceding spread): neighbouring warps then touch neighbouring
  // nodes and the grid is served from L1/L2 instead of being gathered at random from HBM
  using T = decltype(T4::x);
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
  T v[3] = {T(0), T(0), T(0)};
  if (SPREAD) {
    const T *vp = reinterpret_cast<const T *>(val) + (size_t)i * valStride;
    v[0] = vp[0]; v[1] = vp[1]; v[2] = vp[2];
  }
  T ax = T(0), ay = T(0), az = T(0);
  const int total = S * S * S;
  for (int t = lane; t < total; t += 32) {
    const int ii = t % S, jj = (t / S) % S, kk = t / (S * S);
    const int cx = wrapCell(g, 0, o[0] + ii), cy = wrapCell(g, 1, o[1] + jj), cz = wrapCell(g, 2, o[2] + kk);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= g.n[0] || cy >= g.n[1] || cz >= g.n[2]) continue;
    T *gp = grid3 + 3 * ((size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g.n[1] * cz));
    const T wx = w[ii], wy = w[kMaxSupport + jj], wz = w[2 * kMaxSupport + kk];
    if (SPREAD) {
      // real3 atomicAdd of the reference skips zero components (utils/atomics.cuh)
      const T c0 = v[0] * wx * wy * wz, c1 = v[1] * wx * wy * wz, c2 = v[2] * wx * wy * wz;
      if (c0 != T(0)) atomicAdd(gp, c0);
      if (c1 != T(0)) atomicAdd(gp + 1, c1);
      if (c2 != T(0)) atomicAdd(gp + 2, c2);
    } else {
      ax += g.cellVolume * (gp[0] * wx * wy * wz);
      ay += g.cellVolume * (gp[1] * wx * wy * wz);
      az += g.cellVolume * (gp[2] * wx * wy * wz);
    }
  }
  if (!SPREAD) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, off);
      ay += __shfl_xor_sync(0xffffffffu, ay, off);
      az += __shfl_xor_sync(0xffffffffu, az, off);
    }
    if (lane == 0) {
      T *op = out3 + 3 * (size_t)i;
      if (ACCUMULATE) { op[0] += ax; op[1] += ay; op[2] += az; }
      else { op[0] = ax; op[1] = ay; op[2] = az; }
    }
  }
}

} // namespace ub200
