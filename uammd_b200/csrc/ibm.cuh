// Immersed boundary spreading / interpolation for sm_100a.
//
// Replaces IBM<Kernel>::spread / gather (misc/IBM.cuh:117-184) and their kernels particles2GridD /
// grid2ParticlesDTPP (misc/IBM.cu:83-147,168-235): same support-cell rule (computeSupportShift, IBM.cu:11-31),
// same cell-centred distances and weights (fillSharedWeights, IBM.cu:33-66), same products
// value*phiX*phiY*phiZ (DefaultWeightCompute, IBM.cuh:88-97) and quadrature weight = cell volume (:80-86).
//
// Small supports (Peskin 3/4 point): spreading is turned into an atomic-free, write-once NODE gather:
//   binByCell -> stable order -> per-particle stencil records (support origin + 1-D weights + value) ->
//   one thread per grid node sums the records of the particles whose support covers it and writes the node
//   exactly once (no memset, no atomics, deterministic). The reference launches one 32-thread block per
//   particle issuing 3*support^3 scalar atomics into a zero-filled grid.
// Large supports (Gaussian, up to 32 points per dimension): one warp per particle, weights in shared memory,
//   atomics for spreading / shuffle reduction for interpolation.
#pragma once
#include "common.cuh"

namespace ub200 {

enum { kKernelPeskin3 = 0, kKernelPeskin4 = 1, kKernelGaussian = 2 };
constexpr int kMaxSupport = 32;

template <class T> struct IbmKernel {
  int kind, support;
  T invh;                // Peskin
  T prefactor, tau, rmax; // Gaussian
};

// window functions: misc/IBM_kernels.cuh:118-137 (3 point), :140-157 (4 point), :28-40 + FCM_kernels.cuh:54-56
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
  return rr >= k.rmax ? T(0) : k.prefactor * exp(k.tau * rr * rr);
}

// Box + Grid in precision T (utils/Box.cuh, utils/Grid.cuh)
template <class T> struct GridT {
  T L[3], m[3], cs[3], ics[3]; // box, minusInvBoxSize (0 = non periodic), cellSize, invCellSize
  int n[3];
  T cellVolume;
};

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

// ---------------- small supports: sorted records + node-centric spread ----------------
constexpr int kSmallSupport = 4;
template <class T> struct StencilRec {
  int ox, oy, oz, pad; // support origin (unwrapped)
  T w[3 * kSmallSupport];
  T v[3];
  T pad2;
};

template <class T4>
__global__ void __launch_bounds__(256)
ibmBinByCell(const T4 *__restrict__ pos, int N, GridT<decltype(T4::x)> g, uint32_t *__restrict__ binCount,
             uint2 *__restrict__ codeSlot) {
  using T = decltype(T4::x);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const T4 p = pos[i];
  int cx = cellOfT(g, 0, p.x), cy = cellOfT(g, 1, p.y), cz = cellOfT(g, 2, p.z);
  cx = min(max(cx, 0), g.n[0] - 1);
  cy = min(max(cy, 0), g.n[1] - 1);
  cz = min(max(cz, 0), g.n[2] - 1);
  const uint32_t code = (uint32_t)cx + (uint32_t)g.n[0] * ((uint32_t)cy + (uint32_t)g.n[1] * (uint32_t)cz);
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

// stable order inside each cell + stencil record of the particle in its sorted slot
template <class T4, class V>
__global__ void __launch_bounds__(256)
ibmOrderAndStencil(const int *__restrict__ unstable, const uint2 *__restrict__ codeSlot,
                   const uint32_t *__restrict__ binStart, const T4 *__restrict__ pos, const V *__restrict__ val,
                   int valStride, int N, GridT<decltype(T4::x)> g, IbmKernel<decltype(T4::x)> k,
                   int *__restrict__ sortedIndex, StencilRec<decltype(T4::x)> *__restrict__ recs) {
  using T = decltype(T4::x);
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  const int i = unstable[slot];
  const uint32_t code = codeSlot[i].x;
  const int s = (int)binStart[code], e = (int)binStart[code + 1];
  int rank = 0;
  for (int j = s; j < e; j++) rank += (__ldg(unstable + j) < i);
  const int dst = s + rank;
  sortedIndex[dst] = i;
  const T4 p = pos[i];
  StencilRec<T> r;
  const T pr[3] = {p.x, p.y, p.z};
  int o[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int c = cellOfT(g, d, pr[d]);
    o[d] = supportOrigin(g, k, d, pr[d], c);
#pragma unroll
    for (int q = 0; q < kSmallSupport; q++)
      r.w[d * kSmallSupport + q] = q < k.support ? supportWeight(g, k, d, pr[d], o[d], q) : T(0);
  }
  r.ox = o[0]; r.oy = o[1]; r.oz = o[2]; r.pad = 0;
  if (val) {
    const T *vp = reinterpret_cast<const T *>(val) + (size_t)i * valStride;
    r.v[0] = vp[0]; r.v[1] = vp[1]; r.v[2] = vp[2];
  } else {
    r.v[0] = r.v[1] = r.v[2] = T(0);
  }
  r.pad2 = T(0);
  recs[dst] = r;
}

// One thread per grid node (x fastest, padded pitch): sums the particles whose support covers the node.
// Particle cells that can reach node X in one dimension: X - S + 1 + Pmin .. X + Pmax with Pmax = S/2 and
// Pmin = S/2 - (S even) (computeSupportShift only ever lowers P by one, and only for even supports).
template <class T>
__global__ void __launch_bounds__(128)
ibmSpreadNodes(const StencilRec<T> *__restrict__ recs, const uint32_t *__restrict__ binStart, GridT<T> g, int support,
               int nxPad, T *__restrict__ grid3) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, Z = blockIdx.z;
  if (X >= nxPad) return;
  T ax = T(0), ay = T(0), az = T(0);
  if (X < g.n[0]) {
    const int S = support;
    const int lo = -(S - 1) + (S / 2 - ((S & 1) ? 0 : 1)), hi = S / 2; // particle cell offsets relative to the node
    for (int dz = lo; dz <= hi; dz++) {
      int cz = Z + dz;
      if (cz < 0 || cz >= g.n[2]) {
        if (g.m[2] == T(0)) continue;
        cz += cz < 0 ? g.n[2] : -g.n[2];
      }
      for (int dy = lo; dy <= hi; dy++) {
        int cy = Y + dy;
        if (cy < 0 || cy >= g.n[1]) {
          if (g.m[1] == T(0)) continue;
          cy += cy < 0 ? g.n[1] : -g.n[1];
        }
        const uint32_t row = (uint32_t)g.n[0] * ((uint32_t)cy + (uint32_t)g.n[1] * (uint32_t)cz);
        // particle cells X+lo .. X+hi of this row; a periodic wrap splits the range in two segments
        // (the host guarantees n >= support + 1 in periodic dimensions, so the segments never overlap)
        const int a = X + lo, b = X + hi;
        int segLo[2], segHi[2], nseg = 1;
        if (g.m[0] == T(0)) { segLo[0] = max(a, 0); segHi[0] = min(b, g.n[0] - 1); }
        else if (a < 0) { segLo[0] = a + g.n[0]; segHi[0] = g.n[0] - 1; segLo[1] = 0; segHi[1] = b; nseg = 2; }
        else if (b >= g.n[0]) { segLo[0] = a; segHi[0] = g.n[0] - 1; segLo[1] = 0; segHi[1] = b - g.n[0]; nseg = 2; }
        else { segLo[0] = a; segHi[0] = b; }
        for (int seg = 0; seg < nseg; seg++) {
          const int x0 = segLo[seg], x1 = segHi[seg];
          if (x0 > x1) continue;
          const int pb = (int)__ldg(binStart + row + x0), pe = (int)__ldg(binStart + row + x1 + 1);
          for (int q = pb; q < pe; q++) {
            const StencilRec<T> *r = recs + q;
            int ix = X - r->ox, iy = Y - r->oy, iz = Z - r->oz;
            // periodic images of the node relative to the (unwrapped) support origin
            if (g.m[0] != T(0)) { if (ix < 0) ix += g.n[0]; else if (ix >= g.n[0]) ix -= g.n[0]; }
            if (g.m[1] != T(0)) { if (iy < 0) iy += g.n[1]; else if (iy >= g.n[1]) iy -= g.n[1]; }
            if (g.m[2] != T(0)) { if (iz < 0) iz += g.n[2]; else if (iz >= g.n[2]) iz -= g.n[2]; }
            if ((unsigned)ix < (unsigned)S && (unsigned)iy < (unsigned)S && (unsigned)iz < (unsigned)S) {
              const T wx = r->w[ix], wy = r->w[kSmallSupport + iy], wz = r->w[2 * kSmallSupport + iz];
              ax += r->v[0] * wx * wy * wz;
              ay += r->v[1] * wx * wy * wz;
              az += r->v[2] * wx * wy * wz;
            }
          }
        }
      }
    }
  }
  T *out = grid3 + 3 * ((size_t)X + (size_t)nxPad * ((size_t)Y + (size_t)g.n[1] * Z));
  out[0] = ax; out[1] = ay; out[2] = az;
}

// Interpolation with the sorted stencil records: one thread per particle slot, S^3 nodes.
template <class T, bool ACCUMULATE>
__global__ void __launch_bounds__(128)
ibmGatherSorted(const StencilRec<T> *__restrict__ recs, const int *__restrict__ sortedIndex, int N, GridT<T> g,
                int support, int nxPad, const T *__restrict__ grid3, T *__restrict__ out3) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  const StencilRec<T> r = recs[slot];
  T ax = T(0), ay = T(0), az = T(0);
  for (int kk = 0; kk < support; kk++) {
    const int cz = wrapCell(g, 2, r.oz + kk);
    if (cz < 0 || cz >= g.n[2]) continue;
    for (int jj = 0; jj < support; jj++) {
      const int cy = wrapCell(g, 1, r.oy + jj);
      if (cy < 0 || cy >= g.n[1]) continue;
      for (int ii = 0; ii < support; ii++) {
        const int cx = wrapCell(g, 0, r.ox + ii);
        if (cx < 0 || cx >= g.n[0]) continue;
        const T *gp = grid3 + 3 * ((size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g.n[1] * cz));
        const T wx = r.w[ii], wy = r.w[kSmallSupport + jj], wz = r.w[2 * kSmallSupport + kk];
        ax += g.cellVolume * (__ldg(gp) * wx * wy * wz);
        ay += g.cellVolume * (__ldg(gp + 1) * wx * wy * wz);
        az += g.cellVolume * (__ldg(gp + 2) * wx * wy * wz);
      }
    }
  }
  T *o = out3 + 3 * (size_t)sortedIndex[slot];
  if (ACCUMULATE) { o[0] += ax; o[1] += ay; o[2] += az; }
  else { o[0] = ax; o[1] = ay; o[2] = az; }
}

// ---------------- any support: one warp per particle ----------------
template <class T4, class V, bool SPREAD, bool ACCUMULATE>
__global__ void __launch_bounds__(128)
ibmWarpPerParticle(const T4 *__restrict__ pos, const V *__restrict__ val, int valStride, int N,
                   GridT<decltype(T4::x)> g, IbmKernel<decltype(T4::x)> k, int nxPad,
                   decltype(T4::x) *__restrict__ grid3, decltype(T4::x) *__restrict__ out3) {
  using T = decltype(T4::x);
  __shared__ T wsh[4][3 * kMaxSupport];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + warp;
  if (i >= N) return;
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
