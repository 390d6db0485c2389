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
constexpr int kSmallSupport = 4;

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
  const int lz = windowZ(g, cz);
  if (lz >= g.zwinN) { codeSlot[i] = make_uint2(0xffffffffu, 0u); return; }
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

// stable order inside each cell; cell-sorted copies of the spread value, of the support origin and of the 3*S
// one-dimensional weights (window evaluations happen once per particle, here)
template <class T4, int S>
__global__ void __launch_bounds__(256)
ibmOrderSorted(const int *__restrict__ unstable, const uint2 *__restrict__ codeSlot,
               const uint32_t *__restrict__ binStart, const T4 *__restrict__ pos,
               const decltype(T4::x) *__restrict__ val, int valStride, int N, GridT<decltype(T4::x)> g,
               IbmKernel<decltype(T4::x)> k, int *__restrict__ sortedIndex, T4 *__restrict__ sortedPos,
               decltype(T4::x) *__restrict__ sortedVal, int4 *__restrict__ sortedOrigin,
               decltype(T4::x) *__restrict__ sortedW) {
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
  T4 p = pos[i];
  T v0 = T(0), v1 = T(0), v2 = T(0);
  if (val) {
    const T *vp = val + (size_t)i * valStride;
    v0 = vp[0]; v1 = vp[1]; v2 = vp[2];
  }
  const T pr[3] = {p.x, p.y, p.z};
  int o[3];
  T *wout = sortedW + (size_t)dst * (3 * S);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int cell = cellOfT(g, d, pr[d]);
    o[d] = supportOrigin(g, k, d, pr[d], cell);
#pragma unroll
    for (int i = 0; i < S; i++) wout[d * S + i] = supportWeight(g, k, d, pr[d], o[d], i);
  }
  sortedOrigin[dst] = make_int4(o[0], o[1], o[2], cellOfT(g, 2, pr[2])); // .w: the particle's own z plane (slab ownership)
  p.w = v0; // {x, y, z, v.x} in one record, {v.y, v.z} next to it
  sortedPos[dst] = p;
  sortedVal[2 * (size_t)dst] = v1;
  sortedVal[2 * (size_t)dst + 1] = v2;
}

// Brick of kBrickX x kBrickY x kBrickZ grid nodes per CTA; every THREAD owns one short x-row of kBrickX nodes
// and keeps their 3*kBrickX accumulators in registers, so there are no atomics at all. The particles of the
// brick's halo region (cells that can reach a node of the brick) are staged in shared memory once (origin
// relative to the brick, 3*S precomputed weights, value); a thread then walks the particles of the (S or S+1)^2
// region rows around its own row - contiguous in staged order for fixed z - and adds the contributions that
// land on its nodes. The brick is written out once: no zero fill of the grid, deterministic summation order.
// Particle cells reaching node X in one dimension: X + lo .. X + hi with hi = S/2, lo = -(S-1) + S/2 - (S even)
// (computeSupportShift lowers P by at most one).
constexpr int kBrickX = 4, kBrickY = 16, kBrickZ = 16, kBrickThreads = kBrickY * kBrickZ;
template <class T, int S> struct BrickGeom {
  static constexpr int hi = S / 2, lo = -(S - 1) + S / 2 - ((S & 1) ? 0 : 1);
  static constexpr int W = hi - lo; // extra region cells per dimension
  static constexpr int RX = kBrickX + W, RY = kBrickY + W, RZ = kBrickZ + W;
  static constexpr int ncells = RX * RY * RZ, nrows = RY * RZ;
  static constexpr int cap = 512;                 // staged particles per pass
  static constexpr int REC = 3 * S + 3;           // weights + value per particle (T)
  static constexpr size_t smemBytes = (size_t)cap * (REC * sizeof(T) + sizeof(int)) + (size_t)(2 * ncells + 1) * sizeof(int) + 64;
};

template <class T4, int S>
__global__ void __launch_bounds__(kBrickThreads)
ibmSpreadBricks(const T4 *__restrict__ sortedPos, const decltype(T4::x) *__restrict__ sortedVal,
                const int4 *__restrict__ sortedOrigin, const decltype(T4::x) *__restrict__ sortedW,
                const uint32_t *__restrict__ binStart, GridT<decltype(T4::x)> g, int nxPad,
                decltype(T4::x) *__restrict__ grid3, int zPlane0, int nzLocal) {
  // zPlane0 / nzLocal: the z planes [zPlane0, zPlane0 + nzLocal) this launch writes; grid3 starts at plane zPlane0
  // (the whole grid on one GPU: 0, n[2])
  using T = decltype(T4::x);
  using G = BrickGeom<T, S>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  T *recs = reinterpret_cast<T *>(smemRaw);                        // [cap][REC]
  int *org = reinterpret_cast<int *>(recs + (size_t)G::cap * G::REC); // [cap] packed origin relative to the brick
  int *cellOff = org + G::cap;                                     // [ncells+1] staged-order prefix of region cells
  int *cellGStart = cellOff + G::ncells + 1;                       // [ncells] first sorted slot of each region cell
  __shared__ int warpTot[kBrickThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid % kBrickY, tz = tid / kBrickY;
  const int bx0 = blockIdx.x * kBrickX, by0 = blockIdx.y * kBrickY, bz0 = zPlane0 + blockIdx.z * kBrickZ;

  // ---- region cell populations -> exclusive prefix (block scan); thread t owns cells [t*per, (t+1)*per) ----
  constexpr int per = (G::ncells + kBrickThreads - 1) / kBrickThreads;
  int cnt[per];
  int mine = 0;
#pragma unroll
  for (int q = 0; q < per; q++) {
    const int c = tid * per + q;
    cnt[q] = 0;
    if (c < G::ncells) {
      int gs = 0;
      const int rx = c % G::RX, ry = (c / G::RX) % G::RY, rz = c / (G::RX * G::RY);
      int cx = bx0 + G::lo + rx, cy = by0 + G::lo + ry, cz = bz0 + G::lo + rz;
      bool ok = true;
      // unwrapped -> actual cell (periodic image) or nothing (non periodic / beyond one wrap)
      if (cx < 0 || cx >= g.n[0]) { if (g.m[0] != T(0)) cx += cx < 0 ? g.n[0] : -g.n[0]; else ok = false; }
      if (cy < 0 || cy >= g.n[1]) { if (g.m[1] != T(0)) cy += cy < 0 ? g.n[1] : -g.n[1]; else ok = false; }
      if (cz < 0 || cz >= g.n[2]) { if (g.m[2] != T(0)) cz += cz < 0 ? g.n[2] : -g.n[2]; else ok = false; }
      ok = ok && cx >= 0 && cx < g.n[0] && cy >= 0 && cy < g.n[1] && cz >= 0 && cz < g.n[2];
      int lz = 0;
      if (ok) { lz = windowZ(g, cz); ok = lz < g.zwinN; }
      if (ok) {
        const uint32_t cell = (uint32_t)cx + (uint32_t)g.n[0] * ((uint32_t)cy + (uint32_t)g.n[1] * (uint32_t)lz);
        const uint32_t s0 = __ldg(binStart + cell), s1 = __ldg(binStart + cell + 1);
        gs = (int)s0;
        cnt[q] = (int)(s1 - s0);
      }
      cellGStart[c] = gs;
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
  for (int q = 0; q < kBrickThreads / 32; q++) {
    if (q < warp) warpOff += warpTot[q];
    total += warpTot[q];
  }
  int run = warpOff + inc - mine;
#pragma unroll
  for (int q = 0; q < per; q++) {
    const int c = tid * per + q;
    if (c < G::ncells) cellOff[c] = run;
    run += cnt[q];
  }
  if (tid == 0) cellOff[G::ncells] = total;
  __syncthreads();

  T acc[kBrickX][3];
#pragma unroll
  for (int q = 0; q < kBrickX; q++) acc[q][0] = acc[q][1] = acc[q][2] = T(0);

  for (int chunk0 = 0; chunk0 < total; chunk0 += G::cap) {
    // ---- stage: one thread per staged particle; its region row by binary search, its cell by a short walk ----
    const int nstage = min(G::cap, total - chunk0);
    for (int slot = tid; slot < nstage; slot += kBrickThreads) {
      const int t = chunk0 + slot;
      int loc = 0, hic = G::ncells; // largest c with cellOff[c] <= t (cellOff[ncells] = total > t)
      while (hic - loc > 1) {
        const int mid = (loc + hic) >> 1;
        if (cellOff[mid] <= t) loc = mid; else hic = mid;
      }
      const int c = loc;
      const int src = cellGStart[c] + (t - cellOff[c]);
      const int rx = c % G::RX, ry = (c / G::RX) % G::RY, rz = c / (G::RX * G::RY);
      const int rc[3] = {rx, ry, rz};
      const int b0[3] = {bx0, by0, bz0};
      const int4 og = sortedOrigin[src];
      const int oabs[3] = {og.x, og.y, og.z};
      int packed = 0;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        // actual cell from the region coordinate; support shift P = cell - origin; origin relative to the brick
        // from the (unwrapped) region coordinate, valid for any periodic image
        int cell = b0[d] + G::lo + rc[d];
        if (cell < 0) cell += g.n[d]; else if (cell >= g.n[d]) cell -= g.n[d];
        const int rel = G::lo + rc[d] - (cell - oabs[d]);
        packed |= ((rel + 64) & 0xff) << (8 * d);
      }
      org[slot] = packed;
      const T *wsrc = sortedW + (size_t)src * (3 * S);
#pragma unroll
      for (int q = 0; q < 3 * S; q++) recs[q * G::cap + slot] = wsrc[q]; // field-major: lanes hit distinct banks
      recs[(3 * S) * G::cap + slot] = sortedPos[src].w;
      recs[(3 * S + 1) * G::cap + slot] = sortedVal[2 * (size_t)src];
      recs[(3 * S + 2) * G::cap + slot] = sortedVal[2 * (size_t)src + 1];
    }
    __syncthreads();
    // ---- my row of nodes: for each dz the region rows ty .. ty+W are contiguous in staged order ----
    for (int dz = 0; dz <= G::W; dz++) {
      const int r0 = ty + G::RY * (tz + dz);
      int b = cellOff[r0 * G::RX] - chunk0, e = cellOff[(r0 + G::W + 1) * G::RX] - chunk0;
      b = max(b, 0); e = min(e, G::cap);
      for (int s = b; s < e; s++) {
        const int pk = org[s];
        const int iy = ty - (((pk >> 8) & 0xff) - 64), iz = tz - (((pk >> 16) & 0xff) - 64);
        if ((unsigned)iy >= (unsigned)S || (unsigned)iz >= (unsigned)S) continue;
        const T *rec = recs + s;
        const T wyz = rec[(S + iy) * G::cap] * rec[(2 * S + iz) * G::cap];
        const T v0 = rec[(3 * S) * G::cap], v1 = rec[(3 * S + 1) * G::cap], v2 = rec[(3 * S + 2) * G::cap];
        const int relx = (pk & 0xff) - 64;
#pragma unroll
        for (int lx = 0; lx < kBrickX; lx++) {
          const int ix = lx - relx;
          if ((unsigned)ix < (unsigned)S) {
            const T wgt = rec[ix * G::cap] * wyz;
            acc[lx][0] += v0 * wgt;
            acc[lx][1] += v1 * wgt;
            acc[lx][2] += v2 * wgt;
          }
        }
      }
    }
    __syncthreads();
  }
  const int Y = by0 + ty, Z = bz0 + tz;
  if (Y < g.n[1] && Z < zPlane0 + nzLocal) {
#pragma unroll
    for (int lx = 0; lx < kBrickX; lx++) {
      const int X = bx0 + lx;
      if (X < nxPad) {
        T *out = grid3 + 3 * ((size_t)X + (size_t)nxPad * ((size_t)Y + (size_t)g.n[1] * (Z - zPlane0)));
        const bool real = X < g.n[0];
        out[0] = real ? acc[lx][0] : T(0);
        out[1] = real ? acc[lx][1] : T(0);
        out[2] = real ? acc[lx][2] : T(0);
      }
    }
  }
}

// Interpolation: one thread per particle slot (cell-sorted, so neighbouring threads read neighbouring nodes),
// support origin and weights precomputed by ibmOrderSorted, S^3 node loads fully unrolled.
template <class T, int S, bool ACCUMULATE>
__global__ void __launch_bounds__(128)
ibmGatherSorted(const int4 *__restrict__ sortedOrigin, const T *__restrict__ sortedW,
                const int *__restrict__ sortedIndex, int N, GridT<T> g, int nxPad, const T *__restrict__ grid3,
                T *__restrict__ out3) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= N) return;
  const int4 og = sortedOrigin[slot];
  const int o[3] = {og.x, og.y, og.z};
  T w[3][S];
  int cidx[3][S];
  const T *wsrc = sortedW + (size_t)slot * (3 * S);
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
// cell lies in them; the sorted window also holds the neighbours' boundary particles (needed by the spread), which
// are skipped here. Grid planes are read through a table of peer-mapped slab pointers (NVLink loads for the
// support planes that belong to the neighbouring slabs) and the result row is pushed to every rank's copy of the
// output (peer-mapped stores), so that all ranks hold the full result after the closing barrier.
constexpr int kMaxPeers = 8;
template <class T> struct PeerTable { T *p[kMaxPeers]; };

template <class T, int S>
__global__ void __launch_bounds__(128)
ibmGatherSortedDist(const int4 *__restrict__ sortedOrigin, const T *__restrict__ sortedW,
                    const int *__restrict__ sortedIndex, const uint32_t *__restrict__ binStart, GridT<T> g, int nxPad,
                    PeerTable<T> slabs, int z0, int nzl, int world, PeerTable<T> outs) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= (int)binStart[g.n[0] * g.n[1] * g.zwinN]) return;
  const int4 og = sortedOrigin[slot];
  if (og.w < z0 || og.w >= z0 + nzl) return; // a neighbour's particle
  const int o[3] = {og.x, og.y, og.z};
  T w[3][S];
  int cidx[3][S];
  const T *wsrc = sortedW + (size_t)slot * (3 * S);
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
  for (int kk = 0; kk < S; kk++) {
    if (cidx[2][kk] < 0) continue;
    const int owner = cidx[2][kk] / nzl;
    const T *plane = slabs.p[owner] + 3 * (size_t)nxPad * g.n[1] * (size_t)(cidx[2][kk] - owner * nzl);
#pragma unroll
    for (int jj = 0; jj < S; jj++)
#pragma unroll
      for (int ii = 0; ii < S; ii++) {
        if (cidx[0][ii] < 0 || cidx[1][jj] < 0) continue;
        const T *gp = plane + 3 * ((size_t)cidx[0][ii] + (size_t)nxPad * (size_t)cidx[1][jj]);
        const T wx = w[0][ii], wy = w[1][jj], wz = w[2][kk];
        ax += g.cellVolume * (gp[0] * wx * wy * wz);
        ay += g.cellVolume * (gp[1] * wx * wy * wz);
        az += g.cellVolume * (gp[2] * wx * wy * wz);
      }
  }
  const size_t row = 3 * (size_t)sortedIndex[slot];
  for (int r = 0; r < world; r++) {
    T *op = outs.p[r] + row;
    op[0] = ax; op[1] = ay; op[2] = az;
  }
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
