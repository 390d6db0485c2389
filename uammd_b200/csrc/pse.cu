// Positively Split Ewald RPY hydrodynamics (BDHI::PSE) behind the C ABI, sm_100a.
//
// Far field  (PSE/FarField.cuh:535-553): Gaussian spread -> FFT x, y -> fused [FFT z, Hasimoto-split RPY Green's
//            function x projector, Brownian noise, inverse FFT z] -> inverse FFT y, x -> gather (accumulating into MF).
//            Same kernels as FCM; only the spectral operator (greensFunction :85-119, fourierBrownianNoise :235-308)
//            and the window (pse_ns::Kernel :25-41, support 2P+1) differ. One grid buffer, in place, no per-call
//            allocation (the reference re-allocates both grids and the cuFFT work area from its pool every call).
// Near field (PSE/NearField.cuh:120-196,243-282): RPY mat-vec over our cell list with the tabulated F(r), G(r)
//            (RPY_PSE.cuh:45-128, TabulatedFunction.cuh:63-158), warp per home cell with the candidates staged in
//            shared memory (positions AND the vector v), sheared minimum image exactly as the reference computes it.
// Near-field noise (NearField.cuh:254-282): Lanczos sqrt(M) z (misc/LanczosAlgorithm/LanczosAlgorithm.cu) with our
//            own reduction kernels and a host QL eigen-solver for the small tridiagonal matrix (no cuBLAS / LAPACKE).
#include "fft3d.cuh"
#include "ibm_state.cuh"
#include "pse_op.cuh"
#include "pair_common.cuh"
#include "saru.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace ub200 {

// ------------------------------------------------------------------------------------------------------------
// Near field
// ------------------------------------------------------------------------------------------------------------
template <class T> struct TableView { // TabulatedFunction<real2, LinearInterpolation> (misc/TabulatedFunction.cuh:78-158)
  const typename Vec2<T>::type *table;
  int Ntable;
  T rmin, rmax, interval, dr;
};

template <class T> __device__ __forceinline__ typename Vec2<T>::type tableGet(const TableView<T> &tb, T rs) {
  using C = typename Vec2<T>::type;
  const T r = (rs - tb.rmin) * tb.interval;
  if (rs >= tb.rmax) return mk2<T>(T(0), T(0));
  if (r <= T(0)) return __ldg(tb.table);
  const int i = (int)(r * tb.Ntable);
  const T r0 = i * tb.dr;
  const C v0 = __ldg(tb.table + i), v1 = __ldg(tb.table + i + 1);
  const T t = (r - r0) * (T)tb.Ntable;
  return mk2<T>(fma(t, v1.x, fma(-t, v0.x, v0.x)), fma(t, v1.y, fma(-t, v0.y, v0.y)));
}

template <class T> struct NearGeom {
  T Lx, Ly, Lz, iLx, iLy, iLz, shear, rcut2;
};

// RPYNearTransverser::compute (NearField.cuh:131-182). The image shifts use multiplications by 1/L where the
// reference divides: the two can only disagree when a separation component is within an ulp of L/2, far beyond the
// cut-off (rcut <= L/2 is enforced at construction), where both images are rejected.
template <class T, bool SHEAR>
__device__ __forceinline__ void rpyPair(const NearGeom<T> &q, const TableView<T> &tb, T pix, T piy, T piz, T pjx, T pjy,
                                        T pjz, T vjx, T vjy, T vjz, T &ax, T &ay, T &az) {
  T rx = pjx - pix, ry = pjy - piy, rz = pjz - piz;
  // rint (one instruction) where the reference calls round(): they differ only at exact half-integers = L/2 apart
  const T s1 = rint(ry * q.iLy);
  if (SHEAR) { rx += q.shear * ry; rx -= q.shear * q.Ly * s1; }
  ry -= q.Ly * s1;
  rz -= q.Lz * rint(rz * q.iLz);
  rx -= q.Lx * rint(rx * q.iLx);
  const T r2 = rx * rx + ry * ry + rz * rz;
  if (r2 >= q.rcut2) return;
  const auto fg = tableGet(tb, sqrt(r2));
  const T f = fg.x, g = fg.y;
  if (r2 == T(0)) { ax += f * vjx; ay += f * vjy; az += f * vjz; return; }
  const T invr2 = T(1.0) / r2;
  const T gmfv = (g - f) * (rx * vjx + ry * vjy + rz * vjz) * invr2;
  ax += f * vjx + gmfv * rx;
  ay += f * vjy + gmfv * ry;
  az += f * vjz + gmfv * rz;
}

// sorted copies: position (xyz) and the vector v (xyz) of the particle in sorted slot k
template <class T4, class T>
__global__ void __launch_bounds__(256)
pseGatherSorted(const int *__restrict__ groupIndex, const T4 *__restrict__ pos, const T *__restrict__ v, int vStride, int N,
                T *__restrict__ sortedPos3, T *__restrict__ sortedV3) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = groupIndex[k];
  if (sortedPos3) {
    const T4 p = pos[i];
    sortedPos3[3 * (size_t)k] = p.x; sortedPos3[3 * (size_t)k + 1] = p.y; sortedPos3[3 * (size_t)k + 2] = p.z;
  }
  if (sortedV3) {
    const T *vp = v + (size_t)i * vStride;
    sortedV3[3 * (size_t)k] = vp[0]; sortedV3[3 * (size_t)k + 1] = vp[1]; sortedV3[3 * (size_t)k + 2] = vp[2];
  }
}

__global__ void __launch_bounds__(256) pseToFloat4(const double4 *__restrict__ in, float4 *__restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double4 p = in[i];
  out[i] = make_float4((float)p.x, (float)p.y, (float)p.z, 0.f);
}

constexpr int kNearCap = 224; // staged candidates per warp

// One warp per home cell: the (up to) 27 neighbour cells are staged once per cell into two flat shared arrays
// (positions and v, 3 reals per candidate, stride 3 words: conflict free), then the home particles are served TWO at a
// time by the 32 lanes striding over the flat candidate list (each staged candidate feeds two pair evaluations) and a
// butterfly reduction. Replaces transverseWithNeighbourContainer (NeighbourList/common.cuh:10-34) for the
// RPYNearTransverser. ACCUMULATE: Mv[i] += (Transverser::set, NearField.cuh:184-186); otherwise Mv[i] = (the
// Dotctor's fill + traverse, :212-218, in one pass).
template <class T, bool ACCUMULATE, bool SHEAR>
__global__ void __launch_bounds__(kPairThreads)
rpyNearTraversal(const T *__restrict__ sortedPos3, const T *__restrict__ sortedV3, const int *__restrict__ groupIndex,
                 const uint32_t *__restrict__ binStart, GridF g, int ncells, TableView<T> tb, NearGeom<T> q,
                 T *__restrict__ Mv3) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *candP = reinterpret_cast<T *>(smemRaw) + (size_t)warp * kNearCap * 6;
  T *candV = candP + (size_t)kNearCap * 3;
  const int warpsTotal = gridDim.x * kPairWarps;
  for (int cell = blockIdx.x * kPairWarps + warp; cell < ncells; cell += warpsTotal) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue;
    const bool staged = nc.total <= kNearCap;
    __syncwarp();
    if (staged) {
      for (int c = 0; c < 27; c++) {
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        if (cnt == 0) continue;
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        const int off = __shfl_sync(0xffffffffu, nc.off, c);
        for (int t = lane; t < 3 * cnt; t += 32) {
          candP[3 * off + t] = sortedPos3[3 * (size_t)st + t];
          candV[3 * off + t] = sortedV3[3 * (size_t)st + t];
        }
      }
    }
    __syncwarp();
    for (int h = 0; h < hCount; h += 2) {
      const bool two = h + 1 < hCount;
      const size_t i0 = hStart + h, i1 = hStart + h + (two ? 1 : 0);
      const T p0x = sortedPos3[3 * i0], p0y = sortedPos3[3 * i0 + 1], p0z = sortedPos3[3 * i0 + 2];
      const T p1x = sortedPos3[3 * i1], p1y = sortedPos3[3 * i1 + 1], p1z = sortedPos3[3 * i1 + 2];
      T a0x = T(0), a0y = T(0), a0z = T(0), a1x = T(0), a1y = T(0), a1z = T(0);
      if (staged) {
        for (int t = lane; t < nc.total; t += 32) {
          const T pjx = candP[3 * t], pjy = candP[3 * t + 1], pjz = candP[3 * t + 2];
          const T vjx = candV[3 * t], vjy = candV[3 * t + 1], vjz = candV[3 * t + 2];
          rpyPair<T, SHEAR>(q, tb, p0x, p0y, p0z, pjx, pjy, pjz, vjx, vjy, vjz, a0x, a0y, a0z);
          rpyPair<T, SHEAR>(q, tb, p1x, p1y, p1z, pjx, pjy, pjz, vjx, vjy, vjz, a1x, a1y, a1z);
        }
      } else {
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          if (cnt == 0) continue;
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          for (int t = lane; t < cnt; t += 32) {
            const T *pp = sortedPos3 + 3 * (size_t)(st + t), *vp = sortedV3 + 3 * (size_t)(st + t);
            rpyPair<T, SHEAR>(q, tb, p0x, p0y, p0z, pp[0], pp[1], pp[2], vp[0], vp[1], vp[2], a0x, a0y, a0z);
            rpyPair<T, SHEAR>(q, tb, p1x, p1y, p1z, pp[0], pp[1], pp[2], vp[0], vp[1], vp[2], a1x, a1y, a1z);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0x += __shfl_xor_sync(0xffffffffu, a0x, o); a0y += __shfl_xor_sync(0xffffffffu, a0y, o); a0z += __shfl_xor_sync(0xffffffffu, a0z, o);
        a1x += __shfl_xor_sync(0xffffffffu, a1x, o); a1y += __shfl_xor_sync(0xffffffffu, a1y, o); a1z += __shfl_xor_sync(0xffffffffu, a1z, o);
      }
      if (lane == 0 || (lane == 1 && two)) {
        T *out = Mv3 + 3 * (size_t)groupIndex[lane ? i1 : i0];
        const T rx = lane ? a1x : a0x, ry = lane ? a1y : a0y, rz = lane ? a1z : a0z;
        if (ACCUMULATE) { out[0] += rx; out[1] += ry; out[2] += rz; }
        else { out[0] = rx; out[1] = ry; out[2] = rz; }
      }
    }
  }
}

// ---- list-based near-field mat-vec (Lanczos): positions are fixed over the 5-30 products of one square-root, so the
// pairs inside the cut-off are found ONCE (our Verlet-list build with multiplier 1, 146 candidates -> ~22 neighbours
// per particle at config 3) and every product only walks its neighbours. The reference walks the 27 cells again for
// every product (Dotctor, NearField.cuh:201-220). Sorted {x,y,z,vx,vy,vz,-,-} records: one 32-byte sector per
// neighbour in fp32.
template <class T4, class T>
__global__ void __launch_bounds__(256)
pseGatherPV(const int *__restrict__ groupIndex, const T4 *__restrict__ pos, const T *__restrict__ v, int vStride, int N,
            T *__restrict__ pv8) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = groupIndex[k];
  T *o = pv8 + 8 * (size_t)k;
  if (pos) { const T4 p = pos[i]; o[0] = p.x; o[1] = p.y; o[2] = p.z; }
  if (v) { const T *vp = v + (size_t)i * vStride; o[3] = vp[0]; o[4] = vp[1]; o[5] = vp[2]; }
}

__device__ __forceinline__ float4 ldgV4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ double4 ldgV4(const double *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// one row of the near-field product: the neighbours of sorted particle id, in list order
template <class T, bool SHEAR>
__device__ __forceinline__ void rpyListRow(const T *__restrict__ pv8, const int *__restrict__ neighbourList,
                                           const int *__restrict__ numberNeighbours, int N, int id, const TableView<T> &tb,
                                           const NearGeom<T> &q, T &ax, T &ay, T &az) {
  using V4 = typename Real4<T>::type;
  const V4 a = ldgV4(pv8 + 8 * (size_t)id);
  const int nn = numberNeighbours[id];
  const int *lp = neighbourList + id;
  int k = 0;
  for (; k + 2 <= nn; k += 2) {
    const int j0 = __ldg(lp + (size_t)k * N), j1 = __ldg(lp + (size_t)(k + 1) * N);
    const V4 b0 = ldgV4(pv8 + 8 * (size_t)j0), c0 = ldgV4(pv8 + 8 * (size_t)j0 + 4);
    const V4 b1 = ldgV4(pv8 + 8 * (size_t)j1), c1 = ldgV4(pv8 + 8 * (size_t)j1 + 4);
    rpyPair<T, SHEAR>(q, tb, a.x, a.y, a.z, b0.x, b0.y, b0.z, b0.w, c0.x, c0.y, ax, ay, az);
    rpyPair<T, SHEAR>(q, tb, a.x, a.y, a.z, b1.x, b1.y, b1.z, b1.w, c1.x, c1.y, ax, ay, az);
  }
  for (; k < nn; k++) {
    const int j0 = __ldg(lp + (size_t)k * N);
    const V4 b0 = ldgV4(pv8 + 8 * (size_t)j0), c0 = ldgV4(pv8 + 8 * (size_t)j0 + 4);
    rpyPair<T, SHEAR>(q, tb, a.x, a.y, a.z, b0.x, b0.y, b0.z, b0.w, c0.x, c0.y, ax, ay, az);
  }
}

template <class T, bool ACCUMULATE, bool SHEAR>
__global__ void __launch_bounds__(128)
rpyNearList(const T *__restrict__ pv8, const int *__restrict__ groupIndex, const int *__restrict__ neighbourList,
            const int *__restrict__ numberNeighbours, int N, TableView<T> tb, NearGeom<T> q, T *__restrict__ Mv3) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  T ax = T(0), ay = T(0), az = T(0);
  rpyListRow<T, SHEAR>(pv8, neighbourList, numberNeighbours, N, id, tb, q, ax, ay, az);
  T *out = Mv3 + 3 * (size_t)groupIndex[id];
  if (ACCUMULATE) { out[0] += ax; out[1] += ay; out[2] += az; }
  else { out[0] = ax; out[1] = ay; out[2] = az; }
}

// ---- near field over ranks: rank r owns the rows [lo, hi) of the SORTED order (contiguous, spatially compact); the
// positions are replicated, the vector of the product lives in the v half of the pv8 records of every rank (peer
// stores over NVLink, distPublishV), the result rows in a slice indexed by k - lo
template <class T, bool SHEAR>
__global__ void __launch_bounds__(128)
rpyNearListRows(const T *__restrict__ pv8, const int *__restrict__ neighbourList, const int *__restrict__ numberNeighbours,
                int N, int lo, int hi, TableView<T> tb, NearGeom<T> q, T *__restrict__ wRows) {
  const int id = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= hi) return;
  T ax = T(0), ay = T(0), az = T(0);
  rpyListRow<T, SHEAR>(pv8, neighbourList, numberNeighbours, N, id, tb, q, ax, ay, az);
  T *out = wRows + 3 * (size_t)(id - lo);
  out[0] = ax; out[1] = ay; out[2] = az;
}
// vNext = alpha x on the owned rows, written to the local Krylov vector and into the records of every rank
template <class T>
__global__ void __launch_bounds__(256)
distPublishV(PeerTable<T> pv8, int world, int lo, int hi, T alpha, const T *__restrict__ x, T *__restrict__ vNext) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hi - lo) return;
  const T a = alpha * x[3 * (size_t)r], b = alpha * x[3 * (size_t)r + 1], c = alpha * x[3 * (size_t)r + 2];
  vNext[3 * (size_t)r] = a; vNext[3 * (size_t)r + 1] = b; vNext[3 * (size_t)r + 2] = c;
  for (int p = 0; p < world; p++) {
    T *o = pv8.p[p] + 8 * (size_t)(lo + r) + 3;
    o[0] = a; o[1] = b; o[2] = c;
  }
}
// result rows (sorted slice) -> the result array (particle order) of every rank
template <class T>
__global__ void __launch_bounds__(256)
distPublishOut(PeerTable<T> out3, int world, int lo, int hi, const int *__restrict__ groupIndex, const T *__restrict__ rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hi - lo) return;
  const size_t i = (size_t)groupIndex[lo + r];
  const T a = rows[3 * (size_t)r], b = rows[3 * (size_t)r + 1], c = rows[3 * (size_t)r + 2];
  for (int p = 0; p < world; p++) {
    T *o = out3.p[p] + 3 * i;
    o[0] = a; o[1] = b; o[2] = c;
  }
}
// SaruTransform (NearField.cuh:222-232) of the owned rows: the stream of a particle is keyed by its index, not by its row
template <class T>
__global__ void __launch_bounds__(256)
distNoiseRows(T *__restrict__ zRows, const int *__restrict__ groupIndex, int lo, int hi, T variance, uint32_t seed1, uint32_t seed2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hi - lo) return;
  Saru rng((uint32_t)groupIndex[lo + r], seed1, seed2);
  const float2 a = rng.gauss2(1.0f), b = rng.gauss2(1.0f);
  zRows[3 * (size_t)r] = (T)a.x * variance; zRows[3 * (size_t)r + 1] = (T)a.y * variance; zRows[3 * (size_t)r + 2] = (T)b.x * variance;
}

// One warp: all-reduce of one double over the ranks fused with a barrier (value == nullptr: barrier only).
// arena of rank r: flags[p] = last epoch rank p announced to r | slots[row][p] = the addend of rank p. Lane p stores this
// rank's addend into slot [row][rank] of rank p, announces the epoch with a system-scope release, and spins (bounded)
// on what rank p announced here; lane 0 then adds the slots in rank order, so every rank obtains the same bits.
constexpr unsigned long long kNearWaitNs = 10000000000ull; // 10 s: a lost peer must not hang the GPU
constexpr int kNearSlotRows = 4;
__device__ __forceinline__ unsigned long long globalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(32)
nearPeerReduce(PeerTable<char> arena, size_t slotsOff, int row, int rank, int world, uint32_t epoch, const double *value,
               double *out, int *err) { // value may alias out
  const int p = threadIdx.x;
  if (*reinterpret_cast<volatile int *>(err) != 0) { // a barrier timed out before: no more waiting, the host sees NaN and stops
    if (p == 0 && value) *out = nan("");
    return;
  }
  if (p < world) {
    if (value) {
      volatile double *slot = reinterpret_cast<double *>(arena.p[p] + slotsOff) + row * kMaxPeers + rank;
      *slot = *value;
    }
    __threadfence_system();
    uint32_t *remote = reinterpret_cast<uint32_t *>(arena.p[p]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(arena.p[rank]) + p;
    const unsigned long long t0 = globalTimerNs();
    unsigned spins = 0;
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if ((++spins & 1023u) == 0 && globalTimerNs() - t0 > kNearWaitNs) { *err = (int)epoch; break; } // which barrier it was
    }
    __threadfence_system();
  }
  __syncwarp();
  if (p == 0 && value) {
    const volatile double *slots = reinterpret_cast<const double *>(arena.p[rank] + slotsOff) + row * kMaxPeers;
    double s = 0;
    for (int q = 0; q < world; q++) s += slots[q];
    *out = s;
  }
}
template <class T>
__global__ void __launch_bounds__(256) vecAdd(const T *__restrict__ x, T *__restrict__ y, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// ------------------------------------------------------------------------------------------------------------
// Lanczos helpers (vectors of n reals on the device, scalars through a small pinned-free host read)
// ------------------------------------------------------------------------------------------------------------
constexpr int kRedBlocks = 592, kRedThreads = 256;

template <class T>
__global__ void __launch_bounds__(kRedThreads) redDotPartial(const T *__restrict__ a, const T *__restrict__ b, size_t n,
                                                             double *__restrict__ partial) {
  __shared__ double sh[kRedThreads / 32];
  double acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (double)a[i] * (double)b[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < kRedThreads / 32; w++) s += sh[w];
    partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(kRedThreads) redFinal(const double *__restrict__ partial, int nb, double *__restrict__ out) {
  __shared__ double sh[kRedThreads / 32];
  double acc = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < kRedThreads / 32; w++) s += sh[w];
    *out = s;
  }
}
// y = alpha x + beta y (beta == 0: y is not read)
template <class T>
__global__ void __launch_bounds__(256) vecAxpby(T alpha, const T *__restrict__ x, T beta, T *__restrict__ y, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = beta == T(0) ? alpha * x[i] : alpha * x[i] + beta * y[i];
}
// out = scale * V[:, 0:m] c   (V column major, leading dimension n), the coefficients as a kernel argument (no host-to-device
// copy on the stream)
constexpr int kLanczosMaxSteps = 200;
template <class T> struct LanczosCoeff { T c[kLanczosMaxSteps]; };
template <class T>
__global__ void __launch_bounds__(256) vecGemvArg(const T *__restrict__ V, size_t n, int m, const LanczosCoeff<T> c, T scale,
                                                  T *__restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  T acc = T(0);
  for (int j = 0; j < m; j++) acc += V[i + n * (size_t)j] * c.c[j];
  out[i] = scale * acc;
}
// y = 0, or the first unit vector
template <class T>
__global__ void __launch_bounds__(256) vecFillUnit(T *__restrict__ y, size_t n, bool unit) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = (unit && i == 0) ? T(1) : T(0);
}
// SaruTransform (NearField.cuh:222-232)
template <class T>
__global__ void __launch_bounds__(256) pseNearNoise(T *__restrict__ z3, int N, T variance, uint32_t seed1, uint32_t seed2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  Saru rng((uint32_t)i, seed1, seed2);
  const float2 a = rng.gauss2(1.0f), b = rng.gauss2(1.0f);
  z3[3 * (size_t)i] = (T)a.x * variance; z3[3 * (size_t)i + 1] = (T)a.y * variance; z3[3 * (size_t)i + 2] = (T)b.x * variance;
}

// eigen-decomposition of a symmetric tridiagonal matrix (implicit QL, double precision): d[n] diagonal ->
// eigenvalues, e[n-1] sub-diagonal, Z (column major n x n) -> eigenvectors. Stands in for LAPACKE_steqr('I').
inline bool tridiagEigen(int n, std::vector<double> &d, std::vector<double> e, std::vector<double> &Z) {
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Z[(size_t)i * n + i] = 1.0;
  e.resize(n, 0.0);
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; k < n; k++) { // rotate eigenvector columns i and i+1
            f = Z[(size_t)(i + 1) * n + k];
            Z[(size_t)(i + 1) * n + k] = s * Z[(size_t)i * n + k] + c * f;
            Z[(size_t)i * n + k] = c * Z[(size_t)i * n + k] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

struct GrowBuf { // device buffer that keeps its contents when it grows
  void *p = nullptr;
  size_t cap = 0;
  int grow(size_t bytes, cudaStream_t st) {
    if (bytes <= cap) return UB200_OK;
    size_t want = std::max(bytes, 2 * cap);
    void *np = nullptr;
    if (cudaMalloc(&np, want) != cudaSuccess) return UB200_ERR_ALLOC;
    if (p) {
      cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
      cudaStreamSynchronize(st);
      cudaFree(p);
    }
    p = np; cap = want;
    return UB200_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// nextFFTWiseSize3D (utils/Grid.cuh:142-213), one dimension: smallest 2^i 3^j 5^k 7^l 11^m >= n with i >= 1,
// k <= 5, l <= 4, m <= 3 (the reference's "forbidden sizes" loop has no effect: its `continue` only leaves the inner loop)
inline int nextFFTWiseSize(int n) {
  long long best = -1;
  for (long long p11 = 1, m = 0; m <= 3; m++, p11 *= 11)
    for (long long p7 = 1, l = 0; l <= 4; l++, p7 *= 7)
      for (long long p5 = 1, k = 0; k <= 5; k++, p5 *= 5)
        for (long long p3 = 1; p3 * p5 * p7 * p11 <= (1LL << 40); p3 *= 3)
          for (long long p2 = 2; p2 * p3 * p5 * p7 * p11 <= (1LL << 40); p2 *= 2) {
            const long long v = p2 * p3 * p5 * p7 * p11;
            if (v >= n && (best < 0 || v < best)) best = v;
            if (v >= n) break;
          }
  return best >= (1LL << 31) ? -1 : (int)best;
}

int nextFFTWiseSizeOf(int n) { return nextFFTWiseSize(n); } // for poisson.cu

// RPYPSE_near::FandG (PSE/RPY_PSE.cuh:45-128): host, double precision
inline void rpyNearFG(double r, double rh, double psi, double rcut, double &F, double &G) {
  if (r >= rcut) { F = G = 0; return; }
  if (r <= 0.0) {
    const double pi = M_PI;
    F = (1.0 / (4 * sqrt(pi) * psi * rh)) * (1 - exp(-4 * rh * rh * psi * psi) + 4 * sqrt(pi) * rh * psi * erfc(2 * rh * psi));
    G = 0;
    return;
  }
  const double r2 = r * r, a2mr = 2 * rh - r, a2pr = 2 * rh + r, rh2 = rh * rh, rh4 = rh2 * rh2, psi2 = psi * psi,
               psi3 = psi2 * psi, psi4 = psi2 * psi2, r3 = r2 * r, r4 = r3 * r, sp = sqrt(M_PI);
  double f0, f1, f2, f3, f4, f5, f6, f7, g0, g1, g2, g3, g4, g5, g6, g7;
  if (r > 2 * rh) {
    f0 = (64.0 * rh4 * psi4 + 96.0 * rh2 * r2 * psi4 - 128.0 * rh * r3 * psi4 + 36.0 * r4 * psi4 - 3.0) / (128.0 * rh * r3 * psi4);
    f4 = (3.0 - 4.0 * psi4 * a2mr * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2)) / (256.0 * rh * r3 * psi4);
    f5 = 0;
    g0 = (-64.0 * rh4 * psi4 + 96.0 * rh2 * r2 * psi4 - 64.0 * rh * r3 * psi4 + 12.0 * r4 * psi4 + 3.0) / (64.0 * rh * r3 * psi4);
    g4 = (4.0 * psi4 * a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r) - 3.0) / (128.0 * rh * r3 * psi4);
    g5 = 0;
  } else {
    f0 = (-16.0 * rh4 - 24.0 * rh2 * r2 + 32.0 * rh * r3 - 9.0 * r4) / (32.0 * rh * r3);
    f4 = 0;
    f5 = (4.0 * psi4 * a2mr * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2) - 3.0) / (256.0 * rh * r3 * psi4);
    g0 = a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r) / (16.0 * rh * r3);
    g4 = 0;
    g5 = (3.0 - 4.0 * psi4 * a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r)) / (128.0 * rh * r3 * psi4);
  }
  f1 = (-2.0 * psi2 * a2pr * (4.0 * rh2 - 4.0 * rh * r + 9.0 * r2) + 2.0 * rh - 3.0 * r) / (128.0 * rh * r3 * psi3 * sp);
  f2 = (2.0 * psi2 * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2) - 2.0 * rh - 3.0 * r) / (128.0 * rh * r3 * psi3 * sp);
  f3 = 3.0 * (6.0 * r2 * psi2 + 1.0) / (64.0 * sp * rh * r2 * psi3);
  f6 = (4.0 * psi4 * a2pr * a2pr * (4.0 * rh2 - 4.0 * rh * r + 9.0 * r2) - 3.0) / (256.0 * rh * r3 * psi4);
  f7 = 3.0 * (1.0 - 12.0 * r4 * psi4) / (128.0 * rh * r3 * psi4);
  g1 = (2.0 * psi2 * a2pr * a2pr * (2.0 * rh - 3.0 * r) - 2.0 * rh + 3.0 * r) / (64.0 * sp * rh * r3 * psi3);
  g2 = (-2.0 * psi2 * a2mr * a2mr * (2.0 * rh + 3.0 * r) + 2.0 * rh + 3.0 * r) / (64.0 * sp * rh * r3 * psi3);
  g3 = (3.0 * (2.0 * r2 * psi2 - 1.0)) / (32.0 * sp * rh * r2 * psi3);
  g6 = (3.0 - 4.0 * psi4 * (2.0 * rh - 3.0 * r) * a2pr * a2pr * a2pr) / (128.0 * rh * r3 * psi4);
  g7 = -3.0 * (4.0 * r4 * psi4 + 1.0) / (64.0 * rh * r3 * psi4);
  auto comb = [&](double c0, double c1, double c2, double c3, double c4, double c5, double c6, double c7) {
    return c0 + c1 * exp(-psi2 * a2pr * a2pr) + c2 * exp(-a2mr * a2mr * psi2) + c3 * exp(-psi2 * r2) + c4 * erfc(a2mr * psi) +
           c5 * erfc(-a2mr * psi) + c6 * erfc(a2pr * psi) + c7 * erfc(r * psi);
  };
  F = comb(f0, f1, f2, f3, f4, f5, f6, f7);
  G = comb(g0, g1, g2, g3, g4, g5, g6, g7);
}

template <class T> struct PseState {
  using C = typename Vec2<T>::type;
  using T4 = typename Real4<T>::type;
  // parameters (stored in `real` like the reference's members)
  T Lb[3];
  T viscosity, rh, psi, tolerance, shear;
  uint32_t seedNear = 0, seedFar = 0;
  // far field
  Fft3dPlan<T> plan;
  IbmState<T> ibm;
  DevBuf grid;
  T eta = 0;
  int support = 0;
  ub200_ibm_kernel farKernel;
  // near field
  T rcut = 0;
  DevBuf table;
  int nTable = 0;
  ub200_celllist *cl = nullptr;
  DevBuf posF, sortedPos, sortedV;
  int nearN = -1;
  ub200_verletlist *vl = nullptr; // neighbours inside the cut-off, built once per Lanczos square root
  DevBuf sortedPV;
  T *pv8 = nullptr; // the records the list products read: sortedPV, or the arena of the rank decomposition
  int listN = -1;
  const void *listPos = nullptr; // positions the list was built from; set when a caller may reuse it (near_mdot_list)
  // Lanczos
  GrowBuf V;
  DevBuf w, oldBz, z, partial, scalar;
  int checkConvergenceSteps = 3, iterationHardLimit = kLanczosMaxSteps, lastRunRequiredSteps = 0;

  int init(const ub200_pse_params &par, uint32_t seedNear_, uint32_t seedFar_) {
    for (int d = 0; d < 3; d++) Lb[d] = (T)par.L[d];
    viscosity = (T)par.viscosity; rh = (T)par.hydrodynamicRadius; psi = (T)par.psi; tolerance = (T)par.tolerance;
    shear = (T)par.shearStrain;
    seedNear = seedNear_; seedFar = seedFar_;
    if (Lb[0] <= 0 || Lb[1] <= 0 || Lb[2] <= 0 || par.tolerance <= 0 || par.tolerance > 0.1 || par.psi <= 0) return UB200_ERR_INVALID_ARGUMENT;
    int rc;
    // ---- near field: NearField::initializeDeterministicPart (NearField.cuh:65-102) ----
    const double split = psi;
    rcut = (T)(sqrt(-log((double)tolerance)) / split);
    if (0.5 * Lb[0] < rcut) return UB200_ERR_INVALID_ARGUMENT; // "Cut off is too large, try increasing psi"
    const double a = rh;
    const T textureTolerance = (T)(a * tolerance);
    const unsigned maximumTextureElements = 1u << 22;
    unsigned nPointsTable = (unsigned)std::min((double)(rcut / textureTolerance + 0.5), 4e9);
    nPointsTable = std::min(maximumTextureElements, std::max(1u << 14, nPointsTable));
    nTable = (int)nPointsTable;
    {
      // RPYPSE_near(rh, psi, 6 pi a vis, rcut) takes `real` arguments; TabulatedFunction(table, N = nPointsTable, 0, rcut)
      const double rhd = (double)(T)a, psid = (double)(T)split, norm = (double)(T)(6 * M_PI * a * viscosity), rcd = (double)rcut;
      const int Ntable = nTable - 1;
      std::vector<C> h((size_t)Ntable + 2);
      for (int i = 0; i <= Ntable; i++) {
        const double x = (i / (double)Ntable) * ((double)rcut - 0.0) + 0.0;
        double F, G;
        rpyNearFG(x, rhd, psid, rcd, F, G);
        h[i] = mk2<T>((T)(F / norm), (T)(G / norm));
      }
      h[Ntable + 1] = mk2<T>(T(0), T(0));
      if ((rc = table.reserve(sizeof(C) * h.size()))) return rc;
      UB200_CUDA(cudaMemcpy(table.p, h.data(), sizeof(C) * h.size(), cudaMemcpyHostToDevice));
    }
    if ((rc = ub200_celllist_create(&cl))) return rc;
    if ((rc = ub200_verletlist_create(&vl))) return rc;
    vl->refOnly = true; // rpyNearList walks the reference-layout arrays
    if ((rc = ub200_verletlist_set_cutoff_multiplier(vl, 1.0f))) return rc;
    // ---- far field: FarField::initializeGrid / initializeKernel (FarField.cuh:605-654) ----
    const T kcut = T(2) * psi * (T)sqrt(-log(tolerance));
    const double hgrid = 2 * M_PI / kcut;
    int cells[3];
    for (int d = 0; d < 3; d++) {
      cells[d] = (int)(T(2) * Lb[d] / (T)hgrid) + 1;
      cells[d] = nextFFTWiseSize(cells[d]);
      if (cells[d] < 0) return UB200_ERR_GRID_TOO_LARGE;
    }
    if (par.cellsOverride[0] > 0) for (int d = 0; d < 3; d++) cells[d] = par.cellsOverride[d];
    const double C0 = 0.976;
    double m = 1;
    while (erfc(m / sqrt(2)) > 0.1 * tolerance) m += 0.01;
    int sup;
    while ((sup = int(pow(m / C0, 2) / M_PI + 0.5) + 1) % 2 == 0) m += tolerance;
    int P = sup / 2;
    const int minCellDim = std::min({cells[0], cells[1], cells[2]});
    if (sup > minCellDim) {
      sup = minCellDim;
      if (sup % 2 == 0) sup--;
      P = sup / 2;
      m = C0 * sqrt(M_PI * sup);
    }
    const double pw = 2 * P + 1;
    const T cs[3] = {Lb[0] / (T)cells[0], Lb[1] / (T)cells[1], Lb[2] / (T)cells[2]};
    const double h = std::min({cs[0], cs[1], cs[2]});
    const double wgauss = pw * h / 2.0;
    eta = (T)pow(2.0 * psi * wgauss / m, 2);
    support = 2 * P + 1;
    if (support > kMaxSupport) return UB200_ERR_UNSUPPORTED;
    // pse_ns::Kernel(P, width = sqrt(eta)/(2 psi)) (FarField.cuh:25-41), members stored as real
    const T width = (T)(sqrt(eta) / (2.0 * psi));
    ub200_ibm_kernel k;
    k.kind = UB200_KERNEL_GAUSSIAN;
    k.support = support;
    k.h = h;
    k.prefactor = (double)(T)cbrt(1.0 / (width * width * width * pow(2.0 * M_PI, 1.5)));
    k.tau = (double)(T)(-0.5 / (width * width));
    k.rmax = INFINITY; // the PSE window has no cut-off inside its support
    farKernel = k;
    if ((rc = plan.init(cells[0], cells[1], cells[2]))) return rc;
    const int periodic[3] = {1, 1, 1};
    const double Ld[3] = {par.L[0], par.L[1], par.L[2]};
    if ((rc = ibm.init(Ld, periodic, cells, k, plan.nxPad))) return rc;
    if ((rc = grid.reserve(plan.gridBytes()))) return rc;
    return UB200_OK;
  }
  void release() {
    plan.release(); ibm.release(); grid.release(); table.release(); posF.release(); sortedPos.release(); sortedV.release();
    V.release(); w.release(); oldBz.release(); z.release(); partial.release(); scalar.release();
    if (cl) ub200_celllist_destroy(cl);
    cl = nullptr;
    if (vl) ub200_verletlist_destroy(vl);
    vl = nullptr;
    sortedPV.release();
    distRelease();
    if (hScalar) cudaFreeHost(hScalar);
    hScalar = nullptr; hScalarDev = nullptr;
  }

  // ---------------- far field ----------------
  int farMdot(const void *pos, const void *force, int N, double temperature, double prefactor, uint32_t seed2, void *MF3,
              cudaStream_t st) {
    int rc;
    T *g = grid.as<T>();
    const bool det = force != nullptr;
    if (!det && !(temperature > 0)) return UB200_OK; // nothing to add
    if (det) {
      if ((rc = ibm.spread(pos, force, 4, N, g, false, st))) return rc;
      if ((rc = launchPassX<T, true>(plan, g, st))) return rc;
      if ((rc = launchPassY<T, -1>(plan, g, st))) return rc;
    } else {
      UB200_CUDA(cudaMemsetAsync(g, 0, plan.gridBytes(), st));
    }
    PseSpectralOp<T> op;
    op.nx = plan.nx; op.ny = plan.ny; op.nz = plan.nz; op.nkx = plan.nkx;
    op.kfx = T(2.0) * T(M_PI) / Lb[0]; op.kfy = T(2.0) * T(M_PI) / Lb[1]; op.kfz = T(2.0) * T(M_PI) / Lb[2];
    op.shear = shear; op.rh = rh; op.vis = viscosity; op.split = psi; op.eta = eta;
    op.nTot = (T)(plan.nx * plan.ny * plan.nz);
    op.deterministic = det;
    op.noise = temperature > 0;
    op.seed1 = seedFar; op.seed2 = seed2;
    // addBrownianNoise (FarField.cuh:467-492): prefactor * sqrt(2 T / dV)
    op.noisePrefactor = op.noise ? (T)prefactor * (T)sqrt(2 * (T)temperature / ibm.grid.cellVolume) : T(0);
    if ((rc = launchPassZ<T, 0, PseSpectralOp<T>>(plan, g, st, op))) return rc;
    if ((rc = launchPassY<T, +1>(plan, g, st))) return rc;
    if ((rc = launchPassX<T, false>(plan, g, st))) return rc;
    return ibm.gather(pos, N, g, (T *)MF3, true, det, st); // IBM::gather accumulates (misc/IBM.cu:231-233); reuses the spread's sort
  }

  // ---------------- near field ----------------
  // NearField::updateNeighbourList (NearField.cuh:236-241) + sorted copy of the positions
  int nearPrepare(const void *pos, int N, cudaStream_t st) {
    int rc;
    const double gs = shear;
    const T safety = (T)(1 + 0.5 * gs * gs + 0.5 * sqrt(gs * gs * (gs * gs + 4.0))); // cutOffShearedSafetyFactor :24-27
    const float Lf[3] = {(float)Lb[0], (float)Lb[1], (float)Lb[2]};
    int cd[3];
    if ((rc = ub200_neighbour_celldim_f32(Lf, (float)(rcut * safety), cd))) return rc;
    const int periodic[3] = {1, 1, 1};
    const void *posf = pos;
    if (sizeof(T) == 8) {
      if ((rc = posF.reserve(sizeof(float4) * (size_t)N))) return rc;
      pseToFloat4<<<(N + 255) / 256, 256, 0, st>>>((const double4 *)pos, posF.as<float4>(), N);
      UB200_LAUNCHED();
      posf = posF.p;
    }
    if ((rc = ub200_celllist_build_f32(cl, posf, nullptr, N, Lf, periodic, cd, st))) return rc;
    if ((rc = sortedPos.reserve(sizeof(T) * 3 * (size_t)N))) return rc;
    if ((rc = sortedV.reserve(sizeof(T) * 3 * (size_t)N))) return rc;
    pseGatherSorted<T4, T><<<(N + 255) / 256, 256, 0, st>>>(cl->groupIndex.as<int>(), (const T4 *)pos, (const T *)nullptr, 0, N,
                                                           sortedPos.as<T>(), (T *)nullptr);
    UB200_LAUNCHED();
    nearN = N;
    return UB200_OK;
  }
  // Mv (+)= M_near v with the list of the last nearPrepare
  int nearDot(const T *v, int vStride, int N, T *Mv3, bool accumulate, cudaStream_t st) {
    if (nearN != N) return UB200_ERR_NOT_BUILT;
    pseGatherSorted<T4, T><<<(N + 255) / 256, 256, 0, st>>>(cl->groupIndex.as<int>(), (const T4 *)nullptr, v, vStride, N,
                                                           (T *)nullptr, sortedV.as<T>());
    UB200_LAUNCHED();
    TableView<T> tb;
    tb.table = table.as<C>(); tb.Ntable = nTable - 1; tb.rmin = T(0); tb.rmax = rcut;
    tb.interval = (T)(1.0 / (rcut - T(0))); tb.dr = (T)(1.0 / (T)(nTable - 1));
    NearGeom<T> q;
    q.Lx = Lb[0]; q.Ly = Lb[1]; q.Lz = Lb[2]; q.shear = shear; q.rcut2 = rcut * rcut;
    q.iLx = T(1.0) / Lb[0]; q.iLy = T(1.0) / Lb[1]; q.iLz = T(1.0) / Lb[2];
    const size_t smem = (size_t)kPairWarps * kNearCap * 6 * sizeof(T);
    const int needed = (cl->ncells + kPairWarps - 1) / kPairWarps;
    const int gridSize = std::min(needed, kNumSMs * 8);
#define UB200_NEAR(ACC, SH)                                                                                               \
  rpyNearTraversal<T, ACC, SH><<<gridSize, kPairThreads, smem, st>>>(sortedPos.as<T>(), sortedV.as<T>(), cl->groupIndex.as<int>(), \
                                                                    cl->binStart.as<uint32_t>(), cl->grid, cl->ncells, tb, q, Mv3)
    const bool sh = shear != T(0);
    if (accumulate) { if (sh) UB200_NEAR(true, true); else UB200_NEAR(true, false); }
    else { if (sh) UB200_NEAR(false, true); else UB200_NEAR(false, false); }
#undef UB200_NEAR
    UB200_LAUNCHED();
    return UB200_OK;
  }

  TableView<T> tableView() const {
    TableView<T> tb;
    tb.table = table.as<C>(); tb.Ntable = nTable - 1; tb.rmin = T(0); tb.rmax = rcut;
    tb.interval = (T)(1.0 / (rcut - T(0))); tb.dr = (T)(1.0 / (T)(nTable - 1));
    return tb;
  }
  NearGeom<T> nearGeom() const {
    NearGeom<T> q;
    q.Lx = Lb[0]; q.Ly = Lb[1]; q.Lz = Lb[2]; q.shear = shear; q.rcut2 = rcut * rcut;
    q.iLx = T(1.0) / Lb[0]; q.iLy = T(1.0) / Lb[1]; q.iLz = T(1.0) / Lb[2];
    return q;
  }
  // neighbour list of the current positions: every pair the exact test of rpyPair can accept is in it (the list is
  // built from fp32 coordinates with the unsheared minimum image, hence the shear safety factor and a rounding margin)
  int nearPrepareList(const void *pos, int N, cudaStream_t st) {
    int rc;
    const double gs = shear;
    const double safety = 1 + 0.5 * gs * gs + 0.5 * sqrt(gs * gs * (gs * gs + 4.0));
    const float Lf[3] = {(float)Lb[0], (float)Lb[1], (float)Lb[2]};
    const float Lmax = std::max({Lf[0], Lf[1], Lf[2]});
    const float rcList = (float)((double)rcut * safety) * (1.0f + 1e-5f) + 16.0f * Lmax * 1.2e-7f;
    const int periodic[3] = {1, 1, 1};
    const void *posf = pos;
    if (sizeof(T) == 8) {
      if ((rc = posF.reserve(sizeof(float4) * (size_t)N))) return rc;
      pseToFloat4<<<(N + 255) / 256, 256, 0, st>>>((const double4 *)pos, posF.as<float4>(), N);
      UB200_LAUNCHED();
      posf = posF.p;
    }
    if ((rc = ub200_verletlist_update_f32(vl, posf, nullptr, N, Lf, periodic, rcList, 1, nullptr, (void *)st))) return rc;
    if (dArena) { // rank decomposition: the records live in the arena the peers store into
      if (N > dMaxN) return UB200_ERR_INVALID_ARGUMENT;
      pv8 = reinterpret_cast<T *>(static_cast<char *>(dArena) + dOffPV);
    } else {
      if ((rc = sortedPV.reserve(sizeof(T) * 8 * (size_t)N))) return rc;
      pv8 = sortedPV.as<T>();
    }
    pseGatherPV<T4, T><<<(N + 255) / 256, 256, 0, st>>>(vl->cl->groupIndex.as<int>(), (const T4 *)pos, (const T *)nullptr, 0, N, pv8);
    UB200_LAUNCHED();
    listN = N;
    return UB200_OK;
  }
  int nearDotList(const T *v, int vStride, int N, T *Mv3, bool accumulate, cudaStream_t st) {
    if (listN != N) return UB200_ERR_NOT_BUILT;
    pseGatherPV<T4, T><<<(N + 255) / 256, 256, 0, st>>>(vl->cl->groupIndex.as<int>(), (const T4 *)nullptr, v, vStride, N, pv8);
    UB200_LAUNCHED();
    const TableView<T> tb = tableView();
    const NearGeom<T> q = nearGeom();
    const int nb = (N + 127) / 128;
#define UB200_NEARL(ACC, SH)                                                                                              \
  rpyNearList<T, ACC, SH><<<nb, 128, 0, st>>>(pv8, vl->cl->groupIndex.as<int>(), vl->neighbourList.as<int>(),               \
                                              vl->numberNeighbours.as<int>(), N, tb, q, Mv3)
    const bool sh = shear != T(0);
    if (accumulate) { if (sh) UB200_NEARL(true, true); else UB200_NEARL(true, false); }
    else { if (sh) UB200_NEARL(false, true); else UB200_NEARL(false, false); }
#undef UB200_NEARL
    UB200_LAUNCHED();
    return UB200_OK;
  }

  // ---------------- Lanczos ----------------
  // a . b on the host; dist: over the rows of every rank (nearPeerReduce, same bits on every rank)
  // the scalars of the iteration are written by the reduction kernels straight into mapped pinned host memory and the
  // coefficients of the estimate travel as a kernel argument: the iteration enqueues kernels only - no copy-engine work,
  // which is shared by all streams of a device and would order the virtual ranks of a one-process test behind each other
  double *hScalar = nullptr, *hScalarDev = nullptr;
  int ensureHost() {
    if (hScalar) return UB200_OK;
    UB200_CUDA(cudaHostAlloc((void **)&hScalar, sizeof(double) * 2, cudaHostAllocMapped | cudaHostAllocPortable));
    UB200_CUDA(cudaHostGetDevicePointer((void **)&hScalarDev, hScalar, 0));
    return UB200_OK;
  }
  int fillUnit(T *y, size_t n, bool unit, cudaStream_t st) {
    if (n == 0) return UB200_OK;
    vecFillUnit<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, n, unit);
    UB200_LAUNCHED();
    return UB200_OK;
  }
  // every buffer of a square root over vectors of n numbers (dist: sized once, before the first barrier)
  int reserveLanczos(size_t n, int basisVectors, cudaStream_t st) {
    const size_t ld = std::max(n, (size_t)1);
    int rc;
    if ((rc = ensureHost())) return rc;
    if ((rc = w.reserve(sizeof(T) * ld)) || (rc = oldBz.reserve(sizeof(T) * ld)) || (rc = partial.reserve(sizeof(double) * kRedBlocks)) ||
        (rc = scalar.reserve(sizeof(double))))
      return rc;
    return V.grow(sizeof(T) * ld * (size_t)basisVectors, st);
  }
  int dotHost(const T *a, const T *b, size_t n, double *out, cudaStream_t st, bool dist = false) {
    redDotPartial<T><<<kRedBlocks, kRedThreads, 0, st>>>(a, b, n, partial.as<double>());
    UB200_LAUNCHED();
    int rc;
    if (dist) {
      redFinal<<<1, kRedThreads, 0, st>>>(partial.as<double>(), kRedBlocks, scalar.as<double>());
      UB200_LAUNCHED();
      if ((rc = peerReduce(scalar.as<double>(), hScalarDev, st))) return rc;
    } else {
      redFinal<<<1, kRedThreads, 0, st>>>(partial.as<double>(), kRedBlocks, hScalarDev);
      UB200_LAUNCHED();
    }
    UB200_CUDA(cudaStreamSynchronize(st));
    *out = *static_cast<volatile double *>(hScalar);
    return UB200_OK;
  }
  int axpby(T alpha, const T *x, T beta, T *y, size_t n, cudaStream_t st) {
    if (n == 0) return UB200_OK;
    vecAxpby<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(alpha, x, beta, y, n);
    UB200_LAUNCHED();
    return UB200_OK;
  }

  // ---- rank decomposition of the near field (ub200_pse_dist_*) ----
  int dRank = 0, dWorld = 1, dMaxN = 0, dLo = 0, dHi = 0;
  void *dArena = nullptr; // flags | reduction slots | pv8 records | result (particle order); exported to the peers
  size_t dOffSlots = 0, dOffPV = 0, dOffOut = 0, dArenaBytes = 0;
  char *dPeer[kMaxPeers] = {};
  bool dOpened[kMaxPeers] = {}, dAttached = false;
  uint32_t dEpoch = 0, dSeq = 0;
  DevBuf dErr, dRows;

  int distCreate(int rank, int world, int maxParticles) {
    if (dArena || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || maxParticles < 1) return UB200_ERR_INVALID_ARGUMENT;
    dRank = rank; dWorld = world; dMaxN = maxParticles;
    dOffSlots = 256;
    dOffPV = dOffSlots + ((sizeof(double) * kNearSlotRows * kMaxPeers + 255) / 256) * 256;
    dOffOut = dOffPV + ((sizeof(T) * 8 * (size_t)maxParticles + 255) / 256) * 256;
    dArenaBytes = dOffOut + sizeof(T) * 3 * (size_t)maxParticles;
    if (cudaMalloc(&dArena, dArenaBytes) != cudaSuccess) { dArena = nullptr; return UB200_ERR_ALLOC; }
    UB200_CUDA(cudaMemset(dArena, 0, dArenaBytes));
    int rc;
    if ((rc = dErr.reserve(sizeof(int)))) return rc;
    UB200_CUDA(cudaMemset(dErr.p, 0, sizeof(int)));
    dPeer[rank] = (char *)dArena;
    dAttached = world == 1;
    return UB200_OK;
  }
  void distRelease() {
    for (int p = 0; p < kMaxPeers; p++)
      if (dOpened[p]) { cudaIpcCloseMemHandle(dPeer[p]); dOpened[p] = false; }
    if (dArena) cudaFree(dArena);
    dArena = nullptr;
    dErr.release(); dRows.release();
  }
  PeerTable<char> peerArenas() const {
    PeerTable<char> t;
    for (int p = 0; p < kMaxPeers; p++) t.p[p] = p < dWorld ? dPeer[p] : nullptr;
    return t;
  }
  template <class U> PeerTable<U> peerAt(size_t off) const {
    PeerTable<U> t;
    for (int p = 0; p < kMaxPeers; p++) t.p[p] = p < dWorld ? reinterpret_cast<U *>(dPeer[p] + off) : nullptr;
    return t;
  }
  // value != nullptr: *out = sum over the ranks of *value; in any case a barrier of the ranks on this stream
  int peerReduce(const double *value, double *out, cudaStream_t st) {
    if (!dAttached) return UB200_ERR_NOT_BUILT;
    nearPeerReduce<<<1, 32, 0, st>>>(peerArenas(), dOffSlots, (int)(dSeq++ % kNearSlotRows), dRank, dWorld, ++dEpoch, value, out,
                                     dErr.as<int>());
    UB200_LAUNCHED();
    return UB200_OK;
  }
  // list of the replicated positions and the rows of this rank (the list build is not decomposed: every rank builds it)
  int distPrepare(const void *pos, int N, cudaStream_t st) {
    if (!dArena) return UB200_ERR_NOT_BUILT;
    int rc;
    if ((rc = nearPrepareList(pos, N, st))) return rc;
    dLo = (int)(((long long)dRank * N) / dWorld);
    dHi = (int)(((long long)(dRank + 1) * N) / dWorld);
    const size_t nloc3 = 3 * (size_t)std::max(dHi - dLo, 1);
    if ((rc = dRows.reserve(sizeof(T) * nloc3)) || (rc = z.reserve(sizeof(T) * nloc3)) || (rc = noiseOut.reserve(sizeof(T) * nloc3))) return rc;
    return reserveLanczos(3 * (size_t)(dHi - dLo), 40, st);
  }
  // w (rows of this rank) = M_near v, v = the v half of the records
  int distDotRows(T *wRows, cudaStream_t st) {
    const int nloc = dHi - dLo;
    if (nloc == 0) return UB200_OK;
    const TableView<T> tb = tableView();
    const NearGeom<T> q = nearGeom();
    const int nb = (nloc + 127) / 128;
    if (shear != T(0))
      rpyNearListRows<T, true><<<nb, 128, 0, st>>>(pv8, vl->neighbourList.as<int>(), vl->numberNeighbours.as<int>(), listN, dLo, dHi, tb, q, wRows);
    else
      rpyNearListRows<T, false><<<nb, 128, 0, st>>>(pv8, vl->neighbourList.as<int>(), vl->numberNeighbours.as<int>(), listN, dLo, dHi, tb, q, wRows);
    UB200_LAUNCHED();
    return UB200_OK;
  }
  int distPublishVec(T alpha, const T *x, T *vNext, cudaStream_t st) {
    const int nloc = dHi - dLo;
    if (nloc > 0) {
      distPublishV<T><<<(nloc + 255) / 256, 256, 0, st>>>(peerAt<T>(dOffPV), dWorld, dLo, dHi, alpha, x, vNext);
      UB200_LAUNCHED();
    }
    return peerReduce(nullptr, nullptr, st); // every rank's rows have landed before the next product reads them
  }
  // out3 += the full vector whose rows (sorted slices) the ranks hold; callers placed a barrier of this call before it
  int distCollect(const T *rows, int N, T *out3, cudaStream_t st) {
    const int nloc = dHi - dLo;
    int rc;
    if (nloc > 0) {
      distPublishOut<T><<<(nloc + 255) / 256, 256, 0, st>>>(peerAt<T>(dOffOut), dWorld, dLo, dHi, vl->cl->groupIndex.as<int>(), rows);
      UB200_LAUNCHED();
    }
    if ((rc = peerReduce(nullptr, nullptr, st))) return rc;
    const size_t n = 3 * (size_t)N;
    vecAdd<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const T *>(static_cast<char *>(dArena) + dOffOut), out3, n);
    UB200_LAUNCHED();
    return UB200_OK;
  }
  // Mv3 += M_near v on every rank (v replicated); distPrepare first
  int distMdot(const T *v, int vStride, int N, T *Mv3, cudaStream_t st) {
    if (!dAttached || listN != N) return UB200_ERR_NOT_BUILT;
    int rc;
    pseGatherPV<T4, T><<<(N + 255) / 256, 256, 0, st>>>(vl->cl->groupIndex.as<int>(), (const T4 *)nullptr, v, vStride, N, pv8);
    UB200_LAUNCHED();
    if ((rc = distDotRows(dRows.as<T>(), st))) return rc;
    if ((rc = peerReduce(nullptr, nullptr, st))) return rc; // the peers have consumed the result area of the previous call
    return distCollect(dRows.as<T>(), N, Mv3, st);
  }
  // out3 += prefactor sqrt(2 T) M_near^1/2 dW on every rank; distPrepare first
  int distNoise(int N, double temperature, double prefactor, uint32_t seed2, T *out3, int *iterations, cudaStream_t st) {
    if (iterations) *iterations = 0;
    if (!dAttached || listN != N) return UB200_ERR_NOT_BUILT;
    if (temperature == 0.0) return UB200_OK;
    const int nloc = dHi - dLo;
    int rc;
    if ((rc = z.reserve(sizeof(T) * 3 * (size_t)std::max(nloc, 1))) || (rc = noiseOut.reserve(sizeof(T) * 3 * (size_t)std::max(nloc, 1)))) return rc;
    const T noisePrefactor = (T)prefactor * (T)sqrt(2 * (T)temperature);
    if (nloc > 0) {
      distNoiseRows<T><<<(nloc + 255) / 256, 256, 0, st>>>(z.as<T>(), vl->cl->groupIndex.as<int>(), dLo, dHi, noisePrefactor, seedNear, seed2);
      UB200_LAUNCHED();
    }
    if ((rc = lanczosSqrt(z.as<T>(), noiseOut.as<T>(), N, (double)tolerance, iterations, st, true))) return rc;
    return distCollect(noiseOut.as<T>(), N, out3, st);
  }

  // lanczos::Solver::run (LanczosAlgorithm.cu:202-228) with KrylovSubspace (:27-173); matrix = near-field mobility.
  // dist: the vectors are the rows of this rank (zin, Bz and the Krylov basis hold 3 (hi - lo) numbers), the products
  // read the basis vector from the records every rank published into, the scalars are sums over the ranks - every rank
  // takes the same decisions from the same bits
  int lanczosSqrt(const T *zin, T *Bz, int N, double tol, int *iterations, cudaStream_t st, bool dist = false) {
    const size_t n = dist ? 3 * (size_t)(dHi - dLo) : 3 * (size_t)N;
    const size_t ld = std::max(n, (size_t)1);
    int rc;
    // dist: the basis is sized for 40 vectors at once (no allocation, hence no device-wide synchronisation, between barriers)
    if ((rc = reserveLanczos(n, dist ? 40 : 8, st))) return rc;
    if ((rc = fillUnit(oldBz.as<T>(), ld, false, st))) return rc;
    std::vector<double> hdiag, hsup;
    double nz2;
    if ((rc = dotHost(zin, zin, n, &nz2, st, dist))) return rc;
    const T normz = (T)sqrt(nz2);
    if (dist) { if ((rc = distPublishVec((T)(1.0 / normz), zin, (T *)V.p, st))) return rc; }
    else if ((rc = axpby((T)(1.0 / normz), zin, T(0), (T *)V.p, n, st))) return rc;
    const int checkSteps = std::min(checkConvergenceSteps, iterationHardLimit - 2);
    for (int i = 0; i < iterationHardLimit; i++) {
      if ((rc = V.grow(sizeof(T) * ld * (size_t)(i + 2), st))) return rc;
      T *Vm = (T *)V.p, *dw = w.as<T>();
      if (dist) { if ((rc = distDotRows(dw, st))) return rc; }                           // w = M v_i
      else if ((rc = nearDotList(Vm + ld * i, 3, N, dw, false, st))) return rc;
      if (i > 0 && (rc = axpby((T)(-hsup[i - 1]), Vm + ld * (i - 1), T(1), dw, n, st))) return rc;
      double hd;
      if ((rc = dotHost(dw, Vm + ld * i, n, &hd, st, dist))) return rc;
      hdiag.push_back((double)(T)hd);
      if ((rc = axpby((T)(-hdiag[i]), Vm + ld * i, T(1), dw, n, st))) return rc;
      double hs2;
      if ((rc = dotHost(dw, dw, n, &hs2, st, dist))) return rc;
      double hs = (double)(T)sqrt(hs2);
      const T tolw = (T)(1e-3 * hdiag[i] / normz);
      if (hs < tolw) hs = 0.0;
      hsup.push_back(hs);
      if (hs > 0.0) {
        if (dist) { if ((rc = distPublishVec((T)(1.0 / hs), dw, Vm + ld * (i + 1), st))) return rc; }
        else if ((rc = axpby((T)(1.0 / hs), dw, T(0), Vm + ld * (i + 1), n, st))) return rc;
      } else { // w = e1
        if ((rc = fillUnit(Vm + ld * (i + 1), ld, !dist || (dLo == 0 && dHi > 0), st))) return rc;
        if (dist && (rc = distPublishVec(T(1), Vm + ld * (i + 1), Vm + ld * (i + 1), st))) return rc;
      }
      if (i >= checkSteps) {
        // Bz = ||z|| V_m H^1/2 e1 (computeCurrentResultEstimation :163-172, computeSquareRoot :63-80)
        const int m = i + 1;
        std::vector<double> d(hdiag.begin(), hdiag.begin() + m), e(hsup.begin(), hsup.begin() + m), Z;
        if (!tridiagEigen(m, d, e, Z)) return UB200_ERR_UNSUPPORTED;
        std::vector<double> tmp(m);
        for (int j = 0; j < m; j++) tmp[j] = sqrt(std::max(d[j], 0.0)) * Z[(size_t)j * m];
        LanczosCoeff<T> c;
        for (int r = 0; r < m; r++) {
          double s = 0;
          for (int j = 0; j < m; j++) s += Z[(size_t)j * m + r] * tmp[j];
          c.c[r] = (T)s;
        }
        if (n > 0) {
          vecGemvArg<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Vm, ld, m, c, normz, Bz);
          UB200_LAUNCHED();
        }
        if (i > 0) { // computeError (:231-250): ||Bz_i - Bz_{i-1}|| / ||Bz_{i-1}||
          double prev2, yy2;
          if ((rc = dotHost(oldBz.as<T>(), oldBz.as<T>(), n, &prev2, st, dist))) return rc;
          if ((rc = axpby(T(-1), Bz, T(1), oldBz.as<T>(), n, st))) return rc;
          if ((rc = dotHost(oldBz.as<T>(), oldBz.as<T>(), n, &yy2, st, dist))) return rc;
          const double err = fabs(sqrt(yy2) / sqrt(prev2));
          if (std::isnan(err)) return UB200_ERR_UNSUPPORTED;
          if (err <= tol) {
            lastRunRequiredSteps = i;
            if (i - 2 > checkConvergenceSteps) checkConvergenceSteps += 1;
            else checkConvergenceSteps = std::max(1, checkConvergenceSteps - 2);
            if (iterations) *iterations = i;
            return UB200_OK;
          }
        }
        if ((rc = axpby(T(1), Bz, T(0), oldBz.as<T>(), n, st))) return rc;
      }
    }
    return UB200_ERR_UNSUPPORTED; // "[Lanczos] Could not converge"
  }

  // NearField::Mdot over the Verlet list instead of the cell list: for callers that need the list anyway (a step with
  // noise: the Lanczos iteration walks it 5 - 10 times) - the list is built here and may be reused by the nearNoise that follows
  int nearMdotList(const void *pos, const T *v, int vStride, int N, T *Mv3, cudaStream_t st) {
    int rc;
    if ((rc = nearPrepareList(pos, N, st))) return rc;
    listPos = pos;
    return nearDotList(v, vStride, N, Mv3, true, st);
  }

  // NearField::computeStochasticDisplacements (NearField.cuh:254-282): BdW = prefactor sqrt(2 T) M_near^1/2 dW
  // reuseList: the caller states that the positions are the ones of the preceding nearMdotList (same array, unchanged)
  int nearNoise(const void *pos, int N, double temperature, double prefactor, uint32_t seed2, void *BdW3, int *iterations,
                cudaStream_t st, bool reuseList = false) {
    if (iterations) *iterations = 0;
    if (temperature == 0.0) return UB200_OK;
    int rc;
    const bool reuse = reuseList && listPos == pos && listN == N;
    listPos = nullptr;
    if (!reuse && (rc = nearPrepareList(pos, N, st))) return rc;
    if ((rc = z.reserve(sizeof(T) * 3 * (size_t)N))) return rc;
    const T noisePrefactor = (T)prefactor * (T)sqrt(2 * (T)temperature);
    pseNearNoise<T><<<(N + 255) / 256, 256, 0, st>>>(z.as<T>(), N, noisePrefactor, seedNear, seed2);
    UB200_LAUNCHED();
    return lanczosSqrt(z.as<T>(), (T *)BdW3, N, (double)tolerance, iterations, st);
  }
  // out += prefactor sqrt(2T) Mr^1/2 dW: the square root runs into a scratch vector (the Lanczos iteration rewrites its
  // output at every convergence check), which is then added
  int nearNoiseAdd(const void *pos, int N, double temperature, double prefactor, uint32_t seed2, void *out3, int *iterations,
                   cudaStream_t st) {
    if (iterations) *iterations = 0;
    if (temperature == 0.0) return UB200_OK;
    int rc;
    if ((rc = noiseOut.reserve(sizeof(T) * 3 * (size_t)N))) return rc;
    if ((rc = nearNoise(pos, N, temperature, prefactor, seed2, noiseOut.p, iterations, st))) return rc;
    return axpby(T(1), noiseOut.as<T>(), T(1), (T *)out3, 3 * (size_t)N, st);
  }
  DevBuf noiseOut;
};

} // namespace ub200

using namespace ub200;

struct ub200_pse {
  int precision;
  PseState<float> f;
  PseState<double> d;
};

#define PSE_DISPATCH(h, call) ((h)->precision == 4 ? (h)->f.call : (h)->d.call)

extern "C" {

int ub200_pse_create(ub200_pse **out, int precisionBytes, const ub200_pse_params *par, uint32_t seedNear, uint32_t seedFar) {
  if (!out || !par || (precisionBytes != 4 && precisionBytes != 8)) return UB200_ERR_INVALID_ARGUMENT;
  ub200_pse *h = new (std::nothrow) ub200_pse();
  if (!h) return UB200_ERR_ALLOC;
  h->precision = precisionBytes;
  const int rc = precisionBytes == 4 ? h->f.init(*par, seedNear, seedFar) : h->d.init(*par, seedNear, seedFar);
  if (rc) { h->f.release(); h->d.release(); delete h; return rc; }
  *out = h;
  return UB200_OK;
}
int ub200_pse_destroy(ub200_pse *h) {
  if (!h) return UB200_OK;
  h->f.release(); h->d.release();
  delete h;
  return UB200_OK;
}
int ub200_pse_info(ub200_pse *h, ub200_pse_info_t *info) {
  if (!h || !info) return UB200_ERR_INVALID_ARGUMENT;
  const bool f = h->precision == 4;
  info->cells[0] = f ? h->f.plan.nx : h->d.plan.nx;
  info->cells[1] = f ? h->f.plan.ny : h->d.plan.ny;
  info->cells[2] = f ? h->f.plan.nz : h->d.plan.nz;
  info->support = f ? h->f.support : h->d.support;
  info->eta = f ? (double)h->f.eta : h->d.eta;
  info->rcut = f ? (double)h->f.rcut : h->d.rcut;
  info->nTable = f ? h->f.nTable : h->d.nTable;
  info->lastLanczosIterations = f ? h->f.lastRunRequiredSteps : h->d.lastRunRequiredSteps;
  info->d_table = f ? h->f.table.p : h->d.table.p;
  info->d_grid = f ? h->f.grid.p : h->d.grid.p;
  info->kernel = f ? h->f.farKernel : h->d.farKernel;
  return UB200_OK;
}
int ub200_pse_set_shear_strain(ub200_pse *h, double strain) {
  if (!h) return UB200_ERR_INVALID_ARGUMENT;
  h->f.shear = (float)strain; h->d.shear = strain;
  return UB200_OK;
}
int ub200_pse_far_mdot(ub200_pse *h, const void *d_pos, const void *d_force, int N, double temperature, double prefactor,
                       uint32_t seed2, void *d_MF3, void *stream) {
  if (!h || !d_pos || !d_MF3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, farMdot(d_pos, d_force, N, temperature, prefactor, seed2, d_MF3, (cudaStream_t)stream));
}
int ub200_pse_near_mdot(ub200_pse *h, const void *d_pos, const void *d_v, int vStride, int N, void *d_Mv3, void *stream) {
  if (!h || !d_pos || !d_Mv3 || N <= 0 || (vStride != 3 && vStride != 4)) return UB200_ERR_INVALID_ARGUMENT;
  if (!d_v) return UB200_OK; // NearField::Mdot skips when there are no forces
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (h->precision == 4) {
    if ((rc = h->f.nearPrepare(d_pos, N, st))) return rc;
    return h->f.nearDot((const float *)d_v, vStride, N, (float *)d_Mv3, true, st);
  }
  if ((rc = h->d.nearPrepare(d_pos, N, st))) return rc;
  return h->d.nearDot((const double *)d_v, vStride, N, (double *)d_Mv3, true, st);
}
int ub200_pse_near_mdot_list(ub200_pse *h, const void *d_pos, const void *d_v, int vStride, int N, void *d_Mv3, void *stream) {
  if (!h || !d_pos || !d_Mv3 || N <= 0 || (vStride != 3 && vStride != 4)) return UB200_ERR_INVALID_ARGUMENT;
  if (!d_v) return UB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->precision == 4) return h->f.nearMdotList(d_pos, (const float *)d_v, vStride, N, (float *)d_Mv3, st);
  return h->d.nearMdotList(d_pos, (const double *)d_v, vStride, N, (double *)d_Mv3, st);
}
int ub200_pse_near_noise_reuse(ub200_pse *h, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                               void *d_BdW3, int *iterations, void *stream) {
  if (!h || !d_pos || !d_BdW3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, nearNoise(d_pos, N, temperature, prefactor, seed2, d_BdW3, iterations, (cudaStream_t)stream, true));
}
int ub200_pse_near_noise(ub200_pse *h, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                         void *d_BdW3, int *iterations, void *stream) {
  if (!h || !d_pos || !d_BdW3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, nearNoise(d_pos, N, temperature, prefactor, seed2, d_BdW3, iterations, (cudaStream_t)stream));
}
int ub200_pse_near_noise_add(ub200_pse *h, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                             void *d_out3, int *iterations, void *stream) {
  if (!h || !d_pos || !d_out3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, nearNoiseAdd(d_pos, N, temperature, prefactor, seed2, d_out3, iterations, (cudaStream_t)stream));
}

/* ---- near field over ranks (PSE near field + Lanczos decomposed by rows of the sorted order) ---- */
int ub200_pse_dist_create(ub200_pse *h, int rank, int world, int maxParticles) {
  if (!h) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, distCreate(rank, world, maxParticles));
}
int ub200_pse_dist_ipc_export(ub200_pse *h, void *blob) {
  if (!h || !blob) return UB200_ERR_INVALID_ARGUMENT;
  void *arena = h->precision == 4 ? h->f.dArena : h->d.dArena;
  if (!arena) return UB200_ERR_NOT_BUILT;
  cudaIpcMemHandle_t m;
  UB200_CUDA(cudaIpcGetMemHandle(&m, arena));
  memcpy(blob, &m, sizeof(m));
  return UB200_OK;
}
} // extern "C"
template <class S> static int pseDistImport(S &s, const void *blobs) {
  if (!s.dArena) return UB200_ERR_NOT_BUILT;
  for (int p = 0; p < s.dWorld; p++) {
    if (p == s.dRank) continue;
    cudaIpcMemHandle_t m;
    memcpy(&m, static_cast<const char *>(blobs) + (size_t)p * sizeof(m), sizeof(m));
    void *ptr = nullptr;
    UB200_CUDA(cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
    s.dPeer[p] = (char *)ptr;
    s.dOpened[p] = true;
  }
  s.dAttached = true;
  return UB200_OK;
}
extern "C" {
int ub200_pse_dist_ipc_import(ub200_pse *h, const void *blobsOfAllRanks) {
  if (!h || !blobsOfAllRanks) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? pseDistImport(h->f, blobsOfAllRanks) : pseDistImport(h->d, blobsOfAllRanks);
}
int ub200_pse_dist_arena(ub200_pse *h, void **arena) {
  if (!h || !arena) return UB200_ERR_INVALID_ARGUMENT;
  *arena = h->precision == 4 ? h->f.dArena : h->d.dArena;
  return *arena ? UB200_OK : UB200_ERR_NOT_BUILT;
}
} // extern "C"
template <class S> static int pseDistAttach(S &s, void *const *arenas) {
  if (!s.dArena) return UB200_ERR_NOT_BUILT;
  for (int p = 0; p < s.dWorld; p++) {
    if (!arenas[p]) return UB200_ERR_INVALID_ARGUMENT;
    if (p != s.dRank) s.dPeer[p] = (char *)arenas[p];
  }
  s.dAttached = true;
  return UB200_OK;
}
extern "C" {
int ub200_pse_dist_attach_local(ub200_pse *h, void *const *arenasOfAllRanks) {
  if (!h || !arenasOfAllRanks) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? pseDistAttach(h->f, arenasOfAllRanks) : pseDistAttach(h->d, arenasOfAllRanks);
}
int ub200_pse_dist_near_prepare(ub200_pse *h, const void *d_pos, int N, void *stream) {
  if (!h || !d_pos || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return PSE_DISPATCH(h, distPrepare(d_pos, N, (cudaStream_t)stream));
}
int ub200_pse_dist_near_mdot(ub200_pse *h, const void *d_v, int vStride, int N, void *d_Mv3, void *stream) {
  if (!h || !d_Mv3 || N <= 0 || (vStride != 3 && vStride != 4)) return UB200_ERR_INVALID_ARGUMENT;
  if (!d_v) return UB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->precision == 4) return h->f.distMdot((const float *)d_v, vStride, N, (float *)d_Mv3, st);
  return h->d.distMdot((const double *)d_v, vStride, N, (double *)d_Mv3, st);
}
int ub200_pse_dist_near_noise_add(ub200_pse *h, int N, double temperature, double prefactor, uint32_t seed2, void *d_out3,
                                  int *iterations, void *stream) {
  if (!h || !d_out3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->precision == 4) return h->f.distNoise(N, temperature, prefactor, seed2, (float *)d_out3, iterations, st);
  return h->d.distNoise(N, temperature, prefactor, seed2, (double *)d_out3, iterations, st);
}
int ub200_pse_dist_error_flag(ub200_pse *h, void *stream, int *flag) {
  if (!h || !flag) return UB200_ERR_INVALID_ARGUMENT;
  const void *p = h->precision == 4 ? h->f.dErr.p : h->d.dErr.p;
  if (!p) return UB200_ERR_NOT_BUILT;
  UB200_CUDA(cudaMemcpyAsync(flag, p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  UB200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return UB200_OK;
}
}
