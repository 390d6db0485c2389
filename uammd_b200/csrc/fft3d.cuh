// 3-D FFT plan + pass kernels (see fft.cuh for the design). Header so that the FCM / PSE pipelines can
// instantiate the fused z pass with their own spectral operators.
#pragma once
#include "fft.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace ub200 {

constexpr int kFftThreads = 256;

// NC: interleaved components per grid node (3: the real3 grids of the hydrodynamic solvers; 1 and 4: the Poisson solver)
template <class T, int NC = 3> struct Fft3dPlan {
  using C = typename Vec2<T>::type;
  int nx = 0, ny = 0, nz = 0, nkx = 0, nxPad = 0;
  FftAxis ax, ay, az;
  DevBuf twx, twy, twz;
  int linesPerCta = 0; // X pass: real lines per CTA (even)
  int tileY = 0, tileZ = 0; // Y/Z pass: kx per tile
  size_t smemX = 0, smemY = 0, smemZ = 0;

  static size_t smemBudget() { return 75 * 1024; } // 3 CTAs per SM

  int init(int nx_, int ny_, int nz_) {
    nx = nx_; ny = ny_; nz = nz_;
    nkx = nx / 2 + 1;
    nxPad = 2 * nkx;
    if (nx < 2 || ny < 1 || nz < 1) return UB200_ERR_INVALID_ARGUMENT;
    if (!factorize(nx, ax) || !factorize(ny, ay) || !factorize(nz, az)) return UB200_ERR_UNSUPPORTED;
    int rc;
    if ((rc = upload(twx, nx)) || (rc = upload(twy, ny)) || (rc = upload(twz, nz))) return rc;
    // X: pairs of lines, 3 components -> (lines/2)*3 complex transforms of length nx, two buffers
    auto fit = [&](int n, int perUnit, int maxUnits, int nbuf = 3) {
      int units = maxUnits;
      while (units > 1 && nbuf * (size_t)units * perUnit * (n + 1) * sizeof(C) > smemBudget()) units--;
      return units;
    };
    // experiment knobs (environment): UB200_FFT_TILE = kx per strided tile (1, 2, 4), UB200_FFT_PAIRS = line pairs per x CTA
    const char *envTile = getenv("UB200_FFT_TILE"), *envPairs = getenv("UB200_FFT_PAIRS");
    const int maxTile = envTile ? atoi(envTile) : 4, maxPairs = envPairs ? std::min(atoi(envPairs), 4) : 4;
    const int pairs = fit(nx, NC, maxPairs, 2);
    linesPerCta = 2 * pairs;
    smemX = 2 * (size_t)pairs * NC * (nx + 1) * sizeof(C);
    auto pow2 = [](int v) { return v >= 4 ? 4 : (v >= 2 ? 2 : 1); };
    tileY = pow2(fit(ny, NC, maxTile));
    smemY = 3 * (size_t)tileY * NC * (ny + 1) * sizeof(C);
    tileZ = pow2(fit(nz, NC, maxTile));
    smemZ = 3 * (size_t)tileZ * NC * (nz + 1) * sizeof(C);
    if (smemX > 200 * 1024 || smemY > 200 * 1024 || smemZ > 200 * 1024) return UB200_ERR_UNSUPPORTED;
    return UB200_OK;
  }
  void release() { twx.release(); twy.release(); twz.release(); }
  size_t gridBytes() const { return (size_t)nz * ny * nkx * NC * sizeof(C); }

private:
  static int upload(DevBuf &buf, int n) {
    std::vector<C> h(n);
    for (int j = 0; j < n; j++) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
      h[j].x = (T)cosl(a);
      h[j].y = (T)sinl(a);
    }
    int rc = buf.reserve(sizeof(C) * (size_t)n);
    if (rc) return rc;
    UB200_CUDA(cudaMemcpy(buf.p, h.data(), sizeof(C) * (size_t)n, cudaMemcpyHostToDevice));
    return UB200_OK;
  }
};

// 16- or 8-byte asynchronous global->shared copy (LDGSTS): no registers, completion through commit groups
template <class C> __device__ __forceinline__ void cpAsync(C *smemDst, const C *gmemSrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  if (sizeof(C) == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmemSrc) : "memory");
}
// 4- or 8-byte variant for real-valued lines (x pass)
template <class T> __device__ __forceinline__ void cpAsyncReal(T *smemDst, const T *gmemSrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmemSrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------- X pass: real <-> complex along the contiguous axis, in place ----------------
template <class T, bool FORWARD, int NFIX, int NC = 3>
__global__ void __launch_bounds__(kFftThreads)
fftPassX(T *__restrict__ grid, int nxRuntime, int nkxRuntime, int nlines, int linesPerCta, FftAxis ax,
         const typename Vec2<T>::type *__restrict__ tw) {
  using C = typename Vec2<T>::type;
  const int nx = NFIX > 0 ? NFIX : nxRuntime, nkx = NFIX > 0 ? NFIX / 2 + 1 : nkxRuntime;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int fstride = nx + 1;
  const int pairs = linesPerCta / 2;
  const int nf = pairs * NC;
  C *buf0 = reinterpret_cast<C *>(smemRaw);
  C *buf1 = buf0 + (size_t)nf * fstride;
  const int line0 = blockIdx.x * linesPerCta;
  const int nl = min(linesPerCta, nlines - line0);
  const size_t lineReals = (size_t)2 * nkx * NC; // reals per line (padded) == 2 * complex per line
  T *base = grid + (size_t)line0 * lineReals;
  C *cbase = reinterpret_cast<C *>(base);
  if (FORWARD) {
    // load real lines: line l, sample x, component c -> transform (l/2)*3+c, real (l even) or imaginary part
    // (asynchronous copies: every load of the CTA's lines is in flight at once instead of one dependent load -> store
    //  round trip per element; measured 1.8 TB/s with plain loads)
    for (int l = 0; l < linesPerCta; l++)
      for (int rem = threadIdx.x; rem < nx * NC; rem += blockDim.x) {
        const int x = rem / NC, c = rem - NC * x;
        T *dst = reinterpret_cast<T *>(buf0 + ((l >> 1) * NC + c) * fstride + x) + (l & 1);
        if (l < nl) cpAsyncReal(dst, base + (size_t)l * lineReals + rem);
        else *dst = T(0);
      }
    cpAsyncCommit();
    cpAsyncWait<0>();
    __syncthreads();
    C *res = fftShared<T, -1, NFIX>(buf0, buf1, ax, fstride, nf, tw);
    // untangle the two real transforms and store the Hermitian halves
    for (int p = 0; p < pairs; p++)
    for (int rem = threadIdx.x; rem < nkx * NC; rem += blockDim.x) {
      const int k = rem / NC, c = rem - NC * k;
      const C zk = res[(p * NC + c) * fstride + k];
      const C zn = res[(p * NC + c) * fstride + (k == 0 ? 0 : nx - k)];
      const C A = mk2<T>(T(0.5) * (zk.x + zn.x), T(0.5) * (zk.y - zn.y));
      const C B = mk2<T>(T(0.5) * (zk.y + zn.y), T(0.5) * (zn.x - zk.x));
      if (2 * p < nl) cbase[(size_t)(2 * p) * nkx * NC + rem] = A;
      if (2 * p + 1 < nl) cbase[(size_t)(2 * p + 1) * nkx * NC + rem] = B;
    }
  } else {
    // build Z_k = A_k + i B_k for all k from the stored halves: every stored mode is loaded ONCE and also written to its
    // Hermitian mirror k' = nx - k (A_k' = conj A_k, B_k' = conj B_k); like a C2R transform, the imaginary parts of the
    // self-conjugate modes (k = 0, and k = nx/2 for even nx) are ignored
    for (int rem = threadIdx.x; rem < nkx * NC; rem += blockDim.x) {
      const int k = rem / NC, c = rem - NC * k;
      C A[4], B[4]; // pairs <= 4: all loads of this thread are issued back to back
#pragma unroll
      for (int p = 0; p < 4; p++) {
        A[p] = mk2<T>(T(0), T(0)); B[p] = A[p];
        if (p < pairs && 2 * p < nl) A[p] = cbase[(size_t)(2 * p) * nkx * NC + rem];
        if (p < pairs && 2 * p + 1 < nl) B[p] = cbase[(size_t)(2 * p + 1) * nkx * NC + rem];
      }
#pragma unroll
      for (int p = 0; p < 4; p++) {
        if (p >= pairs) break;
        C a = A[p], b = B[p];
        if (k == 0 || 2 * k == nx) { a.y = T(0); b.y = T(0); }
        C *row = buf0 + (p * NC + c) * fstride;
        row[k] = mk2<T>(a.x - b.y, a.y + b.x);
        if (k > 0 && 2 * k < nx) row[nx - k] = mk2<T>(a.x + b.y, b.x - a.y);
      }
    }
    __syncthreads();
    C *res = fftShared<T, +1, NFIX>(buf0, buf1, ax, fstride, nf, tw);
    for (int l = 0; l < nl; l++)
      for (int rem = threadIdx.x; rem < nx * NC; rem += blockDim.x) {
        const int x = rem / NC, c = rem - NC * x;
        const T *src = reinterpret_cast<const T *>(res + ((l >> 1) * NC + c) * fstride + x);
        base[(size_t)l * lineReals + rem] = src[l & 1];
      }
  }
}

// ---------------- Y / Z pass: complex transforms along a strided axis, in place ----------------
// A tile is `tile` consecutive kx (x 3 components) times the whole axis of length n. elemStride = distance
// (in complex3 nodes) between consecutive points of the axis; tiles are enumerated by (tx, other) where
// `other` runs over the remaining axis with stride otherStride.
struct NoSpectralOp {
  template <class C> __device__ __forceinline__ void operator()(int, int, int, C &, C &, C &) const {}
};

// Address policies of the strided pass: where axis point i of tile (other, kx0) is loaded from / stored to.
// In place (one GPU): the same grid for both.
template <class C, int NC = 3> struct AddrInPlace {
  static constexpr int kComp = NC;
  C *grid;
  size_t elemStride, otherStride; // in nodes of NC complex numbers
  __device__ __forceinline__ const C *ld(int other, int i, int kx0) const {
    return grid + ((size_t)other * otherStride + kx0 + (size_t)i * elemStride) * NC;
  }
  __device__ __forceinline__ void store(int other, int i, int kx0, int gf, const C &v) const {
    grid[((size_t)other * otherStride + kx0 + (size_t)i * elemStride) * NC + gf] = v;
  }
};
// Slab-decomposed 3-D FFT over `world` GPUs (NVLink peer stores, no staging copy, no NCCL): rank r owns the z planes
// [r nzl, (r+1) nzl) in S = [nzl][ny][nkx][3] and, after the transpose, the ky rows [r nyl, (r+1) nyl) in
// T = [nz][nyl][nkx][3]. The forward y pass loads a line from the local S and scatters its output straight into
// the owners' T buffers (the all-to-all transpose fused into the pass's store); the fused z pass loads from the
// local T and stores the inverse-transformed lines straight back into the owners' S buffers.
constexpr int kFftMaxPeers = 8;
template <class C> struct AddrSlabYForward { // axis = y, other = local z plane
  static constexpr int kComp = 3;
  C *S;
  C *peerT[kFftMaxPeers];
  int ny, nyl, nkx, z0; // z0: first global plane of this rank
  __device__ __forceinline__ const C *ld(int other, int i, int kx0) const {
    return S + (((size_t)other * ny + i) * nkx + kx0) * 3;
  }
  __device__ __forceinline__ void store(int other, int i, int kx0, int gf, const C &v) const {
    const int r = i / nyl;
    peerT[r][(((size_t)(z0 + other) * nyl + (i - r * nyl)) * nkx + kx0) * 3 + gf] = v;
  }
};
// The slabs carry `halo` extra planes below and above the owned ones ([nzl + 2 halo][ny][nkx][3], owned planes start at
// index halo): the boundary planes of a slab are ALSO stored into the neighbours' halo planes, so that the inverse y / x
// passes and the interpolation of the neighbours never touch remote memory (and need no barrier of their own).
template <class C> struct AddrSlabZFused { // axis = z, other = local ky row
  static constexpr int kComp = 3;
  C *T;
  C *peerS[kFftMaxPeers];
  int ny, nyl, nkx, nzl, y0, halo, world; // y0: first global ky row of this rank
  __device__ __forceinline__ const C *ld(int other, int i, int kx0) const {
    return T + (((size_t)i * nyl + other) * nkx + kx0) * 3;
  }
  __device__ __forceinline__ void store(int other, int i, int kx0, int gf, const C &v) const {
    const int r = i / nzl, lz = i - r * nzl;
    const size_t inPlane = ((size_t)(y0 + other) * nkx + kx0) * 3 + gf, plane = (size_t)ny * nkx * 3;
    peerS[r][(size_t)(lz + halo) * plane + inPlane] = v;
    if (lz < halo) { // upper halo of the slab below
      const int rb = r == 0 ? world - 1 : r - 1;
      peerS[rb][(size_t)(nzl + halo + lz) * plane + inPlane] = v;
    }
    if (lz >= nzl - halo) { // lower halo of the slab above
      const int ra = r == world - 1 ? 0 : r + 1;
      peerS[ra][(size_t)(lz - (nzl - halo)) * plane + inPlane] = v;
    }
  }
};

// MODE: -1 forward, +1 inverse, 0 fused (forward, op, inverse).
// Persistent CTAs walk the tiles; three shared buffers rotate as {current, ping-pong scratch, prefetch}: the
// cp.async loads of the NEXT tile are in flight while the current tile is transformed and stored, so HBM/L2
// latency is hidden even at 3 CTAs per SM.
template <class T, int MODE, bool AXIS_IS_Z, class Op, int NFIX, class Addr>
__global__ void __launch_bounds__(kFftThreads)
fftPassStrided(Addr addr, int nRuntime, int nkx, int nOther, int tile, FftAxis ax,
               const typename Vec2<T>::type *__restrict__ tw, Op op) {
  using C = typename Vec2<T>::type;
  constexpr int NC = Addr::kComp; // interleaved components per node
  static_assert(MODE != 0 || NC == 3, "the fused spectral operators work on three components");
  const int n = NFIX > 0 ? NFIX : nRuntime;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int fstride = n + 1;
  const int ntx = (nkx + tile - 1) / tile;
  const int ntiles = ntx * nOther;
  const int nf = tile * NC;
  C *const bufBase = reinterpret_cast<C *>(smemRaw);
  const int bufStride = nf * fstride;
#define bufs(k) (bufBase + (k) * bufStride)
  // threads in groups of gsz = 2^glog >= nf: lane f of a group moves complex number f of axis point i
  const int glog = nf <= 4 ? 2 : (nf <= 8 ? 3 : 4), gsz = 1 << glog;
  const int gf = threadIdx.x & (gsz - 1), gi = threadIdx.x >> glog, gstep = blockDim.x >> glog;
  const int tlog = tile == 4 ? 2 : (tile == 2 ? 1 : 0); // tile is 1, 2 or 4

  auto issueLoad = [&](int t, C *dst) {
    const int tx = t % ntx, other = t / ntx;
    const int kx0 = tx * tile;
    const int w = min(tile, nkx - kx0) * NC;
    if (gf < nf) {
      if (gf < w) for (int i = gi; i < n; i += gstep) cpAsync(dst + gf * fstride + i, addr.ld(other, i, kx0) + gf);
      else for (int i = gi; i < n; i += gstep) dst[gf * fstride + i] = mk2<T>(T(0), T(0));
    }
    cpAsyncCommit();
  };

  int t = blockIdx.x, cur = 0;
  if (t < ntiles) issueLoad(t, bufs(0));
  for (; t < ntiles; t += gridDim.x) {
    const int nxt = t + gridDim.x;
    const int scratch = cur == 2 ? 0 : cur + 1, pre = scratch == 2 ? 0 : scratch + 1;
    if (nxt < ntiles) { issueLoad(nxt, bufs(pre)); cpAsyncWait<1>(); } else { cpAsyncWait<0>(); }
    __syncthreads();
    const int tx = t % ntx, other = t / ntx;
    const int kx0 = tx * tile;
    const int w = min(tile, nkx - kx0) * NC;
    C *res;
    if (MODE <= 0) res = fftShared<T, -1, NFIX>(bufs(cur), bufs(scratch), ax, fstride, nf, tw);
    else res = fftShared<T, +1, NFIX>(bufs(cur), bufs(scratch), ax, fstride, nf, tw);
    if (MODE == 0) {
      // spectral operator on the three components of every Fourier node of the tile
      for (int idx = threadIdx.x; idx < n * tile; idx += blockDim.x) {
        const int i = idx >> tlog, tt = idx & (tile - 1);
        if (kx0 + tt < nkx) {
          C *v = res + (tt * NC) * fstride + i;
          C vx = v[0], vy = v[fstride], vz = v[2 * fstride];
          if (AXIS_IS_Z) op(kx0 + tt, other, i, vx, vy, vz);
          else op(kx0 + tt, i, other, vx, vy, vz);
          v[0] = vx; v[fstride] = vy; v[2 * fstride] = vz;
        }
      }
      __syncthreads();
      C *other1 = res == bufs(cur) ? bufs(scratch) : bufs(cur);
      res = fftShared<T, +1, NFIX>(res, other1, ax, fstride, nf, tw);
    }
    if (gf < w)
      for (int i = gi; i < n; i += gstep) addr.store(other, i, kx0, gf, res[gf * fstride + i]);
    __syncthreads(); // result consumed: current + scratch may be overwritten by the next iterations
    cur = pre;
  }
#undef bufs
}

template <class T> int fftEnsureSmem(const void *kern, size_t bytes) {
  if (bytes > 48 * 1024) UB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return UB200_OK;
}

// one CTA wave that fills the GPU (occupancy is a property of (kernel, smem, threads); cached per kernel pointer)
inline int persistentGrid(const void *kern, size_t smem, int ntiles, int threads) {
  static const void *cachedKern[64];
  static size_t cachedSmem[64];
  static int cachedBlocks[64];
  static int ncached = 0;
  int perSM = 0;
  for (int i = 0; i < ncached; i++)
    if (cachedKern[i] == kern && cachedSmem[i] == smem) perSM = cachedBlocks[i];
  if (!perSM) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, threads, smem) != cudaSuccess || perSM < 1) perSM = 1;
    if (ncached < 64) { cachedKern[ncached] = kern; cachedSmem[ncached] = smem; cachedBlocks[ncached++] = perSM; }
  }
  const int g = kNumSMs * perSM;
  return g < ntiles ? g : ntiles;
}

// CTA size: the specialised stages split 12 transforms * n/R butterflies evenly over 192 threads
template <int NFIX> inline int fftThreads() {
  static const int envThreads = getenv("UB200_FFT_THREADS") ? atoi(getenv("UB200_FFT_THREADS")) : 0;
  if (envThreads >= 32 && envThreads <= kFftThreads && envThreads % 16 == 0) return envThreads;
  return NFIX > 0 ? 192 : kFftThreads;
}

// axis lengths with compile-time specialised kernels (anything else takes the generic mixed-radix path)
#define UB200_FFT_DISPATCH(n, CALL)                                                                         \
  switch (n) {                                                                                              \
  case 32: { constexpr int NFIX = 32; CALL; } break;                                                        \
  case 64: { constexpr int NFIX = 64; CALL; } break;                                                        \
  case 128: { constexpr int NFIX = 128; CALL; } break;                                                      \
  case 256: { constexpr int NFIX = 256; CALL; } break;                                                      \
  case 512: { constexpr int NFIX = 512; CALL; } break;                                                      \
  default: { constexpr int NFIX = 0; CALL; } break;                                                         \
  }

template <class T, bool FORWARD, int NFIX, int NC>
int launchPassXFixed(const Fft3dPlan<T, NC> &p, void *grid, cudaStream_t st, int nzLocal) {
  auto kern = fftPassX<T, FORWARD, NFIX, NC>;
  int rc = fftEnsureSmem<T>((const void *)kern, p.smemX);
  if (rc) return rc;
  const int nlines = p.ny * (nzLocal > 0 ? nzLocal : p.nz);
  const int nb = (nlines + p.linesPerCta - 1) / p.linesPerCta;
  kern<<<nb, fftThreads<NFIX>(), p.smemX, st>>>((T *)grid, p.nx, p.nkx, nlines, p.linesPerCta, p.ax,
                                                p.twx.template as<typename Vec2<T>::type>());
  UB200_LAUNCHED();
  return UB200_OK;
}
// nzLocal > 0: only that many z planes are held in `grid` (slab decomposition)
template <class T, bool FORWARD, int NC> int launchPassX(const Fft3dPlan<T, NC> &p, void *grid, cudaStream_t st, int nzLocal = 0) {
  int rc = UB200_OK;
  UB200_FFT_DISPATCH(p.nx, (rc = launchPassXFixed<T, FORWARD, NFIX, NC>(p, grid, st, nzLocal)));
  return rc;
}

template <class T, int MODE, bool AXIS_IS_Z, class Op, int NFIX, int NC>
int launchPassStridedFixed(const Fft3dPlan<T, NC> &p, void *grid, cudaStream_t st, Op op) {
  using C = typename Vec2<T>::type;
  auto kern = fftPassStrided<T, MODE, AXIS_IS_Z, Op, NFIX, AddrInPlace<C, NC>>;
  const size_t smem = AXIS_IS_Z ? p.smemZ : p.smemY;
  int rc = fftEnsureSmem<T>((const void *)kern, smem);
  if (rc) return rc;
  const int tile = AXIS_IS_Z ? p.tileZ : p.tileY;
  const int ntx = (p.nkx + tile - 1) / tile;
  const int nOther = AXIS_IS_Z ? p.ny : p.nz;
  const int threads = fftThreads<NFIX>();
  const int nblocks = persistentGrid((const void *)kern, smem, ntx * nOther, threads);
  AddrInPlace<C, NC> addr;
  addr.grid = (C *)grid;
  if (AXIS_IS_Z) {
    addr.elemStride = (size_t)p.nkx * p.ny; addr.otherStride = (size_t)p.nkx;
    kern<<<nblocks, threads, smem, st>>>(addr, p.nz, p.nkx, p.ny, tile, p.az, p.twz.template as<C>(), op);
  } else {
    addr.elemStride = (size_t)p.nkx; addr.otherStride = (size_t)p.nkx * p.ny;
    kern<<<nblocks, threads, smem, st>>>(addr, p.ny, p.nkx, nOther, tile, p.ay, p.twy.template as<C>(), op);
  }
  UB200_LAUNCHED();
  return UB200_OK;
}

// strided pass with an explicit address policy (slab-decomposed transforms); nOther = lines of the other axis held
// locally, n = full axis length
template <class T, int MODE, bool AXIS_IS_Z, class Op, int NFIX, class Addr>
int launchPassAddrFixed(const Fft3dPlan<T> &p, const Addr &addr, int nOther, cudaStream_t st, Op op) {
  using C = typename Vec2<T>::type;
  auto kern = fftPassStrided<T, MODE, AXIS_IS_Z, Op, NFIX, Addr>;
  const size_t smem = AXIS_IS_Z ? p.smemZ : p.smemY;
  int rc = fftEnsureSmem<T>((const void *)kern, smem);
  if (rc) return rc;
  const int tile = AXIS_IS_Z ? p.tileZ : p.tileY;
  const int ntx = (p.nkx + tile - 1) / tile;
  const int threads = fftThreads<NFIX>();
  const int nblocks = persistentGrid((const void *)kern, smem, ntx * nOther, threads);
  if (AXIS_IS_Z) kern<<<nblocks, threads, smem, st>>>(addr, p.nz, p.nkx, nOther, tile, p.az, p.twz.template as<C>(), op);
  else kern<<<nblocks, threads, smem, st>>>(addr, p.ny, p.nkx, nOther, tile, p.ay, p.twy.template as<C>(), op);
  UB200_LAUNCHED();
  return UB200_OK;
}
template <class T, int MODE, bool AXIS_IS_Z, class Op, class Addr>
int launchPassAddr(const Fft3dPlan<T> &p, const Addr &addr, int nOther, cudaStream_t st, Op op = Op()) {
  int rc = UB200_OK;
  UB200_FFT_DISPATCH((AXIS_IS_Z ? p.nz : p.ny), (rc = launchPassAddrFixed<T, MODE, AXIS_IS_Z, Op, NFIX, Addr>(p, addr, nOther, st, op)));
  return rc;
}

template <class T, int MODE, class Op = NoSpectralOp, int NC = 3>
int launchPassY(const Fft3dPlan<T, NC> &p, void *grid, cudaStream_t st, Op op = Op()) {
  int rc = UB200_OK;
  UB200_FFT_DISPATCH(p.ny, (rc = launchPassStridedFixed<T, MODE, false, Op, NFIX, NC>(p, grid, st, op)));
  return rc;
}

template <class T, int MODE, class Op = NoSpectralOp, int NC = 3>
int launchPassZ(const Fft3dPlan<T, NC> &p, void *grid, cudaStream_t st, Op op = Op()) {
  int rc = UB200_OK;
  UB200_FFT_DISPATCH(p.nz, (rc = launchPassStridedFixed<T, MODE, true, Op, NFIX, NC>(p, grid, st, op)));
  return rc;
}

} // namespace ub200
