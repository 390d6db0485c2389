// uammd_b200 internal helpers shared by all kernels (sm_100a only).
#pragma once
#include "../../include/uammd_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <new>
#include <cmath>
#include <vector>

namespace ub200 {

extern thread_local int g_lastCudaError;
extern unsigned long long g_launchCount;

inline int cudaFail(cudaError_t e) {
  g_lastCudaError = (int)e;
  return UB200_ERR_CUDA;
}
#define UB200_CUDA(call)                                                                                     \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) return ::ub200::cudaFail(e__);                                                   \
  } while (0)
// count + check a kernel launch (no sync)
#define UB200_LAUNCHED()                                                                                     \
  do {                                                                                                       \
    ::ub200::g_launchCount++;                                                                                \
    cudaError_t e__ = cudaPeekAtLastError();                                                                 \
    if (e__ != cudaSuccess) return ::ub200::cudaFail(e__);                                                   \
  } while (0)

constexpr int kNumSMs = 148; // B200

// Device buffer that only ever grows (scratch owned by a handle; no allocation on the steady-state path).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return UB200_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      g_lastCudaError = (int)e;
      return UB200_ERR_ALLOC;
    }
    cap = want;
    return UB200_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Box + Grid in the reference's single precision arithmetic (utils/Box.cuh:16-36, utils/Grid.cuh:21-48).
// All derived quantities are computed on the host in fp32 exactly like the reference constructors do.
struct GridF {
  float Lx, Ly, Lz;
  float mx, my, mz; // minusInvBoxSize, 0 when the dimension is not periodic
  float hLx, hLy, hLz; // 0.5*L (exact)
  float ix, iy, iz; // invCellSize
  float csx, csy, csz; // cellSize
  int nx, ny, nz;
  // cell list grids: rank[linear cell] = position of the cell in Morton order (the bin of its particles); filled by the
  // cell list build, null elsewhere
  const uint32_t *rank;
};

inline GridF makeGridF(const float L[3], const int periodic[3], const int cellDim[3]) {
  GridF g;
  float m[3], inv[3], csz[3];
  int n[3];
  for (int d = 0; d < 3; d++) {
    m[d] = -1.0f / L[d];
    if (L[d] == 0.0f || isinf(L[d]) || !periodic[d]) m[d] = 0.0f;
    n[d] = cellDim[d];
    if (d == 2 && n[d] == 0) n[d] = 1;
    const float cs = L[d] / (float)n[d];
    inv[d] = 1.0f / cs;
    csz[d] = cs;
  }
  if (L[2] == 0.0f) inv[2] = 0.0f;
  g.Lx = L[0]; g.Ly = L[1]; g.Lz = L[2];
  g.mx = m[0]; g.my = m[1]; g.mz = m[2];
  g.hLx = 0.5f * L[0]; g.hLy = 0.5f * L[1]; g.hLz = 0.5f * L[2];
  g.ix = inv[0]; g.iy = inv[1]; g.iz = inv[2];
  g.csx = csz[0]; g.csy = csz[1]; g.csz = csz[2];
  g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
  g.rank = nullptr;
  return g;
}

// Box::apply_pbc for one coordinate (utils/Box.cuh:51-58). The reference is compiled with nvcc's default
// -fmad=true, which contracts r*minusInvL+0.5 and r+offset*L into FMAs; we spell the FMAs out so that the
// cell of a particle is decided by the same roundings.
__device__ __forceinline__ float foldCoord(float r, float L, float minusInvL) {
  const float offset = floorf(__fmaf_rn(r, minusInvL, 0.5f));
  return (minusInvL != 0.0f) ? __fmaf_rn(offset, L, r) : r;
}
// Grid::getCell for one coordinate (utils/Grid.cuh:49-71): trunc((fold(r) + 0.5 L) * invCellSize), n -> 0.
__device__ __forceinline__ int cellCoord(float r, float L, float minusInvL, float halfL, float invCell, int n) {
  const float rf = foldCoord(r, L, minusInvL);
  int c = __float2int_rz(__fmul_rn(__fadd_rn(rf, halfL), invCell));
  return (c == n) ? 0 : c;
}

// Sorter::MortonHash (utils/ParticleSorter.cuh:51-76): 10 bits per dimension, x in the lowest bit.
__host__ __device__ __forceinline__ uint32_t spreadBits10(uint32_t v) {
  uint32_t x = v & 0x3ffu;
  x = (x | (x << 16)) & 0x30000ffu;
  x = (x | (x << 8)) & 0x300f00fu;
  x = (x | (x << 4)) & 0x30c30c3u;
  x = (x | (x << 2)) & 0x9249249u;
  return x;
}
__host__ __device__ __forceinline__ uint32_t mortonCode(int cx, int cy, int cz) {
  return spreadBits10((uint32_t)cx) | (spreadBits10((uint32_t)cy) << 1) | (spreadBits10((uint32_t)cz) << 2);
}
// Bin of cell (cx, cy, cz) of a built cell list: the rank of its Morton code among the cells of the grid. Particles sorted
// by bin are sorted by Morton hash (ParticleSorter.cuh:102-111) while the bin table has exactly one entry per cell,
// whatever the shape of the grid.
__device__ __forceinline__ uint32_t cellBin(const GridF &g, int cx, int cy, int cz) {
  return __ldg(g.rank + (cx + g.nx * (cy + g.ny * cz)));
}
__host__ __device__ __forceinline__ uint32_t compactBits10(uint32_t x) {
  x &= 0x9249249u;
  x = (x | (x >> 2)) & 0x30c30c3u;
  x = (x | (x >> 4)) & 0x300f00fu;
  x = (x | (x >> 8)) & 0x30000ffu;
  x = (x | (x >> 16)) & 0x3ffu;
  return x;
}

// Periodic image of coordinate r closest to the reference point c (valid whenever the true separation is
// below L/2, i.e. for neighbour-cell candidates on grids with >= 4 cells per periodic dimension). Unlike
// "fold into the box, then add the image shift of the cell", this is robust for particles whose stored
// coordinate sits outside the geometric bounds of their cell (x == +L/2 is assigned to cell 0 by Grid::getCell).
__device__ __forceinline__ float imageNear(float r, float c, float L, float minusInvL) {
  const float off = floorf(__fmaf_rn(r - c, minusInvL, 0.5f));
  return __fmaf_rn(off, L, r);
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

struct LJPar {
  float cutOff2, sigma2, epsDivSigma2, shift; // LJFunctor::PairParameters (Potential/Potential.cuh:31-35)
};
struct LJTableCache {
  DevBuf dev;
  std::vector<float> host;
};

// defined in celllist.cu
int exclusiveScanAndClear(uint32_t *counts, int M, uint32_t *out, uint32_t *tileSums, cudaStream_t st);
int scatterToBinsLaunch(const uint2 *codeSlot, const uint32_t *binStart, int N, int *unstable, cudaStream_t st);

} // namespace ub200

struct ub200_celllist;
struct ub200_verletlist;
namespace ub200 {
// DPD forces over a built cell list (pair_dpd.cu); ownerHiDev: optional device-side upper bound of the owned index range
int dpdSum(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut, uint32_t seed, uint32_t step,
           int idStride, void *d_force, const int *d_globalIdx, int ownerLo, int ownerHi, int accumulate, void *stream,
           const int *d_noiseId = nullptr, const int *ownerHiDev = nullptr);
// ub200_celllist_build_f32 with an optional device-side particle count (N = launch bound) and an optional key deciding the
// order inside a cell (multi-GPU bricks: global particle ids)
int celllistBuildEx(ub200_celllist *cl, const void *d_pos, const int *d_groupIdx, int N, const int *nDev, const int *sortKey,
                    const float L[3], const int periodic[3], const int cellDim[3], void *stream);
// lj_vlist.cu: row list over the half-cell columns
bool vlistApplies(const float L[3], const int periodic[3], float rcut, int N);
int vlistRebuild(ub200_verletlist *v, cudaStream_t st);
int vlistRefreshPositions(ub200_verletlist *v, const float4 *pos, const int *groupIdx, cudaStream_t st);
int vlistRefreshAndCheck(ub200_verletlist *v, const float4 *pos, const int *groupIdx, float maxDist, bool *over, cudaStream_t st);
int vlistSum(ub200_verletlist *v, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
             const int *globalIdx, bool accumulate, cudaStream_t st);
// LJ forces over a Verlet list handle, whichever list it holds (pair_lj.cu)
int ljVerletSum(ub200_verletlist *vl, const float *params, int ntypes, float4 *force, float *energy, float *virial,
                const int *globalIdx, bool accumulate, cudaStream_t st);
} // namespace ub200

// Opaque handle behind ub200_celllist
struct ub200_celllist {
  ub200::GridF grid;
  int N = 0;
  int ncells = 0;
  int nbins = 0;       // = ncells: one bin per cell, in Morton order
  int built = 0;
  uint32_t validCell = 0; // VALID_CELL epoch (CellListBase.cuh:210-230)
  int validCounter = -1;
  int lastN = -1;
  int cellDim[3] = {0, 0, 0};
  ub200::DevBuf sortPos, groupIndex, cellStart, cellEnd; // reference-layout outputs
  ub200::DevBuf binCount, binStart, blockSums;           // Morton-code space histogram + scan
  ub200::DevBuf codeSlot;                                // per particle {code, slot}
  ub200::DevBuf unstable;                                // scatter target before the stable fix-up
  ub200::DevBuf errorFlag;
  ub200::DevBuf cellRank, cellOfRank;                    // uint32[ncells]: linear cell -> Morton rank and back (per grid shape)
  int rankDims[3] = {0, 0, 0};
  size_t cellStartCells = 0;
  ub200::LJTableCache ljTable;                           // parameter table of the LJ traversals over this list (per handle:
                                                         // two interactors never share or re-upload each other's table)
};

struct ub200_ljengine;
// Opaque handle behind ub200_verletlist (VerletList / VerletListBase / BasicNeighbourListBase of the reference)
struct ub200_verletlist {
  ub200_celllist *cl = nullptr;     // cell list over the stored positions, cell size >= cutOff * multiplier
  ub200::DevBuf storedPos, sortPos; // float4[N]: positions at the last rebuild (group order) / current positions, sorted order
  ub200::DevBuf numberNeighbours;   // int[N]
  ub200::DevBuf neighbourList;      // int[(maxNeighbours + 1) * N], entry k of sorted particle i at [k * N + i]
  ub200::DevBuf flags;              // uint32[2]: {particles over the drift threshold, largest overflowing neighbour count}
  int N = 0;
  int maxNeighbours = 32;           // BasicNeighbourListBase::maxNeighboursPerParticle, grows by 32
  float multiplier = 1.08f;         // VerletListBase::verletRadiusMultiplier
  float cutOff = 0.f;
  float L[3] = {0.f, 0.f, 0.f};
  int periodic[3] = {1, 1, 1};
  bool forceNext = true;
  int stepsSinceLastUpdate = 0;
  int rebuilds = 0;
  // ---- row list of the built-in LJ traversal (lj_vlist.cu); the reference-layout arrays above are then built on demand
  ub200_ljengine *eng = nullptr;    // half-cell list of the stored positions, cells >= cutOff * multiplier / 2
  ub200::DevBuf fastPos;            // float4[N] current positions in half-cell order, image nearest the build-time one
  ub200::DevBuf fastList, fastCount;// int[N * fastStride] rows, int[N]
  int fastStride = 64;
  bool fast = false;                // the row list is the one the last rebuild made
  bool refValid = false;            // the reference-layout list matches the stored positions
  bool wantRef = false;             // somebody read the reference-layout list: keep it current from now on
  uint32_t driftEpoch = 0;          // value the fused refresh + drift check writes into flags[0] when a particle is over
  bool refOnly = false;             // owner reads the reference-layout arrays directly (PSE near field): no row list
  const void *lastPos = nullptr;    // arguments of the last update (lazy build of the reference layout in view_get)
  const int *lastGroupIdx = nullptr;
  cudaStream_t lastStream = nullptr;
};
