// Cell list build for sm_100a: warp-aggregated counting sort over the cells in Morton order.
//
// Replaces the reference chain assignHash -> cub::DeviceRadixSort::SortPairs -> permutation gather ->
// fillCellList (utils/ParticleSorter.cuh:102-111,243-274,179-187; CellList/CellListBase.cuh:68-95) by
//   1. binParticles   : cell of each particle (reference-exact fp32 arithmetic), Morton code, slot in the
//                       bin from a warp-aggregated atomicAdd                       R 16 B  W 8 B / particle
//   2. scan (3 tiny kernels over the ncells bins)                                   ~4 B / cell
//   3. scatterToBins  : unstable[binStart[code]+slot] = i                           R 8 B   W 4 B
//   4. orderAndGather : restores the STABLE order inside each bin (rank = #smaller indices in the bin,
//                       bins hold ~N/ncells entries and sit in L1), writes groupIndex, gathers sortPos and
//                       emits cellStart/cellEnd in the reference layout            R 12+16 B W 20 B
// The result is bit-identical to the reference's stable radix sort by Morton hash. A bin is a CELL: its index is the
// rank of the cell's Morton code among the cells of the grid (a table per grid shape, made once with a radix sort of the
// ncells codes), so the bin table has ncells entries for any grid - 129^3 or 512 x 4 x 4 alike - instead of the
// 2^(bits of the largest code) of a table indexed by the code itself.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace ub200 {

thread_local int g_lastCudaError = 0;
unsigned long long g_launchCount = 0;

// out[k] = in[order[k]] for rows of `words` 32-bit words (ParticleSorter::applyCurrentOrder)
__global__ void __launch_bounds__(256)
applyOrderRows(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, const int *__restrict__ order, int N, int words) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= (size_t)N * words) return;
  const int k = (int)(t / words), w = (int)(t - (size_t)k * words);
  out[t] = in[(size_t)order[k] * words + w];
}

__global__ void __launch_bounds__(256)
binParticles(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, int N, const int *__restrict__ nDev, GridF g,
             uint32_t *__restrict__ binCount, uint2 *__restrict__ codeSlot, int *__restrict__ errorFlag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev); // optional device-side particle count (multi-GPU bricks)
  if (i >= N) return;
  const float4 p = ldg4(pos + (groupIdx ? groupIdx[i] : i));
  int cx = cellCoord(p.x, g.Lx, g.mx, g.hLx, g.ix, g.nx);
  int cy = cellCoord(p.y, g.Ly, g.my, g.hLy, g.iy, g.ny);
  int cz = cellCoord(p.z, g.Lz, g.mz, g.hLz, g.iz, g.nz);
  if ((unsigned)cx >= (unsigned)g.nx || (unsigned)cy >= (unsigned)g.ny || (unsigned)cz >= (unsigned)g.nz) {
    // outside a non periodic box (the reference raises errorFlag in fillCellList): flag and clamp
    *errorFlag = 1;
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
  }
  const uint32_t code = cellBin(g, cx, cy, cz);
  // warp-aggregated increment: one atomic per distinct bin per warp
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

// ---- exclusive scan over the bins: block sums -> top scan -> apply ----
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems; // 4096

__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, total in *total
template <int THREADS> __device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warpSums[THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t inc = warpInclusiveScan(v, lane);
  if (lane == 31) warpSums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < THREADS / 32 ? warpSums[lane] : 0;
    s = warpInclusiveScan(s, lane);
    if (lane < THREADS / 32) warpSums[lane] = s;
  }
  __syncthreads();
  const uint32_t warpOff = w ? warpSums[w - 1] : 0;
  *total = warpSums[THREADS / 32 - 1];
  return warpOff + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scanTileSums(const uint32_t *__restrict__ in, int M,
                                                             uint32_t *__restrict__ tileSums) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
  if (base + kScanItems <= M) {
    const uint4 a = *reinterpret_cast<const uint4 *>(in + base);
    const uint4 b = *reinterpret_cast<const uint4 *>(in + base + 4);
    s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  } else {
    for (int k = 0; k < kScanItems; k++)
      if (base + k < M) s += in[base + k];
  }
  uint32_t total;
  blockExclusiveScan<kScanThreads>(s, &total);
  if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// single block: exclusive scan of up to 4096 tile sums in place
__global__ void __launch_bounds__(1024) scanTop(uint32_t *__restrict__ tileSums, int ntiles) {
  uint32_t v[4];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int idx = threadIdx.x * 4 + k;
    v[k] = idx < ntiles ? tileSums[idx] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t ex = blockExclusiveScan<1024>(s, &total);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int idx = threadIdx.x * 4 + k;
    if (idx < ntiles) tileSums[idx] = ex;
    ex += v[k];
  }
}

// out[i] = exclusive prefix; out[M] = grand total. Also re-zeroes `in` for the next build.
__global__ void __launch_bounds__(kScanThreads) scanApply(uint32_t *__restrict__ in, int M,
                                                          const uint32_t *__restrict__ tileOffsets,
                                                          uint32_t *__restrict__ out) {
  const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  static_assert(kScanItems == 8, "vector path below moves 2 x uint4 per thread");
  uint32_t v[kScanItems];
  uint32_t s = 0;
  const bool full = base + kScanItems <= M; // whole 32-byte run of this thread inside the array: 128-bit accesses
  if (full) {
    const uint4 a = *reinterpret_cast<const uint4 *>(in + base), b = *reinterpret_cast<const uint4 *>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) s += v[k];
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      v[k] = (base + k < M) ? in[base + k] : 0;
      s += v[k];
    }
  }
  uint32_t total;
  uint32_t ex = blockExclusiveScan<kScanThreads>(s, &total) + tileOffsets[blockIdx.x];
  if (full) {
    uint32_t o[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { o[k] = ex; ex += v[k]; }
    *reinterpret_cast<uint4 *>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4 *>(out + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    *reinterpret_cast<uint4 *>(in + base) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(in + base + 4) = make_uint4(0u, 0u, 0u, 0u);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      if (base + k < M) {
        out[base + k] = ex;
        in[base + k] = 0;
      }
      ex += v[k];
    }
  }
  if (base <= M - 1 && M - 1 < base + kScanItems) out[M] = ex; // thread owning the last item: ex == grand total
}

__global__ void __launch_bounds__(256)
scatterToBins(const uint2 *__restrict__ codeSlot, const uint32_t *__restrict__ binStart, int N,
              int *__restrict__ unstable, const int *__restrict__ nDev = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev);
  if (i >= N) return;
  const uint2 cs = codeSlot[i];
  if (cs.x == 0xffffffffu) return; // particle outside the caller's window (slab-decomposed IBM)
  unstable[binStart[cs.x] + cs.y] = i;
}

__global__ void __launch_bounds__(256)
orderAndGather(const int *__restrict__ unstable, const uint2 *__restrict__ codeSlot,
               const uint32_t *__restrict__ binStart, const float4 *__restrict__ pos,
               const int *__restrict__ groupIdx, int N, GridF g, uint32_t validCell,
               float4 *__restrict__ sortPos, int *__restrict__ groupIndex, uint32_t *__restrict__ cellStart,
               int *__restrict__ cellEnd, const uint32_t *__restrict__ cellOfRank, const int *__restrict__ nDev = nullptr,
               const int *__restrict__ sortKey = nullptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev);
  if (k >= N) return;
  const int i = unstable[k];
  const uint32_t code = codeSlot[i].x;
  const int s = (int)binStart[code], e = (int)binStart[code + 1];
  int rank = 0;
  if (sortKey) { // order inside a cell by a caller-supplied key (global particle ids of a multi-GPU brick)
    const int ki = sortKey[i];
    for (int j = s; j < e; j++) rank += (sortKey[__ldg(unstable + j)] < ki);
  } else {
    for (int j = s; j < e; j++) rank += (__ldg(unstable + j) < i);
  }
  const int dst = s + rank;
  groupIndex[dst] = i;
  sortPos[dst] = ldg4(pos + (groupIdx ? groupIdx[i] : i));
  if (rank == 0) {
    const int lin = (int)cellOfRank[code];
    cellStart[lin] = (uint32_t)s + validCell;
    cellEnd[lin] = e;
  }
}

// Exclusive prefix sum of counts[0..M) into out[0..M], out[M] = total; counts are zeroed again on the way.
// tileSums: scratch of >= 4096 uint32. Three tiny launches, no host synchronisation.
int exclusiveScanAndClear(uint32_t *counts, int M, uint32_t *out, uint32_t *tileSums, cudaStream_t st) {
  const int ntiles = (M + kScanTile - 1) / kScanTile;
  if (ntiles > 4096) return UB200_ERR_GRID_TOO_LARGE;
  scanTileSums<<<ntiles, kScanThreads, 0, st>>>(counts, M, tileSums);
  UB200_LAUNCHED();
  scanTop<<<1, 1024, 0, st>>>(tileSums, ntiles);
  UB200_LAUNCHED();
  scanApply<<<ntiles, kScanThreads, 0, st>>>(counts, M, tileSums, out);
  UB200_LAUNCHED();
  return UB200_OK;
}

int scatterToBinsLaunch(const uint2 *codeSlot, const uint32_t *binStart, int N, int *unstable, cudaStream_t st) {
  scatterToBins<<<(N + 255) / 256, 256, 0, st>>>(codeSlot, binStart, N, unstable);
  UB200_LAUNCHED();
  return UB200_OK;
}

static int highestBit(uint32_t v) { // position of MSB + 1 (0 for v == 0)
  int b = 0;
  while (v) { b++; v >>= 1; }
  return b;
}

} // namespace ub200

using namespace ub200;

extern "C" {

int ub200_celllist_create(ub200_celllist **out) {
  if (!out) return UB200_ERR_INVALID_ARGUMENT;
  *out = new (std::nothrow) ub200_celllist();
  return *out ? UB200_OK : UB200_ERR_ALLOC;
}

int ub200_celllist_destroy(ub200_celllist *cl) {
  if (!cl) return UB200_OK;
  DevBuf *bufs[] = {&cl->sortPos, &cl->groupIndex, &cl->cellStart, &cl->cellEnd, &cl->binCount,
                    &cl->binStart, &cl->blockSums, &cl->codeSlot, &cl->unstable, &cl->errorFlag, &cl->ljTable.dev, &cl->cellRank, &cl->cellOfRank};
  for (DevBuf *b : bufs) b->release();
  delete cl;
  return UB200_OK;
}

int ub200_neighbour_celldim_f32(const float L[3], float rc, int cellDim[3]) {
  if (!L || !cellDim || !(rc > 0)) return UB200_ERR_INVALID_ARGUMENT;
  for (int d = 0; d < 3; d++) {
    int c = (int)(L[d] / rc); // Grid(Box, real3 minCellSize): make_int3(boxSize/minCellSize), utils/Grid.cuh:31-33
    if (c <= 3) c = 1;        // CellList.cuh:117-122
    cellDim[d] = c;
  }
  return UB200_OK;
}

// ParticleData::sortParticles (ParticleData/ParticleData.cuh:492-522): the order of the stable sort by the Morton hash of the
// cell a particle lies in, cells of hash_box / hash_cutOff per dimension (truncated) - the sort the cell list performs
int ub200_particles_sort_order_f32(ub200_celllist *scratch, const void *d_pos, int N, const float L[3], const int periodic[3],
                                   float hashCutOff, int *d_order, void *stream) {
  if (!scratch || !d_pos || !d_order || N <= 0 || !L || !periodic || !(hashCutOff > 0.f)) return UB200_ERR_INVALID_ARGUMENT;
  int cd[3];
  for (int d = 0; d < 3; d++) {
    cd[d] = (int)(L[d] / hashCutOff); // make_int3(hints.hash_box.boxSize / hints.hash_cutOff)
    if (cd[d] < 1) cd[d] = 1;
    if (cd[d] > 1024) return UB200_ERR_GRID_TOO_LARGE;
  }
  const int rc = ub200_celllist_build_f32(scratch, d_pos, nullptr, N, L, periodic, cd, stream);
  if (rc) return rc;
  UB200_CUDA(cudaMemcpyAsync(d_order, scratch->groupIndex.p, sizeof(int) * (size_t)N, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return UB200_OK;
}

// ParticleSorter::applyCurrentOrder (utils/ParticleSorter.cuh:177-187): out[k] = in[order[k]], rows of rowBytes (a multiple of 4)
int ub200_apply_order(const void *d_in, void *d_out, const int *d_order, int N, int rowBytes, void *stream) {
  if (!d_in || !d_out || !d_order || N <= 0 || rowBytes <= 0 || rowBytes % 4 || d_in == d_out) return UB200_ERR_INVALID_ARGUMENT;
  const int words = rowBytes / 4;
  const size_t total = (size_t)N * words;
  ub200::applyOrderRows<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t *)d_in, (uint32_t *)d_out,
                                                                                          d_order, N, words);
  UB200_LAUNCHED();
  return UB200_OK;
}

int ub200_celllist_build_f32(ub200_celllist *cl, const void *d_pos, const int *d_groupIdx, int N,
                             const float L[3], const int periodic[3], const int cellDim[3], void *stream) {
  return ub200::celllistBuildEx(cl, d_pos, d_groupIdx, N, nullptr, nullptr, L, periodic, cellDim, stream);
}
}

// N: launch bound; nDev: optional device-side particle count (<= N); sortKey: optional order key inside a cell
namespace ub200 {
__global__ void __launch_bounds__(256) cellCodes(GridF g, int ncells, uint32_t *__restrict__ code, uint32_t *__restrict__ cell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  code[c] = mortonCode(c % g.nx, (c / g.nx) % g.ny, c / (g.nx * g.ny));
  cell[c] = (uint32_t)c;
}
__global__ void __launch_bounds__(256) cellRanks(const uint32_t *__restrict__ cellOfRank, int ncells, uint32_t *__restrict__ rank) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < ncells) rank[cellOfRank[r]] = (uint32_t)r;
}
// rank tables of a grid shape (Sorter::MortonHash order of the cells, utils/ParticleSorter.cuh:51-76); runs when the shape
// changes, not per build
static int buildCellRanks(ub200_celllist *cl, const GridF &g, int ncells, cudaStream_t st) {
  int rc;
  if ((rc = cl->cellRank.reserve(sizeof(uint32_t) * (size_t)ncells)) || (rc = cl->cellOfRank.reserve(sizeof(uint32_t) * (size_t)ncells)))
    return rc;
  DevBuf codeIn, codeOut, cellIn, temp;
  if ((rc = codeIn.reserve(sizeof(uint32_t) * (size_t)ncells)) || (rc = codeOut.reserve(sizeof(uint32_t) * (size_t)ncells)) ||
      (rc = cellIn.reserve(sizeof(uint32_t) * (size_t)ncells)))
    return rc;
  cellCodes<<<(ncells + 255) / 256, 256, 0, st>>>(g, ncells, codeIn.as<uint32_t>(), cellIn.as<uint32_t>());
  UB200_LAUNCHED();
  size_t bytes = 0;
  UB200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, codeIn.as<uint32_t>(), codeOut.as<uint32_t>(), cellIn.as<uint32_t>(),
                                             cl->cellOfRank.as<uint32_t>(), ncells, 0, 30, st));
  if ((rc = temp.reserve(bytes ? bytes : 16))) return rc;
  UB200_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, bytes, codeIn.as<uint32_t>(), codeOut.as<uint32_t>(), cellIn.as<uint32_t>(),
                                             cl->cellOfRank.as<uint32_t>(), ncells, 0, 30, st));
  cellRanks<<<(ncells + 255) / 256, 256, 0, st>>>(cl->cellOfRank.as<uint32_t>(), ncells, cl->cellRank.as<uint32_t>());
  UB200_LAUNCHED();
  UB200_CUDA(cudaStreamSynchronize(st)); // the scratch buffers go out of scope
  codeIn.release(); codeOut.release(); cellIn.release(); temp.release();
  return UB200_OK;
}
} // namespace ub200

int ub200::celllistBuildEx(ub200_celllist *cl, const void *d_pos, const int *d_groupIdx, int N, const int *nDev, const int *sortKey,
                           const float L[3], const int periodic[3], const int cellDim[3], void *stream) {
  if (!cl || !d_pos || N <= 0 || !L || !periodic || !cellDim) return UB200_ERR_INVALID_ARGUMENT;
  for (int d = 0; d < 3; d++)
    if (cellDim[d] < 1 || cellDim[d] > 1024) return UB200_ERR_INVALID_ARGUMENT; // 10 bit Morton fields
  cudaStream_t st = (cudaStream_t)stream;
  GridF g = makeGridF(L, periodic, cellDim);
  const long long ncellsLong = (long long)g.nx * g.ny * g.nz;
  // one bin per cell; the two-level scan covers 4096 tiles of 4096 bins
  if (ncellsLong > (long long)kScanTile * 4096) return UB200_ERR_GRID_TOO_LARGE;
  const int ncells = (int)ncellsLong;
  const int nbins = ncells;

  int rc;
  if (cl->rankDims[0] != g.nx || cl->rankDims[1] != g.ny || cl->rankDims[2] != g.nz || !cl->cellRank.p) {
    if ((rc = buildCellRanks(cl, g, ncells, st))) return rc;
    cl->rankDims[0] = g.nx; cl->rankDims[1] = g.ny; cl->rankDims[2] = g.nz;
  }
  g.rank = cl->cellRank.as<uint32_t>();
  if ((rc = cl->sortPos.reserve(sizeof(float4) * (size_t)N))) return rc;
  if ((rc = cl->groupIndex.reserve(sizeof(int) * (size_t)N))) return rc;
  if ((rc = cl->codeSlot.reserve(sizeof(uint2) * (size_t)N))) return rc;
  if ((rc = cl->unstable.reserve(sizeof(int) * (size_t)N))) return rc;
  if ((rc = cl->blockSums.reserve(sizeof(uint32_t) * 4096))) return rc;
  if (!cl->errorFlag.p) {
    if ((rc = cl->errorFlag.reserve(sizeof(int)))) return rc;
    UB200_CUDA(cudaMemsetAsync(cl->errorFlag.p, 0, sizeof(int), st));
  }
  if (cl->nbins != nbins || !cl->binCount.p) {
    if ((rc = cl->binCount.reserve(sizeof(uint32_t) * (size_t)nbins))) return rc;
    if ((rc = cl->binStart.reserve(sizeof(uint32_t) * ((size_t)nbins + 1)))) return rc;
    // binCount is re-zeroed by scanApply at every build; zero it once here
    UB200_CUDA(cudaMemsetAsync(cl->binCount.p, 0, sizeof(uint32_t) * (size_t)nbins, st));
  }
  // CellListBase::tryToResizeCellListToCurrentGrid (CellListBase.cuh:186-200): cellStart zero-filled on resize
  bool resized = false;
  if (cl->cellStartCells != (size_t)ncells) {
    if ((rc = cl->cellStart.reserve(sizeof(uint32_t) * (size_t)ncells))) return rc;
    if ((rc = cl->cellEnd.reserve(sizeof(int) * (size_t)ncells))) return rc;
    UB200_CUDA(cudaMemsetAsync(cl->cellStart.p, 0, sizeof(uint32_t) * (size_t)ncells, st));
    cl->cellStartCells = (size_t)ncells;
    resized = true;
  }
  (void)resized;
  // CellListBase::updateCurrentValidCell (CellListBase.cuh:210-230)
  if (N != cl->lastN) cl->validCounter = -1;
  const unsigned long long nextMax = (unsigned long long)N * (unsigned long long)(cl->validCounter + 2);
  if (cl->validCounter < 0 || nextMax >= 0xFFFFFFFFull - 1ull) {
    cl->validCell = (uint32_t)N;
    cl->validCounter = 1;
    UB200_CUDA(cudaMemsetAsync(cl->cellStart.p, 0, sizeof(uint32_t) * (size_t)ncells, st));
  } else {
    cl->validCounter++;
    cl->validCell = (uint32_t)N * (uint32_t)cl->validCounter;
  }
  cl->lastN = N;
  cl->grid = g;
  cl->N = N;
  cl->ncells = ncells;
  cl->nbins = nbins;
  for (int d = 0; d < 3; d++) cl->cellDim[d] = cellDim[d];

  const int nb = (N + 255) / 256;
  binParticles<<<nb, 256, 0, st>>>((const float4 *)d_pos, d_groupIdx, N, nDev, g, cl->binCount.as<uint32_t>(),
                                   cl->codeSlot.as<uint2>(), cl->errorFlag.as<int>());
  UB200_LAUNCHED();
  if ((rc = exclusiveScanAndClear(cl->binCount.as<uint32_t>(), nbins, cl->binStart.as<uint32_t>(),
                                  cl->blockSums.as<uint32_t>(), st)))
    return rc;
  scatterToBins<<<nb, 256, 0, st>>>(cl->codeSlot.as<uint2>(), cl->binStart.as<uint32_t>(), N, cl->unstable.as<int>(), nDev);
  UB200_LAUNCHED();
  orderAndGather<<<nb, 256, 0, st>>>(cl->unstable.as<int>(), cl->codeSlot.as<uint2>(), cl->binStart.as<uint32_t>(),
                                     (const float4 *)d_pos, d_groupIdx, N, g, cl->validCell,
                                     cl->sortPos.as<float4>(), cl->groupIndex.as<int>(),
                                     cl->cellStart.as<uint32_t>(), cl->cellEnd.as<int>(), cl->cellOfRank.as<uint32_t>(), nDev,
                                     sortKey);
  UB200_LAUNCHED();
  cl->built = 1;
  return UB200_OK;
}

extern "C" {

int ub200_celllist_view_get(ub200_celllist *cl, ub200_celllist_view *v) {
  if (!cl || !v) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl->built) return UB200_ERR_NOT_BUILT;
  v->d_cellStart = cl->cellStart.as<uint32_t>();
  v->d_cellEnd = cl->cellEnd.as<int>();
  v->d_sortPos = cl->sortPos.p;
  v->d_groupIndex = cl->groupIndex.as<int>();
  v->VALID_CELL = cl->validCell;
  for (int d = 0; d < 3; d++) v->cellDim[d] = cl->cellDim[d];
  v->numberParticles = cl->N;
  v->d_binStart = cl->binStart.as<uint32_t>();
  v->nbins = cl->nbins;
  return UB200_OK;
}

int ub200_celllist_error_flag(ub200_celllist *cl, void *stream, int *flag) {
  if (!cl || !flag) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl->built) return UB200_ERR_NOT_BUILT;
  UB200_CUDA(cudaMemcpyAsync(flag, cl->errorFlag.p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  UB200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return UB200_OK;
}

const char *ub200_error_string(int code) {
  switch (code) {
  case UB200_OK: return "ok";
  case UB200_ERR_INVALID_ARGUMENT: return "invalid argument";
  case UB200_ERR_CUDA: return "CUDA runtime error (see ub200_last_cuda_error)";
  case UB200_ERR_ALLOC: return "device allocation failed";
  case UB200_ERR_GRID_TOO_LARGE: return "cell grid too large for the Morton bin table";
  case UB200_ERR_NOT_BUILT: return "cell list has not been built";
  case UB200_ERR_UNSUPPORTED: return "unsupported configuration";
  default: return "unknown error";
  }
}
int ub200_last_cuda_error(void) { return g_lastCudaError; }
const char *ub200_version(void) { return "uammd_b200 0.1 sm_100a"; }
unsigned long long ub200_launch_count(void) { return g_launchCount; }
}
