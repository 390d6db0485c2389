// Verlet (skin) neighbour list for sm_100a. Replaces VerletList::update (Interactor/NeighbourList/VerletList.cuh:111-124),
// VerletListBase::{update,needsRebuild,isParticleDriftOverThreshold,updateSortedPositions}
// (VerletList/VerletListBase.cuh:107-199) and BasicNeighbourListBase::{update,fillBasicNeighbourList}
// (BasicList/BasicListBase.cuh:42-71,131-215). The list is BIT-IDENTICAL to the reference's: same cell list, same
// visiting order (27 cells x fastest, ascending sorted index inside a cell, self included), same "<= cutOff^2" test on the
// minimum-image separation, same [k*N + i] layout - so the reference's own VerletListBase_ns::NeighbourContainer and
// user transversers run on it unchanged, while the LJ fast path (ljVerletTraversal, pair_lj.cu) reads it directly.
//
// Build: one WARP per home cell; the candidates of the 27 neighbour cells are staged once in shared memory, every
// home particle is tested by the 32 lanes in visiting order and the hits are compacted with ballot/popc, so the
// list comes out ordered without any sort (the reference walks the 27 cells with one thread per particle).
#include "pair_common.cuh"
#include "lj_engine.cuh"

namespace ub200 {

// VerletListBase_ns::checkMaximumDrift (VerletListBase.cuh:57-69)
__global__ void __launch_bounds__(256)
verletDriftCheck(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, const float4 *__restrict__ stored, int N,
                 GridF g, float maxDist2, uint32_t *__restrict__ flag) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  const float4 c = ldg4(pos + (groupIdx ? groupIdx[id] : id)), p = ldg4(stored + id);
  const float dx = foldCoord(c.x - p.x, g.Lx, g.mx), dy = foldCoord(c.y - p.y, g.Ly, g.my), dz = foldCoord(c.z - p.z, g.Lz, g.mz);
  const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
  if (r2 >= maxDist2) atomicAdd(flag, 1u);
}

// storedPos[i] = pos[group[i]] (storeCurrentPos :143-150)
__global__ void __launch_bounds__(256)
verletStore(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, int N, float4 *__restrict__ stored) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  stored[id] = ldg4(pos + (groupIdx ? groupIdx[id] : id));
}
// sortPos[k] = pos[group[groupIndex[k]]] (updateSortedPositions :152-163), every step
__global__ void __launch_bounds__(256)
verletSortedPositions(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, const int *__restrict__ groupIndex, int N,
                      float4 *__restrict__ sortPos) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = groupIndex[k];
  sortPos[k] = ldg4(pos + (groupIdx ? groupIdx[i] : i));
}

constexpr int kVerletCap = 640;   // staged candidates per warp
constexpr int kVerletHome = 20;   // home particles whose lists are assembled in shared memory per pass
constexpr int kVerletK = 96;      // list entries per home particle assembled in shared memory
constexpr size_t kVerletWarpBytes = kVerletCap * sizeof(float4) + kVerletHome * kVerletK * sizeof(unsigned short) + kVerletHome * sizeof(int);
constexpr size_t kVerletSmem = kPairWarps * ((kVerletWarpBytes + 15) / 16 * 16);

// exact test of the reference: dot(apply_pbc(pj - pi)) <= cutOff2 with its roundings (BasicListBase.cuh:56-58)
__device__ __forceinline__ bool verletHitExact(const float4 pi, const float4 pj, const GridF &g, float cutOff2) {
  const float dx = foldCoord(pj.x - pi.x, g.Lx, g.mx), dy = foldCoord(pj.y - pi.y, g.Ly, g.my), dz = foldCoord(pj.z - pi.z, g.Lz, g.mz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) <= cutOff2;
}

// BasicNeighbourList_ns::fillBasicNeighbourList (BasicListBase.cuh:42-71). One warp per home cell. The candidates of
// the 27 neighbour cells are staged once (flat, in the reference's visiting order, already moved to the periodic image
// nearest the home cell, their sorted index in .w); every home particle is tested against 32 candidates per iteration
// and the hits are compacted with ballot/popc, which keeps the visiting order without a sort. The staged image makes
// the common test 7 instructions; a pair within 1e-4 (relative) of the cut-off is re-tested with the reference's exact
// arithmetic on the original coordinates, so the list stays bit-identical. The lists of the cell's home particles
// are assembled in shared memory as staged indices and written out TRANSPOSED: in the reference's [k*N + i] layout
// the entries k of the cell's consecutive home particles are contiguous, so a k-row goes out as one run instead of
// one 4-byte store per 32-byte sector.
__global__ void __launch_bounds__(kPairThreads)
verletFill(const float4 *__restrict__ sortPos, const uint32_t *__restrict__ binStart, GridF g, int ncells, float cutOff2, int N,
           int maxNeighbours, int *__restrict__ neighbourList, int *__restrict__ numberNeighbours, uint32_t *__restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *mine = smemRaw + (size_t)warp * ((kVerletWarpBytes + 15) / 16 * 16);
  float4 *cand = reinterpret_cast<float4 *>(mine);
  unsigned short(*lbuf)[kVerletK] = reinterpret_cast<unsigned short(*)[kVerletK]>(mine + kVerletCap * sizeof(float4));
  int *cntBuf = reinterpret_cast<int *>(mine + kVerletCap * sizeof(float4) + kVerletHome * kVerletK * sizeof(unsigned short));
  const int warpsTotal = gridDim.x * kPairWarps;
  const bool pairMic = (g.mx != 0.0f && g.nx < 4) || (g.my != 0.0f && g.ny < 4) || (g.mz != 0.0f && g.nz < 4);
  // rounding band around the cut-off inside which the reference's exact arithmetic decides: the staged image of a
  // coordinate carries an error of a few ulp(L / 2), i.e. a relative error ~ 2^-21 L / cutOff in r2; 1e-4 covers boxes up to
  // ~200 cut-offs, larger ones widen the band
  const float band = fmaxf(1e-4f, 4.8e-7f * fmaxf(g.Lx, fmaxf(g.Ly, g.Lz)) * rsqrtf(cutOff2));
  const float cutLo = cutOff2 * (1.0f - band), cutHi = cutOff2 * (1.0f + band);
  for (int cell = blockIdx.x * kPairWarps + warp; cell < ncells; cell += warpsTotal) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue;
    const bool staged = nc.total <= kVerletCap && !pairMic;
    const float3 hc = cellCentre(g, cx, cy, cz);
    __syncwarp();
    if (staged) {
      for (int c = 0; c < 27; c++) {
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        const int off = __shfl_sync(0xffffffffu, nc.off, c);
        for (int t = lane; t < cnt; t += 32) {
          float4 p = ldg4(sortPos + st + t);
          toHomeImage(p, g, hc);
          p.w = __int_as_float(st + t);
          cand[off + t] = p;
        }
      }
      __syncwarp();
      for (int h0 = 0; h0 < hCount; h0 += kVerletHome) {
        const int nh = min(kVerletHome, hCount - h0);
        for (int h = 0; h < nh; h += 2) { // two home particles per pass: one shared-memory load feeds two tests
          const bool two = h + 1 < nh;
          const int id0 = hStart + h0 + h, id1 = id0 + (two ? 1 : 0);
          const float4 raw0 = ldg4(sortPos + id0), raw1 = ldg4(sortPos + id1);
          float4 p0 = raw0, p1 = raw1;
          toHomeImage(p0, g, hc);
          toHomeImage(p1, g, hc);
          int n0 = 0, n1 = 0; // warp uniform
          bool over0 = false, over1 = !two;
          for (int t0 = 0; t0 < nc.total && !(over0 && over1); t0 += 32) {
            const int t = t0 + lane;
            bool hit0 = false, hit1 = false;
            if (t < nc.total) {
              const float4 pj = cand[t];
              const float ax = pj.x - p0.x, ay = pj.y - p0.y, az = pj.z - p0.z;
              const float bx = pj.x - p1.x, by = pj.y - p1.y, bz = pj.z - p1.z;
              const float ra = __fmaf_rn(az, az, __fmaf_rn(ay, ay, ax * ax)), rb = __fmaf_rn(bz, bz, __fmaf_rn(by, by, bx * bx));
              hit0 = ra <= cutLo;
              hit1 = rb <= cutLo;
              if ((!hit0 && ra <= cutHi) || (!hit1 && rb <= cutHi)) { // rare: within 1e-4 of the cut-off -> the reference's arithmetic
                const float4 pjRaw = ldg4(sortPos + __float_as_int(pj.w));
                if (!hit0 && ra <= cutHi) hit0 = verletHitExact(raw0, pjRaw, g, cutOff2);
                if (!hit1 && rb <= cutHi) hit1 = verletHitExact(raw1, pjRaw, g, cutOff2);
              }
            }
            hit0 = hit0 && !over0;
            hit1 = hit1 && !over1;
            const unsigned m0 = __ballot_sync(0xffffffffu, hit0), m1 = __ballot_sync(0xffffffffu, hit1);
            const unsigned below = (1u << lane) - 1u;
            const int s0 = n0 + __popc(m0 & below), s1 = n1 + __popc(m1 & below);
            // the reference stops a particle as soon as its count reaches maxNeighboursPerParticle (:60-63)
            if (hit0 && s0 + 1 < maxNeighbours) {
              if (s0 < kVerletK) lbuf[h][s0] = (unsigned short)t;
              else neighbourList[(size_t)s0 * N + id0] = __float_as_int(cand[t].w); // rare: longer than the shared buffer
            }
            if (hit1 && s1 + 1 < maxNeighbours) {
              if (s1 < kVerletK) lbuf[h + 1][s1] = (unsigned short)t;
              else neighbourList[(size_t)s1 * N + id1] = __float_as_int(cand[t].w);
            }
            n0 += __popc(m0);
            n1 += __popc(m1);
            if (n0 >= maxNeighbours) over0 = true;
            if (n1 >= maxNeighbours) over1 = true;
          }
          if (lane == 0) {
            cntBuf[h] = over0 ? -1 : n0;
            if (over0) atomicMax(overflow, (uint32_t)n0);
            else numberNeighbours[id0] = n0;
            if (two) {
              cntBuf[h + 1] = over1 ? -1 : n1;
              if (over1) atomicMax(overflow, (uint32_t)n1);
              else numberNeighbours[id1] = n1;
            }
          }
        }
        __syncwarp();
        // transposed write-out: lanes = home particles of this pass, one k-row per iteration
        int myCnt = lane < nh ? cntBuf[lane] : 0;
        if (myCnt < 0) myCnt = maxNeighbours - 1;
        myCnt = min(myCnt, kVerletK);
        int maxCnt = myCnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxCnt = max(maxCnt, __shfl_xor_sync(0xffffffffu, maxCnt, o));
        for (int k = 0; k < maxCnt; k++)
          if (k < myCnt) neighbourList[(size_t)k * N + hStart + h0 + lane] = __float_as_int(cand[lbuf[lane][k]].w);
        __syncwarp();
      }
    } else {
      // dense neighbourhood (or a periodic dimension with < 4 cells): walk the cells straight from global memory
      for (int h = 0; h < hCount; h++) {
        const int id = hStart + h;
        const float4 pi = ldg4(sortPos + id);
        int nneigh = 0;
        bool over = false;
        for (int c = 0; c < 27 && !over; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          if (cnt == 0) continue;
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          for (int t0 = 0; t0 < cnt; t0 += 32) {
            const int t = t0 + lane;
            const bool hit = t < cnt && verletHitExact(pi, ldg4(sortPos + st + min(t, cnt - 1)), g, cutOff2);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const int slot = nneigh + __popc(m & ((1u << lane) - 1u));
            if (hit && slot + 1 < maxNeighbours) neighbourList[(size_t)slot * N + id] = st + t;
            nneigh += __popc(m);
            if (nneigh >= maxNeighbours) { over = true; break; }
          }
        }
        if (lane == 0) {
          if (over) atomicMax(overflow, (uint32_t)nneigh);
          else numberNeighbours[id] = nneigh;
        }
      }
    }
  }
}

} // namespace ub200

using namespace ub200;

// rebuildList (VerletListBase.cuh:165-168) -> BasicNeighbourListBase::update (BasicListBase.cuh:131-141): the
// reference-layout list of the stored positions
static int buildReferenceList(ub200_verletlist *v, cudaStream_t st) {
  const int N = v->N;
  const float *L = v->L;
  const int *periodic = v->periodic;
  int rc;
  {
    const float rcut = v->cutOff * v->multiplier;
    int cd[3];
    if ((rc = ub200_neighbour_celldim_f32(L, rcut, cd))) return rc;
    if ((rc = ub200_celllist_build_f32(v->cl, v->storedPos.p, nullptr, N, L, periodic, cd, st))) return rc;
    const int needed = (v->cl->ncells + kPairWarps - 1) / kPairWarps;
    const int grid = needed < kNumSMs * 4 ? needed : kNumSMs * 4;
    while (true) { // fillBasicNeighbourList: retry with 32 more slots until nothing overflows (:176-181)
      if ((rc = v->neighbourList.reserve(sizeof(int) * (size_t)N * (v->maxNeighbours + 1)))) return rc;
      UB200_CUDA(cudaMemsetAsync(v->flags.as<uint32_t>() + 1, 0, sizeof(uint32_t), st));
      UB200_CUDA(cudaFuncSetAttribute(verletFill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVerletSmem));
      verletFill<<<grid, kPairThreads, kVerletSmem, st>>>(v->cl->sortPos.as<float4>(), v->cl->binStart.as<uint32_t>(), v->cl->grid, v->cl->ncells,
                                               rcut * rcut, N, v->maxNeighbours, v->neighbourList.as<int>(),
                                               v->numberNeighbours.as<int>(), v->flags.as<uint32_t>() + 1);
      UB200_LAUNCHED();
      uint32_t over = 0;
      UB200_CUDA(cudaMemcpyAsync(&over, v->flags.as<uint32_t>() + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      UB200_CUDA(cudaStreamSynchronize(st));
      if (!over) break;
      v->maxNeighbours += 32;
    }
  }
  v->refValid = true;
  return UB200_OK;
}

extern "C" {

int ub200_verletlist_create(ub200_verletlist **out) {
  if (!out) return UB200_ERR_INVALID_ARGUMENT;
  ub200_verletlist *v = new (std::nothrow) ub200_verletlist();
  if (!v) return UB200_ERR_ALLOC;
  int rc = ub200_celllist_create(&v->cl);
  if (rc) { delete v; return rc; }
  if ((rc = ub200_ljengine_create(&v->eng))) { ub200_celllist_destroy(v->cl); delete v; return rc; }
  if ((rc = v->flags.reserve(2 * sizeof(uint32_t)))) { ub200_ljengine_destroy(v->eng); ub200_celllist_destroy(v->cl); delete v; return rc; }
  *out = v;
  return UB200_OK;
}
int ub200_verletlist_destroy(ub200_verletlist *v) {
  if (!v) return UB200_OK;
  ub200_celllist_destroy(v->cl);
  ub200_ljengine_destroy(v->eng);
  v->storedPos.release(); v->sortPos.release(); v->numberNeighbours.release(); v->neighbourList.release(); v->flags.release();
  v->fastPos.release(); v->fastList.release(); v->fastCount.release();
  delete v;
  return UB200_OK;
}
int ub200_verletlist_set_cutoff_multiplier(ub200_verletlist *v, float multiplier) {
  if (!v || !(multiplier >= 1.0f)) return UB200_ERR_INVALID_ARGUMENT;
  v->forceNext = true; // VerletListBase::setCutOffMultiplier (:126-129)
  v->multiplier = multiplier;
  return UB200_OK;
}

int ub200_verletlist_update_f32(ub200_verletlist *v, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                                const int periodic[3], float cutOff, int forceRebuild, int *rebuilt, void *stream) {
  if (!v || !d_pos || N <= 0 || !L || !periodic || !(cutOff > 0.f)) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const float4 *pos = (const float4 *)d_pos;
  const int nb = (N + 255) / 256;
  int rc;
  // ---- VerletListBase::needsRebuild (:172-190) ----
  bool rebuild = v->forceNext || forceRebuild != 0, refreshed = false;
  v->forceNext = false;
  if (!rebuild) {
    rebuild = N != v->N || cutOff != v->cutOff;
    for (int d = 0; d < 3; d++) rebuild = rebuild || L[d] != v->L[d] || (periodic[d] != 0) != (v->periodic[d] != 0);
  }
  if (!rebuild) {
    // isParticleDriftOverThreshold (:192-218): host-synchronous flag read, like the reference
    const float threshold = (v->multiplier * v->cutOff - v->cutOff) / 2.0f;
    if (threshold <= 1e-6f) rebuild = true;
    else if (v->fast) {
      // the row list refreshes its positions every call anyway: the same pass measures the drift
      bool over = false;
      if ((rc = vlistRefreshAndCheck(v, pos, d_groupIdx, threshold, &over, st))) return rc;
      rebuild = over;
      refreshed = !over;
    } else {
      const int cd1[3] = {1, 1, 1};
      const GridF g = makeGridF(L, periodic, cd1);
      UB200_CUDA(cudaMemsetAsync(v->flags.p, 0, sizeof(uint32_t), st));
      verletDriftCheck<<<nb, 256, 0, st>>>(pos, d_groupIdx, v->storedPos.as<float4>(), N, g, threshold * threshold, v->flags.as<uint32_t>());
      UB200_LAUNCHED();
      uint32_t over = 0;
      UB200_CUDA(cudaMemcpyAsync(&over, v->flags.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      UB200_CUDA(cudaStreamSynchronize(st));
      rebuild = over > 0;
    }
  }
  if (rebuild) {
    v->stepsSinceLastUpdate = 0;
    v->rebuilds++;
    v->N = N; v->cutOff = cutOff;
    for (int d = 0; d < 3; d++) { v->L[d] = L[d]; v->periodic[d] = periodic[d]; }
    if ((rc = v->storedPos.reserve(sizeof(float4) * (size_t)N)) || (rc = v->sortPos.reserve(sizeof(float4) * (size_t)N)) ||
        (rc = v->numberNeighbours.reserve(sizeof(int) * (size_t)N)))
      return rc;
    verletStore<<<nb, 256, 0, st>>>(pos, d_groupIdx, N, v->storedPos.as<float4>());
    UB200_LAUNCHED();
    v->fast = !v->refOnly && vlistApplies(L, periodic, v->cutOff * v->multiplier, N);
    v->refValid = false;
    if (v->fast && (rc = vlistRebuild(v, st))) return rc;
    if ((!v->fast || v->wantRef) && (rc = buildReferenceList(v, st))) return rc;
  }
  v->lastPos = d_pos; v->lastGroupIdx = d_groupIdx; v->lastStream = st;
  if (v->fast && !refreshed && (rc = vlistRefreshPositions(v, pos, d_groupIdx, st))) return rc;
  if (v->refValid) {
    verletSortedPositions<<<nb, 256, 0, st>>>(pos, d_groupIdx, v->cl->groupIndex.as<int>(), N, v->sortPos.as<float4>());
    UB200_LAUNCHED();
  }
  v->stepsSinceLastUpdate++;
  if (rebuilt) *rebuilt = rebuild ? 1 : 0;
  return UB200_OK;
}

int ub200_verletlist_view_get(ub200_verletlist *v, ub200_verletlist_view *view) {
  if (!v || !view) return UB200_ERR_INVALID_ARGUMENT;
  if (!v->N) return UB200_ERR_NOT_BUILT;
  if (!v->refValid) {
    // first reader of the reference layout: build it from the positions stored at the last rebuild and keep it current
    // from now on (the positions given to the last update are read for sortPos)
    cudaStream_t st = v->lastStream;
    if (const int rc = buildReferenceList(v, st)) return rc;
    verletSortedPositions<<<(v->N + 255) / 256, 256, 0, st>>>((const float4 *)v->lastPos, v->lastGroupIdx,
                                                             v->cl->groupIndex.as<int>(), v->N, v->sortPos.as<float4>());
    UB200_LAUNCHED();
    UB200_CUDA(cudaStreamSynchronize(st));
    v->wantRef = true;
  }
  view->d_neighbourList = v->neighbourList.as<int>();
  view->d_numberNeighbours = v->numberNeighbours.as<int>();
  view->d_sortPos = v->sortPos.p;
  view->d_groupIndex = v->cl->groupIndex.as<int>();
  view->particleStride = v->N;
  view->numberParticles = v->N;
  view->maxNeighboursPerParticle = v->maxNeighbours;
  view->stepsSinceLastUpdate = v->stepsSinceLastUpdate - 1; // VerletListBase::getNumberOfStepsSinceLastUpdate (:131)
  view->rebuilds = v->rebuilds;
  return UB200_OK;
}

int ub200_verletlist_stats(ub200_verletlist *v, int *stepsSinceLastUpdate, int *rebuilds) {
  if (!v) return UB200_ERR_INVALID_ARGUMENT;
  if (stepsSinceLastUpdate) *stepsSinceLastUpdate = v->stepsSinceLastUpdate - 1;
  if (rebuilds) *rebuilds = v->rebuilds;
  return UB200_OK;
}
}
