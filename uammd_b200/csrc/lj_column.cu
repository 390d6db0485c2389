// LJ pair forces over an engine-private half-cell list, sm_100a: TMA-staged column traversal.
//
// Replaces PairForces<Potential::LJ, CellList>::sum (Interactor/PairForces.cu:43-78): CellList::update
// (NeighbourList/CellList.cuh:145-163) + transverseWithNeighbourContainer (NeighbourList/common.cuh:10-34) with
// Radial<LJFunctor>::Transverser (Potential/RadialPotential.cuh:107-127). The forces are the reference's (every pair
// within the cut-off exactly once, minimum image); the list that produces them is not: the reference-layout
// CellListData stays available through ub200_celllist_* for callers that read it.
//
// Why a second list. The reference's grid has cells >= cutOff, so a particle tests the 27 cutOff^3 around it: 340
// candidates at rho = 0.8, rc = 2.5, of which 52 are in range. Half cells (edge >= cutOff/2) cover the same sphere with
// 5^3 cells = 15.6 cutOff^3: 196 candidates. The pair kernel is bound by instruction issue, i.e. by candidates.
//
// Traversal (colgeom.h). One WARP per column of kColTZ half cells along z. The 5 x 5 x (kColTZ + 4) halo of the column
// consists of 5 (kColTZ + 4) x-rows, each contiguous in the sorted array: every lane describes up to two rows (4 loads
// of the cell table), a warp scan places them back to back in the warp's slice of shared memory, and the rows are
// fetched by the TMA engine (cp.async.bulk, one bulk copy per row piece, completion on the warp's mbarrier) - no
// thread moves a candidate. Planes are staged in z order, so the neighbourhood of home cell hz is ONE contiguous range
// of the slice: the inner loop is a flat, conflict-free LDS.128 stride over ~196 candidates for two home particles at
// a time (one when the cell holds an odd one), branch-free LJ body, butterfly reduction. Columns touching a periodic
// boundary add the image shift of each row piece in a short pass over the slice; all other columns use the staged
// coordinates as they are (they were folded when the list was built).
#include "lj_engine.cuh"
#include "lj_pair.cuh"
#include <cstdlib>
#include <cstring>

namespace ub200 {

int ljSum(ub200_celllist *cl, const float *params, int ntypes, float4 *force, float *energy, float *virial,
          const int *globalIdx, bool accumulate, LJTableCache *cache, cudaStream_t st, int ownerLo, int ownerHi);

constexpr int kColTZ = 6;                  // home half cells per column
constexpr int kColPlanes = kColTZ + 4;
constexpr int kColCap = 480;               // staged candidates per warp (7.5 KB -> 7 CTAs of 4 warps per SM)
constexpr int kColWarps = 4;
constexpr int kColCTAs = 7;
constexpr int kColThreads = 32 * kColWarps;
constexpr int kColMeta = (kColPlanes + 1) + (kColTZ + 1) + 2 * kColTZ;
static_assert(kColTZ <= 8, "home cell prefix is an 8-lane scan");
constexpr int kNoShift = 2 | (2 << 3) | (2 << 6) | (2 << 9);
static_assert(5 * kColPlanes <= 64, "two rows per lane");

// ---- mbarrier / bulk copy (PTX ISA 8.0, sm_90+) ----
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbarArriveExpectTx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy by the TMA engine (SASS: UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}

__device__ __forceinline__ int packShift(const ColRow &r) {
  return (r.sx[0] + 2) | ((r.sx[1] + 2) << 3) | ((r.sy + 2) << 6) | ((r.sz + 2) << 9);
}

// NH home particles of one half cell against a stream of candidates
template <int NH, bool ENERGY, bool VIRIAL, bool MULTITYPE> struct HomeEval {
  float4 pi[NH];
  Acc a[NH];
  LJPar par[NH];
  uint32_t rcb[NH];
  int trow[NH];
  const LJPar *tab;
  int ntypes;
  __device__ __forceinline__ void init(const LJPar *table, int nt, const LJPar &par0) {
    tab = table;
    ntypes = nt;
#pragma unroll
    for (int k = 0; k < NH; k++) {
      a[k] = Acc{0.f, 0.f, 0.f, 0.f, 0.f};
      par[k] = par0;
      rcb[k] = __float_as_uint(par0.cutOff2) - 1u;
      // types outside the table fall back to entry 0 like BasicParameterHandler::Iterator (ParameterHandler.cuh:55-57)
      const int ty = (int)pi[k].w;
      trow[k] = MULTITYPE && (unsigned)ty < (unsigned)nt ? ty * nt : -1;
    }
  }
  __device__ __forceinline__ void add(const float4 pj) {
#pragma unroll
    for (int k = 0; k < NH; k++) {
      if (MULTITYPE) {
        const int tj = (int)pj.w;
        par[k] = tab[((unsigned)tj < (unsigned)ntypes && trow[k] >= 0) ? trow[k] + tj : 0];
        rcb[k] = __float_as_uint(par[k].cutOff2) - 1u;
      }
      ljPair<ENERGY, VIRIAL>(pj.x - pi[k].x, pj.y - pi[k].y, pj.z - pi[k].z, par[k], rcb[k], a[k]);
    }
  }
  // after reduce(): particle 0 in lane 0, particle 1 (NH == 2) in lane 16, both in a[0]
  __device__ __forceinline__ void reduce(int lane) {
    if (NH == 2) {
      reducePair(a[0], a[NH - 1], lane, ENERGY || VIRIAL);
    } else {
      a[0].fx = warpSum(a[0].fx); a[0].fy = warpSum(a[0].fy); a[0].fz = warpSum(a[0].fz);
      if (ENERGY) a[0].e = warpSum(a[0].e);
      if (VIRIAL) a[0].v = warpSum(a[0].v);
    }
  }
};

template <bool ENERGY, bool VIRIAL, bool ACCUMULATE>
__device__ __forceinline__ void storeHome(const Acc &a, int ori, float4 *__restrict__ force, float *__restrict__ energy,
                                          float *__restrict__ virial) {
  if (force) {
    if (ACCUMULATE) {
      float4 f = force[ori];
      f.x += a.fx; f.y += a.fy; f.z += a.fz;
      force[ori] = f;
    } else {
      force[ori] = make_float4(a.fx, a.fy, a.fz, 0.0f);
    }
  }
  if (ENERGY) energy[ori] += a.e;
  if (VIRIAL) virial[ori] += a.v;
}

// One pass over the staged slice: 32 / W home particles, W lanes each. src = lane of the warp that holds the particle's
// record (c0, c1: candidate range, slot: its own position in the slice, gi: group index); W is a compile-time constant so
// that the candidate loads are base + immediate and the reduction unrolls.
template <int W, bool ENERGY, bool VIRIAL, bool MULTITYPE, bool ACCUMULATE, bool OWNED>
__device__ __forceinline__ void columnPass(const float4 *cand, int lane, int q0, int nHomeP, int recC0, int recC1, int recSlot,
                                           int recGi, const LJPar *__restrict__ parTable, int ntypes, const LJPar &par0,
                                           const LJFold &fold, float4 *__restrict__ force, float *__restrict__ energy,
                                           float *__restrict__ virial, const int *__restrict__ globalIdx, int ownerLo,
                                           int ownerHi) {
  const int sub = lane & (W - 1), q = q0 + lane / W;
  const bool act = q < nHomeP;
  const int src = act ? q : 0;
  int c0 = __shfl_sync(0xffffffffu, recC0, src), c1 = __shfl_sync(0xffffffffu, recC1, src);
  const int slot = __shfl_sync(0xffffffffu, recSlot, src), gi = __shfl_sync(0xffffffffu, recGi, src);
  if (!act) c1 = c0;
  Acc a;
  if (!ENERGY && !VIRIAL && !MULTITYPE) {
    const float4 pi = cand[slot];
    a = Acc{0.f, 0.f, 0.f, 0.f, 0.f};
    const float4 *p = cand + c0 + sub;
    const int n = c1 - c0 - sub; // candidates left for this lane: n, n - W, ...
#pragma unroll 4
    for (int t = 0; t < n; t += W) {
      const float4 pj = p[t];
      ljPairFolded(pj.x - pi.x, pj.y - pi.y, pj.z - pi.z, fold, a);
    }
  } else {
    HomeEval<1, ENERGY, VIRIAL, MULTITYPE> ev;
    ev.pi[0] = cand[slot];
    ev.init(parTable, ntypes, par0);
#pragma unroll 2
    for (int t = c0 + sub; t < c1; t += W) ev.add(cand[t]);
    a = ev.a[0];
  }
#pragma unroll
  for (int o = W >> 1; o > 0; o >>= 1) {
    a.fx += __shfl_xor_sync(0xffffffffu, a.fx, o);
    a.fy += __shfl_xor_sync(0xffffffffu, a.fy, o);
    a.fz += __shfl_xor_sync(0xffffffffu, a.fz, o);
    if (ENERGY) a.e += __shfl_xor_sync(0xffffffffu, a.e, o);
    if (VIRIAL) a.v += __shfl_xor_sync(0xffffffffu, a.v, o);
  }
  if (act && sub == 0 && (!OWNED || (gi >= ownerLo && gi < ownerHi)))
    storeHome<ENERGY, VIRIAL, ACCUMULATE>(a, globalIdx ? globalIdx[gi] : gi, force, energy, virial);
}

template <bool ENERGY, bool VIRIAL, bool MULTITYPE, bool ACCUMULATE, bool OWNED, bool TMA>
__global__ void __launch_bounds__(kColThreads, kColCTAs)
ljColumnTraversal(const float4 *__restrict__ finePos, const int *__restrict__ fineIdx,
                  const uint32_t *__restrict__ binStart, ColGrid cg, float Lx, float Ly, float Lz,
                  const LJPar *__restrict__ parTable, int ntypes, float4 *__restrict__ force, float *__restrict__ energy,
                  float *__restrict__ virial, const int *__restrict__ globalIdx, int ownerLo, int ownerHi,
                  const int *__restrict__ ownerHiDev, int *__restrict__ errFlag, int *__restrict__ nextColumn, int widen) {
  __shared__ __align__(16) float4 candAll[kColWarps][kColCap];
  if (OWNED && ownerHiDev) ownerHi = *ownerHiDev; // multi-GPU bricks: the owned block [0, nOwned) is counted on the device
  __shared__ __align__(8) unsigned long long barAll[kColWarps];
  __shared__ int metaAll[kColWarps][kColMeta];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *cand = candAll[warp];
  int *planeOff = metaAll[warp];              // [kColPlanes + 1] first slice slot of plane p; [nPlanes] = total
  int *homePre = planeOff + kColPlanes + 1;   // [kColTZ + 1] home particles of the column before home cell hz
  int *homeOff = homePre + kColTZ + 1;        // [kColTZ] slice slot of the first particle of home cell hz
  int *homeG = homeOff + kColTZ;              // [kColTZ] its sorted index
  const uint32_t bar = smemAddr(&barAll[warp]), candAddr = smemAddr(cand);
  if (TMA && lane == 0) {
    mbarInit(bar, 1);
    fenceProxyAsync(); // make the initialised barrier visible to the async proxy
  }
  __syncwarp();
  uint32_t phase = 0;
  const LJPar par0 = parTable[0]; // single type: BasicParameterHandler::Iterator returns entry 0 (ParameterHandler.cuh:49-50)
  const LJFold fold = foldLJ(par0);
  const int nzc = (cg.nz + kColTZ - 1) / kColTZ;
  const int ncols = cg.nx * cg.ny * nzc;
  // columns are handed out through a global counter (x fastest, so that the warps running at any moment work on
  // neighbouring columns and share their rows in L2); the next index is fetched while the current column is computed
  const int firstDynamic = gridDim.x * kColWarps;
  int col = blockIdx.x * kColWarps + warp, nextCol = 0;
  for (; col < ncols; col = nextCol) {
    if (lane == 0) nextCol = firstDynamic + atomicAdd(nextColumn, 1);
    nextCol = __shfl_sync(0xffffffffu, nextCol, 0);
    const int x0 = col % cg.nx, t1 = col / cg.nx, y0 = t1 % cg.ny, z0 = (t1 / cg.ny) * kColTZ;
    const int nHome = min(kColTZ, cg.nz - z0);
    const int nRows = 5 * (nHome + 4);
    // bricks: the two outer cell layers of a windowed dimension hold ghosts - nothing to compute for them
    if (OWNED && ownerHiDev && ((cg.wx && (x0 < 2 || x0 >= cg.nx - 2)) || (cg.wy && (y0 < 2 || y0 >= cg.ny - 2)))) continue;
    // ---- two rows per lane: global ranges of the row pieces, image shifts, the home cell of the row (dy == 0 rows)
    int g0[2][2], cn[2][2], shp[2], hG[2], hC[2], hRel[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int r = lane + 32 * q;
      g0[q][0] = g0[q][1] = 0; cn[q][0] = cn[q][1] = 0;
      shp[q] = kNoShift; hG[q] = 0; hC[q] = 0; hRel[q] = 0;
      if (r < nRows) {
        const ColRow row = columnRow(cg, x0, y0, z0, r);
#pragma unroll
        for (int s = 0; s < 2; s++)
          if (row.n[s] > 0) {
            const uint32_t a = __ldg(binStart + row.c0[s]), b = __ldg(binStart + row.c0[s] + row.n[s]);
            g0[q][s] = (int)a;
            cn[q][s] = (int)(b - a);
          }
        shp[q] = packShift(row);
        const int p = r / 5;
        if (r - 5 * p == 2 && p >= 2 && p < 2 + nHome) {
          const int cc = x0 + cg.nx * (y0 + cg.ny * (z0 + p - 2));
          const uint32_t a = __ldg(binStart + cc), b = __ldg(binStart + cc + 1);
          const bool ghostCell = OWNED && ownerHiDev && cg.wz && (z0 + p - 2 < 2 || z0 + p - 2 >= cg.nz - 2);
          hG[q] = (int)a;
          hC[q] = ghostCell ? 0 : (int)(b - a);
          // offset of the home cell inside its row (piece row.hs)
          hRel[q] = row.hs ? cn[q][0] + ((int)a - g0[q][1]) : (int)a - g0[q][0];
        }
      }
    }
    if (!__any_sync(0xffffffffu, hC[0] > 0 || hC[1] > 0)) continue; // no home particle in this column
    if (OWNED && !ownerHiDev) { // owner restriction by index range: skip columns without an owned home particle
      bool mine = false;
#pragma unroll
      for (int q = 0; q < 2; q++)
        for (int k = 0; k < hC[q]; k++) {
          const int gi = fineIdx[hG[q] + k];
          mine |= gi >= ownerLo && gi < ownerHi;
        }
      if (!__any_sync(0xffffffffu, mine)) continue;
    }
    const int cq0 = cn[0][0] + cn[0][1], cq1 = cn[1][0] + cn[1][1];
    int inc0 = cq0, inc1 = cq1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u0 = __shfl_up_sync(0xffffffffu, inc0, o), u1 = __shfl_up_sync(0xffffffffu, inc1, o);
      if (lane >= o) { inc0 += u0; inc1 += u1; }
    }
    const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const int total = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
    const int off[2] = {inc0 - cq0, tot0 + inc1 - cq1};
    // home cell hz sits in row 5 (hz + 2) + 2: bring its data to lane hz and count the home particles before it
    const int hr = 5 * lane + 12, hsrc = hr & 31;
    const int a0 = __shfl_sync(0xffffffffu, hC[0], hsrc), a1 = __shfl_sync(0xffffffffu, hC[1], hsrc);
    const int b0 = __shfl_sync(0xffffffffu, hG[0], hsrc), b1 = __shfl_sync(0xffffffffu, hG[1], hsrc);
    const int c0s = __shfl_sync(0xffffffffu, off[0] + hRel[0], hsrc), c1s = __shfl_sync(0xffffffffu, off[1] + hRel[1], hsrc);
    const bool isHome = lane < nHome;
    const int myCnt = isHome ? (hr >= 32 ? a1 : a0) : 0;
    int pre = myCnt;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += u;
    }
    __syncwarp(); // every lane is done with the previous column's slice and tables
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int r = lane + 32 * q;
      if (r < nRows && r % 5 == 0) planeOff[r / 5] = off[q];
    }
    if (lane == 0) planeOff[nHome + 4] = total;
    if (lane <= kColTZ) homePre[lane] = isHome ? pre - myCnt : 0x3fffffff; // entries past the last home cell never match
    if (isHome) {
      homeOff[lane] = hr >= 32 ? c1s : c0s;
      homeG[lane] = hr >= 32 ? b1 : b0;
    }
    const int nHomeP = __shfl_sync(0xffffffffu, pre, kColTZ - 1 < 31 ? kColTZ - 1 : 31); // cells past nHome add 0
    const bool staged = total <= kColCap && nHomeP <= 32; // warp uniform
    if (staged && TMA) {
      fenceProxyAsync(); // earlier generic-proxy accesses of the slice are ordered before the bulk copies below
      if (lane == 0) mbarArriveExpectTx(bar, (uint32_t)total * 16u);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (cn[q][0] > 0) bulkLoad(candAddr + 16u * (uint32_t)off[q], finePos + g0[q][0], 16u * (uint32_t)cn[q][0], bar);
        if (cn[q][1] > 0)
          bulkLoad(candAddr + 16u * (uint32_t)(off[q] + cn[q][0]), finePos + g0[q][1], 16u * (uint32_t)cn[q][1], bar);
      }
      bool landed = false;
      for (int spin = 0; spin < (1 << 22) && !(landed = mbarTryWait(bar, phase)); spin++) {}
      if (!landed) { // never observed; bounded so that a lost copy cannot hang the device
        if (lane == 0) atomicExch(errFlag, 2);
        return;
      }
      phase ^= 1u;
      if (__any_sync(0xffffffffu, shp[0] != kNoShift || shp[1] != kNoShift)) {
        // column touching a periodic boundary: move the wrapped row pieces to their image
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (shp[q] == kNoShift) continue;
          const float dy = (float)(((shp[q] >> 6) & 7) - 2) * Ly, dz = (float)(((shp[q] >> 9) & 7) - 2) * Lz;
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const float dx = (float)(((shp[q] >> (3 * s)) & 7) - 2) * Lx;
            const int b = off[q] + (s ? cn[q][0] : 0);
            for (int t = b; t < b + cn[q][s]; t++) {
              float4 p = cand[t];
              p.x += dx; p.y += dy; p.z += dz;
              cand[t] = p;
            }
          }
        }
      }
    } else if (staged) {
      // the same staging with ordinary loads: every lane copies its own rows, image shift applied on the way
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int cq = q ? cq1 : cq0;
        const int mq = __reduce_max_sync(0xffffffffu, cq);
        const float dy = (float)(((shp[q] >> 6) & 7) - 2) * Ly, dz = (float)(((shp[q] >> 9) & 7) - 2) * Lz;
        const float dx0 = (float)((shp[q] & 7) - 2) * Lx, dx1 = (float)(((shp[q] >> 3) & 7) - 2) * Lx;
#pragma unroll 4
        for (int k = 0; k < mq; k++) {
          if (k < cq) {
            const bool second = k >= cn[q][0];
            float4 p = ldg4(finePos + (second ? g0[q][1] + (k - cn[q][0]) : g0[q][0] + k));
            p.x += second ? dx1 : dx0; p.y += dy; p.z += dz;
            cand[off[q] + k] = p;
          }
        }
      }
    }
    __syncwarp();
    if (staged) {
      // ---- home particles of the column. Lane q < nHomeP first works out the record of home particle q: its candidate
      // range (the planes hz .. hz + 4 of its home cell: ONE contiguous range of the slice), its own slot and its index.
      int recC0 = 0, recC1 = 0, recSlot = 0, recGi = -1;
      if (lane < nHomeP) {
        int hz = 0;
#pragma unroll
        for (int k = 1; k < kColTZ; k++) hz += lane >= homePre[k];
        const int hrel = lane - homePre[hz];
        recC0 = planeOff[hz];
        recC1 = planeOff[hz + 5];
        recSlot = homeOff[hz] + hrel;
        recGi = fineIdx[homeG[hz] + hrel];
      }
      // Passes of four particles, eight lanes each; the last pass of a column widens to 16 or 32 lanes per particle when
      // only two or one are left.
      // (widen = 0 keeps eight lanes per particle throughout: the summation order of a particle then depends on its own
      // neighbourhood only, not on what else shares its column - the multi-GPU bricks rely on it to reproduce the
      // single-GPU forces bit for bit)
      int q0 = 0;
      for (; nHomeP - q0 >= (widen ? 3 : 1); q0 += 4)
        columnPass<8, ENERGY, VIRIAL, MULTITYPE, ACCUMULATE, OWNED>(cand, lane, q0, nHomeP, recC0, recC1, recSlot, recGi, parTable,
                                                                    ntypes, par0, fold, force, energy, virial, globalIdx,
                                                                    ownerLo, ownerHi);
      if (nHomeP - q0 == 2)
        columnPass<16, ENERGY, VIRIAL, MULTITYPE, ACCUMULATE, OWNED>(cand, lane, q0, nHomeP, recC0, recC1, recSlot, recGi, parTable,
                                                                     ntypes, par0, fold, force, energy, virial, globalIdx,
                                                                     ownerLo, ownerHi);
      else if (nHomeP - q0 == 1)
        columnPass<32, ENERGY, VIRIAL, MULTITYPE, ACCUMULATE, OWNED>(cand, lane, q0, nHomeP, recC0, recC1, recSlot, recGi, parTable,
                                                                     ntypes, par0, fold, force, energy, virial, globalIdx,
                                                                     ownerLo, ownerHi);
    } else {
      // ---- dense column (more candidates than the slice holds): one home particle at a time, the whole warp walks the 25
      // rows of its home cell straight from global memory
      for (int q = 0; q < nHomeP; q++) {
        int hz = 0;
#pragma unroll
        for (int k = 1; k < kColTZ; k++) hz += q >= homePre[k];
        const int gs = homeG[hz] + q - homePre[hz];
        const int gi = fineIdx[gs];
        if (OWNED && !(gi >= ownerLo && gi < ownerHi)) continue;
        HomeEval<1, ENERGY, VIRIAL, MULTITYPE> ev;
        ev.pi[0] = ldg4(finePos + gs);
        ev.init(parTable, ntypes, par0);
        for (int rr = 5 * hz; rr < 5 * hz + 25; rr++) {
          const int src = rr & 31, hi = rr >> 5;
          const int sp = __shfl_sync(0xffffffffu, hi ? shp[1] : shp[0], src);
          const float dy = (float)(((sp >> 6) & 7) - 2) * Ly, dz = (float)(((sp >> 9) & 7) - 2) * Lz;
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int g = __shfl_sync(0xffffffffu, hi ? g0[1][s] : g0[0][s], src);
            const int n = __shfl_sync(0xffffffffu, hi ? cn[1][s] : cn[0][s], src);
            const float dx = (float)(((sp >> (3 * s)) & 7) - 2) * Lx;
            for (int t = lane; t < n; t += 32) {
              float4 pj = ldg4(finePos + g + t);
              pj.x += dx; pj.y += dy; pj.z += dz;
              ev.add(pj);
            }
          }
        }
        ev.reduce(lane);
        if (lane == 0) storeHome<ENERGY, VIRIAL, ACCUMULATE>(ev.a[0], globalIdx ? globalIdx[gi] : gi, force, energy, virial);
      }
    }
  }
}

// ---- half-cell list build: counting sort by linear cell index, stable inside a cell ----

// Cell and stored coordinate of one dimension. The fold and the cell follow Box::apply_pbc / Grid::getCell
// (utils/Box.cuh:51-58, utils/Grid.cuh:49-71) on the half-cell grid; where the reference wraps cell n to 0 and leaves the
// coordinate alone, the coordinate moves down one box length with it, so that stored coordinates and cells agree.
__device__ __forceinline__ void canonicalCoord(float r, float L, float m, float hL, float inv, int n, int &c, float &rf,
                                               bool &bad) {
  rf = foldCoord(r, L, m);
  c = __float2int_rz(__fmul_rn(__fadd_rn(rf, hL), inv));
  if (m != 0.0f) {
    if (c >= n) { c -= n; rf -= L; }
    else if (c < 0) { c += n; rf += L; }
  }
  if ((unsigned)c >= (unsigned)n) {
    bad = true; // outside a non periodic box (the reference raises errorFlag in fillCellList), NaN, inf
    c = min(max(c, 0), n - 1);
  }
}

// global half cell -> cell of the (windowed) local grid; false when it falls outside the window
__device__ __forceinline__ bool localCell(int c, int o, int g, int n, bool globallyPeriodic, int &l) {
  l = c - o;
  if (globallyPeriodic) {
    if (l < 0) l += g;
    else if (l >= g) l -= g;
  }
  if ((unsigned)l >= (unsigned)n) {
    l = min(max(l, 0), n - 1);
    return false;
  }
  return true;
}

// nDev: optional device-side particle count (multi-GPU bricks: the arrays are sized for the worst case, the count of the
// step lives on the device); the launch covers N particles and the excess threads leave.
__global__ void __launch_bounds__(256)
fineBin(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, int N, const int *__restrict__ nDev, GridF g,
        ColGrid cg, uint32_t *__restrict__ binCount, uint2 *__restrict__ codeSlot, int *__restrict__ errorFlag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev);
  if (i >= N) return;
  const float4 p = ldg4(pos + (groupIdx ? groupIdx[i] : i));
  int cx, cy, cz;
  float x, y, z;
  bool bad = false;
  canonicalCoord(p.x, g.Lx, g.mx, g.hLx, g.ix, g.nx, cx, x, bad);
  canonicalCoord(p.y, g.Ly, g.my, g.hLy, g.iy, g.ny, cy, y, bad);
  canonicalCoord(p.z, g.Lz, g.mz, g.hLz, g.iz, g.nz, cz, z, bad);
  if (bad) *errorFlag = 1;
  int lx, ly, lz;
  const bool in = localCell(cx, cg.ox, cg.gx, cg.nx, g.mx != 0.0f, lx) & localCell(cy, cg.oy, cg.gy, cg.ny, g.my != 0.0f, ly) &
                  localCell(cz, cg.oz, cg.gz, cg.nz, g.mz != 0.0f, lz);
  if (!in) *errorFlag = 3; // a particle outside the rank's window: the halo exchange missed it
  const uint32_t cell = (uint32_t)(lx + cg.nx * (ly + cg.ny * lz));
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, cell);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  const int rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(binCount + cell, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  codeSlot[i] = make_uint2(cell, base + rank);
}

__global__ void __launch_bounds__(256)
fineScatter(const uint2 *__restrict__ codeSlot, const uint32_t *__restrict__ binStart, int N, const int *__restrict__ nDev,
            int *__restrict__ unstable) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev);
  if (i >= N) return;
  const uint2 cs = codeSlot[i];
  unstable[binStart[cs.x] + cs.y] = i;
}

// sortKey: optional per-particle key deciding the order inside a cell (multi-GPU bricks pass the global particle ids, so
// that every cell lists its particles in the single-GPU order whatever the order of the local arrays); default: the index.
__global__ void __launch_bounds__(256)
fineOrder(const int *__restrict__ unstable, const uint2 *__restrict__ codeSlot, const uint32_t *__restrict__ binStart,
          const float4 *__restrict__ pos, const int *__restrict__ groupIdx, const int *__restrict__ sortKey, int N,
          const int *__restrict__ nDev, GridF g, float4 *__restrict__ finePos, int *__restrict__ fineIdx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (nDev) N = min(N, *nDev);
  if (k >= N) return;
  const int i = unstable[k];
  const uint32_t cell = codeSlot[i].x;
  const int s = (int)binStart[cell], e = (int)binStart[cell + 1];
  int rank = 0;
  if (sortKey) {
    const int ki = sortKey[i];
    for (int j = s; j < e; j++) rank += (sortKey[__ldg(unstable + j)] < ki);
  } else {
    for (int j = s; j < e; j++) rank += (__ldg(unstable + j) < i);
  }
  const float4 p = ldg4(pos + (groupIdx ? groupIdx[i] : i));
  int c;
  bool bad = false;
  float4 o;
  canonicalCoord(p.x, g.Lx, g.mx, g.hLx, g.ix, g.nx, c, o.x, bad);
  canonicalCoord(p.y, g.Ly, g.my, g.hLy, g.iy, g.ny, c, o.y, bad);
  canonicalCoord(p.z, g.Lz, g.mz, g.hLz, g.iz, g.nz, c, o.z, bad);
  o.w = p.w;
  finePos[s + rank] = o;
  fineIdx[s + rank] = i;
}

// half-cell grid for this box and cut-off; false when the column traversal does not apply
bool ljEngineDims(const float L[3], const int periodic[3], float rc, int dims[3], int per[3]) {
  double cells = 1.0;
  for (int d = 0; d < 3; d++) {
    if (isinf(L[d]) || isnan(L[d]) || L[d] < 0.0f) return false;
    per[d] = periodic[d] && L[d] != 0.0f;
    dims[d] = L[d] == 0.0f ? 1 : colCellsFor((double)L[d], (double)rc);
    if (per[d] && dims[d] < 5) return false; // the five cells of a stencil row must be distinct images
    cells *= dims[d];
  }
  return cells <= (double)(1 << 24);
}

// pos [N] (N = launch bound; nDev = optional device-side count <= N), binned on the window cg of the global half-cell
// grid `dims` (a whole grid for single-GPU use)
static int buildFine(ub200_ljengine *e, const float4 *pos, const int *groupIdx, const int *sortKey, int N, const int *nDev,
                     const float L[3], const int per[3], const int dims[3], const ColGrid &cg, cudaStream_t st) {
  const GridF g = makeGridF(L, per, dims);
  const int ncells = cg.nx * cg.ny * cg.nz;
  int rc;
  if ((rc = e->pos.reserve(sizeof(float4) * (size_t)N))) return rc;
  if ((rc = e->idx.reserve(sizeof(int) * (size_t)N))) return rc;
  if ((rc = e->codeSlot.reserve(sizeof(uint2) * (size_t)N))) return rc;
  if ((rc = e->unstable.reserve(sizeof(int) * (size_t)N))) return rc;
  if ((rc = e->blockSums.reserve(sizeof(uint32_t) * 4096))) return rc;
  if (!e->errorFlag.p) {
    if ((rc = e->errorFlag.reserve(2 * sizeof(int)))) return rc; // {error flag, column counter of the traversal}
    UB200_CUDA(cudaMemsetAsync(e->errorFlag.p, 0, 2 * sizeof(int), st));
  }
  if (e->binCells != ncells || !e->binCount.p) {
    if ((rc = e->binCount.reserve(sizeof(uint32_t) * (size_t)ncells))) return rc;
    if ((rc = e->binStart.reserve(sizeof(uint32_t) * ((size_t)ncells + 1)))) return rc;
    // the scan re-zeroes the histogram at every build; zero it once here
    UB200_CUDA(cudaMemsetAsync(e->binCount.p, 0, sizeof(uint32_t) * (size_t)ncells, st));
    e->binCells = ncells;
  }
  e->grid = g;
  e->cg = cg;
  e->N = N;
  e->ncells = ncells;
  const int nb = (N + 255) / 256;
  fineBin<<<nb, 256, 0, st>>>(pos, groupIdx, N, nDev, g, cg, e->binCount.as<uint32_t>(), e->codeSlot.as<uint2>(),
                              e->errorFlag.as<int>());
  UB200_LAUNCHED();
  if ((rc = exclusiveScanAndClear(e->binCount.as<uint32_t>(), ncells, e->binStart.as<uint32_t>(), e->blockSums.as<uint32_t>(), st)))
    return rc;
  fineScatter<<<nb, 256, 0, st>>>(e->codeSlot.as<uint2>(), e->binStart.as<uint32_t>(), N, nDev, e->unstable.as<int>());
  UB200_LAUNCHED();
  fineOrder<<<nb, 256, 0, st>>>(e->unstable.as<int>(), e->codeSlot.as<uint2>(), e->binStart.as<uint32_t>(), pos, groupIdx, sortKey,
                                N, nDev, g, e->pos.as<float4>(), e->idx.as<int>());
  UB200_LAUNCHED();
  return UB200_OK;
}

template <bool E, bool V, bool M, bool A, bool O, bool T>
static int launchColumnT(ub200_ljengine *e, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
                        const int *globalIdx, int ownerLo, int ownerHi, const int *ownerHiDev, cudaStream_t st) {
  auto kern = ljColumnTraversal<E, V, M, A, O, T>;
  static int blocksPerSM = 0; // per instantiation
  if (!blocksPerSM) {
    UB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kColThreads, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
  }
  const ColGrid &cg = e->cg;
  const char *wsel = getenv("UB200_LJ_WIDEN"); // "0": eight lanes per particle in every pass (decomposition-independent sums)
  const int widen = ownerHiDev ? 0 : !(wsel && wsel[0] == '0');
  const int ncols = cg.nx * cg.ny * ((cg.nz + kColTZ - 1) / kColTZ);
  int grid = kNumSMs * blocksPerSM;
  const int needed = (ncols + kColWarps - 1) / kColWarps;
  if (grid > needed) grid = needed;
  UB200_CUDA(cudaMemsetAsync(e->errorFlag.as<int>() + 1, 0, sizeof(int), st)); // the column counter
  kern<<<grid, kColThreads, 0, st>>>(e->pos.as<float4>(), e->idx.as<int>(), e->binStart.as<uint32_t>(), cg, e->grid.Lx,
                                     e->grid.Ly, e->grid.Lz, table, ntypes, force, energy, virial, globalIdx, ownerLo,
                                     ownerHi, ownerHiDev, e->errorFlag.as<int>(), e->errorFlag.as<int>() + 1, widen);
  UB200_LAUNCHED();
  return UB200_OK;
}

// staging of the column halo: TMA bulk copies (default) or per-lane row copies (UB200_LJ_STAGE=ldg), same results
template <bool E, bool V, bool M, bool A, bool O>
static int launchColumn(ub200_ljengine *e, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
                        const int *globalIdx, int ownerLo, int ownerHi, const int *ownerHiDev, cudaStream_t st) {
  const char *sel = getenv("UB200_LJ_STAGE");
  if (sel && strcmp(sel, "ldg") == 0)
    return launchColumnT<E, V, M, A, O, false>(e, table, ntypes, force, energy, virial, globalIdx, ownerLo, ownerHi, ownerHiDev, st);
  return launchColumnT<E, V, M, A, O, true>(e, table, ntypes, force, energy, virial, globalIdx, ownerLo, ownerHi, ownerHiDev, st);
}

static int columnSum(ub200_ljengine *e, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
                     const int *globalIdx, bool accumulate, int ownerLo, int ownerHi, cudaStream_t st,
                     const int *ownerHiDev = nullptr) {
  const bool E = energy != nullptr, V = virial != nullptr, M = ntypes > 1, O = ownerLo > 0 || ownerHi < 0x7fffffff || ownerHiDev;
  const bool A = accumulate;
#define UB200_COL(ee, vv, mm, aa, oo)                                                                                   \
  if (E == ee && V == vv && M == mm && A == aa && O == oo)                                                              \
    return launchColumn<ee, vv, mm, aa, oo>(e, table, ntypes, force, energy, virial, globalIdx, ownerLo, ownerHi, ownerHiDev, st);
  // forces only: every combination of multi-type / accumulate / owner restriction
  UB200_COL(false, false, false, false, false) UB200_COL(false, false, false, true, false)
  UB200_COL(false, false, true, false, false) UB200_COL(false, false, true, true, false)
  UB200_COL(false, false, false, false, true) UB200_COL(false, false, false, true, true)
  UB200_COL(false, false, true, false, true) UB200_COL(false, false, true, true, true)
#undef UB200_COL
  // energy and/or virial requested: they accumulate like Transverser::set, and so do the forces here
  if (!A || O) return UB200_ERR_UNSUPPORTED;
#define UB200_COL_EV(mm)                                                                                                \
  if (M == mm) {                                                                                                        \
    if (E && V) return launchColumn<true, true, mm, true, false>(e, table, ntypes, force, energy, virial, globalIdx, 0, 0x7fffffff, nullptr, st); \
    if (E) return launchColumn<true, false, mm, true, false>(e, table, ntypes, force, energy, virial, globalIdx, 0, 0x7fffffff, nullptr, st);     \
    return launchColumn<false, true, mm, true, false>(e, table, ntypes, force, energy, virial, globalIdx, 0, 0x7fffffff, nullptr, st);            \
  }
  UB200_COL_EV(false)
  UB200_COL_EV(true)
#undef UB200_COL_EV
  return UB200_ERR_UNSUPPORTED;
}

static int uploadTable(ub200_ljengine *e, const float *params, int ntypes, cudaStream_t st) {
  const size_t n = (size_t)ntypes * ntypes * 4;
  LJTableCache *cache = &e->table;
  if (cache->host.size() != n || memcmp(cache->host.data(), params, n * sizeof(float)) != 0) {
    if (const int rc = cache->dev.reserve(n * sizeof(float))) return rc;
    cache->host.assign(params, params + n);
    UB200_CUDA(cudaMemcpyAsync(cache->dev.p, cache->host.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  return UB200_OK;
}

// Multi-GPU bricks: forces of the owned block [0, *nOwnedDev) of a rank's local arrays [owned | ghosts] (*nLocalDev
// particles, at most maxN), binned on the rank's window cg of the global half-cell grid globalDims. sortKey = global ids.
int ljEngineBuildWindow(ub200_ljengine *e, const float4 *pos, const int *sortKey, int maxN, const int *nLocalDev,
                        const float L[3], const int periodic[3], const int globalDims[3], const ColGrid &cg, cudaStream_t st) {
  if (!e || !pos || maxN <= 0 || !nLocalDev) return UB200_ERR_INVALID_ARGUMENT;
  return buildFine(e, pos, nullptr, sortKey, maxN, nLocalDev, L, periodic, globalDims, cg, st);
}
int ljEngineTraverseWindow(ub200_ljengine *e, const int *nOwnedDev, const float *params, int ntypes, float4 *force,
                           bool accumulate, cudaStream_t st) {
  if (!e || !nOwnedDev || !params || ntypes < 1 || !force) return UB200_ERR_INVALID_ARGUMENT;
  if (const int rcode = uploadTable(e, params, ntypes, st)) return rcode;
  e->lastPath = 0;
  return columnSum(e, e->table.dev.as<LJPar>(), ntypes, force, nullptr, nullptr, nullptr, accumulate, 0, 0x7fffffff, st, nOwnedDev);
}

int ljEngineSum(ub200_ljengine *e, const float4 *pos, const int *groupIdx, int N, const float L[3], const int periodic[3],
                const float *params, int ntypes, float4 *force, float *energy, float *virial, const int *globalIdx,
                bool accumulate, int ownerLo, int ownerHi, cudaStream_t st) {
  if (!e || !pos || N <= 0 || !L || !periodic || !params || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  if (ownerLo < 0 || ownerHi < ownerLo) return UB200_ERR_INVALID_ARGUMENT;
  if (!force && !energy && !virial) return UB200_OK;
  const bool owned = ownerLo > 0 || ownerHi < 0x7fffffff;
  if ((energy || virial) && (!accumulate || owned)) return UB200_ERR_UNSUPPORTED;
  float rc2 = 0.0f;
  for (int k = 0; k < ntypes * ntypes; k++) rc2 = fmaxf(rc2, params[4 * k]); // Radial::getCutOff: largest pair cut-off
  const float rc = sqrtf(rc2);
  if (!(rc > 0.0f)) return UB200_ERR_INVALID_ARGUMENT;
  int rcode;
  // PairForces.cu:49-53: a box no larger than 3 cut-offs in every dimension takes the all-pairs path
  if (L[0] <= 3.0f * rc && L[1] <= 3.0f * rc && L[2] <= 3.0f * rc && !owned) {
    if (groupIdx != globalIdx) return UB200_ERR_UNSUPPORTED; // NBody reads and writes through ONE index list
    if (force && !accumulate) UB200_CUDA(cudaMemsetAsync(force, 0, sizeof(float4) * (size_t)N, st));
    e->lastPath = 2;
    return ub200_lj_nbody_f32(pos, groupIdx, N, L, periodic, params, ntypes, force, energy, virial, (void *)st);
  }
  int dims[3], per[3];
  const char *sel = getenv("UB200_LJ_ENGINE"); // "cell" forces the reference-layout traversal (A/B runs, tests)
  const bool wantColumn = !(sel && strcmp(sel, "cell") == 0);
  if (wantColumn && ljEngineDims(L, periodic, rc, dims, per)) {
    if ((rcode = uploadTable(e, params, ntypes, st))) return rcode;
    LJTableCache *cache = &e->table;
    if ((rcode = buildFine(e, pos, groupIdx, nullptr, N, nullptr, L, per, dims, makeWholeColGrid(dims, per), st))) return rcode;
    e->lastPath = 0;
    return columnSum(e, cache->dev.as<LJPar>(), ntypes, force, energy, virial, globalIdx, accumulate, ownerLo, ownerHi, st);
  }
  // grids the column traversal does not take (a periodic dimension under five half cells, huge sparse grids)
  int cellDim[3];
  if ((rcode = ub200_neighbour_celldim_f32(L, rc, cellDim))) return rcode;
  if ((rcode = ub200_celllist_build_f32(e->cl, pos, groupIdx, N, L, periodic, cellDim, (void *)st))) return rcode;
  e->lastPath = 1;
  return ljSum(e->cl, params, ntypes, force, energy, virial, globalIdx, accumulate, &e->table, st, ownerLo, ownerHi);
}

} // namespace ub200

using namespace ub200;

extern "C" {

int ub200_ljengine_create(ub200_ljengine **out) {
  if (!out) return UB200_ERR_INVALID_ARGUMENT;
  ub200_ljengine *e = new (std::nothrow) ub200_ljengine();
  if (!e) return UB200_ERR_ALLOC;
  const int rc = ub200_celllist_create(&e->cl);
  if (rc) { delete e; return rc; }
  *out = e;
  return UB200_OK;
}

int ub200_ljengine_destroy(ub200_ljengine *e) {
  if (!e) return UB200_OK;
  ub200_celllist_destroy(e->cl);
  DevBuf *bufs[] = {&e->pos, &e->idx, &e->binCount, &e->binStart, &e->blockSums, &e->codeSlot, &e->unstable, &e->errorFlag,
                    &e->table.dev};
  for (DevBuf *b : bufs) b->release();
  delete e;
  return UB200_OK;
}

int ub200_ljengine_sum_f32(ub200_ljengine *e, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                           const int periodic[3], const float *params, int ntypes, void *d_force, float *d_energy,
                           float *d_virial, const int *d_globalIdx, int accumulate, int ownerLo, int ownerHi, void *stream) {
  return ljEngineSum(e, (const float4 *)d_pos, d_groupIdx, N, L, periodic, params, ntypes, (float4 *)d_force, d_energy,
                     d_virial, d_globalIdx, accumulate != 0, ownerLo, ownerHi, (cudaStream_t)stream);
}

// traversal alone over the half-cell list of the last ub200_ljengine_sum_f32 (positions unchanged): kernel timing and
// profiling; forces only
int ub200_ljengine_traverse_f32(ub200_ljengine *e, void *d_force, int accumulate, void *stream) {
  if (!e || !d_force) return UB200_ERR_INVALID_ARGUMENT;
  if (e->lastPath != 0 || !e->table.dev.p) return UB200_ERR_NOT_BUILT;
  const int ntypes = (int)lround(sqrt((double)(e->table.host.size() / 4)));
  return columnSum(e, e->table.dev.as<LJPar>(), ntypes, (float4 *)d_force, nullptr, nullptr, nullptr, accumulate != 0, 0,
                   0x7fffffff, (cudaStream_t)stream);
}

// the half-cell list alone (neighbour search without a traversal): what b200::ColumnList::update calls before it runs a
// user Transverser through the header-template traversal over ub200_ljengine_view
int ub200_ljengine_build_f32(ub200_ljengine *e, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                             const int periodic[3], float cutOff, void *stream) {
  if (!e || !d_pos || N <= 0 || !L || !periodic || !(cutOff > 0)) return UB200_ERR_INVALID_ARGUMENT;
  int dims[3], per[3];
  if (!ljEngineDims(L, periodic, cutOff, dims, per)) return UB200_ERR_UNSUPPORTED;
  const int rc = buildFine(e, (const float4 *)d_pos, d_groupIdx, nullptr, N, nullptr, L, per, dims, makeWholeColGrid(dims, per),
                           (cudaStream_t)stream);
  if (!rc) e->lastPath = 0;
  return rc;
}
int ub200_ljengine_view_get(ub200_ljengine *e, ub200_ljengine_view *v) {
  if (!e || !v) return UB200_ERR_INVALID_ARGUMENT;
  if (e->lastPath != 0 || !e->pos.p) return UB200_ERR_NOT_BUILT;
  v->d_pos = e->pos.p;
  v->d_index = e->idx.as<int>();
  v->d_cellStart = e->binStart.as<uint32_t>();
  v->cells[0] = e->cg.nx; v->cells[1] = e->cg.ny; v->cells[2] = e->cg.nz;
  v->periodic[0] = e->cg.px; v->periodic[1] = e->cg.py; v->periodic[2] = e->cg.pz;
  v->L[0] = e->grid.Lx; v->L[1] = e->grid.Ly; v->L[2] = e->grid.Lz;
  v->numberParticles = e->N;
  return UB200_OK;
}

int ub200_ljengine_last_path(ub200_ljengine *e) { return e ? e->lastPath : -1; }

int ub200_ljengine_error_flag(ub200_ljengine *e, void *stream, int *flag) {
  if (!e || !flag) return UB200_ERR_INVALID_ARGUMENT;
  *flag = 0;
  if (!e->errorFlag.p) return UB200_OK;
  UB200_CUDA(cudaMemcpyAsync(flag, e->errorFlag.p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  UB200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return UB200_OK;
}

int ub200_ljengine_grid(ub200_ljengine *e, int cells[3]) {
  if (!e || !cells) return UB200_ERR_INVALID_ARGUMENT;
  cells[0] = e->cg.nx; cells[1] = e->cg.ny; cells[2] = e->cg.nz;
  return UB200_OK;
}
}
