// Verlet (skin) list of the built-in LJ traversal, sm_100a: filled from the half-cell columns, row per particle.
//
// PairForces<LJ, VerletList>::sum (Interactor/PairForces.cu:43-78 over NeighbourList/VerletList.cuh:111-159) only needs
// the FORCES the reference's list produces; the reference-layout list ([k * N + i], neighbour order of the 27-cell walk,
// BasicList/BasicListBase.cuh:42-71) stays available through ub200_verletlist_view_get for callers that read it, and is
// built only when somebody asks. The list the built-in traversal walks is this one:
//
//  * search on the engine's half-cell grid (colgeom.h) with cells >= r_list / 2: 5^3 half cells = 15.6 r_list^3 instead
//    of 27 r_list^3 tested per particle. One warp per column of kVlTZ half cells; the halo is staged once in shared
//    memory with the periodic image shift of each row piece applied; eight lanes per home particle test eight candidates
//    per iteration into per-lane bit masks; a scan over the eight lanes places the hits in the row, which is assembled
//    in shared memory and written out in 16-byte pieces (padded with the particle itself to a multiple of 16 entries);
//  * one ROW per particle ([i * stride + k], 16-byte aligned): the eight lanes of a home particle write and later read
//    consecutive words. An entry is the neighbour's slot in the half-cell order plus, in the five top bits, the periodic
//    image of that neighbour as seen from the particle's column - so the traversal needs no minimum-image arithmetic:
//    13 (no shift) for all but the pairs that straddle the box boundary;
//  * positions are kept in half-cell order (refreshed every step, moved to the image nearest their build-time coordinate,
//    so that the stored image codes stay valid between rebuilds);
//  * traversal: four particles per warp, eight lanes each, four entries per lane and iteration (two 8-byte list loads,
//    four 16-byte position gathers in flight; neighbours of the lanes of a group are mostly consecutive slots of one
//    x-row: few L1 lines per gather), folded-constant LJ body, three-level butterfly per particle.
#include "lj_engine.cuh"
#include "lj_pair.cuh"
#include <cstdlib>
#include <cstring>

namespace ub200 {

constexpr int kVlTZ = 4;                  // home half cells per column (the list radius makes cells larger than the engine's)
constexpr int kVlPlanes = kVlTZ + 4;
constexpr int kVlCap = 480;               // staged candidates per warp
constexpr int kVlWarps = 4;
constexpr int kVlThreads = 32 * kVlWarps;
constexpr int kVlSlack = 32;              // slice entries the unrolled distance loop may read past a particle's range
constexpr int kVlRowBuf = 128;            // row entries per home particle assembled in shared memory
constexpr int kVlMeta = (kVlPlanes + 1) + (kVlTZ + 1) + 2 * kVlTZ;
constexpr int kVlIndexBits = 27;
constexpr uint32_t kVlIndexMask = (1u << kVlIndexBits) - 1u;
constexpr uint32_t kVlNoShift = 13u;      // (sx + 1) + 3 (sy + 1) + 9 (sz + 1) with no shift
static_assert(5 * kVlPlanes <= 64, "two rows per lane");
static_assert(5 * (kVlTZ + 1) + 2 < 32, "home rows of a column live in the first row of every lane");

__device__ __forceinline__ uint32_t shiftCode(int sx, int sy, int sz) {
  return (uint32_t)((sx + 1) + 3 * (sy + 1) + 9 * (sz + 1));
}
__device__ __forceinline__ void applyShiftCode(float4 &p, uint32_t code, float Lx, float Ly, float Lz) {
  const int c = (int)code;
  const int sz = c / 9, sy = (c - 9 * sz) / 3, sx = c - 9 * sz - 3 * sy;
  p.x += (float)(sx - 1) * Lx;
  p.y += (float)(sy - 1) * Ly;
  p.z += (float)(sz - 1) * Lz;
}

// fastPos[k] = current position of the particle in half-cell slot k, at the periodic image nearest canon[k] (its folded
// coordinate when the list was built): the image codes of the list refer to that frame
__global__ void __launch_bounds__(256)
vlistPositions(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, const int *__restrict__ fineIdx,
               const float4 *__restrict__ canon, int N, GridF g, float4 *__restrict__ fastPos) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = fineIdx[k];
  const float4 cur = ldg4(pos + (groupIdx ? groupIdx[i] : i)), c = canon[k];
  fastPos[k] = make_float4(c.x + foldCoord(cur.x - c.x, g.Lx, g.mx), c.y + foldCoord(cur.y - c.y, g.Ly, g.my),
                           c.z + foldCoord(cur.z - c.z, g.Lz, g.mz), cur.w);
}

// The same refresh fused with VerletListBase_ns::checkMaximumDrift (VerletListBase.cuh:57-69): the displacement since the
// rebuild is the folded difference to the build-time coordinate, already at hand. A particle at or over the threshold writes
// `epoch` into the flag word (no memset between calls: the host compares with the epoch it passed).
__global__ void __launch_bounds__(256)
vlistPositionsCheck(const float4 *__restrict__ pos, const int *__restrict__ groupIdx, const int *__restrict__ fineIdx,
                    const float4 *__restrict__ canon, int N, GridF g, float maxDist2, uint32_t epoch, float4 *__restrict__ fastPos,
                    uint32_t *__restrict__ flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int i = fineIdx[k];
  const float4 cur = ldg4(pos + (groupIdx ? groupIdx[i] : i)), c = canon[k];
  const float dx = foldCoord(cur.x - c.x, g.Lx, g.mx), dy = foldCoord(cur.y - c.y, g.Ly, g.my), dz = foldCoord(cur.z - c.z, g.Lz, g.mz);
  fastPos[k] = make_float4(c.x + dx, c.y + dy, c.z + dz, cur.w);
  if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) >= maxDist2) *flag = epoch;
}

// One warp per column of kVlTZ half cells; see the header. list rows hold at most `stride` entries: a particle with more
// neighbours reports its count through `overflow` and the host retries with longer rows (BasicListBase.cuh:176-181).
__global__ void __launch_bounds__(kVlThreads, 5)
vlistColumnFill(const float4 *__restrict__ finePos, const uint32_t *__restrict__ binStart, ColGrid cg, float Lx, float Ly,
                float Lz, float rcut2, int stride, int *__restrict__ list, int *__restrict__ count,
                uint32_t *__restrict__ overflow, int *__restrict__ nextColumn) {
  __shared__ __align__(16) float4 candAll[kVlWarps][kVlCap + kVlSlack]; // .w = list entry of the candidate
  __shared__ __align__(16) int rowBufAll[kVlWarps][4][kVlRowBuf];
  __shared__ int metaAll[kVlWarps][kVlMeta];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
  float4 *cand = candAll[warp];
  int *rb = rowBufAll[warp][lane >> 3];
  int *planeOff = metaAll[warp];              // [kVlPlanes + 1] first slice slot of plane p; [nPlanes] = total
  int *homePre = planeOff + kVlPlanes + 1;    // [kVlTZ + 1] home particles of the column before home cell hz
  int *homeOff = homePre + kVlTZ + 1;         // [kVlTZ] slice slot of the first particle of home cell hz
  int *homeG = homeOff + kVlTZ;               // [kVlTZ] its slot in the half-cell order
  const int nzc = (cg.nz + kVlTZ - 1) / kVlTZ;
  const int ncols = cg.nx * cg.ny * nzc;
  const int firstDynamic = gridDim.x * kVlWarps;
  int col = blockIdx.x * kVlWarps + warp, nextCol = 0;
  for (; col < ncols; col = nextCol) {
    if (lane == 0) nextCol = firstDynamic + atomicAdd(nextColumn, 1);
    nextCol = __shfl_sync(0xffffffffu, nextCol, 0);
    const int x0 = col % cg.nx, t1 = col / cg.nx, y0 = t1 % cg.ny, z0 = (t1 / cg.ny) * kVlTZ;
    const int nHome = min(kVlTZ, cg.nz - z0);
    const int nRows = 5 * (nHome + 4);
    // ---- two rows per lane: global ranges of the row pieces, their image, the home cell of the row (dy == 0 rows)
    int g0[2][2], cn[2][2], sxyz[2][4], hG = 0, hC = 0, hRel = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int r = lane + 32 * q;
      g0[q][0] = g0[q][1] = 0; cn[q][0] = cn[q][1] = 0;
      sxyz[q][0] = sxyz[q][1] = sxyz[q][2] = sxyz[q][3] = 0;
      if (r < nRows) {
        const ColRow row = columnRow(cg, x0, y0, z0, r);
#pragma unroll
        for (int s = 0; s < 2; s++)
          if (row.n[s] > 0) {
            const uint32_t a = __ldg(binStart + row.c0[s]), b = __ldg(binStart + row.c0[s] + row.n[s]);
            g0[q][s] = (int)a;
            cn[q][s] = (int)(b - a);
          }
        sxyz[q][0] = row.sx[0]; sxyz[q][1] = row.sx[1]; sxyz[q][2] = row.sy; sxyz[q][3] = row.sz;
        const int p = r / 5;
        if (q == 0 && r - 5 * p == 2 && p >= 2 && p < 2 + nHome) {
          const int cc = x0 + cg.nx * (y0 + cg.ny * (z0 + p - 2));
          const uint32_t a = __ldg(binStart + cc), b = __ldg(binStart + cc + 1);
          hG = (int)a;
          hC = (int)(b - a);
          hRel = row.hs ? cn[0][0] + ((int)a - g0[0][1]) : (int)a - g0[0][0];
        }
      }
    }
    if (!__any_sync(0xffffffffu, hC > 0)) continue; // no home particle in this column
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
    // home cell hz sits in row 5 (hz + 2) + 2 (< 32): bring its data to lane hz, count the home particles before it
    const int hsrc = (5 * lane + 12) & 31;
    const int a0 = __shfl_sync(0xffffffffu, hC, hsrc), b0 = __shfl_sync(0xffffffffu, hG, hsrc);
    const int c0s = __shfl_sync(0xffffffffu, off[0] + hRel, hsrc);
    const bool isHome = lane < nHome;
    const int myCnt = isHome ? a0 : 0;
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
    if (lane <= kVlTZ) homePre[lane] = isHome ? pre - myCnt : 0x3fffffff;
    if (isHome) {
      homeOff[lane] = c0s;
      homeG[lane] = b0;
    }
    const int nHomeP = __shfl_sync(0xffffffffu, pre, kVlTZ - 1);
    const bool staged = total <= kVlCap && nHomeP <= 32; // warp uniform
    if (staged) {
      // every lane copies its own rows into the slice, image shift applied on the way, and notes where each came from
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float dy = (float)sxyz[q][2] * Ly, dz = (float)sxyz[q][3] * Lz;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const float dx = (float)sxyz[q][s] * Lx;
          const uint32_t code = shiftCode(sxyz[q][s], sxyz[q][2], sxyz[q][3]) << kVlIndexBits;
          const int b = off[q] + (s ? cn[q][0] : 0);
          for (int k = 0; k < cn[q][s]; k++) {
            float4 p = ldg4(finePos + g0[q][s] + k);
            p.x += dx; p.y += dy; p.z += dz;
            p.w = __int_as_float((int)(code | (uint32_t)(g0[q][s] + k)));
            cand[b + k] = p;
          }
        }
      }
    }
    __syncwarp();
    if (staged) {
      // record of home particle q (lane q): candidate range = planes hz .. hz + 4 of its home cell (contiguous), own slot
      int recC0 = 0, recC1 = 0, recSlot = 0, recG = 0;
      if (lane < nHomeP) {
        int hz = 0;
#pragma unroll
        for (int k = 1; k < kVlTZ; k++) hz += lane >= homePre[k];
        const int hrel = lane - homePre[hz];
        recC0 = planeOff[hz];
        recC1 = planeOff[hz + 5];
        recSlot = homeOff[hz] + hrel;
        recG = homeG[hz] + hrel;
      }
      // Passes of four home particles, eight lanes each. Lane `sub` of a group owns the candidates c0 + sub + 8 k: it
      // first collects its hits in a bit mask (no warp-level operation in the distance loop), an 8-lane scan of the hit
      // counts then gives every lane its place in the row, the entries go to shared memory and leave as 16-byte stores.
      // (the order of a row is lane-major: any fixed order serves the traversal)
      for (int q0 = 0; q0 < nHomeP; q0 += 4) {
        const int q = q0 + (lane >> 3);
        const bool act = q < nHomeP;
        const int src = act ? q : 0;
        const int c0 = __shfl_sync(0xffffffffu, recC0, src), c1 = __shfl_sync(0xffffffffu, recC1, src);
        const int slot = __shfl_sync(0xffffffffu, recSlot, src), gs = __shfl_sync(0xffffffffu, recG, src);
        const int len = act ? c1 - c0 : 0;
        // lane `sub` owns the candidates c0 + sub chunk .. + chunk - 1: lane-major order is then the staging order (the
        // traversal gathers consecutive slots with consecutive lanes). chunk is odd: the 16-byte loads of the eight lanes
        // of a group start in eight different bank quads
        const int chunk = ((len + 7) >> 3) | 1;
        const int nk = min(max(len - sub * chunk, 0), chunk); // <= kVlCap / 8 + 1 = 61
        const float4 pi = cand[slot];
        const float4 *cp = cand + c0 + sub * chunk;
        uint32_t m0 = 0, m1 = 0;
        auto scan4 = [&](int k0) {
          uint32_t bits = 0;
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const float4 pj = cp[k0 + u];
            const float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
            if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) <= rcut2) bits |= 1u << u;
          }
          return bits;
        };
        const int nkLo = min(nk, 32);
        for (int k0 = 0; k0 < nkLo; k0 += 4) m0 |= scan4(k0) << k0;
        for (int k0 = 32; k0 < nk; k0 += 4) m1 |= scan4(k0) << (k0 - 32);
        // the unrolled loop may have tested up to three entries past the lane's last candidate; the particle itself
        if (nk < 32) m0 &= (1u << nk) - 1u;
        m1 = nk <= 32 ? 0u : (nk < 64 ? m1 & ((1u << (nk - 32)) - 1u) : m1);
        const int ds = slot - c0 - sub * chunk;
        if (act && ds >= 0 && ds < nk) {
          if (ds < 32) m0 &= ~(1u << ds);
          else m1 &= ~(1u << (ds - 32));
        }
        const int mine = __popc(m0) + __popc(m1);
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, inc, o, 8);
          if (sub >= o) inc += u;
        }
        const int n = __shfl_sync(0xffffffffu, inc, 7, 8);
        int *row = list + (size_t)gs * stride;
        if (!__any_sync(0xffffffffu, n > kVlRowBuf)) {
          // every row of the pass fits its shared-memory buffer (the usual case)
          int *dst = rb + (inc - mine);
          const float *wp = &cp[0].w;
          auto emit = [&](uint32_t m, const float *w) {
            while (m) {
              const int k = __ffs((int)m) - 1;
              m &= m - 1u;
              *dst++ = __float_as_int(w[4 * k]);
            }
          };
          emit(m0, wp);
          emit(m1, wp + 4 * 32);
        } else {
          int at = inc - mine;
          auto emit = [&](uint32_t m, int kbase) {
            while (m) {
              const int k = __ffs((int)m) - 1;
              m &= m - 1u;
              const int e = __float_as_int(cp[kbase + k].w);
              if (at < kVlRowBuf) rb[at] = e;
              else if (at < stride) row[at] = e;
              at++;
            }
          };
          emit(m0, 0);
          emit(m1, 32);
        }
        __syncwarp();
        // rows are padded to a multiple of 8 entries with the particle itself (r2 = 0: no force), so that the traversal
        // reads whole 32-byte pieces without a bound check
        const int nw = min(n, stride), npad = (nw + 7) & ~7; // stride is a multiple of 16
        const int selfE = (int)((kVlNoShift << kVlIndexBits) | (uint32_t)gs);
        if (act) {
          const int nb = min(npad, kVlRowBuf);
          for (int k = 4 * sub; k < nb; k += 32) {
            int4 e = *reinterpret_cast<const int4 *>(rb + k);
            if (k + 0 >= nw) e.x = selfE;
            if (k + 1 >= nw) e.y = selfE;
            if (k + 2 >= nw) e.z = selfE;
            if (k + 3 >= nw) e.w = selfE;
            *reinterpret_cast<int4 *>(row + k) = e;
          }
          for (int k = max(nw, kVlRowBuf) + sub; k < npad; k += 8) row[k] = selfE;
          if (sub == 0) {
            count[gs] = nw;
            if (n > stride) atomicMax(overflow, (uint32_t)n);
          }
        }
        __syncwarp();
      }
    } else {
      // ---- dense column: one home particle at a time, the warp walks the 25 rows of its home cell in global memory
      for (int q = 0; q < nHomeP; q++) {
        int hz = 0;
#pragma unroll
        for (int k = 1; k < kVlTZ; k++) hz += q >= homePre[k];
        const int gs = homeG[hz] + q - homePre[hz];
        const float4 pi = ldg4(finePos + gs);
        int *row = list + (size_t)gs * stride;
        int n = 0;
        for (int rr = 5 * hz; rr < 5 * hz + 25; rr++) {
          const int src = rr & 31, hi = rr >> 5;
          const int sy = __shfl_sync(0xffffffffu, hi ? sxyz[1][2] : sxyz[0][2], src);
          const int sz = __shfl_sync(0xffffffffu, hi ? sxyz[1][3] : sxyz[0][3], src);
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int g = __shfl_sync(0xffffffffu, hi ? g0[1][s] : g0[0][s], src);
            const int cnt = __shfl_sync(0xffffffffu, hi ? cn[1][s] : cn[0][s], src);
            const int sx = __shfl_sync(0xffffffffu, hi ? sxyz[1][s] : sxyz[0][s], src);
            const float dx = (float)sx * Lx, dy = (float)sy * Ly, dz = (float)sz * Lz;
            const uint32_t code = shiftCode(sx, sy, sz) << kVlIndexBits;
            for (int t0 = 0; t0 < cnt; t0 += 32) {
              const int t = t0 + lane;
              bool hit = false;
              if (t < cnt) {
                const float4 pj = ldg4(finePos + g + t);
                const float ex = pj.x + dx - pi.x, ey = pj.y + dy - pi.y, ez = pj.z + dz - pi.z;
                hit = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, ex * ex)) <= rcut2 && g + t != gs;
              }
              const unsigned m = __ballot_sync(0xffffffffu, hit);
              const int at = n + __popc(m & ((1u << lane) - 1u));
              if (hit && at < stride) row[at] = (int)(code | (uint32_t)(g + t));
              n += __popc(m);
            }
          }
        }
        const int nw = min(n, stride), npad = (nw + 7) & ~7;
        if (nw + lane < npad) row[nw + lane] = (int)((kVlNoShift << kVlIndexBits) | (uint32_t)gs);
        if (lane == 0) {
          count[gs] = nw;
          if (n > stride) atomicMax(overflow, (uint32_t)n);
        }
      }
    }
  }
}

// LJ over the row list: four particles per warp, eight lanes each (see the header)
template <bool ENERGY, bool VIRIAL, bool MULTITYPE, bool ACCUMULATE>
__global__ void __launch_bounds__(256)
ljListTraversal(const float4 *__restrict__ fastPos, const int *__restrict__ fineIdx, const int *__restrict__ list,
                const int *__restrict__ count, int stride, int N, float Lx, float Ly, float Lz,
                const LJPar *__restrict__ parTable, int ntypes, float4 *__restrict__ force, float *__restrict__ energy,
                float *__restrict__ virial, const int *__restrict__ globalIdx) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int p = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  const bool act = p < N;
  const int ps = act ? p : N - 1;
  const int *row = list + (size_t)ps * stride + sub;
  const int self = (int)((kVlNoShift << kVlIndexBits) | (uint32_t)ps);
  const LJPar par0 = parTable[0];
  Acc a = Acc{0.f, 0.f, 0.f, 0.f, 0.f};
  float4 pi;
  int npad;
  if (!ENERGY && !VIRIAL && !MULTITYPE) {
    // The first 64 entries of the row (rows are never shorter) are requested together with the particle and its count, so
    // the list costs one memory latency; what lies past the padded count is replaced by the particle itself (r2 = 0).
    int e[8];
#pragma unroll
    for (int m = 0; m < 8; m++) e[m] = __ldg(row + 8 * m);
    pi = ldg4(fastPos + ps);
    const int n = act ? __ldg(count + ps) : 0;
    npad = (n + 7) & ~7; // rows are padded with the particle itself to a multiple of 8 entries
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (8 * m >= npad) e[m] = self;
    const LJFold fold = foldLJ(par0);
    auto four = [&](uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3) {
      float4 p0 = ldg4(fastPos + (e0 & kVlIndexMask)), p1 = ldg4(fastPos + (e1 & kVlIndexMask));
      float4 p2 = ldg4(fastPos + (e2 & kVlIndexMask)), p3 = ldg4(fastPos + (e3 & kVlIndexMask));
      if ((e0 >> kVlIndexBits) != kVlNoShift) applyShiftCode(p0, e0 >> kVlIndexBits, Lx, Ly, Lz);
      if ((e1 >> kVlIndexBits) != kVlNoShift) applyShiftCode(p1, e1 >> kVlIndexBits, Lx, Ly, Lz);
      if ((e2 >> kVlIndexBits) != kVlNoShift) applyShiftCode(p2, e2 >> kVlIndexBits, Lx, Ly, Lz);
      if ((e3 >> kVlIndexBits) != kVlNoShift) applyShiftCode(p3, e3 >> kVlIndexBits, Lx, Ly, Lz);
      ljPairFolded(p0.x - pi.x, p0.y - pi.y, p0.z - pi.z, fold, a);
      ljPairFolded(p1.x - pi.x, p1.y - pi.y, p1.z - pi.z, fold, a);
      ljPairFolded(p2.x - pi.x, p2.y - pi.y, p2.z - pi.z, fold, a);
      ljPairFolded(p3.x - pi.x, p3.y - pi.y, p3.z - pi.z, fold, a);
    };
    four((uint32_t)e[0], (uint32_t)e[1], (uint32_t)e[2], (uint32_t)e[3]);
    if (npad > 32) four((uint32_t)e[4], (uint32_t)e[5], (uint32_t)e[6], (uint32_t)e[7]);
    for (int k = 64; k < npad; k += 32) {
      const uint32_t e0 = (uint32_t)__ldg(row + k);
      const uint32_t e1 = (uint32_t)(k + 8 < npad ? __ldg(row + k + 8) : self);
      const uint32_t e2 = (uint32_t)(k + 16 < npad ? __ldg(row + k + 16) : self);
      const uint32_t e3 = (uint32_t)(k + 24 < npad ? __ldg(row + k + 24) : self);
      four(e0, e1, e2, e3);
    }
  } else {
    pi = ldg4(fastPos + ps);
    const int n = act ? __ldg(count + ps) : 0;
    npad = (n + 7) & ~7;
    LJPar par = par0;
    uint32_t rcb = __float_as_uint(par.cutOff2) - 1u;
    const int ti = (int)pi.w;
    const int trow = MULTITYPE && (unsigned)ti < (unsigned)ntypes ? ti * ntypes : -1;
    auto one = [&](float4 pj, uint32_t e) {
      if ((e >> kVlIndexBits) != kVlNoShift) applyShiftCode(pj, e >> kVlIndexBits, Lx, Ly, Lz);
      if (MULTITYPE) {
        const int tj = (int)pj.w;
        par = parTable[((unsigned)tj < (unsigned)ntypes && trow >= 0) ? trow + tj : 0];
        rcb = __float_as_uint(par.cutOff2) - 1u;
      }
      ljPair<ENERGY, VIRIAL>(pj.x - pi.x, pj.y - pi.y, pj.z - pi.z, par, rcb, a);
    };
    for (int k = 0; k < npad; k += 16) {
      const uint32_t e0 = (uint32_t)__ldg(row + k), e1 = (uint32_t)(k + 8 < npad ? __ldg(row + k + 8) : self);
      const float4 p0 = ldg4(fastPos + (e0 & kVlIndexMask)), p1 = ldg4(fastPos + (e1 & kVlIndexMask));
      one(p0, e0);
      one(p1, e1);
    }
  }
  __syncwarp();
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a.fx += __shfl_xor_sync(0xffffffffu, a.fx, o);
    a.fy += __shfl_xor_sync(0xffffffffu, a.fy, o);
    a.fz += __shfl_xor_sync(0xffffffffu, a.fz, o);
    if (ENERGY) a.e += __shfl_xor_sync(0xffffffffu, a.e, o);
    if (VIRIAL) a.v += __shfl_xor_sync(0xffffffffu, a.v, o);
  }
  if (!act || sub != 0) return;
  const int gi = fineIdx[p];
  const int ori = globalIdx ? globalIdx[gi] : gi;
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

// ---- host side ----

bool vlistApplies(const float L[3], const int periodic[3], float rcut, int N) {
  int dims[3], per[3];
  const char *sel = getenv("UB200_VERLET_FAST"); // "0": reference-layout list only (A/B runs, tests)
  if (sel && sel[0] == '0') return false;
  return (unsigned)N < (1u << kVlIndexBits) && ljEngineDims(L, periodic, rcut, dims, per);
}

// half-cell list of the stored positions (group order) + row list; v->storedPos holds the positions
int vlistRebuild(ub200_verletlist *v, cudaStream_t st) {
  const int N = v->N;
  const float rcut = v->cutOff * v->multiplier;
  int rc;
  if ((rc = ub200_ljengine_build_f32(v->eng, v->storedPos.p, nullptr, N, v->L, v->periodic, rcut, (void *)st))) return rc;
  ub200_ljengine *e = v->eng;
  if ((rc = v->fastPos.reserve(sizeof(float4) * (size_t)N)) || (rc = v->fastCount.reserve(sizeof(int) * (size_t)N))) return rc;
  static int blocksPerSM = 0;
  if (!blocksPerSM) {
    UB200_CUDA(cudaFuncSetAttribute(vlistColumnFill, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, vlistColumnFill, kVlThreads, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
  }
  const ColGrid &cg = e->cg;
  const int ncols = cg.nx * cg.ny * ((cg.nz + kVlTZ - 1) / kVlTZ);
  int grid = kNumSMs * blocksPerSM;
  const int needed = (ncols + kVlWarps - 1) / kVlWarps;
  if (grid > needed) grid = needed;
  // a list radius a hair above the nominal one: a pair the staged fp32 coordinates put an ulp outside is kept
  const float rcut2 = rcut * rcut * (1.0f + 4e-6f);
  while (true) {
    if ((rc = v->fastList.reserve(sizeof(int) * (size_t)N * v->fastStride))) return rc;
    UB200_CUDA(cudaMemsetAsync(v->flags.as<uint32_t>() + 1, 0, sizeof(uint32_t), st));
    UB200_CUDA(cudaMemsetAsync(e->errorFlag.as<int>() + 1, 0, sizeof(int), st)); // the column counter
    vlistColumnFill<<<grid, kVlThreads, 0, st>>>(e->pos.as<float4>(), e->binStart.as<uint32_t>(), cg, e->grid.Lx, e->grid.Ly,
                                                 e->grid.Lz, rcut2, v->fastStride, v->fastList.as<int>(), v->fastCount.as<int>(),
                                                 v->flags.as<uint32_t>() + 1, e->errorFlag.as<int>() + 1);
    UB200_LAUNCHED();
    uint32_t over = 0;
    UB200_CUDA(cudaMemcpyAsync(&over, v->flags.as<uint32_t>() + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    UB200_CUDA(cudaStreamSynchronize(st));
    if (!over) break;
    v->fastStride = ((int)over + 15) / 16 * 16 + 16;
  }
  return UB200_OK;
}

int vlistRefreshPositions(ub200_verletlist *v, const float4 *pos, const int *groupIdx, cudaStream_t st) {
  const int cd1[3] = {1, 1, 1};
  const GridF g = makeGridF(v->L, v->periodic, cd1);
  ub200_ljengine *e = v->eng;
  vlistPositions<<<(v->N + 255) / 256, 256, 0, st>>>(pos, groupIdx, e->idx.as<int>(), e->pos.as<float4>(), v->N, g,
                                                     v->fastPos.as<float4>());
  UB200_LAUNCHED();
  return UB200_OK;
}

// refresh + drift check in one pass; *over = some particle moved maxDist or more since the rebuild (synchronises the stream)
int vlistRefreshAndCheck(ub200_verletlist *v, const float4 *pos, const int *groupIdx, float maxDist, bool *over, cudaStream_t st) {
  const int cd1[3] = {1, 1, 1};
  const GridF g = makeGridF(v->L, v->periodic, cd1);
  ub200_ljengine *e = v->eng;
  const uint32_t epoch = ++v->driftEpoch;
  vlistPositionsCheck<<<(v->N + 255) / 256, 256, 0, st>>>(pos, groupIdx, e->idx.as<int>(), e->pos.as<float4>(), v->N, g,
                                                          maxDist * maxDist, epoch, v->fastPos.as<float4>(), v->flags.as<uint32_t>());
  UB200_LAUNCHED();
  uint32_t seen = 0;
  UB200_CUDA(cudaMemcpyAsync(&seen, v->flags.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  UB200_CUDA(cudaStreamSynchronize(st));
  *over = seen == epoch;
  return UB200_OK;
}

int vlistSum(ub200_verletlist *v, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
             const int *globalIdx, bool accumulate, cudaStream_t st) {
  const bool E = energy != nullptr, V = virial != nullptr, M = ntypes > 1;
  if ((E || V) && !accumulate) return UB200_ERR_UNSUPPORTED;
  ub200_ljengine *e = v->eng;
  const int N = v->N, nb = (N + 31) / 32;
#define UB200_VL(ee, vv, mm, aa)                                                                                        \
  if (E == ee && V == vv && M == mm && accumulate == aa) {                                                              \
    ljListTraversal<ee, vv, mm, aa><<<nb, 256, 0, st>>>(v->fastPos.as<float4>(), e->idx.as<int>(), v->fastList.as<int>(), \
                                                        v->fastCount.as<int>(), v->fastStride, N, e->grid.Lx, e->grid.Ly, \
                                                        e->grid.Lz, table, ntypes, force, energy, virial, globalIdx);   \
    UB200_LAUNCHED();                                                                                                   \
    return UB200_OK;                                                                                                    \
  }
  UB200_VL(false, false, false, false) UB200_VL(false, false, true, false) UB200_VL(false, false, false, true)
  UB200_VL(false, false, true, true) UB200_VL(true, false, false, true) UB200_VL(true, false, true, true)
  UB200_VL(false, true, false, true) UB200_VL(false, true, true, true) UB200_VL(true, true, false, true)
  UB200_VL(true, true, true, true)
#undef UB200_VL
  return UB200_ERR_UNSUPPORTED;
}

} // namespace ub200

using namespace ub200;

// the row list itself (tests, tools): entry = slot in half-cell order | image code << 27
extern "C" int ub200_verletlist_rows_get(ub200_verletlist *v, ub200_verletlist_rows *out) {
  if (!v || !out) return UB200_ERR_INVALID_ARGUMENT;
  if (!v->N || !v->fast) return UB200_ERR_NOT_BUILT;
  out->d_list = v->fastList.as<int>();
  out->d_count = v->fastCount.as<int>();
  out->d_pos = v->fastPos.p;
  out->d_index = v->eng->idx.as<int>();
  out->stride = v->fastStride;
  out->numberParticles = v->N;
  out->indexBits = kVlIndexBits;
  return UB200_OK;
}
