// Short-range pair traversal specialised for the Lennard-Jones transverser, sm_100a.
//
// Replaces NeighbourList_ns::transverseWithNeighbourContainer (Interactor/NeighbourList/common.cuh:10-34)
// + CellList_ns::NeighbourIterator (CellList/NeighbourContainer.cuh:95-138) + Radial<LJFunctor>::Transverser
// (Potential/RadialPotential.cuh:107-127, Potential/Potential.cuh:37-65).
//
// Design (B200): persistent warps walk the home cells, one WARP per cell (no block barriers). For each home
// cell the particles of its (up to) 27 neighbour cells are staged ONCE into the warp's slice of shared memory,
// already folded into the primary box and displaced by the periodic image shift of their cell, so the inner
// loop needs no per-pair minimum-image arithmetic. The warp then takes the home particles two at a time: its
// 32 lanes stride over the staged candidates (one conflict-free LDS.128 feeds two pair evaluations), the LJ
// body is branch free, and a half-warp split butterfly reduces both particles at once. The reference instead
// runs one thread per particle through a divergent 27-cell iterator with ~340 dependent global loads.
#include "lj_pair.cuh"
#include <cstdlib>
#include <cstring>

namespace ub200 {

constexpr int kWarpCap = 416; // staged candidates per warp (6.5 KB, 8 CTAs/SM); denser neighbourhoods take the direct path

// One WARP per home cell (no block level barriers). PAIRMIC: per-pair minimum image exactly like
// Radial::Transverser::compute (box.apply_pbc(rj-ri)); needed when a periodic dimension has fewer than 4 cells
// (a collapsed dimension still wraps). Otherwise the cell image shift staged with the candidates is the
// minimum image.
// ownerLo/ownerHi: only home particles whose group index lies in [ownerLo, ownerHi) are computed and written
// (multi-GPU particle decomposition: every rank holds all positions and the full list, and computes its block).
// PACKED (force only, single type, cell image shifts): the two home particles of a pass ride in the two halves of
// f32x2 registers, so every arithmetic instruction of the pair body serves both pairs - the operation sequence per pair is
// the one of ljPair, hence the same bits (checked on a B200: 0 differing words at N = 1e6, profiles/r01e_lj_packed_ab.json).
// It is NOT faster (0.4997 vs 0.4955 ms): FFMA2 does two lanes' work in two pipe cycles, and the kernel is bound by the
// total issue slots, 42 % of which are spent outside this loop (staging, neighbour description, reductions). Kept as an
// experiment behind UB200_LJ_PACKED=1; the default path is the scalar one.
template <bool ENERGY, bool VIRIAL, bool MULTITYPE, bool PAIRMIC, bool ACCUMULATE, bool PACKED = false>
__global__ void __launch_bounds__(kPairThreads, 8)
ljCellTraversal(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex,
                const uint32_t *__restrict__ binStart, GridF g, int ncells,
                const LJPar *__restrict__ parTable, int ntypes, float4 *__restrict__ force,
                float *__restrict__ energy, float *__restrict__ virial, const int *__restrict__ globalIdx,
                int ownerLo, int ownerHi) {
  __shared__ float4 candAll[kPairWarps][kWarpCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *cand = candAll[warp];
  const int warpsTotal = gridDim.x * kPairWarps;
  const LJPar par0 = parTable[0]; // single type: BasicParameterHandler::Iterator returns entry 0 (ParameterHandler.cuh:49-50)
  const uint32_t rc2bitsm1 = __float_as_uint(par0.cutOff2) - 1u;
  for (int cell = blockIdx.x * kPairWarps + warp; cell < ncells; cell += warpsTotal) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue; // warp uniform
    if (ownerLo > 0 || ownerHi < 0x7fffffff) { // any owned home particle in this cell?
      bool mineAny = false;
      for (int h = lane; h < hCount; h += 32) {
        const int gi = groupIndex[hStart + h];
        mineAny |= (gi >= ownerLo && gi < ownerHi);
      }
      if (!__any_sync(0xffffffffu, mineAny)) continue;
    }
    const int hOff = __shfl_sync(0xffffffffu, nc.off, nc.centre);
    const bool staged = nc.total <= kWarpCap;
    const float3 hc = cellCentre(g, cx, cy, cz);
    __syncwarp();
    if (staged) {
      // two neighbour cells per pass, 16 lanes each
      const int half = lane >> 4, l16 = lane & 15;
#pragma unroll 1
      for (int c0 = 0; c0 < 27; c0 += 2) {
        const int c = c0 + half;
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        const int off = __shfl_sync(0xffffffffu, nc.off, c);
        for (int t = l16; t < cnt; t += 16) {
          float4 p = ldg4(sortPos + st + t);
          if (!PAIRMIC) toHomeImage(p, g, hc);
          cand[off + t] = p;
        }
      }
    }
    __syncwarp();
    for (int h = 0; h < hCount; h += 2) {
      const bool two = h + 1 < hCount;
      const int gi0 = groupIndex[hStart + h], gi1 = groupIndex[hStart + h + (two ? 1 : 0)];
      const bool own0 = gi0 >= ownerLo && gi0 < ownerHi, own1 = two && gi1 >= ownerLo && gi1 < ownerHi;
      if (!own0 && !own1) continue;
      float4 pi0, pi1;
      if (staged) {
        pi0 = cand[hOff + h];
        pi1 = cand[hOff + h + (two ? 1 : 0)];
      } else {
        pi0 = ldg4(sortPos + hStart + h);
        pi1 = ldg4(sortPos + hStart + h + (two ? 1 : 0));
        if (!PAIRMIC) {
          toHomeImage(pi0, g, hc);
          toHomeImage(pi1, g, hc);
        }
      }
      Acc a0 = {0.f, 0.f, 0.f, 0.f, 0.f}, a1 = {0.f, 0.f, 0.f, 0.f, 0.f};
      LJPar p0 = par0, p1 = par0;
      uint32_t rcb0 = rc2bitsm1, rcb1 = rc2bitsm1;
      // types outside the table fall back to entry 0 like BasicParameterHandler::Iterator (ParameterHandler.cuh:55-57)
      const int ty0 = (int)pi0.w, ty1 = (int)pi1.w;
      const int t0 = MULTITYPE && (unsigned)ty0 < (unsigned)ntypes ? ty0 * ntypes : -1;
      const int t1 = MULTITYPE && (unsigned)ty1 < (unsigned)ntypes ? ty1 * ntypes : -1;
      if (staged && PACKED && !ENERGY && !VIRIAL && !MULTITYPE && !PAIRMIC) {
        const f32x2 NPX = pack2(-pi0.x, -pi1.x), NPY = pack2(-pi0.y, -pi1.y), NPZ = pack2(-pi0.z, -pi1.z);
        const f32x2 S2 = pack2(par0.sigma2, par0.sigma2), EPS = pack2(par0.epsDivSigma2, par0.epsDivSigma2);
        const f32x2 M48 = pack2(-48.0f, -48.0f), C24 = pack2(24.0f, 24.0f);
        f32x2 AX = pack2(0.f, 0.f), AY = AX, AZ = AX;
#pragma unroll 2
        for (int t = lane; t < nc.total; t += 32) {
          const float4 pj = cand[t];
          const f32x2 DX = add2(pack2(pj.x, pj.x), NPX), DY = add2(pack2(pj.y, pj.y), NPY), DZ = add2(pack2(pj.z, pj.z), NPZ);
          const f32x2 R2 = fma2(DZ, DZ, fma2(DY, DY, mul2(DX, DX)));
          float r2a, r2b;
          unpack2(R2, r2a, r2b);
          const bool ina = (__float_as_uint(r2a) - 1u) < rc2bitsm1, inb = (__float_as_uint(r2b) - 1u) < rc2bitsm1;
          const float sa = ina ? r2a : __int_as_float(0x7f800000), sb = inb ? r2b : __int_as_float(0x7f800000);
          float ia, ib;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(sa));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ib) : "f"(sb));
          const f32x2 U = mul2(S2, pack2(ia, ib));
          const f32x2 U2 = mul2(U, U);
          const f32x2 U3 = mul2(U2, U);
          const f32x2 FM = mul2(mul2(EPS, fma2(M48, U3, C24)), mul2(U2, U2));
          AX = fma2(FM, DX, AX);
          AY = fma2(FM, DY, AY);
          AZ = fma2(FM, DZ, AZ);
        }
        unpack2(AX, a0.fx, a1.fx);
        unpack2(AY, a0.fy, a1.fy);
        unpack2(AZ, a0.fz, a1.fz);
      } else if (staged) {
#pragma unroll 2
        for (int t = lane; t < nc.total; t += 32) {
          const float4 pj = cand[t];
          float dx0 = pj.x - pi0.x, dy0 = pj.y - pi0.y, dz0 = pj.z - pi0.z;
          float dx1 = pj.x - pi1.x, dy1 = pj.y - pi1.y, dz1 = pj.z - pi1.z;
          if (PAIRMIC) {
            dx0 = foldCoord(dx0, g.Lx, g.mx); dy0 = foldCoord(dy0, g.Ly, g.my); dz0 = foldCoord(dz0, g.Lz, g.mz);
            dx1 = foldCoord(dx1, g.Lx, g.mx); dy1 = foldCoord(dy1, g.Ly, g.my); dz1 = foldCoord(dz1, g.Lz, g.mz);
          }
          if (MULTITYPE) {
            const int tj = (int)pj.w;
            const bool ok = (unsigned)tj < (unsigned)ntypes;
            p0 = parTable[(ok && t0 >= 0) ? t0 + tj : 0]; rcb0 = __float_as_uint(p0.cutOff2) - 1u;
            p1 = parTable[(ok && t1 >= 0) ? t1 + tj : 0]; rcb1 = __float_as_uint(p1.cutOff2) - 1u;
          }
          ljPair<ENERGY, VIRIAL>(dx0, dy0, dz0, p0, rcb0, a0);
          ljPair<ENERGY, VIRIAL>(dx1, dy1, dz1, p1, rcb1, a1);
        }
      } else {
        // dense neighbourhood: walk the neighbour cells straight from global memory
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          if (cnt == 0) continue;
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          for (int t = lane; t < cnt; t += 32) {
            float4 pj = ldg4(sortPos + st + t);
            if (!PAIRMIC) toHomeImage(pj, g, hc);
            float dx0 = pj.x - pi0.x, dy0 = pj.y - pi0.y, dz0 = pj.z - pi0.z;
            float dx1 = pj.x - pi1.x, dy1 = pj.y - pi1.y, dz1 = pj.z - pi1.z;
            if (PAIRMIC) {
              dx0 = foldCoord(dx0, g.Lx, g.mx); dy0 = foldCoord(dy0, g.Ly, g.my); dz0 = foldCoord(dz0, g.Lz, g.mz);
              dx1 = foldCoord(dx1, g.Lx, g.mx); dy1 = foldCoord(dy1, g.Ly, g.my); dz1 = foldCoord(dz1, g.Lz, g.mz);
            }
            if (MULTITYPE) {
              const int tj = (int)pj.w;
              const bool ok = (unsigned)tj < (unsigned)ntypes;
              p0 = parTable[(ok && t0 >= 0) ? t0 + tj : 0]; rcb0 = __float_as_uint(p0.cutOff2) - 1u;
              p1 = parTable[(ok && t1 >= 0) ? t1 + tj : 0]; rcb1 = __float_as_uint(p1.cutOff2) - 1u;
            }
            ljPair<ENERGY, VIRIAL>(dx0, dy0, dz0, p0, rcb0, a0);
            ljPair<ENERGY, VIRIAL>(dx1, dy1, dz1, p1, rcb1, a1);
          }
        }
      }
      reducePair(a0, a1, lane, ENERGY || VIRIAL);
      if ((lane == 0 && own0) || (lane == 16 && own1)) {
        const int gi = lane ? gi1 : gi0;
        const int ori = globalIdx ? globalIdx[gi] : gi;
        if (force) {
          if (ACCUMULATE) {
            float4 f = force[ori];
            f.x += a0.fx; f.y += a0.fy; f.z += a0.fz;
            force[ori] = f;
          } else {
            force[ori] = make_float4(a0.fx, a0.fy, a0.fz, 0.0f);
          }
        }
        if (ENERGY) energy[ori] += a0.e;
        if (VIRIAL) virial[ori] += a0.v;
      }
    }
  }
}

// LJ over the Verlet list: one thread per sorted particle (the list is [k*N + i], so a warp reads 32 consecutive
// entries per k), neighbour positions gathered from the sorted array (neighbours of neighbouring particles sit close
// together: L1/L2 resident), four neighbours in flight per thread, per-pair minimum image like
// Radial::Transverser::compute (RadialPotential.cuh:107-127). Self is in the list and contributes 0 (r2 == 0).
template <bool ENERGY, bool VIRIAL, bool MULTITYPE>
__global__ void __launch_bounds__(128)
ljVerletTraversal(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex, const int *__restrict__ neighbourList,
                  const int *__restrict__ numberNeighbours, int N, GridF g, const LJPar *__restrict__ parTable, int ntypes,
                  float4 *__restrict__ force, float *__restrict__ energy, float *__restrict__ virial,
                  const int *__restrict__ globalIdx) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  const float4 pi = ldg4(sortPos + id);
  const int nn = numberNeighbours[id];
  LJPar par = parTable[0];
  uint32_t rcb = __float_as_uint(par.cutOff2) - 1u;
  const int ti = (int)pi.w;
  const int trow = MULTITYPE && (unsigned)ti < (unsigned)ntypes ? ti * ntypes : -1;
  Acc a = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int *lp = neighbourList + id;
  int k = 0;
  auto one = [&](const float4 pj) {
    const float dx = foldCoord(pj.x - pi.x, g.Lx, g.mx), dy = foldCoord(pj.y - pi.y, g.Ly, g.my), dz = foldCoord(pj.z - pi.z, g.Lz, g.mz);
    if (MULTITYPE) {
      const int tj = (int)pj.w;
      par = parTable[((unsigned)tj < (unsigned)ntypes && trow >= 0) ? trow + tj : 0];
      rcb = __float_as_uint(par.cutOff2) - 1u;
    }
    ljPair<ENERGY, VIRIAL>(dx, dy, dz, par, rcb, a);
  };
  for (; k + 4 <= nn; k += 4) {
    const int j0 = __ldg(lp + (size_t)k * N), j1 = __ldg(lp + (size_t)(k + 1) * N), j2 = __ldg(lp + (size_t)(k + 2) * N),
              j3 = __ldg(lp + (size_t)(k + 3) * N);
    const float4 p0 = ldg4(sortPos + j0), p1 = ldg4(sortPos + j1), p2 = ldg4(sortPos + j2), p3 = ldg4(sortPos + j3);
    one(p0); one(p1); one(p2); one(p3);
  }
  for (; k < nn; k++) one(ldg4(sortPos + __ldg(lp + (size_t)k * N)));
  const int gi = groupIndex[id];
  const int ori = globalIdx ? globalIdx[gi] : gi;
  if (force) {
    float4 f = force[ori];
    f.x += a.fx; f.y += a.fy; f.z += a.fz;
    force[ori] = f;
  }
  if (ENERGY) energy[ori] += a.e;
  if (VIRIAL) virial[ori] += a.v;
}

// All-pairs fallback of PairForces for boxes no larger than 3 cut-offs in every dimension (PairForces.cu:49-53 ->
// NBody::transverse, Interactor/NBodyBase.cuh:46-116): one thread per particle, the particles visited tile by tile
// through shared memory in ascending group order and accumulated sequentially like the reference kernel, per-pair
// minimum image like Radial::Transverser::compute (RadialPotential.cuh:107-127). Such boxes hold a few hundred
// particles at most, so this kernel is never on the hot path.
constexpr int kNBodyThreads = 128;
template <bool ENERGY, bool VIRIAL, bool MULTITYPE>
__global__ void __launch_bounds__(kNBodyThreads)
ljNBody(const float4 *__restrict__ pos, const int *__restrict__ globalIdx, int N, GridF g,
        const LJPar *__restrict__ parTable, int ntypes, float4 *__restrict__ force, float *__restrict__ energy,
        float *__restrict__ virial) {
  __shared__ float4 tile[kNBodyThreads];
  const int tid = blockIdx.x * kNBodyThreads + threadIdx.x;
  const bool active = tid < N;
  const int ori = active ? (globalIdx ? globalIdx[tid] : tid) : 0;
  const float4 pi = active ? ldg4(pos + ori) : make_float4(0.f, 0.f, 0.f, 0.f);
  LJPar par = parTable[0];
  uint32_t rcb = __float_as_uint(par.cutOff2) - 1u;
  const int ti = (int)pi.w;
  const int trow = MULTITYPE && (unsigned)ti < (unsigned)ntypes ? ti * ntypes : -1;
  Acc a = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < N; t0 += kNBodyThreads) {
    const int jl = t0 + threadIdx.x;
    if (jl < N) tile[threadIdx.x] = ldg4(pos + (globalIdx ? globalIdx[jl] : jl));
    __syncthreads();
    const int cnt = min(kNBodyThreads, N - t0);
    if (active) {
      for (int c = 0; c < cnt; c++) {
        const float4 pj = tile[c];
        const float dx = foldCoord(pj.x - pi.x, g.Lx, g.mx), dy = foldCoord(pj.y - pi.y, g.Ly, g.my),
                    dz = foldCoord(pj.z - pi.z, g.Lz, g.mz);
        if (MULTITYPE) {
          const int tj = (int)pj.w;
          par = parTable[((unsigned)tj < (unsigned)ntypes && trow >= 0) ? trow + tj : 0];
          rcb = __float_as_uint(par.cutOff2) - 1u;
        }
        ljPair<ENERGY, VIRIAL>(dx, dy, dz, par, rcb, a);
      }
    }
    __syncthreads();
  }
  if (!active) return;
  if (force) {
    float4 f = force[ori];
    f.x += a.fx; f.y += a.fy; f.z += a.fz;
    force[ori] = f;
  }
  if (ENERGY) energy[ori] += a.e;
  if (VIRIAL) virial[ori] += a.v;
}

template <bool E, bool V, bool M, bool P, bool A, bool PK = false>
static int launchLJ(ub200_celllist *cl, const LJPar *table, int ntypes, float4 *force, float *energy, float *virial,
                    const int *globalIdx, cudaStream_t st, int ownerLo, int ownerHi) {
  auto kern = ljCellTraversal<E, V, M, P, A, PK>;
  static int blocksPerSM = 0; // per instantiation
  if (!blocksPerSM) {
    UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kPairThreads, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
  }
  int grid = kNumSMs * blocksPerSM;
  const int needed = (cl->ncells + kPairWarps - 1) / kPairWarps;
  if (grid > needed) grid = needed;
  kern<<<grid, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(), cl->binStart.as<uint32_t>(),
                                      cl->grid, cl->ncells, table, ntypes, force, energy, virial, globalIdx, ownerLo,
                                      ownerHi);
  UB200_LAUNCHED();
  return UB200_OK;
}

// d_table: device table [ntypes*ntypes] of LJPar
int ljSumDev(ub200_celllist *cl, const LJPar *d_table, int ntypes, float4 *force, float *energy, float *virial,
             const int *globalIdx, bool accumulate, cudaStream_t st, int ownerLo = 0, int ownerHi = 0x7fffffff) {
  if (!cl || !d_table || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl->built) return UB200_ERR_NOT_BUILT;
  if (!force && !energy && !virial) return UB200_OK;
  const GridF &g = cl->grid;
  // a periodic dimension with fewer than 4 cells needs the per-pair minimum image
  const bool pairMic = (g.mx != 0.0f && g.nx < 4) || (g.my != 0.0f && g.ny < 4) || (g.mz != 0.0f && g.nz < 4);
  const bool E = energy != nullptr, V = virial != nullptr, M = ntypes > 1;
  // experimental packed-fp32 (FFMA2) body of the force-only, single-type traversal: UB200_LJ_PACKED=1 (same bits, same speed)
  const char *pk = getenv("UB200_LJ_PACKED");
  if (pk && pk[0] == '1' && !E && !V && !M && !pairMic) {
    if (accumulate) return launchLJ<false, false, false, false, true, true>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi);
    return launchLJ<false, false, false, false, false, true>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi);
  }
#define UB200_LJ_DISPATCH(e, v, m, p, a)                                                                     \
  if (E == e && V == v && M == m && pairMic == p && accumulate == a)                                         \
    return launchLJ<e, v, m, p, a>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi);
  UB200_LJ_DISPATCH(false, false, false, false, true)
  UB200_LJ_DISPATCH(false, false, false, false, false)
  UB200_LJ_DISPATCH(false, false, false, true, true)
  UB200_LJ_DISPATCH(false, false, false, true, false)
  UB200_LJ_DISPATCH(false, false, true, false, true)
  UB200_LJ_DISPATCH(false, false, true, false, false)
  UB200_LJ_DISPATCH(false, false, true, true, true)
  UB200_LJ_DISPATCH(false, false, true, true, false)
#undef UB200_LJ_DISPATCH
#define UB200_LJ_DISPATCH_EV(m, p)                                                                           \
  if (M == m && pairMic == p) {                                                                              \
    if (E && V) return launchLJ<true, true, m, p, true>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi); \
    if (E) return launchLJ<true, false, m, p, true>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi);     \
    return launchLJ<false, true, m, p, true>(cl, d_table, ntypes, force, energy, virial, globalIdx, st, ownerLo, ownerHi);      \
  }
  UB200_LJ_DISPATCH_EV(false, false)
  UB200_LJ_DISPATCH_EV(false, true)
  UB200_LJ_DISPATCH_EV(true, false)
  UB200_LJ_DISPATCH_EV(true, true)
#undef UB200_LJ_DISPATCH_EV
  return UB200_ERR_UNSUPPORTED;
}

// host parameter table: uploaded into `cache` only when it changed since the last call
static int uploadLJTable(LJTableCache *cache, const float *params, int ntypes, cudaStream_t st) {
  const size_t n = (size_t)ntypes * ntypes * 4;
  if (cache->host.size() != n || memcmp(cache->host.data(), params, n * sizeof(float)) != 0) {
    int rc = cache->dev.reserve(n * sizeof(float));
    if (rc) return rc;
    cache->host.assign(params, params + n);
    UB200_CUDA(cudaMemcpyAsync(cache->dev.p, cache->host.data(), n * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  return UB200_OK;
}
int ljSum(ub200_celllist *cl, const float *params, int ntypes, float4 *force, float *energy, float *virial,
          const int *globalIdx, bool accumulate, LJTableCache *cache, cudaStream_t st, int ownerLo, int ownerHi) {
  if (!params || ntypes < 1 || !cache) return UB200_ERR_INVALID_ARGUMENT;
  const int rc = uploadLJTable(cache, params, ntypes, st);
  if (rc) return rc;
  return ljSumDev(cl, cache->dev.as<LJPar>(), ntypes, force, energy, virial, globalIdx, accumulate, st, ownerLo, ownerHi);
}

} // namespace ub200

using namespace ub200;

// parameter table of ub200_lj_nbody_f32, the one entry point without a handle (boxes of a few hundred particles). Not
// thread safe; the list-based entry points keep their table in the list handle.
static LJTableCache g_ljTable;

extern "C" int ub200_lj_sum_f32(ub200_celllist *cl, const float *params, int ntypes, void *d_force, float *d_energy,
                                float *d_virial, const int *d_globalIdx, void *stream) {
  if (!cl) return UB200_ERR_INVALID_ARGUMENT;
  return ljSum(cl, params, ntypes, (float4 *)d_force, d_energy, d_virial, d_globalIdx, true, &cl->ljTable,
               (cudaStream_t)stream, 0, 0x7fffffff);
}

// LJ forces over a Verlet list handle: the row list when the last rebuild made one (lj_vlist.cu), else the
// reference-layout list
int ub200::ljVerletSum(ub200_verletlist *vl, const float *params, int ntypes, float4 *force, float *d_energy, float *d_virial,
                       const int *d_globalIdx, bool accumulate, cudaStream_t st) {
  if (!vl || !params || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  if (!vl->N) return UB200_ERR_NOT_BUILT;
  if (!force && !d_energy && !d_virial) return UB200_OK;
  LJTableCache *cache = &vl->cl->ljTable;
  if (const int rc = uploadLJTable(cache, params, ntypes, st)) return rc;
  if (vl->fast) return vlistSum(vl, cache->dev.as<LJPar>(), ntypes, force, d_energy, d_virial, d_globalIdx, accumulate, st);
  if (!accumulate && force) UB200_CUDA(cudaMemsetAsync(force, 0, sizeof(float4) * (size_t)vl->N, st));
  const int cd1[3] = {1, 1, 1};
  const GridF g = makeGridF(vl->L, vl->periodic, cd1);
  const int N = vl->N, nb = (N + 127) / 128;
  const bool E = d_energy != nullptr, V = d_virial != nullptr, M = ntypes > 1;
#define UB200_LJV(e, v, m)                                                                                              \
  if (E == e && V == v && M == m) {                                                                                     \
    ljVerletTraversal<e, v, m><<<nb, 128, 0, st>>>(vl->sortPos.as<float4>(), vl->cl->groupIndex.as<int>(),               \
                                                   vl->neighbourList.as<int>(), vl->numberNeighbours.as<int>(), N, g,     \
                                                   cache->dev.as<LJPar>(), ntypes, force, d_energy, d_virial, d_globalIdx); \
    UB200_LAUNCHED();                                                                                                   \
    return UB200_OK;                                                                                                    \
  }
  UB200_LJV(false, false, false) UB200_LJV(false, false, true) UB200_LJV(true, false, false) UB200_LJV(true, false, true)
  UB200_LJV(false, true, false) UB200_LJV(false, true, true) UB200_LJV(true, true, false) UB200_LJV(true, true, true)
#undef UB200_LJV
  return UB200_ERR_UNSUPPORTED;
}

extern "C" int ub200_lj_sum_verlet_f32(ub200_verletlist *vl, const float *params, int ntypes, void *d_force, float *d_energy,
                                       float *d_virial, const int *d_globalIdx, void *stream) {
  return ljVerletSum(vl, params, ntypes, (float4 *)d_force, d_energy, d_virial, d_globalIdx, true, (cudaStream_t)stream);
}

extern "C" int ub200_lj_sum_owned_f32(ub200_celllist *cl, const float *params, int ntypes, void *d_force, int ownerLo,
                                      int ownerHi, int accumulate, void *stream) {
  if (ownerLo < 0 || ownerHi < ownerLo) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl) return UB200_ERR_INVALID_ARGUMENT;
  return ljSum(cl, params, ntypes, (float4 *)d_force, nullptr, nullptr, nullptr, accumulate != 0, &cl->ljTable,
               (cudaStream_t)stream, ownerLo, ownerHi);
}

extern "C" int ub200_lj_sum_devparams_f32(ub200_celllist *cl, const void *d_params, int ntypes, void *d_force,
                                          float *d_energy, float *d_virial, const int *d_globalIdx, void *stream) {
  return ljSumDev(cl, (const LJPar *)d_params, ntypes, (float4 *)d_force, d_energy, d_virial, d_globalIdx, true,
                  (cudaStream_t)stream);
}

extern "C" int ub200_lj_nbody_f32(const void *d_pos, const int *d_globalIdx, int N, const float L[3], const int periodic[3],
                                  const float *params, int ntypes, void *d_force, float *d_energy, float *d_virial,
                                  void *stream) {
  if (!d_pos || !L || !periodic || !params || ntypes < 1 || N < 0) return UB200_ERR_INVALID_ARGUMENT;
  if (N == 0 || (!d_force && !d_energy && !d_virial)) return UB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LJTableCache *cache = &g_ljTable;
  if (const int rc = uploadLJTable(cache, params, ntypes, st)) return rc;
  const int cd1[3] = {1, 1, 1};
  const GridF g = makeGridF(L, periodic, cd1);
  const int nb = (N + kNBodyThreads - 1) / kNBodyThreads;
  const bool E = d_energy != nullptr, V = d_virial != nullptr, M = ntypes > 1;
#define UB200_LJN(e, v, m)                                                                                          \
  if (E == e && V == v && M == m) {                                                                                 \
    ljNBody<e, v, m><<<nb, kNBodyThreads, 0, st>>>((const float4 *)d_pos, d_globalIdx, N, g, cache->dev.as<LJPar>(), \
                                                   ntypes, (float4 *)d_force, d_energy, d_virial);                   \
    UB200_LAUNCHED();                                                                                               \
    return UB200_OK;                                                                                                \
  }
  UB200_LJN(false, false, false) UB200_LJN(false, false, true) UB200_LJN(true, false, false) UB200_LJN(true, false, true)
  UB200_LJN(false, true, false) UB200_LJN(false, true, true) UB200_LJN(true, true, false) UB200_LJN(true, true, true)
#undef UB200_LJN
  return UB200_ERR_UNSUPPORTED;
}
