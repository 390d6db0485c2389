// Short-range pair traversal specialised for the Lennard-Jones transverser, sm_100a.
//
// Replaces NeighbourList_ns::transverseWithNeighbourContainer (Interactor/NeighbourList/common.cuh:10-34)
// + CellList_ns::NeighbourIterator (CellList/NeighbourContainer.cuh:95-138) + Radial<LJFunctor>::Transverser
// (Potential/RadialPotential.cuh:107-127, Potential/Potential.cuh:37-65).
//
// Design (B200): persistent CTAs walk the home cells. For each home cell the particles of its (up to) 27
// neighbour cells are staged ONCE into shared memory, already folded into the primary box and displaced by
// the periodic image shift of their cell, so the inner loop needs no per-pair minimum-image arithmetic.
// Each warp then owns one home particle at a time: its 32 lanes stride over the staged candidates
// (conflict-free LDS.128), accumulate privately and finish with a shuffle reduction. The reference instead
// runs one thread per particle through a divergent 27-cell iterator with ~340 dependent global loads.
#include "pair_common.cuh"

namespace ub200 {

struct LJPar {
  float cutOff2, sigma2, epsDivSigma2, shift;
};

struct Acc {
  float fx, fy, fz, e, v;
};

template <bool ENERGY, bool VIRIAL>
__device__ __forceinline__ void ljPair(float dx, float dy, float dz, const LJPar &p, Acc &a) {
  const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
  if (r2 < p.cutOff2 && r2 != 0.0f) {
    const float invr2 = __fdividef(p.sigma2, r2);
    const float invr6 = invr2 * invr2 * invr2;
    const float fm = p.epsDivSigma2 * __fmaf_rn(-48.0f, invr6, 24.0f) * invr6 * invr2;
    a.fx = __fmaf_rn(fm, dx, a.fx);
    a.fy = __fmaf_rn(fm, dy, a.fy);
    a.fz = __fmaf_rn(fm, dz, a.fz);
    if (ENERGY) a.e += 0.5f * (p.epsDivSigma2 * p.sigma2 * 4.0f * invr6 * (invr6 - 1.0f) - p.shift);
    if (VIRIAL) a.v = __fmaf_rn(fm, r2, a.v);
  }
}

// PAIRMIC: per-pair minimum image exactly like Radial::Transverser::compute (box.apply_pbc(rj-ri)); needed
// when a periodic dimension has fewer than 4 cells (a collapsed dimension still wraps). Otherwise the cell
// image shift staged with the candidates is the minimum image.
template <bool ENERGY, bool VIRIAL, bool MULTITYPE, bool PAIRMIC, bool ACCUMULATE>
__global__ void __launch_bounds__(kPairThreads)
ljCellTraversal(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex,
                const uint32_t *__restrict__ binStart, GridF g, int ncells, LJPar par0,
                const LJPar *__restrict__ parTable, int ntypes, float4 *__restrict__ force,
                float *__restrict__ energy, float *__restrict__ virial, const int *__restrict__ globalIdx) {
  __shared__ float4 cand[kCandCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue; // CTA uniform
    const int hOff = __shfl_sync(0xffffffffu, nc.off, nc.centre);
    const bool staged = nc.total <= kCandCap;
    if (staged) {
      for (int c = warp; c < 27; c += kPairWarps) {
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        if (cnt == 0) continue;
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        const int off = __shfl_sync(0xffffffffu, nc.off, c);
        const float sx = __shfl_sync(0xffffffffu, nc.sx, c);
        const float sy = __shfl_sync(0xffffffffu, nc.sy, c);
        const float sz = __shfl_sync(0xffffffffu, nc.sz, c);
        for (int t = lane; t < cnt; t += 32) {
          float4 p = ldg4(sortPos + st + t);
          if (!PAIRMIC) {
            p.x = foldCoord(p.x, g.Lx, g.mx) + sx;
            p.y = foldCoord(p.y, g.Ly, g.my) + sy;
            p.z = foldCoord(p.z, g.Lz, g.mz) + sz;
          }
          cand[off + t] = p;
        }
      }
    }
    __syncthreads();
    for (int h = warp; h < hCount; h += kPairWarps) {
      float4 pi;
      if (staged) pi = cand[hOff + h];
      else {
        pi = ldg4(sortPos + hStart + h);
        if (!PAIRMIC) {
          pi.x = foldCoord(pi.x, g.Lx, g.mx);
          pi.y = foldCoord(pi.y, g.Ly, g.my);
          pi.z = foldCoord(pi.z, g.Lz, g.mz);
        }
      }
      Acc a = {0.f, 0.f, 0.f, 0.f, 0.f};
      LJPar p = par0;
      const int ti = MULTITYPE ? (int)pi.w * ntypes : 0;
      if (staged) {
#pragma unroll 2
        for (int t = lane; t < nc.total; t += 32) {
          const float4 pj = cand[t];
          float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
          if (PAIRMIC) {
            dx = foldCoord(dx, g.Lx, g.mx);
            dy = foldCoord(dy, g.Ly, g.my);
            dz = foldCoord(dz, g.Lz, g.mz);
          }
          if (MULTITYPE) p = parTable[ti + (int)pj.w];
          ljPair<ENERGY, VIRIAL>(dx, dy, dz, p, a);
        }
      } else {
        // dense neighbourhood: walk the neighbour cells straight from global memory
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          if (cnt == 0) continue;
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          const float sx = __shfl_sync(0xffffffffu, nc.sx, c);
          const float sy = __shfl_sync(0xffffffffu, nc.sy, c);
          const float sz = __shfl_sync(0xffffffffu, nc.sz, c);
          for (int t = lane; t < cnt; t += 32) {
            const float4 pj = ldg4(sortPos + st + t);
            float dx, dy, dz;
            if (PAIRMIC) {
              dx = foldCoord(pj.x - pi.x, g.Lx, g.mx);
              dy = foldCoord(pj.y - pi.y, g.Ly, g.my);
              dz = foldCoord(pj.z - pi.z, g.Lz, g.mz);
            } else {
              dx = (foldCoord(pj.x, g.Lx, g.mx) + sx) - pi.x;
              dy = (foldCoord(pj.y, g.Ly, g.my) + sy) - pi.y;
              dz = (foldCoord(pj.z, g.Lz, g.mz) + sz) - pi.z;
            }
            if (MULTITYPE) p = parTable[ti + (int)pj.w];
            ljPair<ENERGY, VIRIAL>(dx, dy, dz, p, a);
          }
        }
      }
      a.fx = warpSum(a.fx);
      a.fy = warpSum(a.fy);
      a.fz = warpSum(a.fz);
      if (ENERGY) a.e = warpSum(a.e);
      if (VIRIAL) a.v = warpSum(a.v);
      if (lane == 0) {
        const int gi = groupIndex[hStart + h];
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
    }
    __syncthreads();
  }
}

template <bool E, bool V, bool M, bool P, bool A>
static int launchLJ(ub200_celllist *cl, LJPar par0, const LJPar *table, int ntypes, float4 *force, float *energy,
                    float *virial, const int *globalIdx, cudaStream_t st) {
  auto kern = ljCellTraversal<E, V, M, P, A>;
  static int blocksPerSM = 0; // per instantiation
  if (!blocksPerSM) {
    UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, kern, kPairThreads, 0));
    if (blocksPerSM < 1) blocksPerSM = 1;
  }
  int grid = kNumSMs * blocksPerSM;
  if (grid > cl->ncells) grid = cl->ncells;
  kern<<<grid, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(), cl->binStart.as<uint32_t>(),
                                      cl->grid, cl->ncells, par0, table, ntypes, force, energy, virial, globalIdx);
  UB200_LAUNCHED();
  return UB200_OK;
}

// parameter table cache (device) for multi-type systems
struct ParTableCache {
  DevBuf buf;
  int ntypes = 0;
};

int ljSum(ub200_celllist *cl, const float *params, int ntypes, float4 *force, float *energy, float *virial,
          const int *globalIdx, bool accumulate, DevBuf *tableBuf, cudaStream_t st) {
  if (!cl || !params || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl->built) return UB200_ERR_NOT_BUILT;
  if (!force && !energy && !virial) return UB200_OK;
  const GridF &g = cl->grid;
  // a periodic dimension with fewer than 4 cells needs the per-pair minimum image
  const bool pairMic = (g.mx != 0.0f && g.nx < 4) || (g.my != 0.0f && g.ny < 4) || (g.mz != 0.0f && g.nz < 4);
  LJPar par0 = {params[0], params[1], params[2], params[3]};
  const LJPar *table = nullptr;
  if (ntypes > 1) {
    if (!tableBuf) return UB200_ERR_INVALID_ARGUMENT;
    int rc = tableBuf->reserve(sizeof(LJPar) * (size_t)ntypes * ntypes);
    if (rc) return rc;
    UB200_CUDA(cudaMemcpyAsync(tableBuf->p, params, sizeof(LJPar) * (size_t)ntypes * ntypes, cudaMemcpyHostToDevice, st));
    table = tableBuf->as<LJPar>();
  }
  const bool E = energy != nullptr, V = virial != nullptr, M = ntypes > 1;
#define UB200_LJ_DISPATCH(e, v, m, p, a)                                                                     \
  if (E == e && V == v && M == m && pairMic == p && accumulate == a)                                         \
    return launchLJ<e, v, m, p, a>(cl, par0, table, ntypes, force, energy, virial, globalIdx, st);
  // forces only, single type: the hot configurations
  UB200_LJ_DISPATCH(false, false, false, false, true)
  UB200_LJ_DISPATCH(false, false, false, false, false)
  UB200_LJ_DISPATCH(false, false, false, true, true)
  UB200_LJ_DISPATCH(false, false, false, true, false)
  UB200_LJ_DISPATCH(false, false, true, false, true)
  UB200_LJ_DISPATCH(false, false, true, false, false)
  UB200_LJ_DISPATCH(false, false, true, true, true)
  UB200_LJ_DISPATCH(false, false, true, true, false)
#undef UB200_LJ_DISPATCH
  // anything asking for energy and/or virial: one generic instantiation per (M, P) computing both when needed
  // (unrequested outputs are routed to a null pointer check inside the kernel via template flags)
#define UB200_LJ_DISPATCH_EV(m, p)                                                                           \
  if (M == m && pairMic == p) {                                                                              \
    if (E && V) return launchLJ<true, true, m, p, true>(cl, par0, table, ntypes, force, energy, virial, globalIdx, st); \
    if (E) return launchLJ<true, false, m, p, true>(cl, par0, table, ntypes, force, energy, virial, globalIdx, st);     \
    return launchLJ<false, true, m, p, true>(cl, par0, table, ntypes, force, energy, virial, globalIdx, st);  \
  }
  UB200_LJ_DISPATCH_EV(false, false)
  UB200_LJ_DISPATCH_EV(false, true)
  UB200_LJ_DISPATCH_EV(true, false)
  UB200_LJ_DISPATCH_EV(true, true)
#undef UB200_LJ_DISPATCH_EV
  return UB200_ERR_UNSUPPORTED;
}

} // namespace ub200

using namespace ub200;

static DevBuf g_ljTable; // shared parameter table for the stateless ub200_lj_sum_f32 entry point

extern "C" int ub200_lj_sum_f32(ub200_celllist *cl, const float *params, int ntypes, void *d_force, float *d_energy,
                                float *d_virial, const int *d_globalIdx, void *stream) {
  return ljSum(cl, params, ntypes, (float4 *)d_force, d_energy, d_virial, d_globalIdx, true, &g_ljTable,
               (cudaStream_t)stream);
}
