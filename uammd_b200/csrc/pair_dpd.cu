// Dissipative particle dynamics pair forces over the cell list, sm_100a.
// Replaces transverseWithNeighbourContainer (Interactor/NeighbourList/common.cuh:10-34) driving
// DPD_impl::ForceTransverser (Interactor/Potential/DPD.cuh:92-159). The reference evaluates getInfo(j)
// = {vel[global j], global j} for every CANDIDATE through the unsorted global index (81 scattered 12-byte
// loads per particle at rho=3); here velocities and ids are staged once per neighbour cell next to the
// positions and read conflict-free from shared memory.
//
// Saru PRNG (Afshar et al., Comput. Phys. Commun. 184 (2013) 1119; third_party/saruprng.cuh:257-280 seeding,
// :196-213,:339-351 stepping/output, :115-128 Box-Muller) is restated below; its constants are the algorithm.
#include "pair_common.cuh"
#include "saru.cuh"

namespace ub200 {

struct DPDPar {
  float A, gamma, sigmaSqrtGamma, invrcut;
  uint32_t seed, step;
  int idStride;
};

__device__ __forceinline__ void dpdPair(float rx, float ry, float rz, float vx, float vy, float vz, int idi, int idj,
                                        const DPDPar &p, float &fx, float &fy, float &fz) {
  // rij = ri - rj, vij = vi - vj (DPD.cuh:123-124)
  const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, rx * rx));
  const float rmod = __fsqrt_rn(r2);
  if (rmod == 0.0f) return;
  const float invrmod = __frcp_rn(rmod);
  if (invrmod <= p.invrcut) return;
  int i = idi, j = idj;
  if (i > j) { const int t = i; i = j; j = t; }
  const uint32_t ij = (uint32_t)i + (uint32_t)p.idStride * (uint32_t)j; // int32 wrap of i + N*j (DPD.cuh:128)
  Saru rng(ij, p.seed, p.step);
  const float wr = __fmaf_rn(-rmod, p.invrcut, 1.0f);
  const float Fc = p.A * wr * invrmod;
  const float wd = wr * wr;
  const float rv = __fmaf_rn(rz, vz, __fmaf_rn(ry, vy, rx * vx));
  const float Fd = -p.gamma * wd * invrmod * invrmod * rv;
  const float Fr = rng.gaussX(p.sigmaSqrtGamma * wr * invrmod);
  const float ft = Fc + Fd + Fr;
  fx = __fmaf_rn(ft, rx, fx);
  fy = __fmaf_rn(ft, ry, fy);
  fz = __fmaf_rn(ft, rz, fz);
}

constexpr int kDpdCap = 768; // staged candidates (pos + vel/id): 24 KB

template <bool PAIRMIC>
__global__ void __launch_bounds__(kPairThreads)
dpdCellTraversal(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex,
                 const uint32_t *__restrict__ binStart, GridF g, int ncells, const float *__restrict__ vel, DPDPar par,
                 float4 *__restrict__ force, const int *__restrict__ globalIdx, int ownerLo, int ownerHi, int accumulate,
                 const int *__restrict__ noiseId) {
  // ownerLo/ownerHi: only home particles whose (global) index lies in [ownerLo, ownerHi) are computed and written
  // (multi-GPU particle decomposition, like ub200_lj_sum_owned_f32)
  // noiseId: optional id of every particle used ONLY in the Saru key of a pair (brick decomposition: the local
  // arrays hold owned + ghost particles, the noise stays keyed on the global particle ids)
  __shared__ float4 cand[kDpdCap];
  __shared__ float4 candVel[kDpdCap]; // vx, vy, vz, id (bits)
  __shared__ unsigned short queue[kPairWarps][64]; // per warp: staged indices of the candidates that passed the distance test
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    if (hCount == 0) continue;
    const int hOff = __shfl_sync(0xffffffffu, nc.off, nc.centre);
    const bool staged = nc.total <= kDpdCap;
    const float3 hc = cellCentre(g, cx, cy, cz);
    if (staged) {
      for (int c = warp; c < 27; c += kPairWarps) {
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        if (cnt == 0) continue;
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        const int off = __shfl_sync(0xffffffffu, nc.off, c);
        for (int t = lane; t < cnt; t += 32) {
          float4 p = ldg4(sortPos + st + t);
          if (!PAIRMIC) toHomeImage(p, g, hc);
          const int gi = groupIndex[st + t];
          const int id = globalIdx ? globalIdx[gi] : gi;
          cand[off + t] = p;
          candVel[off + t] = make_float4(__ldg(vel + 3 * (size_t)id), __ldg(vel + 3 * (size_t)id + 1),
                                         __ldg(vel + 3 * (size_t)id + 2), __int_as_float(noiseId ? __ldg(noiseId + id) : id));
        }
      }
    }
    __syncthreads();
    for (int h = warp; h < hCount; h += kPairWarps) {
      float4 pi, vi;
      const int gih = groupIndex[hStart + h];
      const int idi = globalIdx ? globalIdx[gih] : gih; // array index of the home particle (velocity, force, ownership)
      if (idi < ownerLo || idi >= ownerHi) continue; // warp uniform
      if (staged) {
        pi = cand[hOff + h];
        vi = candVel[hOff + h];
      } else {
        pi = ldg4(sortPos + hStart + h);
        if (!PAIRMIC) toHomeImage(pi, g, hc);
        vi = make_float4(vel[3 * (size_t)idi], vel[3 * (size_t)idi + 1], vel[3 * (size_t)idi + 2],
                         __int_as_float(noiseId ? noiseId[idi] : idi));
      }
      const int nidi = __float_as_int(vi.w); // id of the home particle in the Saru key
      float fx = 0.f, fy = 0.f, fz = 0.f;
      if (staged) {
        // Only ~15 % of the candidates are within the cut-off and an in-range pair costs ~10x a rejected one (Saru
        // seeding + Box-Muller), so the warp first filters 32 candidates with the cheap distance test - the very
        // test dpdPair applies - and queues the survivors; the expensive body then runs on full warps.
        unsigned short *q = queue[warp];
        int qn = 0; // warp uniform
        auto body = [&](int t) {
          const float4 pj = cand[t];
          const float4 vj = candVel[t];
          float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
          if (PAIRMIC) {
            rx = foldCoord(rx, g.Lx, g.mx);
            ry = foldCoord(ry, g.Ly, g.my);
            rz = foldCoord(rz, g.Lz, g.mz);
          }
          dpdPair(rx, ry, rz, vi.x - vj.x, vi.y - vj.y, vi.z - vj.z, nidi, __float_as_int(vj.w), par, fx, fy, fz);
        };
        for (int t0 = 0; t0 < nc.total; t0 += 32) {
          const int t = t0 + lane;
          bool in = false;
          if (t < nc.total) {
            const float4 pj = cand[t];
            float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
            if (PAIRMIC) {
              rx = foldCoord(rx, g.Lx, g.mx);
              ry = foldCoord(ry, g.Ly, g.my);
              rz = foldCoord(rz, g.Lz, g.mz);
            }
            const float rmod = __fsqrt_rn(__fmaf_rn(rz, rz, __fmaf_rn(ry, ry, rx * rx)));
            in = rmod != 0.0f && __frcp_rn(rmod) > par.invrcut;
          }
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) q[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)t;
          qn += __popc(m);
          __syncwarp();
          if (qn >= 32) {
            body(q[lane]);
            const int rem = qn - 32;
            const unsigned short mv = lane < rem ? q[32 + lane] : (unsigned short)0;
            __syncwarp();
            if (lane < rem) q[lane] = mv;
            __syncwarp();
            qn = rem;
          }
        }
        if (lane < qn) body(q[lane]);
        __syncwarp();
      } else {
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
          if (cnt == 0) continue;
          const int st = __shfl_sync(0xffffffffu, nc.start, c);
          for (int t = lane; t < cnt; t += 32) {
            float4 pj = ldg4(sortPos + st + t);
            const int gj = groupIndex[st + t];
            const int idj = globalIdx ? globalIdx[gj] : gj;
            float rx, ry, rz;
            if (PAIRMIC) {
              rx = foldCoord(pi.x - pj.x, g.Lx, g.mx);
              ry = foldCoord(pi.y - pj.y, g.Ly, g.my);
              rz = foldCoord(pi.z - pj.z, g.Lz, g.mz);
            } else {
              toHomeImage(pj, g, hc);
              rx = pi.x - pj.x;
              ry = pi.y - pj.y;
              rz = pi.z - pj.z;
            }
            dpdPair(rx, ry, rz, vi.x - vel[3 * (size_t)idj], vi.y - vel[3 * (size_t)idj + 1],
                    vi.z - vel[3 * (size_t)idj + 2], nidi, noiseId ? noiseId[idj] : idj, par, fx, fy, fz);
          }
        }
      }
      fx = warpSum(fx);
      fy = warpSum(fy);
      fz = warpSum(fz);
      if (lane == 0) {
        // ForceTransverser::set: force[pi] += make_real4(total) (make_real4(real3) zero-fills w)
        float4 f = accumulate ? force[idi] : make_float4(0.f, 0.f, 0.f, 0.f);
        f.x += fx; f.y += fy; f.z += fz;
        force[idi] = f;
      }
    }
    __syncthreads();
  }
}

} // namespace ub200

using namespace ub200;

static int dpdSum(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut, uint32_t seed,
                  uint32_t step, int idStride, void *d_force, const int *d_globalIdx, int ownerLo, int ownerHi, int accumulate,
                  void *stream, const int *d_noiseId = nullptr) {
  if (!cl || !d_vel || !d_force || !(rcut > 0)) return UB200_ERR_INVALID_ARGUMENT;
  if (!cl->built) return UB200_ERR_NOT_BUILT;
  cudaStream_t st = (cudaStream_t)stream;
  const GridF &g = cl->grid;
  const bool pairMic = (g.mx != 0.0f && g.nx < 4) || (g.my != 0.0f && g.ny < 4) || (g.mz != 0.0f && g.nz < 4);
  DPDPar par;
  par.A = A;
  par.gamma = gamma;
  par.sigmaSqrtGamma = sigma * sqrtf(gamma); // sigma*sqrt(g), evaluated per pair in the reference (DPD.cuh:148)
  par.invrcut = (float)(1.0 / (double)rcut); // invrcut(1.0/rcut) (DPD.cuh:110)
  par.seed = seed;
  par.step = step;
  par.idStride = idStride;
  static int bps[2] = {0, 0};
  if (!bps[pairMic]) {
    if (pairMic) UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[1], dpdCellTraversal<true>, kPairThreads, 0));
    else UB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps[0], dpdCellTraversal<false>, kPairThreads, 0));
    if (bps[pairMic] < 1) bps[pairMic] = 1;
  }
  int grid = kNumSMs * bps[pairMic];
  if (grid > cl->ncells) grid = cl->ncells;
  if (pairMic)
    dpdCellTraversal<true><<<grid, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(),
                                                          cl->binStart.as<uint32_t>(), g, cl->ncells,
                                                          (const float *)d_vel, par, (float4 *)d_force, d_globalIdx, ownerLo, ownerHi, accumulate,
                                                          d_noiseId);
  else
    dpdCellTraversal<false><<<grid, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(),
                                                           cl->binStart.as<uint32_t>(), g, cl->ncells,
                                                           (const float *)d_vel, par, (float4 *)d_force, d_globalIdx, ownerLo, ownerHi, accumulate,
                                                          d_noiseId);
  UB200_LAUNCHED();
  return UB200_OK;
}

extern "C" int ub200_dpd_sum_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                                 uint32_t seed, uint32_t step, int idStride, void *d_force, const int *d_globalIdx,
                                 void *stream) {
  return dpdSum(cl, d_vel, A, gamma, sigma, rcut, seed, step, idStride, d_force, d_globalIdx, 0, 0x7fffffff, 1, stream);
}
extern "C" int ub200_dpd_sum_owned_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                                       uint32_t seed, uint32_t step, int idStride, void *d_force, int ownerLo, int ownerHi,
                                       int accumulate, void *stream) {
  if (ownerLo < 0 || ownerHi < ownerLo) return UB200_ERR_INVALID_ARGUMENT;
  return dpdSum(cl, d_vel, A, gamma, sigma, rcut, seed, step, idStride, d_force, nullptr, ownerLo, ownerHi, accumulate, stream);
}
extern "C" int ub200_dpd_sum_owned_ids_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                                           uint32_t seed, uint32_t step, int idStride, void *d_force, int ownerLo, int ownerHi,
                                           int accumulate, const int *d_noiseId, void *stream) {
  if (ownerLo < 0 || ownerHi < ownerLo || !d_noiseId) return UB200_ERR_INVALID_ARGUMENT;
  return dpdSum(cl, d_vel, A, gamma, sigma, rcut, seed, step, idStride, d_force, nullptr, ownerLo, ownerHi, accumulate, stream,
                d_noiseId);
}
