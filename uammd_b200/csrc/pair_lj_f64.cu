// PairForces<Potential::LJ, CellList>::sum for `real = double` builds of UAMMD (-DDOUBLE_PRECISION), sm_100a.
//
// The reference is one code base templated on `real` (global/defines.h): with DOUBLE_PRECISION its cell list, its
// Radial<LJFunctor>::Transverser (Potential/RadialPotential.cuh:107-127) and LJFunctor::force / energy
// (Potential/Potential.cuh:37-56) all run in double. BASELINE config 1 is single precision and is served by the column
// engine (lj_column.cu); this file is the double precision sibling of the CELL traversal, built like the PSE near field
// (pse.cu): the neighbour search runs on single precision copies of the positions with a cut-off padded by the rounding
// of that copy (every pair the exact test can accept lies in adjacent cells), the pair arithmetic - minimum image
// (Box::apply_pbc, utils/Box.cuh:50-57), cut-off test, force, energy, virial - is the reference's, in double, on the
// double positions. One warp per home cell, the lanes stride over the particles of the (up to) 27 neighbour cells.
// Sums differ from the reference's in order only.
#include "pair_common.cuh"
#include <algorithm>
#include <cmath>
#include <new>

namespace ub200 {

struct LJPar64 { double cutOff2, sigma2, epsDivSigma2, shift; }; // LJFunctor::PairParameters (Potential.cuh:31-35)

struct Box64 {
  double L[3], mInvL[3]; // boxSize, minusInvBoxSize (0: not periodic)
};

__global__ void __launch_bounds__(256) lj64ToFloat4(const double4 *__restrict__ in, float4 *__restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double4 p = in[i];
  out[i] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
}
__global__ void __launch_bounds__(256)
lj64GatherSorted(const int *__restrict__ groupIndex, const double4 *__restrict__ pos, double4 *__restrict__ sorted, int N) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) sorted[k] = pos[groupIndex[k]];
}

// Box::apply_pbc on one component
__device__ __forceinline__ double pbc64(double r, double L, double mInvL) {
  return mInvL != 0.0 ? r + floor(r * mInvL + 0.5) * L : r;
}

template <bool ENERGY, bool VIRIAL>
__global__ void __launch_bounds__(kPairThreads)
ljCellTraversal64(const double4 *__restrict__ sortedPos, const int *__restrict__ groupIndex, const uint32_t *__restrict__ binStart,
                  GridF g, int ncells, Box64 box, const LJPar64 *__restrict__ params, int ntypes, double4 *__restrict__ force,
                  double *__restrict__ energy, double *__restrict__ virial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warpsTotal = gridDim.x * kPairWarps;
  for (int cell = blockIdx.x * kPairWarps + warp; cell < ncells; cell += warpsTotal) {
    const int cx = cell % g.nx, cy = (cell / g.nx) % g.ny, cz = cell / (g.nx * g.ny);
    const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
    const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
    const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
    for (int h = 0; h < hCount; h++) {
      const double4 pi = sortedPos[hStart + h];
      const int ti = (int)pi.w;
      double fx = 0, fy = 0, fz = 0, e = 0, v = 0;
      for (int c = 0; c < 27; c++) {
        const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
        if (cnt == 0) continue;
        const int st = __shfl_sync(0xffffffffu, nc.start, c);
        for (int t = lane; t < cnt; t += 32) {
          const double4 pj = sortedPos[st + t];
          const double rx = pbc64(pj.x - pi.x, box.L[0], box.mInvL[0]);
          const double ry = pbc64(pj.y - pi.y, box.L[1], box.mInvL[1]);
          const double rz = pbc64(pj.z - pi.z, box.L[2], box.mInvL[2]);
          const double r2 = rx * rx + ry * ry + rz * rz;
          if (r2 == 0.0) continue; // Transverser::compute returns {} (RadialPotential.cuh:112-114)
          const LJPar64 p = params[ti * ntypes + (int)pj.w];
          if (r2 >= p.cutOff2) continue;
          const double invr2 = p.sigma2 / r2;
          const double invr6 = invr2 * invr2 * invr2;
          const double fm = p.epsDivSigma2 * (-48.0 * invr6 + 24.0) * invr6 * invr2;
          fx += fm * rx; fy += fm * ry; fz += fm * rz;
          if (ENERGY) e += 0.5 * (p.epsDivSigma2 * p.sigma2 * 4.0 * invr6 * (invr6 - 1.0) - p.shift);
          if (VIRIAL) v += fm * r2; // dot(F, r12)
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
        if (ENERGY) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (VIRIAL) v += __shfl_xor_sync(0xffffffffu, v, o);
      }
      if (lane == 0) { // Transverser::set: += (RadialPotential.cuh:119-126)
        const int i = groupIndex[hStart + h];
        if (force) { double4 f = force[i]; f.x += fx; f.y += fy; f.z += fz; force[i] = f; }
        if (ENERGY) energy[i] += e;
        if (VIRIAL) virial[i] += v;
      }
    }
  }
}

} // namespace ub200

using namespace ub200;

struct ub200_lj64 {
  ub200_celllist *cl = nullptr;
  DevBuf posF, sorted, params;
};

extern "C" {

int ub200_lj64_create(ub200_lj64 **out) {
  if (!out) return UB200_ERR_INVALID_ARGUMENT;
  ub200_lj64 *h = new (std::nothrow) ub200_lj64();
  if (!h) return UB200_ERR_ALLOC;
  const int rc = ub200_celllist_create(&h->cl);
  if (rc) { delete h; return rc; }
  *out = h;
  return UB200_OK;
}

int ub200_lj64_destroy(ub200_lj64 *h) {
  if (!h) return UB200_OK;
  ub200_celllist_destroy(h->cl);
  h->posF.release(); h->sorted.release(); h->params.release();
  delete h;
  return UB200_OK;
}

int ub200_lj_sum_f64(ub200_lj64 *h, const void *d_pos, int N, const double L[3], const int periodic[3], double cutOff,
                     const double *params, int ntypes, void *d_force, double *d_energy, double *d_virial, void *stream) {
  if (!h || !d_pos || N <= 0 || !L || !periodic || !(cutOff > 0) || !params || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  if (!d_force && !d_energy && !d_virial) return UB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // neighbour search on a single precision copy: cells of at least the cut-off plus the rounding of the copy
  const float Lf[3] = {(float)L[0], (float)L[1], (float)L[2]};
  const float Lmax = std::max({Lf[0], Lf[1], Lf[2]});
  const float rcList = (float)cutOff * (1.0f + 1e-5f) + 16.0f * Lmax * 1.2e-7f;
  int cd[3];
  if ((rc = ub200_neighbour_celldim_f32(Lf, rcList, cd))) return rc;
  if ((rc = h->posF.reserve(sizeof(float4) * (size_t)N)) || (rc = h->sorted.reserve(sizeof(double4) * (size_t)N)) ||
      (rc = h->params.reserve(sizeof(LJPar64) * (size_t)ntypes * ntypes)))
    return rc;
  const int nb = (N + 255) / 256;
  lj64ToFloat4<<<nb, 256, 0, st>>>((const double4 *)d_pos, h->posF.as<float4>(), N);
  UB200_LAUNCHED();
  if ((rc = ub200_celllist_build_f32(h->cl, h->posF.p, nullptr, N, Lf, periodic, cd, stream))) return rc;
  lj64GatherSorted<<<nb, 256, 0, st>>>(h->cl->groupIndex.as<int>(), (const double4 *)d_pos, h->sorted.as<double4>(), N);
  UB200_LAUNCHED();
  UB200_CUDA(cudaMemcpyAsync(h->params.p, params, sizeof(LJPar64) * (size_t)ntypes * ntypes, cudaMemcpyHostToDevice, st));
  Box64 box;
  for (int d = 0; d < 3; d++) { box.L[d] = L[d]; box.mInvL[d] = periodic[d] ? -1.0 / L[d] : 0.0; }
  const int needed = (h->cl->ncells + kPairWarps - 1) / kPairWarps;
  const int gridSize = std::max(1, std::min(needed, kNumSMs * 8));
#define UB200_LJ64(E, V)                                                                                                  \
  ljCellTraversal64<E, V><<<gridSize, kPairThreads, 0, st>>>(h->sorted.as<double4>(), h->cl->groupIndex.as<int>(),          \
                                                            h->cl->binStart.as<uint32_t>(), h->cl->grid, h->cl->ncells, box, \
                                                            h->params.as<LJPar64>(), ntypes, (double4 *)d_force, d_energy, d_virial)
  if (d_energy && d_virial) UB200_LJ64(true, true);
  else if (d_energy) UB200_LJ64(true, false);
  else if (d_virial) UB200_LJ64(false, true);
  else UB200_LJ64(false, false);
#undef UB200_LJ64
  UB200_LAUNCHED();
  return UB200_OK;
}
}
