// Host-side state of the immersed-boundary operators (scratch buffers + launch logic), shared by the FCM and PSE
// pipelines and by the stand-alone ub200_ibm_* entry points.
#pragma once
#include <cstdlib>
#include "ibm.cuh"

namespace ub200 {

template <class T> struct Real4;
template <> struct Real4<float> { using type = float4; };
template <> struct Real4<double> { using type = double4; };

template <class T> struct IbmState {
  GridT<T> grid;
  IbmKernel<T> kern;
  int nxPad = 0;
  bool nodeCentric = false;
  DevBuf binCount, binStart, tileSums, codeSlot, unstable, sortedIndex, sortedRec;
  int sortedValidFor = -1;
  bool sortedGather = false; // thread-per-particle gather over the sorted records (supports 3, 4)
  int recWords() const {
    switch (kern.support) {
    case 3: return RecGeom<T, 3>::REC;
    case 4: return RecGeom<T, 4>::REC;
    case 5: return RecGeom<T, 5>::REC;
    default: return RecGeom<T, 7>::REC;
    }
  }

  int init(const double L[3], const int periodic[3], const int cells[3], const ub200_ibm_kernel &k, int nxPad_) {
    grid = makeGridT<T>(L, periodic, cells);
    kern.kind = k.kind;
    kern.support = k.support;
    kern.invh = (T)(1.0 / k.h);
    kern.prefactor = (T)k.prefactor;
    kern.tau = (T)k.tau;
    kern.rmax = (T)k.rmax;
    nxPad = nxPad_;
    if (k.support < 1 || k.support > kMaxSupport) return UB200_ERR_INVALID_ARGUMENT;
    const long long ncells = (long long)grid.n[0] * grid.n[1] * grid.n[2];
    nodeCentric = (k.support == 3 || k.support == 4 || k.support == 5 || k.support == 7) && ncells <= 4096LL * 4096LL;
    for (int d = 0; d < 3; d++)
      if (grid.m[d] != T(0) && grid.n[d] < k.support + 1) nodeCentric = false; // support would overlap itself
    // the row-brick spread stages ONE periodic image of a particle per brick: a periodic dimension must be longer than
    // a brick plus a support (smaller grids take the generic path); origins are packed in 16 bits
    if (grid.n[0] < kRbX + k.support || grid.n[0] > 65000 || grid.n[1] > 65000 || grid.n[2] > 65000) nodeCentric = false;
    if ((grid.m[1] != T(0) && grid.n[1] < kRbY + k.support) || (grid.m[2] != T(0) && grid.n[2] < kRbZ + k.support)) nodeCentric = false;
    // thread-per-particle interpolation over the cell-sorted records; UB200_IBM_GATHER=warp keeps the warp-per-particle kernel
    // for the wide supports (5, 7) - A/B switch
    const char *gsel = getenv("UB200_IBM_GATHER");
    sortedGather = nodeCentric && (k.support <= 4 || !(gsel && gsel[0] == 'w'));
    if (grid.n[2] == 1) nodeCentric = false; // 2-D grids take the generic path
    return UB200_OK;
  }
  void release() {
    DevBuf *b[] = {&binCount, &binStart, &tileSums, &codeSlot, &unstable, &sortedIndex, &sortedRec};
    for (auto *x : b) x->release();
  }

  // bin, order and build the stencil records (positions + optional values)
  int prepare(const void *pos, const void *val, int valStride, int N, cudaStream_t st) {
    using T4 = typename Real4<T>::type;
    const int ncells = grid.n[0] * grid.n[1] * grid.n[2];
    int rc;
    if (!binCount.p || binCount.cap < sizeof(uint32_t) * (size_t)ncells) {
      if ((rc = binCount.reserve(sizeof(uint32_t) * (size_t)ncells))) return rc;
      if ((rc = binStart.reserve(sizeof(uint32_t) * ((size_t)ncells + 1)))) return rc;
      if ((rc = tileSums.reserve(sizeof(uint32_t) * 4096))) return rc;
      UB200_CUDA(cudaMemsetAsync(binCount.p, 0, binCount.cap, st));
    }
    if ((rc = codeSlot.reserve(sizeof(uint2) * (size_t)N))) return rc;
    if ((rc = unstable.reserve(sizeof(int) * (size_t)N))) return rc;
    if ((rc = sortedIndex.reserve(sizeof(int) * (size_t)N))) return rc;
    if ((rc = sortedRec.reserve(sizeof(T) * recWords() * (size_t)N))) return rc;
    const int nb = (N + 255) / 256;
    ibmBinByCell<T4><<<nb, 256, 0, st>>>((const T4 *)pos, N, grid, binCount.as<uint32_t>(), codeSlot.as<uint2>());
    UB200_LAUNCHED();
    if ((rc = exclusiveScanAndClear(binCount.as<uint32_t>(), ncells, binStart.as<uint32_t>(), tileSums.as<uint32_t>(), st)))
      return rc;
    if ((rc = scatterToBinsLaunch(codeSlot.as<uint2>(), binStart.as<uint32_t>(), N, unstable.as<int>(), st))) return rc;
#define UB200_ORDER(SS)                                                                                                  \
  ibmOrderSorted<T4, SS><<<nb, 256, 0, st>>>(unstable.as<int>(), codeSlot.as<uint2>(), binStart.as<uint32_t>(), (const T4 *)pos, \
                                             (const T *)val, valStride, N, grid, kern, sortedIndex.as<int>(), sortedRec.as<T>())
    switch (kern.support) {
    case 3: UB200_ORDER(3); break;
    case 4: UB200_ORDER(4); break;
    case 5: UB200_ORDER(5); break;
    default: UB200_ORDER(7); break;
    }
#undef UB200_ORDER
    UB200_LAUNCHED();
    sortedValidFor = N;
    return UB200_OK;
  }

  // grid3 is completely overwritten on the node-centric path; the generic path accumulates into it
  int spread(const void *pos, const void *val, int valStride, int N, T *grid3, bool gridIsZero, cudaStream_t st) {
    using T4 = typename Real4<T>::type;
    if (nodeCentric) {
      int rc = prepare(pos, val, valStride, N, st);
      if (rc) return rc;
      dim3 grd((nxPad + kRbX - 1) / kRbX, (grid.n[1] + kRbY - 1) / kRbY, (grid.n[2] + kRbZ - 1) / kRbZ);
      // staging budget per CTA: 52 KB (a whole brick in one pass at FCM densities) for the narrow supports, 26 KB (twice
      // the resident CTAs, bricks with more records take two passes) for supports 5 and 7, whose records are large -
      // measured: FCM 128^3 / support 3: 0.504 ms (52) vs 0.525 (26); PSE far field 256^3 / support 7: 3.68 ms (52) vs
      // 3.54 (26). UB200_IBM_SPREAD_KB=52|26 overrides (A/B switch).
      static const int envKB = getenv("UB200_IBM_SPREAD_KB") ? atoi(getenv("UB200_IBM_SPREAD_KB")) : 0;
      const bool smallStage = envKB ? envKB <= 26 : kern.support >= 5;
#define UB200_SPREAD_KB(SS, KB)                                                                                         \
  {                                                                                                                      \
    auto kfn = ibmSpreadRows<T, SS, KB>;                                                                                 \
    const size_t sm = RowBrickGeom<T, SS, KB>::smemBytes;                                                                \
    UB200_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                         \
    kfn<<<grd, kRbThreads, sm, st>>>(sortedRec.as<T>(), binStart.as<uint32_t>(), grid, nxPad, grid3, 0, grid.n[2]);      \
  }
#define UB200_SPREAD(SS) { if (smallStage) UB200_SPREAD_KB(SS, 26) else UB200_SPREAD_KB(SS, 52) }
      switch (kern.support) {
      case 3: UB200_SPREAD(3) break;
      case 4: UB200_SPREAD(4) break;
      case 5: UB200_SPREAD(5) break;
      default: UB200_SPREAD(7) break;
      }
#undef UB200_SPREAD_KB
#undef UB200_SPREAD
      UB200_LAUNCHED();
      return UB200_OK;
    }
    if (!gridIsZero)
      UB200_CUDA(cudaMemsetAsync(grid3, 0, sizeof(T) * 3 * (size_t)nxPad * grid.n[1] * grid.n[2], st));
    ibmWarpPerParticle<T4, T, true, false><<<(N + 3) / 4, 128, 0, st>>>((const T4 *)pos, (const T *)val, valStride, N,
                                                                         grid, kern, nxPad, grid3, nullptr, nullptr);
    UB200_LAUNCHED();
    sortedValidFor = -1;
    return UB200_OK;
  }

  // reuseRecords: positions are the ones of the preceding spread on this state
  int gather(const void *pos, int N, const T *grid3, T *out3, bool accumulate, bool reuseRecords, cudaStream_t st) {
    using T4 = typename Real4<T>::type;
    if (nodeCentric && sortedGather) {
      if (!(reuseRecords && sortedValidFor == N)) {
        int rc = prepare(pos, nullptr, 0, N, st);
        if (rc) return rc;
      }
      const int nb = (N + 127) / 128;
#define UB200_GATHER(SS, ACC)                                                                                 \
  ibmGatherSorted<T, SS, ACC><<<nb, 128, 0, st>>>(sortedRec.as<T>(), sortedIndex.as<int>(), N, grid, nxPad, grid3, out3)
      switch (kern.support) {
      case 3: if (accumulate) UB200_GATHER(3, true); else UB200_GATHER(3, false); break;
      case 4: if (accumulate) UB200_GATHER(4, true); else UB200_GATHER(4, false); break;
      case 5: if (accumulate) UB200_GATHER(5, true); else UB200_GATHER(5, false); break;
      default: if (accumulate) UB200_GATHER(7, true); else UB200_GATHER(7, false); break;
      }
#undef UB200_GATHER
      UB200_LAUNCHED();
      return UB200_OK;
    }
    // supports 5 / 7: the spread of the same positions left a cell-sorted permutation behind
    const int *order = (nodeCentric && reuseRecords && sortedValidFor == N) ? sortedIndex.as<int>() : nullptr;
    if (accumulate)
      ibmWarpPerParticle<T4, T, false, true><<<(N + 3) / 4, 128, 0, st>>>((const T4 *)pos, (const T *)nullptr, 0, N, grid,
                                                                         kern, nxPad, const_cast<T *>(grid3), out3, order);
    else
      ibmWarpPerParticle<T4, T, false, false><<<(N + 3) / 4, 128, 0, st>>>((const T4 *)pos, (const T *)nullptr, 0, N,
                                                                          grid, kern, nxPad, const_cast<T *>(grid3), out3, order);
    UB200_LAUNCHED();
    return UB200_OK;
  }
};

} // namespace ub200
