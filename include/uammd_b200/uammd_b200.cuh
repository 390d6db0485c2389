/* uammd_b200 - C++14 glue between UAMMD's concepts and the C ABI (include/uammd_b200.h).
 *
 * Include it from a UAMMD program (after "uammd.cuh") and link with -luammd_b200. The classes satisfy the
 * reference's duck-typed concepts, so they drop into its templates:
 *
 *   uammd::b200::CellList         NeighbourList concept (Interactor/NeighbourList/CellList.cuh:66-208):
 *                                 PairForces<AnyPotential, b200::CellList> builds the list with our CUDA path and
 *                                 runs ANY user Transverser through the reference's own traversal kernel, because
 *                                 getCellList() returns the reference's CellListData bit for bit.
 *   uammd::b200::ColumnList       NeighbourList concept over the engine's own half-cell list: PairForces<AnyPotential,
 *                                 b200::ColumnList> runs ANY user Transverser (compute / set / getInfo / zero / accumulate,
 *                                 utils/TransverserUtils.cuh:151-274) through a B200 traversal compiled into the user's
 *                                 translation unit: one warp per column of half cells, the halo staged once in shared
 *                                 memory, eight lanes per home particle - 196 candidates instead of 340 and no
 *                                 thread-per-particle walk through global memory (NeighbourList/common.cuh:10-34).
 *   uammd::b200::LJ               Potential::LJ with access to its device parameter table.
 *   uammd::b200::PairForcesLJ     Interactor (Interactor/Interactor.cuh:56-119) = PairForces<Potential::LJ, CellList>
 *                                 with the specialised LJ traversal (Interactor/PairForces.cu:43-78).
 *   uammd::b200::VerletList       NeighbourList concept with a skin (Interactor/NeighbourList/VerletList.cuh:83-201): the list
 *                                 is built by our CUDA path in the reference's VerletListData layout, bit for bit (from the
 *                                 first getVerletList()/transverseList() on; b200::PairForcesLJ walks the engine's row list), so
 *                                 PairForces<AnyPotential, b200::VerletList> runs user Transversers through the reference's
 *                                 own kernel; b200::PairForcesLJ takes it as its neighbour list for the fast LJ path.
 *   uammd::b200::PSE              BDHI Method concept for BDHI::EulerMaruyama<Method> = BDHI::PSE (BDHI_PSE.cuh:82-176).
 *   uammd::b200::VerletNVTGronbechJensen / VerletNVTBasic   Integrators = VerletNVT::GronbechJensen / VerletNVT::Basic
 *                                 (Integrator/VerletNVT.cuh:59-117), bit-identical half steps.
 *   uammd::b200::BDEulerMaruyama  Integrator = BD::EulerMaruyama (Integrator/BrownianDynamics.cuh:111-126).
 *   uammd::b200::FCM<Kernel>      BDHI Method concept (Integrator/BDHI/BDHI_FCM.cuh:85-153) for
 *                                 BDHI::EulerMaruyama<Method> (Integrator/BDHI/BDHI_EulerMaruyama.cuh:64-98).
 *   uammd::b200::FCM_impl<K, KT>  the class the reference's own tests drive (Integrator/BDHI/FCM/FCM_impl.cuh:36-129):
 *                                 same Parameters, computeHydrodynamicDisplacements returns the reference's pair of
 *                                 cached vectors (linear, angular).
 *   uammd::b200::IBM<Kernel>      spread / gather of misc/IBM.cuh:99-203 for the Peskin and Gaussian windows.
 *   uammd::b200::DPDPotential     Potential::DPD plus the getTransverser PairForces looks for (the stock class only has
 *                                 the pre-v2 getForceTransverser and is silently skipped): drops into
 *                                 PairForces<b200::DPDPotential, AnyNeighbourList>.
 *   uammd::b200::PairForcesDPD    Interactor = PairForces<Potential::DPD, CellList> with the specialised DPD traversal.
 * Error codes of the C ABI are converted into the reference's exception convention (std::runtime_error).
 */
#ifndef UAMMD_B200_GLUE_CUH
#define UAMMD_B200_GLUE_CUH
#include "uammd.cuh"
#include "Interactor/Interactor.cuh"
#include "Interactor/NeighbourList/CellList.cuh"
#include "Interactor/NeighbourList/VerletList.cuh"
#include "Integrator/Integrator.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "Integrator/BDHI/BDHI.cuh"
#include "Integrator/BDHI/FCM/FCM_kernels.cuh"
#include "Integrator/BDHI/FCM/utils.cuh"
#include "Interactor/Potential/DPD.cuh"
#include "misc/IBM_kernels.cuh"
#include <thrust/device_vector.h>
#include <thrust/transform.h>
#include <thrust/iterator/counting_iterator.h>
#include "../uammd_b200.h"
#include "colgeom.h"
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace uammd {
namespace b200 {

inline void check(int code, const char *where) {
  if (code != UB200_OK) {
    std::string msg = std::string("[uammd_b200] ") + where + ": " + ub200_error_string(code);
    if (code == UB200_ERR_CUDA) msg += " (cudaError " + std::to_string(ub200_last_cuda_error()) + ")";
    System::log<System::ERROR>("%s", msg.c_str());
    throw std::runtime_error(msg);
  }
}

#ifndef DOUBLE_PRECISION /* path 1 is single precision (BASELINE config 2); a double build only gets path 2 */
/* ---------------------------------------------------------------- CellList ---------------------------------- */
class CellList {
  shared_ptr<ParticleGroup> pg;
  ub200_celllist *handle = nullptr;
  connection posWriteConnection;
  bool force_next_update = true;
  real3 currentCutOff = real3();
  Box currentBox = Box();
  Grid grid;

  /* CellList::createUpdateGrid (CellList.cuh:100-126): infinite dimensions get 64 non periodic cells of one cut-off,
     cellDim = int(L/rc), dimensions with <= 3 cells collapse to one cell */
  Grid createUpdateGrid(Box box, real3 cutOff) {
    real3 L = box.boxSize;
    constexpr real inf = std::numeric_limits<real>::max();
    const bool fx = L.x < inf, fy = L.y < inf, fz = L.z < inf;
    if (!fx) L.x = 64 * cutOff.x;
    if (!fy) L.y = 64 * cutOff.y;
    if (!fz) L.z = 64 * cutOff.z;
    Box ubox(L);
    ubox.setPeriodicity(box.isPeriodicX() and fx, box.isPeriodicY() and fy, box.isPeriodicZ() and fz);
    int3 cd = make_int3(L / cutOff);
    if (cd.x <= 3) cd.x = 1;
    if (cd.y <= 3) cd.y = 1;
    if (cd.z <= 3) cd.z = 1;
    return Grid(ubox, cd);
  }

public:
  CellList(shared_ptr<ParticleData> pd) : CellList(std::make_shared<ParticleGroup>(pd)) {}
  CellList(shared_ptr<ParticleGroup> pg) : pg(pg) {
    check(ub200_celllist_create(&handle), "celllist_create");
    posWriteConnection = pg->getParticleData()->getPosWriteRequestedSignal()->connect(
        [this]() { this->force_next_update = true; });
  }
  CellList(const CellList &) = delete;
  ~CellList() {
    posWriteConnection.disconnect();
    ub200_celllist_destroy(handle);
  }

  void update(Box box, real cutOff, cudaStream_t st = 0) { update(box, make_real3(cutOff), st); }

  void update(Box box, real3 cutOff, cudaStream_t st = 0) {
    const bool rebuild = force_next_update or cutOff.x != currentCutOff.x or cutOff.y != currentCutOff.y or
                         cutOff.z != currentCutOff.z or box != currentBox;
    if (!rebuild) return;
    currentBox = box;
    currentCutOff = cutOff;
    grid = createUpdateGrid(box, cutOff);
    auto pd = pg->getParticleData();
    const int N = pg->getNumberParticles();
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    const int *gidx = pg->getIndicesRawPtr(access::location::gpu);
    const float L[3] = {grid.box.boxSize.x, grid.box.boxSize.y, grid.box.boxSize.z};
    const int periodic[3] = {grid.box.isPeriodicX(), grid.box.isPeriodicY(), grid.box.isPeriodicZ()};
    const int cd[3] = {grid.cellDim.x, grid.cellDim.y, grid.cellDim.z};
    check(ub200_celllist_build_f32(handle, pos.raw(), gidx, N, L, periodic, cd, (void *)st), "celllist_build");
    force_next_update = false;
  }

  /* same launch as CellList::transverseList (CellList.cuh:165-182): the reference's generic kernel over OUR list */
  template <class Transverser> void transverseList(Transverser &tr, cudaStream_t st = 0) {
    const int N = pg->getNumberParticles();
    const int Nthreads = 128;
    const int Nblocks = N / Nthreads + ((N % Nthreads) ? 1 : 0);
    auto globalIndex = pg->getIndexIterator(access::location::gpu);
    size_t shMemorySize = SFINAE::SharedMemorySizeDelegator<Transverser>().getSharedMemorySize(tr);
    SFINAE::TransverserAdaptor<Transverser>::prepare(tr, pg->getParticleData());
    NeighbourList_ns::transverseWithNeighbourContainer<<<Nblocks, Nthreads, shMemorySize, st>>>(
        tr, globalIndex, this->getNeighbourContainer(), N);
    CudaCheckError();
  }

  CellListBase::CellListData getCellList() {
    ub200_celllist_view v;
    check(ub200_celllist_view_get(handle, &v), "celllist_view_get");
    CellListBase::CellListData cl;
    cl.cellStart = v.d_cellStart;
    cl.cellEnd = v.d_cellEnd;
    cl.sortPos = reinterpret_cast<const real4 *>(v.d_sortPos);
    cl.groupIndex = v.d_groupIndex;
    cl.grid = grid;
    cl.VALID_CELL = v.VALID_CELL;
    return cl;
  }

  CellList_ns::NeighbourContainer getNeighbourContainer() { return CellList_ns::NeighbourContainer(getCellList()); }

  ub200_celllist *getHandle() { return handle; }
  shared_ptr<ParticleGroup> getGroup() { return pg; }
};

/* ---------------------------------------------------------------- ColumnList -------------------------------- */
/* Generic-Transverser traversal over the engine's half-cell list (ub200_ljengine_build_f32 / _view_get). Same geometry as
   the library's LJ column traversal (uammd_b200/csrc/lj_column.cu, include/uammd_b200/colgeom.h): a warp stages the
   5 x 5 x (6 + 4) halo of a column of six half cells (positions and group indices, row by row, image shifts applied), then
   serves the home particles four at a time, eight lanes each; a lane folds its share of the neighbourhood into a private
   quantity with the Transverser's own accumulate, the eight partial quantities are combined with accumulate through
   shuffles and lane 0 calls set. compute() receives the neighbour at the periodic image next to the home particle (a
   Transverser's own box.apply_pbc is then the identity). accumulate must be associative and commutative (a sum, a max,
   ...), which every Transverser of the reference is. */
namespace detail {
constexpr int kColTZ = 6, kColCap = 416, kColWarps = 4;
template <class Q> __device__ inline Q shflXorPod(const Q &q, int o) {
  static_assert(sizeof(Q) % 4 == 0, "the quantity a Transverser accumulates must be made of 32-bit words");
  Q r;
  const unsigned *src = reinterpret_cast<const unsigned *>(&q);
  unsigned *dst = reinterpret_cast<unsigned *>(&r);
#pragma unroll
  for (int w = 0; w < (int)(sizeof(Q) / 4); w++) dst[w] = __shfl_xor_sync(0xffffffffu, src[w], o);
  return r;
}

template <class Transverser, class IndexIterator>
__global__ void __launch_bounds__(32 * kColWarps)
columnTransverse(Transverser tr, IndexIterator globalIndex, const real4 *__restrict__ finePos, const int *__restrict__ fineIdx,
                 const unsigned *__restrict__ cellStart, ub200::ColGrid cg, real3 L) {
  using Adaptor = SFINAE::TransverserAdaptor<Transverser>;
  __shared__ real4 candAll[kColWarps][kColCap];
  __shared__ int cidxAll[kColWarps][kColCap];
  __shared__ int metaAll[kColWarps][(kColTZ + 5) + (kColTZ + 1) + 2 * kColTZ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  real4 *cand = candAll[warp];
  int *cidx = cidxAll[warp];
  int *planeOff = metaAll[warp], *homePre = planeOff + kColTZ + 5, *homeOff = homePre + kColTZ + 1, *homeG = homeOff + kColTZ;
  const int nzc = (cg.nz + kColTZ - 1) / kColTZ;
  const int ncols = cg.nx * cg.ny * nzc;
  for (int col = blockIdx.x * kColWarps + warp; col < ncols; col += gridDim.x * kColWarps) {
    const int x0 = col % cg.nx, t1 = col / cg.nx, y0 = t1 % cg.ny, z0 = (t1 / cg.ny) * kColTZ;
    const int nHome = min(kColTZ, cg.nz - z0);
    const int nRows = 5 * (nHome + 4);
    int g0[2][2], cn[2][2], sh[2][4], hG[2], hC[2], hRel[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int r = lane + 32 * q;
      g0[q][0] = g0[q][1] = 0; cn[q][0] = cn[q][1] = 0;
      sh[q][0] = sh[q][1] = sh[q][2] = sh[q][3] = 0;
      hG[q] = 0; hC[q] = 0; hRel[q] = 0;
      if (r < nRows) {
        const ub200::ColRow row = ub200::columnRow(cg, x0, y0, z0, r);
#pragma unroll
        for (int s = 0; s < 2; s++)
          if (row.n[s] > 0) {
            const unsigned a = cellStart[row.c0[s]], b = cellStart[row.c0[s] + row.n[s]];
            g0[q][s] = (int)a;
            cn[q][s] = (int)(b - a);
          }
        sh[q][0] = row.sx[0]; sh[q][1] = row.sx[1]; sh[q][2] = row.sy; sh[q][3] = row.sz;
        const int p = r / 5;
        if (r - 5 * p == 2 && p >= 2 && p < 2 + nHome) {
          const int cc = x0 + cg.nx * (y0 + cg.ny * (z0 + p - 2));
          const unsigned a = cellStart[cc], b = cellStart[cc + 1];
          hG[q] = (int)a;
          hC[q] = (int)(b - a);
          hRel[q] = row.hs ? cn[q][0] + ((int)a - g0[q][1]) : (int)a - g0[q][0];
        }
      }
    }
    if (!__any_sync(0xffffffffu, hC[0] > 0 || hC[1] > 0)) continue;
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
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int r = lane + 32 * q;
      if (r < nRows && r % 5 == 0) planeOff[r / 5] = off[q];
    }
    if (lane == 0) planeOff[nHome + 4] = total;
    if (lane <= kColTZ) homePre[lane] = isHome ? pre - myCnt : 0x3fffffff;
    if (isHome) {
      homeOff[lane] = hr >= 32 ? c1s : c0s;
      homeG[lane] = hr >= 32 ? b1 : b0;
    }
    const int nHomeP = __shfl_sync(0xffffffffu, pre, kColTZ - 1);
    const bool staged = total <= kColCap;
    if (staged) { // every lane copies its own rows: positions moved to their image, group indices beside them
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int cq = q ? cq1 : cq0;
        const real dy = sh[q][2] * L.y, dz = sh[q][3] * L.z;
        for (int k = 0; k < cq; k++) {
          const bool second = k >= cn[q][0];
          const int src = second ? g0[q][1] + (k - cn[q][0]) : g0[q][0] + k;
          real4 p = finePos[src];
          p.x += sh[q][second ? 1 : 0] * L.x; p.y += dy; p.z += dz;
          cand[off[q] + k] = p;
          cidx[off[q] + k] = fineIdx[src];
        }
      }
    }
    __syncwarp();
    for (int q0 = 0; q0 < nHomeP; q0 += (staged ? 4 : 1)) {
      // staged: four home particles per pass, eight lanes each; dense column: one particle, the whole warp walks global memory
      const int W = staged ? 8 : 32;
      const int sub = lane & (W - 1), q = q0 + lane / W;
      const bool act = q < nHomeP;
      int hz = 0;
#pragma unroll
      for (int k = 1; k < kColTZ; k++) hz += q >= homePre[k];
      const int hrel = q - homePre[hz];
      const int gs = act ? homeG[hz] + hrel : 0;
      const int ori = globalIndex[fineIdx[gs]];
      const real4 pi = finePos[gs]; // a home cell is never an image: its stored coordinates are the staged ones
      Adaptor adaptor;
      auto quantity = Adaptor::zero(tr);
      adaptor.getInfo(tr, ori);
      if (staged) {
        const int c0 = act ? planeOff[hz] : 0, c1 = act ? planeOff[hz + 5] : 0;
        for (int t = c0 + sub; t < c1; t += 8)
          Adaptor::accumulate(tr, quantity, adaptor.compute(tr, globalIndex[cidx[t]], pi, cand[t]));
      } else {
        for (int rr = 5 * hz; rr < 5 * hz + 25; rr++) {
          const int src = rr & 31, hi = rr >> 5;
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int g = __shfl_sync(0xffffffffu, hi ? g0[1][s] : g0[0][s], src);
            const int n = __shfl_sync(0xffffffffu, hi ? cn[1][s] : cn[0][s], src);
            const real dx = __shfl_sync(0xffffffffu, hi ? sh[1][s] : sh[0][s], src) * L.x;
            const real dy = __shfl_sync(0xffffffffu, hi ? sh[1][2] : sh[0][2], src) * L.y;
            const real dz = __shfl_sync(0xffffffffu, hi ? sh[1][3] : sh[0][3], src) * L.z;
            for (int t = lane; t < n; t += 32) {
              real4 pj = finePos[g + t];
              pj.x += dx; pj.y += dy; pj.z += dz;
              Adaptor::accumulate(tr, quantity, adaptor.compute(tr, globalIndex[fineIdx[g + t]], pi, pj));
            }
          }
        }
      }
      for (int o = W >> 1; o > 0; o >>= 1) {
        const auto other = shflXorPod(quantity, o);
        Adaptor::accumulate(tr, quantity, other);
      }
      if (act && sub == 0) tr.set(ori, quantity);
    }
  }
}
} // namespace detail

class ColumnList {
  shared_ptr<ParticleGroup> pg;
  ub200_ljengine *engine = nullptr;
  shared_ptr<CellList> fallback; // grids the column traversal does not take (a periodic dimension under five half cells)
  bool useFallback = false;

public:
  ColumnList(shared_ptr<ParticleData> pd) : ColumnList(std::make_shared<ParticleGroup>(pd)) {}
  ColumnList(shared_ptr<ParticleGroup> pg) : pg(pg) { check(ub200_ljengine_create(&engine), "ljengine_create"); }
  ColumnList(const ColumnList &) = delete;
  ~ColumnList() { ub200_ljengine_destroy(engine); }

  void update(Box box, real3 cutOff, cudaStream_t st = 0) { update(box, std::max({cutOff.x, cutOff.y, cutOff.z}), st); }
  /* rebuilt on every call, like CellList::update in practice (its force_next_update is never cleared) */
  void update(Box box, real cutOff, cudaStream_t st = 0) {
    auto pd = pg->getParticleData();
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    const float L[3] = {(float)box.boxSize.x, (float)box.boxSize.y, (float)box.boxSize.z};
    const int periodic[3] = {box.isPeriodicX(), box.isPeriodicY(), box.isPeriodicZ()};
    const int rc = ub200_ljengine_build_f32(engine, pos.raw(), pg->getIndicesRawPtr(access::location::gpu), pg->getNumberParticles(), L,
                                            periodic, cutOff, (void *)st);
    useFallback = rc == UB200_ERR_UNSUPPORTED;
    if (useFallback) {
      if (!fallback) fallback = std::make_shared<CellList>(pg);
      fallback->update(box, cutOff, st);
    } else {
      check(rc, "ljengine_build");
    }
  }

  template <class Transverser> void transverseList(Transverser &tr, cudaStream_t st = 0) {
    if (useFallback) { fallback->transverseList(tr, st); return; }
    ub200_ljengine_view v;
    check(ub200_ljengine_view_get(engine, &v), "ljengine_view_get");
    ub200::ColGrid cg = ub200::makeWholeColGrid(v.cells, v.periodic);
    auto globalIndex = pg->getIndexIterator(access::location::gpu);
    SFINAE::TransverserAdaptor<Transverser>::prepare(tr, pg->getParticleData());
    const size_t shMemorySize = SFINAE::SharedMemorySizeDelegator<Transverser>().getSharedMemorySize(tr);
    const int ncols = v.cells[0] * v.cells[1] * ((v.cells[2] + detail::kColTZ - 1) / detail::kColTZ);
    const int Nblocks = std::max(1, std::min((ncols + detail::kColWarps - 1) / detail::kColWarps, 148 * 4));
    detail::columnTransverse<<<Nblocks, 32 * detail::kColWarps, shMemorySize, st>>>(
        tr, globalIndex, reinterpret_cast<const real4 *>(v.d_pos), v.d_index, v.d_cellStart, cg, make_real3(v.L[0], v.L[1], v.L[2]));
    CudaCheckError();
  }
  ub200_ljengine *getHandle() { return engine; }
};

/* ---------------------------------------------------------------- VerletList -------------------------------- */
class VerletList {
  shared_ptr<ParticleGroup> pg;
  ub200_verletlist *handle = nullptr;
  connection posWriteConnection, reorderConnection;
  bool forceNextUpdate = true, reordered = true;
  Box currentBox = Box();
  real currentCutOff = 0;
  /* BasicNeighbourListData carries a StrideIterator whose types are protected members of BasicNeighbourListBase */
  struct MakeData : BasicNeighbourListBase {
    static BasicNeighbourListData make(const ub200_verletlist_view &v) {
      BasicNeighbourListData nl;
      nl.neighbourList = v.d_neighbourList;
      nl.numberNeighbours = v.d_numberNeighbours;
      nl.sortPos = reinterpret_cast<const real4 *>(v.d_sortPos);
      nl.groupIndex = v.d_groupIndex;
      nl.particleStride = StrideIterator(CountingIterator(0), NeighbourListOffsetFunctor(v.particleStride));
      return nl;
    }
  };

public:
  VerletList(shared_ptr<ParticleData> pd) : VerletList(std::make_shared<ParticleGroup>(pd, "All")) {}
  VerletList(shared_ptr<ParticleGroup> pg) : pg(pg) {
    check(ub200_verletlist_create(&handle), "verletlist_create");
    auto pd = pg->getParticleData();
    posWriteConnection = pd->getPosWriteRequestedSignal()->connect([this]() { this->forceNextUpdate = true; });
    reorderConnection = pd->getReorderSignal()->connect([this]() { this->forceNextUpdate = true; this->reordered = true; });
  }
  VerletList(const VerletList &) = delete;
  ~VerletList() {
    posWriteConnection.disconnect();
    reorderConnection.disconnect();
    ub200_verletlist_destroy(handle);
  }

  /* VerletList::update (VerletList.cuh:111-124): the wrapper-level flag only gates the call; the drift check inside
     decides about the rebuild, a particle reorder forces it (handleReorder :185-190) */
  void update(Box box, real cutOff, cudaStream_t st = 0) {
    if (!(forceNextUpdate or box != currentBox or cutOff != currentCutOff)) return;
    forceNextUpdate = false;
    auto pd = pg->getParticleData();
    pd->hintSortByHash(box, make_real3(cutOff * 0.5));
    currentBox = box;
    currentCutOff = cutOff;
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    const int *gidx = pg->getIndicesRawPtr(access::location::gpu);
    const float L[3] = {box.boxSize.x, box.boxSize.y, box.boxSize.z};
    const int periodic[3] = {box.isPeriodicX(), box.isPeriodicY(), box.isPeriodicZ()};
    check(ub200_verletlist_update_f32(handle, pos.raw(), gidx, pg->getNumberParticles(), L, periodic, cutOff, reordered, nullptr,
                                      (void *)st),
          "verletlist_update");
    reordered = false;
  }
  void update(Box box, real3 cutOff, cudaStream_t st = 0) {
    if (cutOff.x != cutOff.y or cutOff.x != cutOff.z) throw std::runtime_error("[VerletList] Invalid argument");
    update(box, cutOff.x, st);
  }

  template <class Transverser> void transverseList(Transverser &tr, cudaStream_t st = 0) {
    const int N = pg->getNumberParticles();
    const int Nthreads = 128;
    const int Nblocks = N / Nthreads + ((N % Nthreads) ? 1 : 0);
    auto globalIndex = pg->getIndexIterator(access::location::gpu);
    SFINAE::TransverserAdaptor<Transverser>::prepare(tr, pg->getParticleData());
    size_t shMemorySize = SFINAE::SharedMemorySizeDelegator<Transverser>().getSharedMemorySize(tr);
    NeighbourList_ns::transverseWithNeighbourContainer<<<Nblocks, Nthreads, shMemorySize, st>>>(
        tr, globalIndex, this->getNeighbourContainer(), N);
    CudaCheckError();
  }

  VerletListBase::VerletListData getVerletList() {
    ub200_verletlist_view v;
    check(ub200_verletlist_view_get(handle, &v), "verletlist_view_get");
    return MakeData::make(v);
  }
  VerletListBase_ns::NeighbourContainer getNeighbourContainer() { return VerletListBase_ns::NeighbourContainer(getVerletList()); }
  void setCutOffMultiplier(real m) { check(ub200_verletlist_set_cutoff_multiplier(handle, m), "set_cutoff_multiplier"); }
  int getNumberOfStepsSinceLastUpdate() {
    int steps = 0;
    check(ub200_verletlist_stats(handle, &steps, nullptr), "verletlist_stats");
    return steps;
  }
  ub200_verletlist *getHandle() { return handle; }
};

/* ---------------------------------------------------------------- LJ ---------------------------------------- */
/* Radial<LJFunctor> keeps its PairParameters table {cutOff2, sigma2, epsilonDivSigma2, shift} in a protected
   BasicParameterHandler (Potential/RadialPotential.cuh:55-58, ParameterHandler.cuh:41-65); deriving exposes it. */
class LJ : public Potential::LJ {
public:
  struct DeviceTable {
    const void *d_params;
    int ntypes;
  };
  DeviceTable getDeviceTable() {
    auto it = this->pairParameters->getIterator();
    return {it.globalMem, it.ntypes};
  }
};

class PairForcesLJ : public Interactor {
  shared_ptr<CellList> nl;   // when set, the reference-layout cell list is built and traversed (users reading getNeighbourList())
  shared_ptr<VerletList> vl; // when set, the Verlet list is the neighbour list (PairForces<LJ, VerletList>)
  shared_ptr<LJ> pot;
  Box box;
  ub200_ljengine *engine = nullptr; // default: the engine's own half-cell list + column traversal (ub200_ljengine_sum_f32)
  std::vector<float> hostTable;     // host copy of the pair parameters, refreshed when the potential's table changes size
  const void *hostTableSource = nullptr;

  const std::vector<float> &tableOnHost(const LJ::DeviceTable &table, cudaStream_t st) {
    const size_t n = (size_t)table.ntypes * table.ntypes * 4;
    if (hostTable.size() != n or hostTableSource != table.d_params) {
      hostTable.resize(n);
      CudaSafeCall(cudaMemcpyAsync(hostTable.data(), table.d_params, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      CudaSafeCall(cudaStreamSynchronize(st));
      hostTableSource = table.d_params;
    }
    return hostTable;
  }

public:
  struct Parameters {
    Box box = Box(std::numeric_limits<real>::infinity());
    shared_ptr<CellList> nl = nullptr;
    shared_ptr<VerletList> verletList = nullptr;
  };
  PairForcesLJ(shared_ptr<ParticleData> pd, Parameters par, shared_ptr<LJ> pot)
      : PairForcesLJ(std::make_shared<ParticleGroup>(pd, "All"), par, pot) {}
  PairForcesLJ(shared_ptr<ParticleGroup> pg, Parameters par, shared_ptr<LJ> pot)
      : Interactor(pg, "b200::PairForcesLJ"), nl(par.nl), vl(par.verletList), pot(pot), box(par.box) {
    if (!nl and !vl) check(ub200_ljengine_create(&engine), "ljengine_create");
  }
  ~PairForcesLJ() { ub200_ljengine_destroy(engine); }
  /* setPotParameters after the first sum(): call this so that the host copy of the table is fetched again */
  void parametersChanged() { hostTableSource = nullptr; }

  void updateBox(Box newBox) override { box = newBox; }

  /* Interactor::sum: accumulates into pd's force / energy / virial like Radial::Transverser::set */
  void sum(Computables comp, cudaStream_t st = 0) override {
    const real rcut = pot->getCutOff();
    if (engine) { // neighbour search, traversal and the NBody fallback in one call
      auto force = comp.force ? pd->getForce(access::location::gpu, access::mode::readwrite).raw() : nullptr;
      auto energy = comp.energy ? pd->getEnergy(access::location::gpu, access::mode::readwrite).raw() : nullptr;
      auto virial = comp.virial ? pd->getVirial(access::location::gpu, access::mode::readwrite).raw() : nullptr;
      auto pos = pd->getPos(access::location::gpu, access::mode::read);
      const int *gidx = pg->getIndicesRawPtr(access::location::gpu);
      const auto table = pot->getDeviceTable();
      const auto &host = tableOnHost(table, st);
      const float L[3] = {(float)box.boxSize.x, (float)box.boxSize.y, (float)box.boxSize.z};
      const int periodic[3] = {box.isPeriodicX(), box.isPeriodicY(), box.isPeriodicZ()};
      check(ub200_ljengine_sum_f32(engine, pos.raw(), gidx, pg->getNumberParticles(), L, periodic, host.data(), table.ntypes, force,
                                   energy, virial, gidx, 1, 0, 0x7fffffff, (void *)st),
            "ljengine_sum");
      return;
    }
    // PairForces.cu:49-53: a box no larger than 3 cut-offs in every dimension takes the all-pairs NBody path
    const bool nbody = box.boxSize.x <= 3 * rcut and box.boxSize.y <= 3 * rcut and box.boxSize.z <= 3 * rcut;
    if (!nbody) {
      if (vl) vl->update(box, rcut, st);
      else nl->update(box, rcut, st);
    }
    auto force = comp.force ? pd->getForce(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    auto energy = comp.energy ? pd->getEnergy(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    auto virial = comp.virial ? pd->getVirial(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    const auto table = pot->getDeviceTable();
    const int *gidx = pg->getIndicesRawPtr(access::location::gpu);
    if (nbody) {
      std::vector<float> host((size_t)table.ntypes * table.ntypes * 4);
      CudaSafeCall(cudaMemcpyAsync(host.data(), table.d_params, host.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
      CudaSafeCall(cudaStreamSynchronize(st));
      auto pos = pd->getPos(access::location::gpu, access::mode::read);
      const float L[3] = {(float)box.boxSize.x, (float)box.boxSize.y, (float)box.boxSize.z};
      const int periodic[3] = {box.isPeriodicX(), box.isPeriodicY(), box.isPeriodicZ()};
      check(ub200_lj_nbody_f32(pos.raw(), gidx, pg->getNumberParticles(), L, periodic, host.data(), table.ntypes, force, energy,
                               virial, (void *)st),
            "lj_nbody");
      return;
    }
    if (vl) {
      // the Verlet entry point takes the host copy of the table (tiny; cached on the device by the library)
      std::vector<float> host((size_t)table.ntypes * table.ntypes * 4);
      CudaSafeCall(cudaMemcpyAsync(host.data(), table.d_params, host.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
      CudaSafeCall(cudaStreamSynchronize(st));
      check(ub200_lj_sum_verlet_f32(vl->getHandle(), host.data(), table.ntypes, force, energy, virial, gidx, (void *)st), "lj_sum_verlet");
      return;
    }
    check(ub200_lj_sum_devparams_f32(nl->getHandle(), table.d_params, table.ntypes, force, energy, virial, gidx,
                                     (void *)st),
          "lj_sum");
  }
  shared_ptr<CellList> getNeighbourList() { return nl; }
};

/* ---------------------------------------------------------------- DPD --------------------------------------- */
/* Potential::DPD (Interactor/Potential/DPD.cuh:40-180) with the method PairForces looks for. At this commit the stock class
   only exposes getForceTransverser(box, pd); Potential::has_getTransverser<Potential::DPD> is false and
   PairForces<Potential::DPD> silently sums a null transverser. The arithmetic - ForceTransverser::{getInfo,
   compute, set}, :92-159 - is the reference's own, untouched. */
class DPDPotential : public Potential::DPD {
public:
  using Potential::DPD::DPD;
  auto getTransverser(Interactor::Computables comp, Box box, shared_ptr<ParticleData> pd) {
    return this->getForceTransverser(box, pd);
  }
  /* what the specialised traversal needs; advance() increments the step like getForceTransverser does (:165) */
  struct Step {
    real rcut, gamma, sigma, A;
    int step;
  };
  Step advance() {
    step++;
    return {rcut, gamma.gamma, sigma, A, step};
  }
};
static_assert(Potential::has_getTransverser<DPDPotential>::value, "PairForces does not see b200::DPDPotential's transverser");

/* Interactor = PairForces<Potential::DPD, CellList> (Interactor/PairForces.cu:43-78) with the specialised traversal
   (ub200_dpd_sum_f32: cell blocks staged once, Saru / Box-Muller body on full warps). The Saru seed of the reference is a
   function-local static of getForceTransverser (DPD.cuh:165) and cannot be read: this class draws its own from the system
   generator, once, like the reference does. */
class PairForcesDPD : public Interactor {
  shared_ptr<CellList> nl;
  shared_ptr<DPDPotential> pot;
  Box box;
  uint seed;

public:
  struct Parameters {
    Box box = Box(std::numeric_limits<real>::infinity());
    shared_ptr<CellList> nl = nullptr;
  };
  PairForcesDPD(shared_ptr<ParticleData> pd, Parameters par, shared_ptr<DPDPotential> pot)
      : PairForcesDPD(std::make_shared<ParticleGroup>(pd, "All"), par, pot) {}
  PairForcesDPD(shared_ptr<ParticleGroup> pg, Parameters par, shared_ptr<DPDPotential> pot)
      : Interactor(pg, "b200::PairForcesDPD"), nl(par.nl), pot(pot), box(par.box) {
    if (!nl) nl = std::make_shared<CellList>(pg);
    seed = (uint)sys->rng().next(); /* Saru takes 32-bit seeds: the reference's 64-bit draw is truncated the same way */
  }
  void updateBox(Box newBox) override { box = newBox; }
  void updateTemperature(real T) override { pot->updateTemperature(T); }
  void updateTimeStep(real dt) override { pot->updateTimeStep(dt); }
  void setSeed(uint s) { seed = s; }
  uint getSeed() const { return seed; }

  void sum(Computables comp, cudaStream_t st = 0) override {
    if (!comp.force) return; /* no energy in DPD (DPD.cuh:172-180) */
    nl->update(box, pot->getCutOff(), st);
    const auto p = pot->advance();
    auto vel = pd->getVel(access::location::gpu, access::mode::read);
    auto force = pd->getForce(access::location::gpu, access::mode::readwrite);
    const int *gidx = pg->getIndicesRawPtr(access::location::gpu);
    check(ub200_dpd_sum_f32(nl->getHandle(), vel.raw(), p.A, p.gamma, p.sigma, p.rcut, seed, (uint)p.step, pd->getNumParticles(),
                            force.raw(), gidx, (void *)st),
          "dpd_sum");
  }
  shared_ptr<CellList> getNeighbourList() { return nl; }
};
#endif /* !DOUBLE_PRECISION */

/* ---------------------------------------------------------------- FCM --------------------------------------- */
namespace detail {
template <class Kernel> struct KernelDescriptor;
template <> struct KernelDescriptor<BDHI::FCM_ns::Kernels::Peskin::threePoint> {
  static ub200_ibm_kernel make(real h, real) { return {UB200_KERNEL_PESKIN3, 3, (double)h, 0, 0, 0}; }
  static real radius(real h, real) { return h; }
};
template <> struct KernelDescriptor<BDHI::FCM_ns::Kernels::Peskin::fourPoint> {
  static ub200_ibm_kernel make(real h, real) { return {UB200_KERNEL_PESKIN4, 4, (double)h, 0, 0, 0}; }
  static real radius(real h, real) { return h * BDHI::FCM_ns::Kernels::Peskin::fourPoint::fac; }
};
/* FCM_ns::Kernels::Gaussian keeps width/prefactor private: recompute them with its constructor's rule
   (Integrator/BDHI/FCM/FCM_kernels.cuh:22-46) */
template <> struct KernelDescriptor<BDHI::FCM_ns::Kernels::Gaussian> {
  static double upsampling(double tolerance) {
    const double amin = 0.55, amax = 1.65, x = -std::log10(3 * tolerance) / 10.0;
    return std::min(amin + x * (amax - amin), amax);
  }
  static ub200_ibm_kernel make(real h, real tolerance) {
    const double width = (double)h * upsampling(tolerance);
    ub200_ibm_kernel k;
    k.kind = UB200_KERNEL_GAUSSIAN;
    k.h = h;
    k.prefactor = std::pow(2.0 * M_PI * width * width, -0.5);
    k.tau = -0.5 / (width * width);
    const double dr = 0.5 * h;
    double r = dr;
    while (k.prefactor * std::exp(k.tau * r * r) > tolerance) r += dr;
    k.support = std::max(3, int(2 * r / h + 0.5));
    k.rmax = k.support * (double)h;
    return k;
  }
  static real radius(real h, real tolerance) { return h * upsampling(tolerance) * std::sqrt(M_PI); }
};
} // namespace detail

/* descriptor of a kernel INSTANCE (FCM_impl / IBM receive the objects the caller built). The Gaussians keep width and
   prefactor private: prefactor = phi(0), tau = log(phi(r)/phi(0)) / r^2 from two host evaluations of their own phi. */
inline ub200_ibm_kernel describeKernel(const BDHI::FCM_ns::Kernels::Peskin::threePoint &, real h) {
  return {UB200_KERNEL_PESKIN3, 3, (double)h, 0, 0, 0};
}
inline ub200_ibm_kernel describeKernel(const BDHI::FCM_ns::Kernels::Peskin::fourPoint &, real h) {
  return {UB200_KERNEL_PESKIN4, 4, (double)h, 0, 0, 0};
}
inline ub200_ibm_kernel describeKernel(const IBM_kernels::Peskin::threePoint &k, real) { return {UB200_KERNEL_PESKIN3, 3, 1.0 / (double)k.invh, 0, 0, 0}; }
inline ub200_ibm_kernel describeKernel(const IBM_kernels::Peskin::fourPoint &k, real) { return {UB200_KERNEL_PESKIN4, 4, 1.0 / (double)k.invh, 0, 0, 0}; }
template <class G> inline ub200_ibm_kernel describeGaussian(const G &k, real h) {
  ub200_ibm_kernel d;
  d.kind = UB200_KERNEL_GAUSSIAN;
  d.support = k.support;
  d.h = h;
  d.rmax = k.rmax;
  const double p0 = k.phi(real(0), real3()), r1 = 0.25 * (double)k.rmax, p1 = k.phi(real(r1), real3());
  d.prefactor = p0;
  d.tau = std::log(p1 / p0) / (r1 * r1);
  return d;
}
/* IBM_kernels::BarnettMagland (misc/IBM_kernels.cuh:91-113) keeps its norm private: 1/norm = phi(0). It declares no
   support (the reference leaves that to the kernel that wraps it), so the caller states it. */
inline ub200_ibm_kernel describeBarnettMagland(const IBM_kernels::BarnettMagland &k, int support) {
  return {UB200_KERNEL_BARNETT_MAGLAND, support, 2.0 * (double)k.alpha / support, (double)k.phi(real(0)), (double)k.beta, (double)k.alpha};
}
/* IBM_kernels::GaussianFlexible::sixPoint (misc/IBM_kernels.cuh:163-237): closed form, support 6, grid spacing h */
inline ub200_ibm_kernel describeKernel(const IBM_kernels::GaussianFlexible::sixPoint &, real h) {
  return {UB200_KERNEL_SIXPOINT, 6, (double)h, 0, 0, 0};
}
inline ub200_ibm_kernel describeKernel(const BDHI::FCM_ns::Kernels::Gaussian &k, real h) { return describeGaussian(k, h); }
inline ub200_ibm_kernel describeKernel(const BDHI::FCM_ns::Kernels::GaussianTorque &k, real h) { return describeGaussian(k, h); }

/* ---------------------------------------------------------------- FCM_impl ---------------------------------- */
/* The class the reference's tests construct directly (test/BDHI/FCM/fcm_test.cu:85-144): same Parameters, same return
   type (std::pair of uammd's pool-backed cached vectors: linear and, with torques, angular velocities). */
template <class Kernel, class KernelTorque> class FCM_impl {
public:
  template <class T> using cached_vector = BDHI::cached_vector<T>;
  struct Parameters : BDHI::Parameters {
    int3 cells = make_int3(-1, -1, -1);
    uint seed = 0;
    std::shared_ptr<Kernel> kernel = nullptr;
    std::shared_ptr<KernelTorque> kernelTorque = nullptr;
    bool adaptBoxSize = false;
  };

  FCM_impl(Parameters par) : viscosity(par.viscosity), hydrodynamicRadius(par.hydrodynamicRadius), box(par.box) {
    if (par.box.boxSize.x <= 0 or par.cells.x <= 0 or not par.kernel or not par.kernelTorque) {
      System::log<System::EXCEPTION>("FCM_impl requires a valid box, grid and instances of the spreading kernels");
      throw std::runtime_error("Invalid arguments");
    }
    if (par.seed == 0) par.seed = 0x9e3779b9u;
    Grid grid(par.box, par.cells);
    const real h = std::min({grid.cellSize.x, grid.cellSize.y, grid.cellSize.z});
    const ub200_ibm_kernel k = describeKernel(*par.kernel, h), kt = describeKernel(*par.kernelTorque, h);
    const double L[3] = {(double)box.boxSize.x, (double)box.boxSize.y, (double)box.boxSize.z};
    const int cells[3] = {par.cells.x, par.cells.y, par.cells.z};
    check(ub200_fcm_create(&handle, (int)sizeof(real), L, cells, &k, (double)viscosity, par.seed), "fcm_create");
    check(ub200_fcm_set_torque_kernel(handle, &kt), "fcm_set_torque_kernel");
  }
  FCM_impl(const FCM_impl &) = delete;
  ~FCM_impl() { ub200_fcm_destroy(handle); }

  real getHydrodynamicRadius() { return hydrodynamicRadius; }
  /* FCM_impl::getSelfMobility (FCM_impl.cuh:102-119) */
  real getSelfMobility() {
    long double rh = hydrodynamicRadius, L = box.boxSize.x, a = rh / L, a3 = a * a * a;
    const long double c = 2.83729747948061947666591710460773907l, b = 0.19457l;
    const long double a6pref = 16.0l * M_PIl * M_PIl / 45.0l + 630.0L * b * b;
    return 1.0l / (6.0l * M_PIl * viscosity * rh) * (1.0l - c * a + (4.0l / 3.0l) * M_PIl * a3 - a6pref * a3 * a3);
  }
  Box getBox() { return box; }

  /* FCM_impl::computeHydrodynamicDisplacements (FCM_impl.cuh:652-693): torque == nullptr skips the rotational part and
     leaves the second vector empty; force may be nullptr (noise only) */
  std::pair<cached_vector<real3>, cached_vector<real3>> computeHydrodynamicDisplacements(real4 *pos, real4 *force, real4 *torque,
                                                                                         int numberParticles, real temperature,
                                                                                         real prefactor, cudaStream_t st) {
    cached_vector<real3> linear(numberParticles), angular(torque ? numberParticles : 0);
    real3 *lin = thrust::raw_pointer_cast(linear.data());
    if (torque) {
      check(ub200_fcm_mdot_torque(handle, pos, force, torque, numberParticles, (double)temperature, (double)prefactor, lin,
                                  thrust::raw_pointer_cast(angular.data()), (void *)st),
            "fcm_mdot_torque");
    } else {
      check(ub200_fcm_mdot(handle, pos, force, numberParticles, (double)temperature, (double)prefactor, lin, (void *)st), "fcm_mdot");
    }
    return {std::move(linear), std::move(angular)};
  }

private:
  ub200_fcm *handle = nullptr;
  real viscosity, hydrodynamicRadius;
  Box box;
};

/* ---------------------------------------------------------------- IBM --------------------------------------- */
/* spread / gather of misc/IBM.cuh:99-203 on a regular grid (LinearIndex3D: i + nx (j + ny k)) for the windows the library
   builds (Peskin 3 / 4 point, truncated Gaussian). The library moves real3 quantities between real4 positions and a real3
   grid; other combinations the reference's templates accept (real3 positions, scalar or real4 quantities, scalar grids)
   are converted through scratch vectors here. Both calls ADD into their output like the reference. */
namespace detail {
template <class T> struct Components;
template <> struct Components<real> { static constexpr int n = 1; };
template <> struct Components<real2> { static constexpr int n = 2; };
template <> struct Components<real3> { static constexpr int n = 3; };
template <> struct Components<real4> { static constexpr int n = 4; };
struct ToReal4Pos {
  __host__ __device__ real4 operator()(real3 p) const { return make_real4(p.x, p.y, p.z, 0); }
  __host__ __device__ real4 operator()(real4 p) const { return p; }
};
struct ToReal3 {
  __host__ __device__ real3 operator()(real v) const { return make_real3(v, 0, 0); }
  __host__ __device__ real3 operator()(real3 v) const { return v; }
  __host__ __device__ real3 operator()(real4 v) const { return make_real3(v.x, v.y, v.z); }
};
template <class T> struct AddFromReal3;
template <> struct AddFromReal3<real> { __host__ __device__ real operator()(real o, real3 v) const { return o + v.x; } };
template <> struct AddFromReal3<real3> { __host__ __device__ real3 operator()(real3 o, real3 v) const { return o + v; } };
template <> struct AddFromReal3<real4> { __host__ __device__ real4 operator()(real4 o, real3 v) const { return o + make_real4(v.x, v.y, v.z, 0); } };
} // namespace detail

template <class Kernel> class IBM {
  shared_ptr<Kernel> kernel;
  Grid grid;
  ub200_ibm *handle = nullptr;
  mutable thrust::device_vector<real4> pos4;
  mutable thrust::device_vector<real3> val3, grid3, out3;

  template <class PosIterator> const real4 *positions(PosIterator pos, int N, cudaStream_t st) const {
    pos4.resize(N);
    thrust::transform(thrust::cuda::par.on(st), pos, pos + N, pos4.begin(), detail::ToReal4Pos());
    return thrust::raw_pointer_cast(pos4.data());
  }

public:
  IBM(shared_ptr<Kernel> kern, Grid a_grid) : kernel(kern), grid(a_grid) {
    const real h = std::min({grid.cellSize.x, grid.cellSize.y, grid.cellSize.z});
    const ub200_ibm_kernel k = describeKernel(*kernel, h);
    const double L[3] = {(double)grid.box.boxSize.x, (double)grid.box.boxSize.y, (double)grid.box.boxSize.z};
    const int periodic[3] = {grid.box.isPeriodicX(), grid.box.isPeriodicY(), grid.box.isPeriodicZ()};
    const int cells[3] = {grid.cellDim.x, grid.cellDim.y, grid.cellDim.z};
    check(ub200_ibm_create(&handle, (int)sizeof(real), L, periodic, cells, &k, grid.cellDim.x), "ibm_create");
  }
  IBM(const IBM &) = delete;
  ~IBM() { ub200_ibm_destroy(handle); }

  /* gridData[c] += v_p phi(x) phi(y) phi(z) over the support of every particle (IBM::spread, misc/IBM.cuh:117-138) */
  template <class PosIterator, class QuantityIterator, class GridQuantity>
  void spread(PosIterator pos, QuantityIterator v, GridQuantity *gridData, int numberParticles, cudaStream_t st = 0) const {
    const int ncells = grid.cellDim.x * grid.cellDim.y * grid.cellDim.z;
    const real4 *p = positions(pos, numberParticles, st);
    val3.resize(numberParticles);
    thrust::transform(thrust::cuda::par.on(st), v, v + numberParticles, val3.begin(), detail::ToReal3());
    grid3.assign(ncells, real3());
    check(ub200_ibm_spread(handle, p, thrust::raw_pointer_cast(val3.data()), 3, numberParticles, thrust::raw_pointer_cast(grid3.data()),
                           (void *)st),
          "ibm_spread");
    thrust::transform(thrust::cuda::par.on(st), gridData, gridData + ncells, grid3.begin(), gridData, detail::AddFromReal3<GridQuantity>());
  }

  /* Jq[p] += sum_c gridData[c] phi phi phi dV (IBM::gather with the default quadrature weights, misc/IBM.cuh:140-184) */
  template <class PosIterator, class ResultQuantity, class GridQuantityIterator>
  void gather(PosIterator pos, ResultQuantity *Jq, GridQuantityIterator gridData, int numberParticles, cudaStream_t st = 0) const {
    const int ncells = grid.cellDim.x * grid.cellDim.y * grid.cellDim.z;
    const real4 *p = positions(pos, numberParticles, st);
    grid3.resize(ncells);
    thrust::transform(thrust::cuda::par.on(st), gridData, gridData + ncells, grid3.begin(), detail::ToReal3());
    out3.assign(numberParticles, real3());
    check(ub200_ibm_gather(handle, p, numberParticles, thrust::raw_pointer_cast(grid3.data()), thrust::raw_pointer_cast(out3.data()),
                           (void *)st),
          "ibm_gather");
    thrust::transform(thrust::cuda::par.on(st), Jq, Jq + numberParticles, out3.begin(), Jq, detail::AddFromReal3<ResultQuantity>());
  }
  shared_ptr<Kernel> getKernel() { return kernel; }
};

template <class Kernel = BDHI::FCM_ns::Kernels::Gaussian> class FCM {
  shared_ptr<ParticleGroup> pg;
  ub200_fcm *handle = nullptr;
  real temperature, dt, viscosity, hydrodynamicRadius;
  Box box;

public:
  struct Parameters : BDHI::Parameters {
    int3 cells = make_int3(-1, -1, -1);
    uint seed = 0;
    bool adaptBoxSize = false;
  };

  FCM(shared_ptr<ParticleData> pd, Parameters par) : FCM(std::make_shared<ParticleGroup>(pd, "All"), par) {}

  FCM(shared_ptr<ParticleGroup> pg, Parameters par)
      : pg(pg), temperature(par.temperature), dt(par.dt), viscosity(par.viscosity), box(par.box) {
    if (par.seed == 0) par.seed = pg->getParticleData()->getSystem()->rng().next32();
    /* detail::initializeGrid (BDHI_FCM.cuh:29-48) */
    int3 cd = par.cells;
    if (cd.x <= 0) {
      if (par.hydrodynamicRadius <= 0)
        throw std::runtime_error("[uammd_b200::FCM] hydrodynamic radius needed when cells are not provided");
      const real h = Kernel::adviseGridSize(par.hydrodynamicRadius, par.tolerance);
      cd = nextFFTWiseSize3D(make_int3(box.boxSize / h));
      if (par.adaptBoxSize) box = Box(make_real3(cd) * h);
    }
    Grid grid(box, cd);
    const real h = std::min({grid.cellSize.x, grid.cellSize.y, grid.cellSize.z});
    const ub200_ibm_kernel k = detail::KernelDescriptor<Kernel>::make(h, par.tolerance);
    hydrodynamicRadius = detail::KernelDescriptor<Kernel>::radius(grid.cellSize.x, par.tolerance);
    const double L[3] = {(double)box.boxSize.x, (double)box.boxSize.y, (double)box.boxSize.z};
    const int cells[3] = {cd.x, cd.y, cd.z};
    check(ub200_fcm_create(&handle, (int)sizeof(real), L, cells, &k, (double)viscosity, par.seed), "fcm_create");
  }
  FCM(const FCM &) = delete;
  ~FCM() { ub200_fcm_destroy(handle); }

  void setup_step(cudaStream_t st = 0) {}

  /* BDHI::FCM::computeMF (BDHI_FCM.cuh:131-142): MF = M F + sqrt(2T/dt) M^1/2 dW, written straight into MF */
  void computeMF(real3 *MF, cudaStream_t st = 0) {
    auto pd = pg->getParticleData();
    auto force = pd->getForce(access::gpu, access::read);
    auto pos = pd->getPos(access::gpu, access::read);
    const int N = pg->getNumberParticles();
    check(ub200_fcm_mdot(handle, pos.raw(), force.raw(), N, (double)temperature, 1.0 / std::sqrt((double)dt), MF,
                         (void *)st),
          "fcm_mdot");
  }
  void computeBdW(real3 *BdW, cudaStream_t st = 0) {} // included in MF, like the reference
  void finish_step(cudaStream_t st = 0) {}

  real getHydrodynamicRadius() { return hydrodynamicRadius; }

  /* FCM_impl::getSelfMobility (FCM_impl.cuh:102-119) */
  real getSelfMobility() {
    long double rh = hydrodynamicRadius, L = box.boxSize.x, a = rh / L, a3 = a * a * a;
    const long double c = 2.83729747948061947666591710460773907l, b = 0.19457l;
    const long double a6pref = 16.0l * M_PIl * M_PIl / 45.0l + 630.0L * b * b;
    return 1.0l / (6.0l * M_PIl * viscosity * rh) * (1.0l - c * a + (4.0l / 3.0l) * M_PIl * a3 - a6pref * a3 * a3);
  }
  ub200_fcm *getHandle() { return handle; }
};

/* ---------------------------------------------------------------- PSE --------------------------------------- */
/* BDHI Method concept = BDHI::PSE (Integrator/BDHI/BDHI_PSE.cuh:82-176). Parameter resolution (grid, Gaussian support,
   eta, near-field cut-off and table) happens inside ub200_pse_create with the reference constructors' rules. */
class PSE {
  shared_ptr<ParticleGroup> pg;
  ub200_pse *handle = nullptr;
  real hydrodynamicRadius, M0, temperature, dt;

public:
  struct Parameters : BDHI::Parameters {
    real psi = 0.5;
    real shearStrain = 0;
  };
  PSE(shared_ptr<ParticleData> pd, Parameters par) : PSE(std::make_shared<ParticleGroup>(pd, "All"), par) {}
  PSE(shared_ptr<ParticleGroup> pg, Parameters par)
      : pg(pg), hydrodynamicRadius(par.hydrodynamicRadius), temperature(par.temperature), dt(par.dt) {
    {
      long double rh = par.hydrodynamicRadius, L = par.box.boxSize.x, a = rh / L, a3 = a * a * a;
      const long double c = 2.83729747948061947666591710460773907l, b = 0.19457l;
      const long double a6pref = 16.0l * M_PIl * M_PIl / 45.0l + 630.0L * b * b;
      M0 = 1.0l / (6.0l * M_PIl * par.viscosity * rh) * (1.0l - c * a + (4.0l / 3.0l) * M_PIl * a3 - a6pref * a3 * a3);
    }
    if (par.tolerance > 0.1) throw std::invalid_argument("Tolerance too high");
    auto sys = pg->getParticleData()->getSystem();
    const uint seedNear = sys->rng().next32(), seedFar = sys->rng().next32(); /* same draw order as the reference */
    ub200_pse_params p;
    p.L[0] = par.box.boxSize.x; p.L[1] = par.box.boxSize.y; p.L[2] = par.box.boxSize.z;
    p.viscosity = par.viscosity; p.hydrodynamicRadius = par.hydrodynamicRadius; p.tolerance = par.tolerance;
    p.psi = par.psi; p.shearStrain = par.shearStrain;
    p.cellsOverride[0] = p.cellsOverride[1] = p.cellsOverride[2] = 0;
    check(ub200_pse_create(&handle, (int)sizeof(real), &p, seedNear, seedFar), "pse_create");
  }
  PSE(const PSE &) = delete;
  ~PSE() { ub200_pse_destroy(handle); }

  void setup_step(cudaStream_t st = 0) {}
  void finish_step(cudaStream_t st = 0) {}
  void computeMF(real3 *MF, cudaStream_t st) {
    auto pd = pg->getParticleData();
    const int N = pg->getNumberParticles();
    CudaSafeCall(cudaMemsetAsync(MF, 0, N * sizeof(real3), st));
    auto pos = pd->getPos(access::gpu, access::read);
    auto force = pd->getForce(access::gpu, access::read);
    const uint seed2 = temperature > real(0) ? pd->getSystem()->rng().next32() : 0u;
    check(ub200_pse_far_mdot(handle, pos.raw(), force.raw(), N, temperature, 1.0 / sqrt((double)dt), seed2, MF, (void *)st), "pse_far");
    // with noise the Lanczos iteration of computeBdW needs the near field's Verlet list anyway: the product runs over it and
    // leaves it for the computeBdW that BDHI::EulerMaruyama issues next on the same positions (BDHI_EulerMaruyama.cu:125-166)
    if (temperature > real(0)) check(ub200_pse_near_mdot_list(handle, pos.raw(), force.raw(), 4, N, MF, (void *)st), "pse_near");
    else check(ub200_pse_near_mdot(handle, pos.raw(), force.raw(), 4, N, MF, (void *)st), "pse_near");
  }
  void computeBdW(real3 *BdW, cudaStream_t st) {
    if (temperature == real(0)) return;
    auto pd = pg->getParticleData();
    auto pos = pd->getPos(access::gpu, access::read);
    const uint seed2 = pd->getSystem()->rng().next32();
    check(ub200_pse_near_noise_reuse(handle, pos.raw(), pg->getNumberParticles(), temperature, 1.0, seed2, BdW, nullptr, (void *)st),
          "pse_noise");
  }
  void computeDivM(real3 *divM, cudaStream_t st = 0) {}
  void setShearStrain(real s) { check(ub200_pse_set_shear_strain(handle, s), "pse_shear"); }
  real getHydrodynamicRadius() { return hydrodynamicRadius; }
  real getSelfMobility() { return M0; }
};

/* ---------------------------------------------------------------- BD::EulerMaruyama ------------------------- */
/* Integrator (Integrator/BrownianDynamics.cuh:111-126): forwardTime = steps++, interactor forces, position update
   through ub200_bd_euler_maruyama_step (bit-identical positions, tests/test_bd_gpu.py). */
class BDEulerMaruyama : public Integrator {
  real temperature, dt, selfMobility;
  bool is2D;
  uint seed;
  int steps = 0;
  cudaStream_t st;

public:
  struct Parameters {
    real temperature = 0, viscosity = 1, hydrodynamicRadius = -1, dt = 0;
    bool is2D = false;
  };
  BDEulerMaruyama(shared_ptr<ParticleData> pd, Parameters par) : BDEulerMaruyama(std::make_shared<ParticleGroup>(pd, "All"), par) {}
  BDEulerMaruyama(shared_ptr<ParticleGroup> pg, Parameters par)
      : Integrator(pg, "b200::BDEulerMaruyama"), temperature(par.temperature), dt(par.dt), is2D(par.is2D) {
    sys->rng().next32();
    sys->rng().next32();
    seed = sys->rng().next32();
    selfMobility = 1.0 / (6.0 * M_PI * par.viscosity);
    if (par.hydrodynamicRadius != real(-1.0)) selfMobility /= par.hydrodynamicRadius;
    CudaSafeCall(cudaStreamCreate(&st));
  }
  ~BDEulerMaruyama() { cudaStreamDestroy(st); }
  uint getSeed() const { return seed; }
  void forwardTime() override {
    steps++;
    for (auto u : updatables) u->updateSimulationTime(steps * dt);
    const int N = pg->getNumberParticles();
    const real4 *d_force = nullptr;
    if (!interactors.empty()) {
      {
        auto force = pd->getForce(access::location::gpu, access::mode::write);
        auto fg = pg->getPropertyIterator(force);
        thrust::fill(thrust::cuda::par.on(st), fg, fg + N, real4());
      }
      for (auto f : interactors) f->sum({.force = true, .energy = false, .virial = false}, st);
    }
    auto pos = pd->getPos(access::location::gpu, access::mode::readwrite);
    auto force = pd->getForce(access::location::gpu, access::mode::read);
    if (!interactors.empty()) d_force = force.raw();
    const real *radius = nullptr;
    check(ub200_bd_euler_maruyama_step((int)sizeof(real), pos.raw(), pg->getIndicesRawPtr(access::location::gpu), d_force, nullptr,
                                       selfMobility, radius, dt, is2D, temperature, N, (uint)steps, seed, (void *)st),
          "bd_euler_maruyama_step");
  }
};

/* ---------------------------------------------------------------- Poisson (SpectralEwaldPoisson) ------------ */
/* Interactor = Poisson (Interactor/SpectralEwaldPoisson.cuh:84-184): same Parameters fields, sum(Computables) adds q E to the
   forces and q phi to the energies of the group's particles (charges from pd->getCharge), computeFieldPotentialAtParticles()
   returns (Ex, Ey, Ez, phi). Like the reference it works on the whole ParticleData (positions by particle index). */
class Poisson : public Interactor {
  ub200_poisson *handle = nullptr;

public:
  struct Parameters {
    real upsampling = -1.0;
    int3 cells = make_int3(-1, -1, -1); /* accepted and, like in the reference's constructor, not used */
    Box box;
    real epsilon = -1;
    real tolerance = 1e-5;
    real gw = -1;
    int support = -1;                   /* idem */
    real split = -1;
  };
  Poisson(shared_ptr<ParticleData> pd, Parameters par) : Poisson(std::make_shared<ParticleGroup>(pd, "All"), par) {}
  Poisson(shared_ptr<ParticleGroup> pg, Parameters par) : Interactor(pg, "b200::Poisson") {
    ub200_poisson_params p;
    p.L[0] = par.box.boxSize.x; p.L[1] = par.box.boxSize.y; p.L[2] = par.box.boxSize.z;
    p.epsilon = par.epsilon; p.tolerance = par.tolerance; p.gw = par.gw; p.split = par.split; p.upsampling = par.upsampling;
    const int rc = ub200_poisson_create(&handle, (int)sizeof(real), &p);
    if (rc == UB200_ERR_UNSUPPORTED) throw std::invalid_argument("[Poisson] Kernel support is too large");
    if (rc == UB200_ERR_INVALID_ARGUMENT) throw std::invalid_argument("[Poisson] Near field cut off is too large");
    check(rc, "poisson_create");
    ub200_poisson_info_t info;
    check(ub200_poisson_info(handle, &info), "poisson_info");
    System::log<System::MESSAGE>("[b200::Poisson] cells %d %d %d, support %d, near field cut off %g", info.cells[0], info.cells[1],
                                 info.cells[2], info.support, info.nearFieldCutOff);
  }
  Poisson(const Poisson &) = delete;
  ~Poisson() { ub200_poisson_destroy(handle); }

  void sum(Computables comp, cudaStream_t st = 0) override {
    if (comp.virial) {
      System::log<System::EXCEPTION>("[Poisson] Virial functionality not implemented.");
      throw std::runtime_error("[Poisson] not implemented");
    }
    const int N = pg->getNumberParticles();
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    auto charge = pd->getCharge(access::location::gpu, access::mode::read);
    real4 *force = nullptr;
    real *energy = nullptr;
    // the reference's far field interpolates into BOTH arrays on every call (SpectralEwaldPoisson.cu:561-578)
    auto f = pd->getForce(access::location::gpu, access::mode::readwrite);
    auto e = pd->getEnergy(access::location::gpu, access::mode::readwrite);
    force = f.raw();
    energy = e.raw();
    check(ub200_poisson_sum_ex(handle, pos.raw(), charge.raw(), N, force, energy, comp.force, comp.energy, (void *)st), "poisson_sum");
  }
  thrust::device_vector<real4> computeFieldPotentialAtParticles() {
    const int N = pg->getNumberParticles();
    thrust::device_vector<real4> out(N);
    thrust::fill(out.begin(), out.end(), real4());
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    auto charge = pd->getCharge(access::location::gpu, access::mode::read);
    check(ub200_poisson_field_potential(handle, pos.raw(), charge.raw(), N, thrust::raw_pointer_cast(out.data()), nullptr),
          "poisson_field_potential");
    CudaSafeCall(cudaDeviceSynchronize());
    return out;
  }
};

#ifndef DOUBLE_PRECISION
/* ---------------------------------------------------------------- VerletNVT::{GronbechJensen, Basic} -------- */
/* Integrator (Integrator/VerletNVT.cuh:100-117, VerletNVT/Basic.cu:31-77, VerletNVT/GronbechJensen.cu:96-127): the
   Langevin integrator generic_md and examples/misc/benchmark.cu drive. Same constructor side effects as
   VerletNVT::Basic (Saru seed = third next32() of the system generator, optional initial velocities with the fourth),
   same forwardTime sequence; the two half steps run through ub200_nvt_gj_half_step_f32 (bit-identical).
   VerletNVTBasic = VerletNVT::Basic (Basic.cu:87-172), the scheme GronbechJensen derives from: same class, its half steps
   run through ub200_nvt_basic_half_step_f32. (The reference declares Basic's public constructor, VerletNVT.cuh:92, but
   never defines it - a program that constructs VerletNVT::Basic does not link; this one does.) */
struct VerletNVTParameters { /* VerletNVT::Basic::Parameters (VerletNVT.cuh:64-71) */
  real temperature = 0, dt = 0, friction = 1.0;
  bool is2D = false, initVelocities = true;
  real mass = -1.0;
};
template <bool GRONBECH_JENSEN> class VerletNVTScheme : public Integrator {
  real noiseAmplitude, dt, temperature, friction, defaultMass;
  bool is2D;
  uint seed;
  int steps = 0;
  cudaStream_t st;

  void half(int step) {
    const int N = pg->getNumberParticles();
    auto pos = pd->getPos(access::location::gpu, access::mode::readwrite);
    auto vel = pd->getVel(access::location::gpu, access::mode::readwrite);
    auto force = pd->getForce(access::location::gpu, access::mode::readwrite);
    auto mass = pd->getMassIfAllocated(access::location::gpu, access::mode::read).raw();
    auto fn = GRONBECH_JENSEN ? ub200_nvt_gj_half_step_f32 : ub200_nvt_basic_half_step_f32;
    check(fn(pos.raw(), vel.raw(), force.raw(), defaultMass > 0 ? nullptr : mass, defaultMass > 0 ? defaultMass : 0,
             pg->getIndicesRawPtr(access::location::gpu), N, dt, friction, is2D, noiseAmplitude, (uint)steps, seed, step, (void *)st),
          "nvt_half_step");
  }

public:
  using Parameters = VerletNVTParameters;
  VerletNVTScheme(shared_ptr<ParticleData> pd, Parameters par) : VerletNVTScheme(std::make_shared<ParticleGroup>(pd, "All"), par) {}
  VerletNVTScheme(shared_ptr<ParticleGroup> pg, Parameters par)
      : Integrator(pg, GRONBECH_JENSEN ? "b200::VerletNVTGronbechJensen" : "b200::VerletNVTBasic"), dt(par.dt), temperature(par.temperature), friction(par.friction),
        is2D(par.is2D) {
    sys->rng().next32();
    sys->rng().next32();
    seed = sys->rng().next32();
    noiseAmplitude = sqrt(2 * dt * friction * temperature);
    defaultMass = par.mass;
    if (!pd->isMassAllocated() and defaultMass < 0) defaultMass = 1.0;
    CudaSafeCall(cudaStreamCreate(&st));
    if (par.initVelocities) {
      auto vel = pd->getVel(access::location::gpu, access::mode::write);
      const real velAmplitude = sqrt(3.0 * temperature);
      check(ub200_nvt_initial_velocities_f32(vel.raw(), pg->getIndicesRawPtr(access::location::gpu), pg->getNumberParticles(),
                                             velAmplitude, is2D, sys->rng().next32(), nullptr),
            "nvt_initial_velocities");
      CudaSafeCall(cudaDeviceSynchronize());
    }
  }
  ~VerletNVTScheme() { cudaStreamDestroy(st); }
  uint getSeed() const { return seed; }
  void forwardTime() override {
    for (auto u : updatables) u->updateSimulationTime(steps * dt);
    steps++;
    if (steps == 1) {
      {
        auto force = pd->getForce(access::location::gpu, access::mode::write);
        auto fg = pg->getPropertyIterator(force);
        thrust::fill(thrust::cuda::par.on(st), fg, fg + pg->getNumberParticles(), real4());
      }
      for (auto u : updatables) {
        u->updateTemperature(temperature);
        u->updateTimeStep(dt);
      }
      for (auto f : interactors) f->sum({.force = true, .energy = false, .virial = false}, st);
      CudaSafeCall(cudaDeviceSynchronize());
    }
    half(1);
    for (auto f : interactors) f->sum({.force = true, .energy = false, .virial = false}, st);
    half(2);
  }
};
using VerletNVTGronbechJensen = VerletNVTScheme<true>;
using VerletNVTBasic = VerletNVTScheme<false>;
#endif /* !DOUBLE_PRECISION */

} // namespace b200
} // namespace uammd
#endif
