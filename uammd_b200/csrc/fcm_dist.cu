// Slab-decomposed Force Coupling Method over the GPUs of one NVSwitch box (one process per GPU), sm_100a.
//
// The reference is single-GPU; the oracle of this path is the single-GPU result on the same input,
// which it reproduces bit for bit (same kernels, same per-node summation order, same noise stream keyed on the
// GLOBAL Fourier index).
//
// Rank r owns the z planes [r nzl, (r+1) nzl) of the grid (buffer S) and the particles whose cell lies in them;
// positions / forces are replicated (every rank receives the same arrays), so spreading needs no communication:
// a rank bins the particles of its planes plus one support of halo cells and writes its planes once.
//   bin (window) + scan + scatter + order   local
//   spread bricks -> S                       local, atomic-free
//   FFT x, FFT y                             local lines; the y pass STORES its output into the owners' transposed
//                                            buffers T over NVLink (the all-to-all is the pass's store, no NCCL, no pack)
//   --- barrier ---
//   fused [FFT z, Stokes, noise, iFFT z]     on T = [nz][nyl][nkx] (ky rows of this rank); stores go straight back
//                                            into the owners' S
//                                            into the owners' S - and the boundary planes ALSO into the neighbours' halo planes
//   --- barrier ---
//   iFFT y, iFFT x                           local, on the owned planes plus the halo planes (a few redundant lines
//                                            instead of a halo exchange and its barrier)
//   gather                                   owned particles, all loads local; rows packed in sorted-slot order
//   push                                     the packed {index, row} block goes to every rank's inbox (coalesced stores)
//   --- barrier ---
//   scatter                                  out[index] = row for everything received
// Buffers S, T, the result and the barrier flags of a rank live in ONE cudaMalloc allocation exported to the peers
// through CUDA IPC (ub200_fcm_dist_ipc_export / _import; the caller moves the 64-byte handles, e.g. with
// torch.distributed.all_gather_object). Barriers are one-block kernels that publish an epoch to every peer's flag
// array with system-scope release stores and spin (bounded) on their own array.
#include "fcm_op.cuh"
#include "pse_op.cuh"
#include "fft3d.cuh"
#include "ibm_state.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace ub200 {

constexpr unsigned long long kBarrierSpinLimit = 400000000ull; // ~ seconds; a lost peer must not hang the GPU forever

// flags[p] of rank r = last epoch rank p announced to r. err: set when the spin limit is hit.
// epoch = *epochBase + offset: the base is uploaded before every call, so that the captured graph of a call replays
__global__ void __launch_bounds__(32) peerBarrier(PeerTable<uint32_t> flags, int rank, int world, const uint32_t *epochBase,
                                                  uint32_t offset, int *err) {
  const int p = threadIdx.x;
  if (p >= world) return;
  const uint32_t epoch = *epochBase + offset;
  __threadfence_system();
  volatile uint32_t *remote = flags.p[p] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t *mine = flags.p[rank] + p;
  unsigned long long spins = 0;
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if (++spins > kBarrierSpinLimit) { *err = 1; break; }
  }
  __threadfence_system();
}

template <class T> struct FcmDistState {
  using C = typename Vec2<T>::type;
  using T4 = typename Real4<T>::type;
  int rank = 0, world = 1;
  int nzl = 0, nyl = 0, z0 = 0, y0 = 0, halo = 0;
  int maxParticles = 0;
  Fft3dPlan<T> plan;
  GridT<T> grid;      // with the z window of this rank
  IbmKernel<T> kern;
  double viscosity = 1, L[3];
  uint32_t seed = 0, seed2 = 0;
  bool pseOperator = false; // spectral operator: FCM Stokes (default) or the PSE far-field Green's function
  double pseRh = 0, psePsi = 0, pseEta = 0, pseShear = 0;
  // one exported allocation: [flags | S (owned planes + 2 halo) | T | inboxes: world x {count, index[], rows[]}]
  void *arena = nullptr;
  size_t arenaBytes = 0, offS = 0, offT = 0, offInbox = 0, inboxStride = 0, inboxIdxOff = 0, inboxRowsOff = 0;
  DevBuf packIdx, packRows, packCount;
  void *peerArena[kMaxPeers] = {};
  bool imported = false;
  uint32_t epoch = 0;
  DevBuf errFlag;
  DevBuf callVars;            // device {epoch base of this call, noise seed2 of this call}
  uint32_t callVarsHost[2] = {0, 0};
  int barriersThisCall = 0;
  // CUDA graph of one call (19 launches), replayed while the arguments stay the same; runs on an internal stream joined to
  // the caller's by events (the legacy default stream cannot be captured). UB200_DIST_GRAPH=0 disables it.
  bool useGraph = true;
  cudaStream_t gs = nullptr;
  cudaEvent_t evIn = nullptr, evOut = nullptr;
  cudaGraphExec_t exec = nullptr;
  struct CallKey {
    const void *pos, *force, *out;
    int N;
    double temperature, prefactor;
    bool operator==(const CallKey &o) const {
      return pos == o.pos && force == o.force && out == o.out && N == o.N && temperature == o.temperature && prefactor == o.prefactor;
    }
  } key = {nullptr, nullptr, nullptr, 0, 0.0, 0.0};
  int callsWithKey = 0;
  // optional phase timing (UB200_DIST_PROFILE=1): events between the phases of mdot, summed on the host
  static constexpr int kPhases = 12;
  bool profile = false;
  cudaEvent_t ev[kPhases + 1] = {};
  double phaseMs[kPhases] = {};
  int profiledCalls = 0;
  // particle scratch (window)
  DevBuf binCount, binStart, tileSums, codeSlot, unstable, sortedIndex, sortedRec;

  int recWords() const {
    switch (kern.support) {
    case 3: return RecGeom<T, 3>::REC;
    case 4: return RecGeom<T, 4>::REC;
    case 5: return RecGeom<T, 5>::REC;
    default: return RecGeom<T, 7>::REC;
    }
  }
  size_t planeBytes() const { return (size_t)plan.ny * plan.nkx * 3 * sizeof(C); }
  size_t slabBytes() const { return (size_t)(nzl + 2 * halo) * planeBytes(); }
  size_t tposeBytes() const { return (size_t)plan.nz * nyl * plan.nkx * 3 * sizeof(C); }
  template <class U> U *at(void *base, size_t off) const { return reinterpret_cast<U *>(static_cast<char *>(base) + off); }

  int init(const double L_[3], const int cells[3], const ub200_ibm_kernel &k, double vis, uint32_t seed_, int rank_, int world_,
           int maxParticles_) {
    rank = rank_; world = world_; maxParticles = maxParticles_;
    { const char *g = getenv("UB200_DIST_GRAPH"); useGraph = !(g && g[0] == '0'); }
    if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world || maxParticles < 1) return UB200_ERR_INVALID_ARGUMENT;
    if (cells[2] % world || cells[1] % world) return UB200_ERR_INVALID_ARGUMENT; // equal slabs in z and in ky
    if (k.support != 3 && k.support != 4 && k.support != 5 && k.support != 7) return UB200_ERR_UNSUPPORTED; // row-brick spread
    int rc = plan.init(cells[0], cells[1], cells[2]);
    if (rc) return rc;
    nzl = cells[2] / world; nyl = cells[1] / world; z0 = rank * nzl; y0 = rank * nyl;
    const int periodic[3] = {1, 1, 1};
    grid = makeGridT<T>(L_, periodic, cells);
    const int S = k.support;
    const int hi = S / 2, lo = -(S - 1) + S / 2 - ((S & 1) ? 0 : 1);
    if (nzl + (hi - lo) > cells[2]) return UB200_ERR_INVALID_ARGUMENT;
    if (cells[0] < kRbX + S || cells[1] < kRbY + S || cells[2] < kRbZ + S) return UB200_ERR_INVALID_ARGUMENT; // row-brick spread
    grid.zwin0 = ((z0 + lo) % cells[2] + cells[2]) % cells[2];
    grid.zwinN = nzl + (hi - lo);
    halo = std::max(-lo, hi);
    if (halo > nzl) return UB200_ERR_INVALID_ARGUMENT;
    kern.kind = k.kind; kern.support = k.support; kern.invh = (T)(1.0 / k.h);
    kern.prefactor = (T)k.prefactor; kern.tau = (T)k.tau; kern.rmax = (T)k.rmax;
    for (int d = 0; d < 3; d++) L[d] = L_[d];
    viscosity = vis; seed = seed_;
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    offS = align(4096);
    offT = offS + align(slabBytes());
    offInbox = offT + align(tposeBytes());
    inboxIdxOff = 256;
    inboxRowsOff = inboxIdxOff + align(sizeof(int) * (size_t)maxParticles);
    inboxStride = inboxRowsOff + align(sizeof(T) * 3 * (size_t)maxParticles);
    arenaBytes = offInbox + inboxStride * world;
    if ((rc = packIdx.reserve(sizeof(int) * (size_t)maxParticles)) || (rc = packRows.reserve(sizeof(T) * 3 * (size_t)maxParticles)) ||
        (rc = packCount.reserve(sizeof(int))))
      return rc;
    if (cudaMalloc(&arena, arenaBytes) != cudaSuccess) return UB200_ERR_ALLOC;
    UB200_CUDA(cudaMemset(arena, 0, arenaBytes));
    if ((rc = errFlag.reserve(sizeof(int)))) return rc;
    UB200_CUDA(cudaMemset(errFlag.p, 0, sizeof(int)));
    const int ncw = cells[0] * cells[1] * grid.zwinN;
    if ((rc = binCount.reserve(sizeof(uint32_t) * (size_t)ncw)) || (rc = binStart.reserve(sizeof(uint32_t) * ((size_t)ncw + 1))) ||
        (rc = tileSums.reserve(sizeof(uint32_t) * 4096)))
      return rc;
    UB200_CUDA(cudaMemset(binCount.p, 0, binCount.cap));
    UB200_CUDA(cudaDeviceSynchronize());
    peerArena[rank] = arena;
    profile = getenv("UB200_DIST_PROFILE") != nullptr;
    if (profile) for (auto &e : ev) cudaEventCreate(&e);
    return UB200_OK;
  }
  void mark(int k, cudaStream_t st) { if (profile) cudaEventRecord(ev[k], st); }
  void collect(int nmarks, cudaStream_t st) {
    if (!profile) return;
    cudaStreamSynchronize(st);
    for (int k = 0; k + 1 < nmarks; k++) { float ms = 0; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); phaseMs[k] += ms; }
    profiledCalls++;
  }
  void release() {
    for (int p = 0; p < world; p++)
      if (p != rank && peerArena[p]) cudaIpcCloseMemHandle(peerArena[p]);
    if (arena) cudaFree(arena);
    arena = nullptr;
    if (exec) cudaGraphExecDestroy(exec);
    if (gs) { cudaStreamDestroy(gs); cudaEventDestroy(evIn); cudaEventDestroy(evOut); }
    DevBuf *b[] = {&callVars, &errFlag, &packIdx, &packRows, &packCount, &binCount, &binStart, &tileSums, &codeSlot, &unstable, &sortedIndex, &sortedRec};
    for (auto *x : b) x->release();
    plan.release();
  }
  int ipcExport(void *blob) {
    cudaIpcMemHandle_t h;
    UB200_CUDA(cudaIpcGetMemHandle(&h, arena));
    memcpy(blob, &h, sizeof(h));
    return UB200_OK;
  }
  int ipcImport(const void *blobs) {
    for (int p = 0; p < world; p++) {
      if (p == rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const char *>(blobs) + (size_t)p * sizeof(h), sizeof(h));
      UB200_CUDA(cudaIpcOpenMemHandle(&peerArena[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    imported = true;
    return UB200_OK;
  }

  int barrier(cudaStream_t st) {
    PeerTable<uint32_t> flags;
    for (int p = 0; p < world; p++) flags.p[p] = at<uint32_t>(peerArena[p], 0);
    barriersThisCall++;
    peerBarrier<<<1, 32, 0, st>>>(flags, rank, world, callVars.as<uint32_t>(), (uint32_t)barriersThisCall, errFlag.as<int>());
    UB200_LAUNCHED();
    return UB200_OK;
  }

  int mdot(const void *pos, const void *force, int N, double temperature, double prefactor, void *out3, cudaStream_t st) {
    if (!imported) return UB200_ERR_NOT_BUILT;
    if (N > maxParticles) return UB200_ERR_INVALID_ARGUMENT;
    int rc;
    if ((rc = callVars.reserve(sizeof(uint32_t) * 2))) return rc;
    // per-call variables go to the device first: barrier epochs of this call = base + 1, 2, 3; the noise seed
    const bool noisy = temperature > 0.0;
    if (noisy && !pseOperator) seed2++; // fourierBrownianNoise's call counter (FCM_impl.cuh:517,523)
    callVarsHost[0] = epoch;
    callVarsHost[1] = seed2;
    epoch += 3;
    const CallKey k = {pos, force, out3, N, temperature, prefactor};
    if (!useGraph || profile) {
      UB200_CUDA(cudaMemcpyAsync(callVars.p, callVarsHost, sizeof(callVarsHost), cudaMemcpyHostToDevice, st));
      return enqueue(pos, force, N, temperature, prefactor, out3, st);
    }
    if (!gs) {
      UB200_CUDA(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
      UB200_CUDA(cudaEventCreateWithFlags(&evIn, cudaEventDisableTiming));
      UB200_CUDA(cudaEventCreateWithFlags(&evOut, cudaEventDisableTiming));
    }
    if (!(k == key)) {
      if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
      key = k;
      callsWithKey = 0;
    }
    callsWithKey++;
    // the graph is CAPTURED on the internal stream (the caller's may be the legacy default stream, which cannot capture) and
    // LAUNCHED on the caller's stream
    UB200_CUDA(cudaMemcpyAsync(callVars.p, callVarsHost, sizeof(callVarsHost), cudaMemcpyHostToDevice, st));
    if (callsWithKey == 1) { // first call with these arguments: plain launches (scratch allocation, function attributes)
      rc = enqueue(pos, force, N, temperature, prefactor, out3, st);
    } else {
      if (!exec) {
        cudaGraph_t graph = nullptr;
        UB200_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed));
        rc = enqueue(pos, force, N, temperature, prefactor, out3, gs);
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (rc || ce != cudaSuccess || !graph) {
          if (graph) cudaGraphDestroy(graph);
          useGraph = false;
          if (!rc) rc = enqueue(pos, force, N, temperature, prefactor, out3, st);
        } else {
          const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
          cudaGraphDestroy(graph);
          if (ie != cudaSuccess) return cudaFail(ie);
        }
      }
      if (exec) {
        UB200_CUDA(cudaGraphLaunch(exec, st));
        g_launchCount += 19;
        rc = UB200_OK;
      }
    }
    return rc;
  }

  // the launches of one call; no host state changes, so that the sequence can be captured into a CUDA graph
  int enqueue(const void *pos, const void *force, int N, double temperature, double prefactor, void *out3, cudaStream_t st) {
    int rc;
    barriersThisCall = 0;
    T *Sall = at<T>(arena, offS);                                         // first halo plane
    T *S = reinterpret_cast<T *>(reinterpret_cast<char *>(Sall) + (size_t)halo * planeBytes()); // first owned plane
    C *Tb = at<C>(arena, offT);
    const bool det = force != nullptr;
    const int nb = (N + 255) / 256;
    const int ncw = grid.n[0] * grid.n[1] * grid.zwinN;
    // ---- particles of the window: bin, scan, scatter, order + stencil records ----
    if ((rc = codeSlot.reserve(sizeof(uint2) * (size_t)N)) || (rc = unstable.reserve(sizeof(int) * (size_t)N)) ||
        (rc = sortedIndex.reserve(sizeof(int) * (size_t)N)) ||
        (rc = sortedRec.reserve(sizeof(T) * recWords() * (size_t)N)))
      return rc;
    mark(0, st);
    ibmBinByCell<T4><<<nb, 256, 0, st>>>((const T4 *)pos, N, grid, binCount.as<uint32_t>(), codeSlot.as<uint2>());
    UB200_LAUNCHED();
    if ((rc = exclusiveScanAndClear(binCount.as<uint32_t>(), ncw, binStart.as<uint32_t>(), tileSums.as<uint32_t>(), st))) return rc;
    if ((rc = scatterToBinsLaunch(codeSlot.as<uint2>(), binStart.as<uint32_t>(), N, unstable.as<int>(), st))) return rc;
#define UB200_ORDER(SS)                                                                                                  \
  ibmOrderSorted<T4, SS><<<nb, 256, 0, st>>>(unstable.as<int>(), codeSlot.as<uint2>(), binStart.as<uint32_t>(), (const T4 *)pos, \
                                             (const T *)force, 4, N, grid, kern, sortedIndex.as<int>(), sortedRec.as<T>())
    switch (kern.support) {
    case 3: UB200_ORDER(3); break;
    case 4: UB200_ORDER(4); break;
    case 5: UB200_ORDER(5); break;
    default: UB200_ORDER(7); break;
    }
#undef UB200_ORDER
    UB200_LAUNCHED();
    mark(1, st);
    AddrSlabZFused<C> az;
    az.T = Tb; az.ny = plan.ny; az.nyl = nyl; az.nkx = plan.nkx; az.nzl = nzl; az.y0 = y0; az.halo = halo; az.world = world;
    for (int p = 0; p < world; p++) az.peerS[p] = at<C>(peerArena[p], offS);
    if (det) {
      // ---- spread into the owned planes ----
      dim3 grd((plan.nxPad + kRbX - 1) / kRbX, (grid.n[1] + kRbY - 1) / kRbY, (nzl + kRbZ - 1) / kRbZ);
#define UB200_SPREAD(SS)                                                                                                 \
  {                                                                                                                      \
    auto kfn = ibmSpreadRows<T, SS>;                                                                                     \
    const size_t sm = RowBrickGeom<T, SS>::smemBytes;                                                                    \
    UB200_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                         \
    kfn<<<grd, kRbThreads, sm, st>>>(sortedRec.as<T>(), binStart.as<uint32_t>(), grid, plan.nxPad, S, z0, nzl);          \
  }
      switch (kern.support) {
      case 3: UB200_SPREAD(3) break;
      case 4: UB200_SPREAD(4) break;
      case 5: UB200_SPREAD(5) break;
      default: UB200_SPREAD(7) break;
      }
#undef UB200_SPREAD
      UB200_LAUNCHED();
      mark(2, st);
      if ((rc = launchPassX<T, true>(plan, S, st, nzl))) return rc;
      mark(3, st);
      // ---- forward y pass, output scattered into the owners' transposed buffers (NVLink stores) ----
      AddrSlabYForward<C> ay;
      ay.S = reinterpret_cast<C *>(S); ay.ny = plan.ny; ay.nyl = nyl; ay.nkx = plan.nkx; ay.z0 = z0;
      for (int p = 0; p < world; p++) ay.peerT[p] = at<C>(peerArena[p], offT);
      if ((rc = launchPassAddr<T, -1, false, NoSpectralOp>(plan, ay, nzl, st))) return rc;
    } else {
      UB200_CUDA(cudaMemsetAsync(Tb, 0, tposeBytes(), st));
    }
    mark(4, st);
    if ((rc = barrier(st))) return rc;
    mark(5, st);
    // ---- fused z pass on the local ky rows; output pushed back into the owners' slabs ----
    if (!pseOperator) {
      FcmSpectralOp<T> op;
      op.nx = plan.nx; op.ny = plan.ny; op.nz = plan.nz; op.nkx = plan.nkx;
      op.kfx = (T)(T(2.0) * T(M_PI) / (T)L[0]); op.kfy = (T)(T(2.0) * T(M_PI) / (T)L[1]); op.kfz = (T)(T(2.0) * T(M_PI) / (T)L[2]);
      op.vis = (T)viscosity;
      op.invNorm = T(1.0) / T((double)plan.nx * plan.ny * plan.nz);
      op.deterministic = det;
      op.noise = temperature > 0.0;
      op.noisePrefactor = T(0);
      op.seed1 = seed; op.seed2 = seed2;
      op.seed2Dev = callVars.as<uint32_t>() + 1;
      op.yOff = y0;
      if (op.noise) {
        const T fourierNormalization = (T)(1.0 / ((double)plan.nx * plan.ny * plan.nz));
        op.noisePrefactor = (T)prefactor * (T)sqrt((double)(fourierNormalization * 2 * (T)temperature / grid.cellVolume));
      }
      if ((rc = launchPassAddr<T, 0, true, FcmSpectralOp<T>>(plan, az, nyl, st, op))) return rc;
    } else { // PseState::farMdot
      PseSpectralOp<T> op;
      op.nx = plan.nx; op.ny = plan.ny; op.nz = plan.nz; op.nkx = plan.nkx;
      op.kfx = T(2.0) * T(M_PI) / (T)L[0]; op.kfy = T(2.0) * T(M_PI) / (T)L[1]; op.kfz = T(2.0) * T(M_PI) / (T)L[2];
      op.shear = (T)pseShear; op.rh = (T)pseRh; op.vis = (T)viscosity; op.split = (T)psePsi; op.eta = (T)pseEta;
      op.nTot = (T)(plan.nx * plan.ny * plan.nz);
      op.deterministic = det;
      op.noise = temperature > 0.0;
      op.seed1 = seed; op.seed2 = seed2; // seed2: set by the caller before every noisy call (ub200_fcm_dist_set_noise_seed2)
      op.seed2Dev = callVars.as<uint32_t>() + 1;
      op.yOff = y0;
      op.noisePrefactor = op.noise ? (T)prefactor * (T)sqrt(2 * (T)temperature / grid.cellVolume) : T(0);
      if ((rc = launchPassAddr<T, 0, true, PseSpectralOp<T>>(plan, az, nyl, st, op))) return rc;
    }
    mark(6, st);
    if ((rc = barrier(st))) return rc;
    mark(7, st);
    // ---- inverse y and x passes on the owned planes AND the halo planes the neighbours pushed ----
    AddrInPlace<C> ainv;
    ainv.grid = reinterpret_cast<C *>(Sall); ainv.elemStride = (size_t)plan.nkx; ainv.otherStride = (size_t)plan.nkx * plan.ny;
    if ((rc = launchPassAddr<T, +1, false, NoSpectralOp>(plan, ainv, nzl + 2 * halo, st))) return rc;
    if ((rc = launchPassX<T, false>(plan, Sall, st, nzl + 2 * halo))) return rc;
    mark(8, st);
    // ---- gather for the owned particles (local loads), packed rows ----
    const int ngb = (N + 127) / 128;
#define UB200_GSLAB(SS)                                                                                                       \
  ibmGatherSortedSlab<T, SS><<<ngb, 128, 0, st>>>(sortedRec.as<T>(), sortedIndex.as<int>(), binStart.as<uint32_t>(), grid, plan.nxPad, \
                                                  Sall, z0, nzl, halo, packIdx.as<int>(), packRows.as<T>(), packCount.as<int>())
    switch (kern.support) {
    case 3: UB200_GSLAB(3); break;
    case 4: UB200_GSLAB(4); break;
    case 5: UB200_GSLAB(5); break;
    default: UB200_GSLAB(7); break;
    }
#undef UB200_GSLAB
    UB200_LAUNCHED();
    mark(9, st);
    // ---- push the packed block into inbox[rank] of every rank ----
    PeerTable<int> ibCount, ibIdx;
    PeerTable<T> ibRows;
    for (int p = 0; p < world; p++) {
      char *box = at<char>(peerArena[p], offInbox) + inboxStride * (size_t)rank;
      ibCount.p[p] = reinterpret_cast<int *>(box);
      ibIdx.p[p] = reinterpret_cast<int *>(box + inboxIdxOff);
      ibRows.p[p] = reinterpret_cast<T *>(box + inboxRowsOff);
    }
    slabPushPacked<T><<<dim3(96, world), 256, 0, st>>>(packCount.as<int>(), packIdx.as<int>(), packRows.as<T>(), ibCount, ibIdx, ibRows, world);
    UB200_LAUNCHED();
    mark(10, st);
    if ((rc = barrier(st))) return rc; // every rank's block has landed; slabs and T may be overwritten by the next call
    mark(11, st);
    // ---- scatter what the ranks sent me ----
    for (int p = 0; p < world; p++) {
      char *box = at<char>(arena, offInbox) + inboxStride * (size_t)p;
      ibCount.p[p] = reinterpret_cast<int *>(box);
      ibIdx.p[p] = reinterpret_cast<int *>(box + inboxIdxOff);
      ibRows.p[p] = reinterpret_cast<T *>(box + inboxRowsOff);
    }
    slabScatterInbox<T><<<dim3(64, world), 256, 0, st>>>(ibCount, ibIdx, ibRows, world, (T *)out3);
    UB200_LAUNCHED();
    mark(12, st);
    collect(13, st);
    return UB200_OK;
  }
};

} // namespace ub200

using namespace ub200;

struct ub200_fcm_dist {
  int precision;
  FcmDistState<float> f;
  FcmDistState<double> d;
};

extern "C" {

int ub200_fcm_dist_create(ub200_fcm_dist **out, int precisionBytes, const double L[3], const int cells[3],
                          const ub200_ibm_kernel *kernel, double viscosity, uint32_t seed, int rank, int world, int maxParticles) {
  if (!out || !L || !cells || !kernel || (precisionBytes != 4 && precisionBytes != 8)) return UB200_ERR_INVALID_ARGUMENT;
  ub200_fcm_dist *h = new (std::nothrow) ub200_fcm_dist();
  if (!h) return UB200_ERR_ALLOC;
  h->precision = precisionBytes;
  const int rc = precisionBytes == 4 ? h->f.init(L, cells, *kernel, viscosity, seed, rank, world, maxParticles)
                                     : h->d.init(L, cells, *kernel, viscosity, seed, rank, world, maxParticles);
  if (rc) { h->f.release(); h->d.release(); delete h; return rc; }
  *out = h;
  return UB200_OK;
}
int ub200_fcm_dist_destroy(ub200_fcm_dist *h) {
  if (!h) return UB200_OK;
  h->f.release(); h->d.release();
  delete h;
  return UB200_OK;
}
/* UB200_DIST_PROFILE=1: mean milliseconds of the 12 phases of mdot (sort, spread, fft x, fft y + transpose, barrier,
 * fused z + transpose, barrier, ifft y+x, gather, push, barrier, scatter); returns the number of profiled calls */
int ub200_fcm_dist_profile(ub200_fcm_dist *h, double phases[12]) {
  if (!h || !phases) return 0;
  const int n = h->precision == 4 ? h->f.profiledCalls : h->d.profiledCalls;
  for (int k = 0; k < 12; k++) phases[k] = n ? (h->precision == 4 ? h->f.phaseMs[k] : h->d.phaseMs[k]) / n : 0.0;
  return n;
}
int ub200_fcm_dist_ipc_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }
int ub200_fcm_dist_ipc_export(ub200_fcm_dist *h, void *blob) {
  if (!h || !blob) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? h->f.ipcExport(blob) : h->d.ipcExport(blob);
}
int ub200_fcm_dist_ipc_import(ub200_fcm_dist *h, const void *blobs) {
  if (!h || !blobs) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? h->f.ipcImport(blobs) : h->d.ipcImport(blobs);
}
int ub200_fcm_dist_mdot(ub200_fcm_dist *h, const void *d_pos, const void *d_force, int N, double temperature, double prefactor,
                        void *d_out3, void *stream) {
  if (!h || !d_pos || !d_out3 || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  return h->precision == 4 ? h->f.mdot(d_pos, d_force, N, temperature, prefactor, d_out3, (cudaStream_t)stream)
                           : h->d.mdot(d_pos, d_force, N, temperature, prefactor, d_out3, (cudaStream_t)stream);
}
int ub200_fcm_dist_set_pse_operator(ub200_fcm_dist *h, double hydrodynamicRadius, double psi, double eta, double shearStrain) {
  if (!h || !(hydrodynamicRadius > 0) || !(psi > 0)) return UB200_ERR_INVALID_ARGUMENT;
  auto set = [&](auto &s) { s.pseOperator = true; s.pseRh = hydrodynamicRadius; s.psePsi = psi; s.pseEta = eta; s.pseShear = shearStrain; };
  set(h->f); set(h->d);
  return UB200_OK;
}
int ub200_fcm_dist_set_noise_seed2(ub200_fcm_dist *h, uint32_t seed2) {
  if (!h) return UB200_ERR_INVALID_ARGUMENT;
  h->f.seed2 = seed2; h->d.seed2 = seed2;
  return UB200_OK;
}
/* reads back (synchronising the stream) whether a peer barrier ever timed out */
int ub200_fcm_dist_error_flag(ub200_fcm_dist *h, void *stream, int *flag) {
  if (!h || !flag) return UB200_ERR_INVALID_ARGUMENT;
  void *p = h->precision == 4 ? h->f.errFlag.p : h->d.errFlag.p;
  UB200_CUDA(cudaMemcpyAsync(flag, p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  UB200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return UB200_OK;
}
}
