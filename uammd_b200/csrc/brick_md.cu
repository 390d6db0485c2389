// Brick domain decomposition of the pair path over the GPUs of one NVSwitch box (one process per GPU), sm_100a:
// device-resident particle state, ONE halo exchange per step made of peer-to-peer stores, no host round trip.
// ("3-D brick domain decomposition, owned particles + ghost shell of width rc, full-neighbour scheme => no
// reverse force communication, migration of particles that leave the brick"; BASELINE config 4 "ghost-cell halo exchange".)
// The reference is single-GPU; the oracle of this path is the single-GPU trajectory, which it reproduces bit for bit.
//
// Decomposition. The engine's half-cell grid (lj_column.cu: edge >= cutOff / 2) is cut into px x py x pz bricks of whole
// half cells: brick k of a dimension with g cells holds [floor(k g / p), floor((k + 1) g / p)). A rank owns the
// particles whose half cell - computed with exactly the arithmetic of the list build - lies in its brick, and needs as
// ghosts the particles of the two layers of half cells around it (>= cutOff). Its list is built on that WINDOW of the
// global grid (colgeom.h), so list build and traversal cost scale with the brick, not with the box.
//
// One exchange per step (brickAdvancePush + brickUnpack). After the drift, the CURRENT holder of a particle works out its
// new owner and every rank whose window contains it, and stores the 32-byte row {pos, vel, id} straight into those ranks'
// inboxes over NVLink (slots from warp-aggregated counters): migration and ghost distribution are the same pass, the
// holder distributing ghosts on behalf of the new owner. The last block of the kernel publishes the row counts and
// raises the rank's flag in every peer (system-scope release); the unpack kernel of the receiver waits on its flags
// (acquire), appends migrants to the owned block and ghosts behind it, and leaves the counts ON THE DEVICE: every later
// kernel of the step (list build, traversal, kick) reads them there. Inboxes are double buffered by step parity, which
// makes one flag wait per step sufficient. Arrays are exchanged once through CUDA IPC (or attached directly when several
// virtual ranks live in one process: tests).
//
// Order independence. Atomic slots make the order of the local arrays arbitrary. The list build therefore orders the
// particles of a cell by GLOBAL id (fineOrder's sort key), which is the single-GPU order: every owned particle sums the
// same pairs in the same order with the same image shifts as on one GPU, whatever the decomposition.
#include "lj_engine.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ub200 {

constexpr int kBrickMaxRanks = 8;
constexpr unsigned long long kBrickSpinLimit = 400000000ull; // a lost peer must not hang the GPU forever
constexpr int kHdrBytes = 256;

struct BrickGeom {
  GridF g;      // global half-cell grid (canonical coordinates, same arithmetic as the list build)
  int dims[3];  // global half cells
  int rg[3];    // rank grid
  int per[3];   // global periodicity
  int me, world;
};
struct BrickArena {
  char *p[kBrickMaxRanks];
  size_t inboxOff, segBytes, migBytes; // segment (parity, src): header | migrant rows | ghost rows
  int capMig, capGhost;
};

__host__ __device__ __forceinline__ int brickOf(int c, int n, int p) { return ((c + 1) * p - 1) / n; }
__host__ __device__ __forceinline__ int brickLo(int k, int n, int p) { return (k * n) / p; }

__device__ __forceinline__ uint32_t *arenaFlags(char *arena) { return reinterpret_cast<uint32_t *>(arena); }
__device__ __forceinline__ char *arenaSeg(const BrickArena &ar, int rank, int parity, int src, int world) {
  return ar.p[rank] + ar.inboxOff + (size_t)(parity * world + src) * ar.segBytes;
}

// canonical half cell of a coordinate (lj_column.cu canonicalCoord, cell only)
__device__ __forceinline__ int halfCell(float r, float L, float m, float hL, float inv, int n) {
  const float rf = foldCoord(r, L, m);
  int c = __float2int_rz(__fmul_rn(__fadd_rn(rf, hL), inv));
  if (m != 0.0f) {
    if (c >= n) c -= n;
    else if (c < 0) c += n;
  }
  return min(max(c, 0), n - 1);
}

// is global cell c inside the window (brick k +- 2 cells) of a dimension with n cells cut into p bricks?
__device__ __forceinline__ bool inWindow(int c, int k, int n, int p, int per) {
  if (p == 1) return true;
  const int lo = brickLo(k, n, p) - 2, hi = brickLo(k + 1, n, p) + 2;
  if (c >= lo && c < hi) return true;
  if (per && ((c + n >= lo && c + n < hi) || (c - n >= lo && c - n < hi))) return true;
  return false;
}
// bricks of one dimension whose window holds cell c (the own brick k and its neighbours)
__device__ __forceinline__ int windowsOf(int c, int k, int n, int p, int per, int out[3]) {
  int nc = 0;
  for (int d = -1; d <= 1; d++) {
    int kk = k + d;
    if (kk < 0) { if (!per) continue; kk += p; }
    else if (kk >= p) { if (!per) continue; kk -= p; }
    if (kk < 0 || kk >= p) continue;
    bool dup = false;
    for (int j = 0; j < nc; j++) dup |= out[j] == kk;
    if (!dup && inWindow(c, kk, n, p, per)) out[nc++] = kk;
  }
  return nc;
}

struct BrickClass {
  int owner;
  uint32_t mask; // ranks other than the owner whose window holds the particle
};
__device__ __forceinline__ BrickClass classify(const BrickGeom &b, float x, float y, float z) {
  const GridF &g = b.g;
  const int cx = halfCell(x, g.Lx, g.mx, g.hLx, g.ix, g.nx), cy = halfCell(y, g.Ly, g.my, g.hLy, g.iy, g.ny),
            cz = halfCell(z, g.Lz, g.mz, g.hLz, g.iz, g.nz);
  const int kx = brickOf(cx, b.dims[0], b.rg[0]), ky = brickOf(cy, b.dims[1], b.rg[1]), kz = brickOf(cz, b.dims[2], b.rg[2]);
  BrickClass c;
  c.owner = kx + b.rg[0] * (ky + b.rg[1] * kz);
  int wx[3], wy[3], wz[3];
  const int nx = windowsOf(cx, kx, b.dims[0], b.rg[0], b.per[0], wx), ny = windowsOf(cy, ky, b.dims[1], b.rg[1], b.per[1], wy),
            nz = windowsOf(cz, kz, b.dims[2], b.rg[2], b.per[2], wz);
  c.mask = 0;
  for (int a = 0; a < nz; a++)
    for (int e = 0; e < ny; e++)
      for (int f = 0; f < nx; f++) c.mask |= 1u << (wx[f] + b.rg[0] * (wy[e] + b.rg[1] * wz[a]));
  c.mask &= ~(1u << c.owner);
  return c;
}

// slot for every flagged lane of the warp from one atomic per warp
__device__ __forceinline__ int warpSlots(bool flag, int *counter, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (!m) return -1;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return flag ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ void storeRow(char *rows, int slot, const float4 &p, float vx, float vy, float vz, int id) {
  float4 *dst = reinterpret_cast<float4 *>(rows) + 2 * (size_t)slot;
  dst[0] = p;
  dst[1] = make_float4(vx, vy, vz, __int_as_float(id));
}

// work: [0] kept (new owned block so far), [1] block tickets, [2 + 2 r] migrants for rank r, [3 + 2 r] ghosts for rank r
constexpr int kWorkInts = 2 + 2 * kBrickMaxRanks;

// (optional first kick + drift of velocity Verlet) + classification + push of one rank's owned block.
// The kick and the drift spell out the roundings of VerletNVE_ns::integrateGPU<1> (Integrator/VerletNVE.cu:64-85) exactly
// like nveHalfStep<1> (nve.cu), unit mass.
__global__ void __launch_bounds__(256)
brickAdvancePush(BrickGeom b, BrickArena ar, const float4 *__restrict__ pos,
                 const float *__restrict__ vel, const int *__restrict__ gid, const float4 *__restrict__ force,
                 const int *__restrict__ counts, float dt, int doKick, float4 *__restrict__ posN, float *__restrict__ velN,
                 int *__restrict__ gidN, int *__restrict__ work, int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const int nOwned = counts[0];
  // number of this exchange (counted on the device by brickBegin, so that a captured CUDA graph of the step can be
  // replayed): the flag value the peers wait for; its parity selects the inbox buffer
  const uint32_t epoch = (uint32_t)counts[2];
  const int parity = (int)((epoch - 1u) & 1u);
  const bool active = i < nOwned;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  float vx = 0.f, vy = 0.f, vz = 0.f;
  int id = 0;
  BrickClass c;
  c.owner = -1; c.mask = 0;
  if (active) {
    p = pos[i];
    vx = vel[3 * (size_t)i]; vy = vel[3 * (size_t)i + 1]; vz = vel[3 * (size_t)i + 2];
    id = gid[i];
    if (doKick) {
      const float4 f = force[i];
      if (doKick == 2) { // the closing kick of the previous step first (nveKickKickDrift: same roundings as two passes)
        vx = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.x), dt), 0.5f, vx);
        vy = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.y), dt), 0.5f, vy);
        vz = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.z), dt), 0.5f, vz);
      }
      vx = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.x), dt), 0.5f, vx);
      vy = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.y), dt), 0.5f, vy);
      vz = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.z), dt), 0.5f, vz);
      p.x = __fmaf_rn(vx, dt, p.x);
      p.y = __fmaf_rn(vy, dt, p.y);
      p.z = __fmaf_rn(vz, dt, p.z);
    }
    c = classify(b, p.x, p.y, p.z);
  }
  // stays owned: compacted into the new arrays
  const int ks = warpSlots(active && c.owner == b.me, work + 0, lane);
  if (ks >= 0) {
    posN[ks] = p;
    velN[3 * (size_t)ks] = vx; velN[3 * (size_t)ks + 1] = vy; velN[3 * (size_t)ks + 2] = vz;
    gidN[ks] = id;
  }
  // interior particles (the bulk) have nothing to send: the whole warp skips the per-rank passes
  const bool sends = active && (c.owner != b.me || c.mask != 0u);
  for (int r = 0; __any_sync(0xffffffffu, sends) && r < b.world; r++) {
    char *seg = arenaSeg(ar, r, parity, b.me, b.world);
    if (r != b.me) { // migrates to r
      const int s = warpSlots(active && c.owner == r, work + 2 + 2 * r, lane);
      if (s >= 0) {
        if (s < ar.capMig) storeRow(seg + kHdrBytes, s, p, vx, vy, vz, id);
        else *err = 5;
      }
    }
    // ghost of r (r == me: a particle that just left this brick but stays inside its window)
    const int s = warpSlots(active && ((c.mask >> r) & 1u), work + 3 + 2 * r, lane);
    if (s >= 0) {
      if (s < ar.capGhost) storeRow(seg + kHdrBytes + ar.migBytes, s, p, vx, vy, vz, id);
      else *err = 5;
    }
  }
  // last block: publish the counts, then raise this rank's flag in every peer. One system-scope fence per block, after
  // the block barrier: it orders the rows every thread of the block stored (cumulativity) before the block's ticket.
  __syncthreads();
  __shared__ int isLast;
  if (threadIdx.x == 0) {
    __threadfence_system();
    isLast = atomicAdd(work + 1, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!isLast) return;
  __threadfence();
  if (threadIdx.x < b.world) {
    const int r = threadIdx.x;
    const int nm = r == b.me ? 0 : min(((volatile int *)work)[2 + 2 * r], ar.capMig);
    const int ng = min(((volatile int *)work)[3 + 2 * r], ar.capGhost);
    volatile int *hdr = reinterpret_cast<volatile int *>(arenaSeg(ar, r, parity, b.me, b.world));
    hdr[0] = nm;
    hdr[1] = ng;
    __threadfence_system();
    uint32_t *flag = arenaFlags(ar.p[r]) + b.me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
  }
}

// Receiver side: wait for every rank's flag of this exchange, append the migrants to the kept block and the ghosts behind
// them, publish {nOwned, nLocal} on the device.
__global__ void __launch_bounds__(256)
brickUnpack(BrickArena ar, int me, int world, int cap, float4 *__restrict__ posN,
            float *__restrict__ velN, int *__restrict__ gidN, const int *__restrict__ work, int *__restrict__ counts,
            int *__restrict__ err) {
  __shared__ int sOff[2 * kBrickMaxRanks + 1];
  __shared__ int sKept;
  const uint32_t epoch = (uint32_t)counts[2];
  const int parity = (int)((epoch - 1u) & 1u);
  if (threadIdx.x == 0) {
    const uint32_t *flags = arenaFlags(ar.p[me]);
    for (int s = 0; s < world; s++) {
      unsigned long long spins = 0;
      while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + s) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (++spins > kBrickSpinLimit) { *err = 6; break; }
      }
    }
    __threadfence_system();
    int acc = 0;
    for (int k = 0; k < 2; k++) // migrants of every source first, then ghosts of every source
      for (int s = 0; s < world; s++) {
        sOff[k * world + s] = acc;
        const volatile int *hdr = reinterpret_cast<const volatile int *>(arenaSeg(ar, me, parity, s, world));
        acc += hdr[k];
      }
    sOff[2 * world] = acc;
    sKept = ((const volatile int *)work)[0];
  }
  __syncthreads();
  const int kept = sKept, total = sOff[2 * world], nMig = sOff[world];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[0] = min(kept + nMig, cap);
    counts[1] = min(kept + total, cap);
    if (kept + total > cap) *err = 4;
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int seg = 0;
    while (t >= sOff[seg + 1]) seg++;
    const int k = seg / world, s = seg - k * world;
    const char *rows = arenaSeg(ar, me, parity, s, world) + kHdrBytes + (k ? ar.migBytes : 0);
    const float4 *src = reinterpret_cast<const float4 *>(rows) + 2 * (size_t)(t - sOff[seg]);
    const float4 a = __ldcg(src), v = __ldcg(src + 1); // written by a peer: read past L1
    const int dst = kept + t;
    if (dst < cap) {
      posN[dst] = a;
      velN[3 * (size_t)dst] = v.x; velN[3 * (size_t)dst + 1] = v.y; velN[3 * (size_t)dst + 2] = v.z;
      gidN[dst] = __float_as_int(v.w);
    }
  }
}

// opens an exchange: clears the slot counters and counts the exchange
__global__ void brickBegin(int *__restrict__ work, int *__restrict__ counts) {
  if (threadIdx.x < kWorkInts) work[threadIdx.x] = 0;
  if (threadIdx.x == 0) counts[2] += 1;
}

// second kick of velocity Verlet on the owned block (nveHalfStep<2>, unit mass)
__global__ void __launch_bounds__(256)
brickKick2(float *__restrict__ vel, const float4 *__restrict__ force, const int *__restrict__ counts, float dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[0]) return;
  const float4 f = force[i];
  vel[3 * (size_t)i] = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.x), dt), 0.5f, vel[3 * (size_t)i]);
  vel[3 * (size_t)i + 1] = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.y), dt), 0.5f, vel[3 * (size_t)i + 1]);
  vel[3 * (size_t)i + 2] = __fmaf_rn(__fmul_rn(__fmul_rn(1.0f, f.z), dt), 0.5f, vel[3 * (size_t)i + 2]);
}

// initial condition: every rank sees the full (replicated) arrays and keeps what it owns
__global__ void __launch_bounds__(256)
brickSelectOwned(BrickGeom b, const float4 *__restrict__ posG, const float *__restrict__ velG, int N, int cap,
                 float4 *__restrict__ pos, float *__restrict__ vel, int *__restrict__ gid, int *__restrict__ work,
                 int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  bool mine = false;
  if (i < N) {
    p = posG[i];
    mine = classify(b, p.x, p.y, p.z).owner == b.me;
  }
  const int s = warpSlots(mine, work + 0, lane);
  if (s < 0) return;
  if (s >= cap) { *err = 4; return; }
  pos[s] = p;
  vel[3 * (size_t)s] = velG[3 * (size_t)i]; vel[3 * (size_t)s + 1] = velG[3 * (size_t)i + 1]; vel[3 * (size_t)s + 2] = velG[3 * (size_t)i + 2];
  gid[s] = i;
}

} // namespace ub200

using namespace ub200;

struct ub200_brick {
  BrickGeom geom;
  ColGrid cg;   // this rank's window of the global half-cell grid
  float L[3];
  int periodic[3];
  float rc = 0;
  int cap = 0, N = 0;
  DevBuf pos[2], vel[2], gid[2], force, counts, work, err;
  int cur = 0;
  void *arena = nullptr;
  size_t arenaBytes = 0;
  BrickArena ar;
  bool attached = false, ipcOpened[kBrickMaxRanks] = {};
  bool prepared = false;
  ub200_ljengine *eng = nullptr;
  ub200_celllist *cl = nullptr; // DPD: reference-layout list over the local arrays (cells >= cutOff, global grid)
  uint32_t dpdStep = 0;         // DPD_impl::step: incremented before every force evaluation (DPD.cuh:165)
  // CUDA graph of kGraphSteps consecutive steps (the whole step is device driven: counts, exchange number and inbox parity
  // live on the device, so one captured sequence replays for every step). Runs on an internal stream joined to the caller's
  // by events (the legacy default stream cannot be captured). exec[c]: graph captured with state buffer c current.
  bool useGraph = true;
  cudaStream_t gs = nullptr;
  cudaEvent_t evIn = nullptr, evOut = nullptr;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};  // kGraphSteps steps starting with state buffer c
  cudaGraphExec_t exec1[2] = {nullptr, nullptr}; // one step (callers that step one at a time: the host enqueues one graph
                                                 // instead of a dozen launches)
  std::vector<float> graphParams;
  float graphDt = 0.f;
  // optional phase timing (UB200_BRICK_PROFILE=1; makes every step synchronous): push, unpack (incl. waiting for the
  // peers), list build, traversal, kick
  static constexpr int kPhases = 5;
  bool profile = false;
  cudaEvent_t ev[kPhases + 1] = {};
  double phaseMs[kPhases] = {};
  int profiledSteps = 0;
};

static void brickMark(ub200_brick *h, int k, cudaStream_t st) {
  if (h->profile) cudaEventRecord(h->ev[k], st);
}
static void brickCollect(ub200_brick *h, cudaStream_t st) {
  if (!h->profile) return;
  cudaStreamSynchronize(st);
  for (int k = 0; k < ub200_brick::kPhases; k++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]) == cudaSuccess) h->phaseMs[k] += ms;
  }
  h->profiledSteps++;
}

static int brickPush(ub200_brick *h, float dt, int doKick, cudaStream_t st) {
  const int c = h->cur, n = c ^ 1;
  brickBegin<<<1, 32, 0, st>>>(h->work.as<int>(), h->counts.as<int>());
  UB200_LAUNCHED();
  brickMark(h, 0, st);
  brickAdvancePush<<<(h->cap + 255) / 256, 256, 0, st>>>(h->geom, h->ar, h->pos[c].as<float4>(), h->vel[c].as<float>(),
                                                        h->gid[c].as<int>(), h->force.as<float4>(), h->counts.as<int>(), dt, doKick,
                                                        h->pos[n].as<float4>(), h->vel[n].as<float>(), h->gid[n].as<int>(),
                                                        h->work.as<int>(), h->err.as<int>());
  UB200_LAUNCHED();
  brickMark(h, 1, st);
  return UB200_OK;
}
static int brickPull(ub200_brick *h, cudaStream_t st) {
  const int n = h->cur ^ 1;
  brickUnpack<<<2 * kNumSMs, 256, 0, st>>>(h->ar, h->geom.me, h->geom.world, h->cap, h->pos[n].as<float4>(), h->vel[n].as<float>(),
                                          h->gid[n].as<int>(), h->work.as<int>(), h->counts.as<int>(), h->err.as<int>());
  UB200_LAUNCHED();
  brickMark(h, 2, st);
  h->cur = n;
  return UB200_OK;
}
static int brickExchange(ub200_brick *h, float dt, int doKick, cudaStream_t st) {
  if (!h->attached) return UB200_ERR_NOT_BUILT;
  if (const int rc = brickPush(h, dt, doKick, st)) return rc;
  return brickPull(h, st);
}
// phase 0 of an exchange alone / phase 1 alone (virtual ranks of one process enqueue phase 0 of every rank first)
static int brickExchangePhase(ub200_brick *h, int phase, float dt, int doKick, cudaStream_t st) {
  if (!h->attached) return UB200_ERR_NOT_BUILT;
  return phase == 0 ? brickPush(h, dt, doKick, st) : brickPull(h, st);
}

static int brickForcesLJ(ub200_brick *h, const float *params, int ntypes, cudaStream_t st) {
  const int c = h->cur;
  int rc = ljEngineBuildWindow(h->eng, h->pos[c].as<float4>(), h->gid[c].as<int>(), h->cap, h->counts.as<int>() + 1, h->L,
                               h->periodic, h->geom.dims, h->cg, st);
  if (rc) return rc;
  brickMark(h, 3, st);
  rc = ljEngineTraverseWindow(h->eng, h->counts.as<int>(), params, ntypes, h->force.as<float4>(), false, st);
  brickMark(h, 4, st);
  return rc;
}

extern "C" {

int ub200_brick_create(ub200_brick **out, int rank, const int rankGrid[3], const float L[3], const int periodic[3], float cutOff,
                       int numberParticles, int capacity) {
  if (!out || !rankGrid || !L || !periodic || !(cutOff > 0) || numberParticles < 1) return UB200_ERR_INVALID_ARGUMENT;
  const int world = rankGrid[0] * rankGrid[1] * rankGrid[2];
  if (rankGrid[0] < 1 || rankGrid[1] < 1 || rankGrid[2] < 1 || world > kBrickMaxRanks || rank < 0 || rank >= world)
    return UB200_ERR_INVALID_ARGUMENT;
  int dims[3], per[3];
  if (!ljEngineDims(L, periodic, cutOff, dims, per)) return UB200_ERR_UNSUPPORTED;
  ub200_brick *h = new (std::nothrow) ub200_brick();
  if (!h) return UB200_ERR_ALLOC;
  BrickGeom &b = h->geom;
  b.g = makeGridF(L, per, dims);
  b.me = rank; b.world = world;
  const int k[3] = {rank % rankGrid[0], (rank / rankGrid[0]) % rankGrid[1], rank / (rankGrid[0] * rankGrid[1])};
  int wn[3], wo[3], wp[3];
  double frac = 1.0; // window volume / box volume
  for (int d = 0; d < 3; d++) {
    b.dims[d] = dims[d]; b.rg[d] = rankGrid[d]; b.per[d] = per[d];
    h->L[d] = L[d]; h->periodic[d] = per[d];
    const int lo = brickLo(k[d], dims[d], rankGrid[d]), hi = brickLo(k[d] + 1, dims[d], rankGrid[d]);
    if (rankGrid[d] == 1) { wn[d] = dims[d]; wo[d] = 0; wp[d] = per[d]; }
    else {
      // every brick at least two half cells thick (only adjacent bricks exchange), window no wider than the grid
      if (dims[d] / rankGrid[d] < 2 || hi - lo + 4 > dims[d]) { delete h; return UB200_ERR_UNSUPPORTED; }
      wn[d] = hi - lo + 4; wo[d] = lo - 2; wp[d] = 0;
    }
    frac *= (double)wn[d] / dims[d];
  }
  h->cg = ColGrid{wn[0], wn[1], wn[2], wp[0], wp[1], wp[2], wo[0], wo[1], wo[2], dims[0], dims[1], dims[2],
                  rankGrid[0] > 1, rankGrid[1] > 1, rankGrid[2] > 1};
  h->rc = cutOff;
  h->N = numberParticles;
  // Capacity: the senders address the receivers' inboxes with THEIR OWN layout constants, so every rank must arrive at the
  // same numbers: the default is derived from the LARGEST window of the rank grid, not from this rank's.
  double maxFrac = 0.0;
  for (int r = 0; r < world; r++) {
    const int kr[3] = {r % rankGrid[0], (r / rankGrid[0]) % rankGrid[1], r / (rankGrid[0] * rankGrid[1])};
    double f = 1.0;
    for (int d = 0; d < 3; d++) {
      const int lo = brickLo(kr[d], dims[d], rankGrid[d]), hi = brickLo(kr[d] + 1, dims[d], rankGrid[d]);
      f *= rankGrid[d] == 1 ? 1.0 : (double)(hi - lo + 4) / dims[d];
    }
    maxFrac = std::max(maxFrac, f);
  }
  (void)frac;
  h->cap = capacity > 0 ? capacity : (int)std::min<double>(numberParticles + 4096.0, 1.3 * numberParticles * maxFrac + 16384.0);
  int rc;
  for (int s = 0; s < 2; s++) {
    if ((rc = h->pos[s].reserve(sizeof(float4) * (size_t)h->cap)) || (rc = h->vel[s].reserve(sizeof(float) * 3 * (size_t)h->cap)) ||
        (rc = h->gid[s].reserve(sizeof(int) * (size_t)h->cap))) { delete h; return rc; }
  }
  if ((rc = h->force.reserve(sizeof(float4) * (size_t)h->cap)) || (rc = h->counts.reserve(sizeof(int) * 4)) ||
      (rc = h->work.reserve(sizeof(int) * kWorkInts)) || (rc = h->err.reserve(sizeof(int)))) { delete h; return rc; }
  cudaMemset(h->counts.p, 0, sizeof(int) * 4);
  cudaMemset(h->err.p, 0, sizeof(int));
  cudaMemset(h->force.p, 0, sizeof(float4) * (size_t)h->cap);
  BrickArena &ar = h->ar;
  ar.capGhost = h->cap;
  ar.capMig = std::max(4096, h->cap / 8);
  ar.migBytes = (size_t)ar.capMig * 32;
  ar.segBytes = kHdrBytes + ar.migBytes + (size_t)ar.capGhost * 32;
  ar.inboxOff = 256;
  h->arenaBytes = ar.inboxOff + ar.segBytes * 2 * world;
  if (cudaMalloc(&h->arena, h->arenaBytes) != cudaSuccess) { delete h; return UB200_ERR_ALLOC; }
  cudaMemset(h->arena, 0, ar.inboxOff + 0);
  for (int p = 0; p < 2 * world; p++) cudaMemset((char *)h->arena + ar.inboxOff + p * ar.segBytes, 0, kHdrBytes);
  for (int p = 0; p < kBrickMaxRanks; p++) ar.p[p] = nullptr;
  ar.p[rank] = (char *)h->arena;
  if ((rc = ub200_ljengine_create(&h->eng)) || (rc = ub200_celllist_create(&h->cl))) { cudaFree(h->arena); delete h; return rc; }
  h->attached = world == 1;
  const char *gr = getenv("UB200_BRICK_GRAPH"); // "0": plain launches instead of the captured step graph
  h->useGraph = !(gr && gr[0] == '0');
  const char *pf = getenv("UB200_BRICK_PROFILE");
  if (pf && pf[0] == '1') {
    h->profile = true;
    for (auto &e : h->ev) cudaEventCreate(&e);
  }
  *out = h;
  return UB200_OK;
}

int ub200_brick_destroy(ub200_brick *h) {
  if (!h) return UB200_OK;
  for (int p = 0; p < h->geom.world; p++)
    if (h->ipcOpened[p]) cudaIpcCloseMemHandle(h->ar.p[p]);
  if (h->arena) cudaFree(h->arena);
  for (auto &e : h->exec)
    if (e) cudaGraphExecDestroy(e);
  for (auto &e : h->exec1)
    if (e) cudaGraphExecDestroy(e);
  if (h->gs) { cudaStreamDestroy(h->gs); cudaEventDestroy(h->evIn); cudaEventDestroy(h->evOut); }
  ub200_ljengine_destroy(h->eng);
  ub200_celllist_destroy(h->cl);
  DevBuf *b[] = {&h->pos[0], &h->pos[1], &h->vel[0], &h->vel[1], &h->gid[0], &h->gid[1], &h->force, &h->counts, &h->work, &h->err};
  for (auto *x : b) x->release();
  delete h;
  return UB200_OK;
}

int ub200_comm_ipc_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int ub200_brick_ipc_export(ub200_brick *h, void *blob) {
  if (!h || !blob) return UB200_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t m;
  UB200_CUDA(cudaIpcGetMemHandle(&m, h->arena));
  memcpy(blob, &m, sizeof(m));
  return UB200_OK;
}

int ub200_brick_ipc_import(ub200_brick *h, const void *blobsOfAllRanks) {
  if (!h || !blobsOfAllRanks) return UB200_ERR_INVALID_ARGUMENT;
  for (int p = 0; p < h->geom.world; p++) {
    if (p == h->geom.me) continue;
    cudaIpcMemHandle_t m;
    memcpy(&m, static_cast<const char *>(blobsOfAllRanks) + (size_t)p * sizeof(m), sizeof(m));
    void *ptr = nullptr;
    UB200_CUDA(cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
    h->ar.p[p] = (char *)ptr;
    h->ipcOpened[p] = true;
  }
  h->attached = true;
  return UB200_OK;
}

int ub200_brick_arena(ub200_brick *h, void **arena) {
  if (!h || !arena) return UB200_ERR_INVALID_ARGUMENT;
  *arena = h->arena;
  return UB200_OK;
}

int ub200_brick_attach_local(ub200_brick *h, void *const *arenasOfAllRanks) {
  if (!h || !arenasOfAllRanks) return UB200_ERR_INVALID_ARGUMENT;
  for (int p = 0; p < h->geom.world; p++) h->ar.p[p] = (char *)arenasOfAllRanks[p];
  h->attached = true;
  return UB200_OK;
}

int ub200_brick_set_global_state_f32(ub200_brick *h, const void *d_pos, const void *d_vel, int N, void *stream) {
  if (!h || !d_pos || !d_vel || N != h->N) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  UB200_CUDA(cudaMemsetAsync(h->work.p, 0, sizeof(int) * kWorkInts, st));
  h->cur = 0;
  brickSelectOwned<<<(N + 255) / 256, 256, 0, st>>>(h->geom, (const float4 *)d_pos, (const float *)d_vel, N, h->cap, h->pos[0].as<float4>(),
                                                   h->vel[0].as<float>(), h->gid[0].as<int>(), h->work.as<int>(), h->err.as<int>());
  UB200_LAUNCHED();
  UB200_CUDA(cudaMemcpyAsync(h->counts.p, h->work.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
  UB200_CUDA(cudaMemcpyAsync(h->counts.as<int>() + 1, h->work.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
  h->prepared = false;
  return UB200_OK;
}

int ub200_halo_exchange_f32(ub200_brick *h, void *stream) {
  if (!h) return UB200_ERR_INVALID_ARGUMENT;
  h->prepared = false;
  return brickExchange(h, 0.0f, 0, (cudaStream_t)stream);
}

int ub200_halo_exchange_phase_f32(ub200_brick *h, int phase, void *stream) {
  if (!h || (phase != 0 && phase != 1)) return UB200_ERR_INVALID_ARGUMENT;
  h->prepared = false;
  return brickExchangePhase(h, phase, 0.0f, 0, (cudaStream_t)stream);
}

int ub200_brick_lj_forces_f32(ub200_brick *h, const float *params, int ntypes, void *stream) {
  if (!h || !params || ntypes < 1) return UB200_ERR_INVALID_ARGUMENT;
  const int rc = brickForcesLJ(h, params, ntypes, (cudaStream_t)stream);
  if (!rc) h->prepared = true;
  return rc;
}

// DPD forces of the owned block (DPD_impl::ForceTransverser, Interactor/Potential/DPD.cuh:92-159): cell list over
// [owned | ghosts] with the cells ordered by global id, ghosts bring their velocities, the pairwise noise is keyed on global
// ids (ij = min + N max with N the GLOBAL particle number), so every pair draws the single-GPU random force.
struct BrickDPD {
  float A, gamma, sigma, rcut;
  uint32_t seed;
};
static int brickForcesDPD(ub200_brick *h, const BrickDPD &p, cudaStream_t st) {
  const int c = h->cur;
  int cd[3], rc;
  if ((rc = ub200_neighbour_celldim_f32(h->L, p.rcut, cd))) return rc;
  if ((rc = celllistBuildEx(h->cl, h->pos[c].p, nullptr, h->cap, h->counts.as<int>() + 1, h->gid[c].as<int>(), h->L, h->periodic, cd,
                            (void *)st)))
    return rc;
  brickMark(h, 3, st);
  h->dpdStep++;
  rc = dpdSum(h->cl, h->vel[c].p, p.A, p.gamma, p.sigma, p.rcut, p.seed, h->dpdStep, h->N, h->force.p, nullptr, 0, 0x7fffffff, 0,
              (void *)st, h->gid[c].as<int>(), h->counts.as<int>());
  brickMark(h, 4, st);
  return rc;
}
static int brickStepsDPD(ub200_brick *h, const BrickDPD &p, float dt, int nsteps, cudaStream_t st) {
  int rc;
  for (int s = 0; s < nsteps; s++) {
    if ((rc = brickExchange(h, dt, s == 0 || h->profile ? 1 : 2, st)) || (rc = brickForcesDPD(h, p, st))) return rc;
    if (s == nsteps - 1 || h->profile) {
      brickKick2<<<(h->cap + 255) / 256, 256, 0, st>>>(h->vel[h->cur].as<float>(), h->force.as<float4>(), h->counts.as<int>(), dt);
      UB200_LAUNCHED();
    }
    brickMark(h, 5, st);
    brickCollect(h, st);
  }
  return UB200_OK;
}

// nsteps consecutive steps: the closing kick of a step is fused with the opening kick + drift of the next one
static int brickSteps(ub200_brick *h, const float *params, int ntypes, float dt, int nsteps, cudaStream_t st) {
  int rc;
  for (int s = 0; s < nsteps; s++) {
    if ((rc = brickExchange(h, dt, s == 0 || h->profile ? 1 : 2, st)) || (rc = brickForcesLJ(h, params, ntypes, st))) return rc;
    if (s == nsteps - 1 || h->profile) {
      brickKick2<<<(h->cap + 255) / 256, 256, 0, st>>>(h->vel[h->cur].as<float>(), h->force.as<float4>(), h->counts.as<int>(), dt);
      UB200_LAUNCHED();
    }
    brickMark(h, 5, st);
    brickCollect(h, st);
  }
  return UB200_OK;
}

constexpr int kGraphSteps = 4; // even: the double-buffered state is back in the same buffer after one graph

int ub200_brick_lj_nve_run_f32(ub200_brick *h, const float *params, int ntypes, float dt, int nsteps, void *stream) {
  if (!h || !params || ntypes < 1 || nsteps < 0) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (!h->prepared) {
    if ((rc = brickExchange(h, 0.0f, 0, st)) || (rc = brickForcesLJ(h, params, ntypes, st))) return rc;
    h->prepared = true;
  }
  if (!h->useGraph || h->profile || nsteps < 1) return brickSteps(h, params, ntypes, dt, nsteps, st);
  if (!h->gs) {
    UB200_CUDA(cudaStreamCreateWithFlags(&h->gs, cudaStreamNonBlocking));
    UB200_CUDA(cudaEventCreateWithFlags(&h->evIn, cudaEventDisableTiming));
    UB200_CUDA(cudaEventCreateWithFlags(&h->evOut, cudaEventDisableTiming));
  }
  const size_t np = (size_t)ntypes * ntypes * 4;
  if (h->graphDt != dt || h->graphParams.size() != np || memcmp(h->graphParams.data(), params, np * sizeof(float)) != 0) {
    for (auto &e : h->exec)
      if (e) { cudaGraphExecDestroy(e); e = nullptr; }
    for (auto &e : h->exec1)
      if (e) { cudaGraphExecDestroy(e); e = nullptr; }
    h->graphParams.assign(params, params + np);
    h->graphDt = dt;
  }
  // graphs are CAPTURED on the internal stream (the caller's may be the legacy default stream, which cannot capture) and
  // LAUNCHED on the caller's stream: no event hops between streams on the step path
  int done = 0;
  while (nsteps - done >= kGraphSteps) {
    const int c = h->cur;
    if (!h->exec[c]) {
      cudaGraph_t graph = nullptr;
      UB200_CUDA(cudaStreamBeginCapture(h->gs, cudaStreamCaptureModeRelaxed));
      rc = brickSteps(h, params, ntypes, dt, kGraphSteps, h->gs);
      const cudaError_t ce = cudaStreamEndCapture(h->gs, &graph);
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        h->useGraph = false; // fall back to plain launches for good
        h->cur = c;
        if ((rc = brickSteps(h, params, ntypes, dt, nsteps - done, st))) return rc;
        done = nsteps;
        break;
      }
      const cudaError_t ie = cudaGraphInstantiate(&h->exec[c], graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) return cudaFail(ie);
      // capturing ran the host side of the steps only (kGraphSteps is even: h->cur is back at c)
    }
    UB200_CUDA(cudaGraphLaunch(h->exec[c], st));
    g_launchCount += 11 * kGraphSteps; // kernels of the replayed steps
    done += kGraphSteps;
  }
  while (h->useGraph && done < nsteps) { // left-over steps one by one, each a graph of its own
    const int c = h->cur;
    if (!h->exec1[c]) {
      cudaGraph_t graph = nullptr;
      UB200_CUDA(cudaStreamBeginCapture(h->gs, cudaStreamCaptureModeRelaxed));
      rc = brickSteps(h, params, ntypes, dt, 1, h->gs);
      const cudaError_t ce = cudaStreamEndCapture(h->gs, &graph);
      h->cur = c; // capturing ran the host side only
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        h->useGraph = false;
        break;
      }
      const cudaError_t ie = cudaGraphInstantiate(&h->exec1[c], graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) return cudaFail(ie);
    }
    UB200_CUDA(cudaGraphLaunch(h->exec1[c], st));
    g_launchCount += 12;
    h->cur = c ^ 1;
    done++;
  }
  if (done < nsteps && (rc = brickSteps(h, params, ntypes, dt, nsteps - done, st))) return rc;
  return UB200_OK;
}

// VerletNVE::forwardTime x nsteps with one PairForces<Potential::DPD> interactor on the bricks (BASELINE config 4)
int ub200_brick_dpd_nve_run_f32(ub200_brick *h, float A, float gamma, float sigma, float rcut, uint32_t seed, float dt, int nsteps,
                                void *stream) {
  if (!h || !(rcut > 0) || nsteps < 0) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const BrickDPD p = {A, gamma, sigma, rcut, seed};
  int rc;
  if (!h->prepared) {
    if ((rc = brickExchange(h, 0.0f, 0, st)) || (rc = brickForcesDPD(h, p, st))) return rc;
    h->prepared = true;
  }
  return brickStepsDPD(h, p, dt, nsteps, st);
}
int ub200_brick_dpd_nve_phase_f32(ub200_brick *h, int phase, float A, float gamma, float sigma, float rcut, uint32_t seed, float dt,
                                  int doKick, void *stream) {
  if (!h || !(rcut > 0) || (phase != 0 && phase != 1)) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const BrickDPD p = {A, gamma, sigma, rcut, seed};
  int rc;
  if ((rc = brickExchangePhase(h, phase, dt, doKick, st))) return rc;
  if (phase == 0) return UB200_OK;
  if ((rc = brickForcesDPD(h, p, st))) return rc;
  h->prepared = true;
  if (doKick) {
    brickKick2<<<(h->cap + 255) / 256, 256, 0, st>>>(h->vel[h->cur].as<float>(), h->force.as<float4>(), h->counts.as<int>(), dt);
    UB200_LAUNCHED();
  }
  return UB200_OK;
}

int ub200_brick_profile(ub200_brick *h, double phases[5]) {
  if (!h || !phases) return UB200_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < ub200_brick::kPhases; k++) phases[k] = h->profiledSteps ? h->phaseMs[k] / h->profiledSteps : 0.0;
  return UB200_OK;
}

// the same step in two halves for virtual ranks sharing one process and one stream: phase 0 (kick + drift + push) of
// every rank must be enqueued before phase 1 (unpack + forces + kick) of any
int ub200_brick_lj_nve_phase_f32(ub200_brick *h, int phase, const float *params, int ntypes, float dt, int doKick, void *stream) {
  if (!h || !params || ntypes < 1 || (phase != 0 && phase != 1)) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = brickExchangePhase(h, phase, dt, doKick, st))) return rc;
  if (phase == 0) return UB200_OK;
  if ((rc = brickForcesLJ(h, params, ntypes, st))) return rc;
  h->prepared = true;
  if (doKick) {
    brickKick2<<<(h->cap + 255) / 256, 256, 0, st>>>(h->vel[h->cur].as<float>(), h->force.as<float4>(), h->counts.as<int>(), dt);
    UB200_LAUNCHED();
  }
  return UB200_OK;
}

int ub200_brick_info(ub200_brick *h, ub200_brick_info_t *info) {
  if (!h || !info) return UB200_ERR_INVALID_ARGUMENT;
  const int c = h->cur;
  info->d_pos = h->pos[c].p; info->d_vel = h->vel[c].p; info->d_gid = h->gid[c].as<int>(); info->d_force = h->force.p;
  info->d_counts = h->counts.as<int>();
  info->capacity = h->cap;
  info->rank = h->geom.me; info->world = h->geom.world;
  info->halfCells[0] = h->geom.dims[0]; info->halfCells[1] = h->geom.dims[1]; info->halfCells[2] = h->geom.dims[2];
  info->window[0] = h->cg.nx; info->window[1] = h->cg.ny; info->window[2] = h->cg.nz;
  info->windowOrigin[0] = h->cg.ox; info->windowOrigin[1] = h->cg.oy; info->windowOrigin[2] = h->cg.oz;
  return UB200_OK;
}

// owned block <-> host buffers (pinned for asynchronous copies): pos real4[n], vel real3[n], ids int[n]. n owned particles
// as reported by ub200_brick_counts; upload replaces the owned block (the next exchange redistributes it).
int ub200_brick_download_owned_f32(ub200_brick *h, void *h_pos, void *h_vel, int *h_gid, int n, void *stream) {
  if (!h || n < 0 || n > h->cap) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int c = h->cur;
  if (h_pos) UB200_CUDA(cudaMemcpyAsync(h_pos, h->pos[c].p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (h_vel) UB200_CUDA(cudaMemcpyAsync(h_vel, h->vel[c].p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (h_gid) UB200_CUDA(cudaMemcpyAsync(h_gid, h->gid[c].p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  return UB200_OK;
}
int ub200_brick_upload_owned_f32(ub200_brick *h, const void *h_pos, const void *h_vel, const int *h_gid, int n, void *stream) {
  if (!h || !h_pos || !h_vel || n < 0 || n > h->cap) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int c = h->cur;
  UB200_CUDA(cudaMemcpyAsync(h->pos[c].p, h_pos, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  UB200_CUDA(cudaMemcpyAsync(h->vel[c].p, h_vel, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (h_gid) UB200_CUDA(cudaMemcpyAsync(h->gid[c].p, h_gid, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
  return UB200_OK;
}

int ub200_brick_counts(ub200_brick *h, void *stream, int *nOwned, int *nLocal, int *errorFlag) {
  if (!h) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int c[2] = {0, 0}, e = 0, ee = 0;
  UB200_CUDA(cudaMemcpyAsync(c, h->counts.p, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  UB200_CUDA(cudaMemcpyAsync(&e, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  UB200_CUDA(cudaStreamSynchronize(st));
  ub200_ljengine_error_flag(h->eng, stream, &ee);
  if (nOwned) *nOwned = c[0];
  if (nLocal) *nLocal = c[1];
  if (errorFlag) *errorFlag = e ? e : ee;
  return UB200_OK;
}
}
