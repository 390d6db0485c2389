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
#include <cstdlib>

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
                 const int *__restrict__ noiseId, const int *__restrict__ ownerHiDev) {
  if (ownerHiDev) ownerHi = *ownerHiDev; // multi-GPU bricks: the owned block [0, nOwned) is counted on the device
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


// ---------------------------------------------------------------------------------------------------------------
// The same traversal with one CTA per 4x4x4 block of cells (UB200_DPD_TILE, see dpdSum). At DPD densities a cell
// holds ~3 particles, so the per-cell kernel above stages 81 candidates, describes 27 neighbours in every warp and
// synchronises twice to serve 3 home particles (9.0 ms at N = 4e6). Here the 6x6x6 halo of the block (~650 particles,
// each cell staged ONCE per CTA instead of once per neighbouring home cell) is staged with its velocities, and every
// warp then walks home cells of the block on its own: the 27 neighbour ranges come from a shared-memory table, the
// candidates are visited through a per-warp index list in exactly the flattened order of the kernel above (same lanes,
// same queue), so the forces are bit-identical to it: 0 differing words and 3.7 ms at N = 4e6 on a B200
// (profiles/r01e_dpd_tile_ab.json, scripts/dpd_ab.cu). Blocks whose halo does not fit the staging area, and grids with
// fewer than 8 cells in a periodic dimension, take the per-cell algorithm.
constexpr int kTileB = 4, kTileH = kTileB + 2, kTileHalo = kTileH * kTileH * kTileH; // 216 halo cells
constexpr int kTileCap = 1024;                                                       // staged particles per block

// the per-cell algorithm of dpdCellTraversal<false> for one home cell, run cooperatively by the whole CTA with
// cand / candVel / queue provided by the caller (used for blocks whose halo overflows the staging area)
__device__ void dpdOneCellBlockwide(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex,
                                    const uint32_t *__restrict__ binStart, const GridF &g, int cx, int cy, int cz,
                                    const float *__restrict__ vel, const DPDPar &par, float4 *__restrict__ force,
                                    const int *__restrict__ globalIdx, int ownerLo, int ownerHi, int accumulate,
                                    const int *__restrict__ noiseId, float4 *cand, float4 *candVel,
                                    unsigned short (*queue)[64]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const NeighbourCells nc = describeNeighbours(g, cx, cy, cz, binStart, lane);
  const int hStart = __shfl_sync(0xffffffffu, nc.start, nc.centre);
  const int hCount = __shfl_sync(0xffffffffu, nc.count, nc.centre);
  const int hOff = __shfl_sync(0xffffffffu, nc.off, nc.centre);
  const bool staged = nc.total <= kDpdCap;
  const float3 hc = cellCentre(g, cx, cy, cz);
  if (hCount > 0 && staged) {
    for (int c = warp; c < 27; c += kPairWarps) {
      const int cnt = __shfl_sync(0xffffffffu, nc.count, c);
      if (cnt == 0) continue;
      const int st = __shfl_sync(0xffffffffu, nc.start, c);
      const int off = __shfl_sync(0xffffffffu, nc.off, c);
      for (int t = lane; t < cnt; t += 32) {
        float4 p = ldg4(sortPos + st + t);
        toHomeImage(p, g, hc);
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
    const int gih = groupIndex[hStart + h];
    const int idi = globalIdx ? globalIdx[gih] : gih;
    if (idi < ownerLo || idi >= ownerHi) continue; // warp uniform
    float4 pi, vi;
    if (staged) {
      pi = cand[hOff + h];
      vi = candVel[hOff + h];
    } else {
      pi = ldg4(sortPos + hStart + h);
      toHomeImage(pi, g, hc);
      vi = make_float4(vel[3 * (size_t)idi], vel[3 * (size_t)idi + 1], vel[3 * (size_t)idi + 2],
                       __int_as_float(noiseId ? noiseId[idi] : idi));
    }
    const int nidi = __float_as_int(vi.w);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (staged) {
      unsigned short *q = queue[warp];
      int qn = 0;
      auto body = [&](int t) {
        const float4 pj = cand[t];
        const float4 vj = candVel[t];
        dpdPair(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z, vi.x - vj.x, vi.y - vj.y, vi.z - vj.z, nidi, __float_as_int(vj.w), par,
                fx, fy, fz);
      };
      for (int t0 = 0; t0 < nc.total; t0 += 32) {
        const int t = t0 + lane;
        bool in = false;
        if (t < nc.total) {
          const float4 pj = cand[t];
          const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
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
          toHomeImage(pj, g, hc);
          dpdPair(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z, vi.x - vel[3 * (size_t)idj], vi.y - vel[3 * (size_t)idj + 1],
                  vi.z - vel[3 * (size_t)idj + 2], nidi, noiseId ? noiseId[idj] : idj, par, fx, fy, fz);
        }
      }
    }
    fx = warpSum(fx);
    fy = warpSum(fy);
    fz = warpSum(fz);
    if (lane == 0) {
      float4 f = accumulate ? force[idi] : make_float4(0.f, 0.f, 0.f, 0.f);
      f.x += fx; f.y += fy; f.z += fz;
      force[idi] = f;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPairThreads)
dpdTileTraversal(const float4 *__restrict__ sortPos, const int *__restrict__ groupIndex,
                 const uint32_t *__restrict__ binStart, GridF g, int tilesX, int tilesY, const float *__restrict__ vel,
                 DPDPar par, float4 *__restrict__ force, const int *__restrict__ globalIdx, int ownerLo, int ownerHi,
                 int accumulate, const int *__restrict__ noiseId, const int *__restrict__ ownerHiDev) {
  if (ownerHiDev) ownerHi = *ownerHiDev;
  __shared__ float4 sPos[kTileCap];
  __shared__ float4 sVel[kTileCap]; // vx, vy, vz, noise id (bits)
  __shared__ int sOff[kTileHalo + 1];  // exclusive prefix of the staged particles per halo cell
  __shared__ int sGlob[kTileHalo];     // first sorted index of each halo cell
  __shared__ unsigned short sIdx[kPairWarps][kDpdCap]; // per warp: staged slots of the current home cell's candidates
  __shared__ unsigned short queue[kPairWarps][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x;
  const int ox = (tile % tilesX) * kTileB, oy = ((tile / tilesX) % tilesY) * kTileB, oz = (tile / (tilesX * tilesY)) * kTileB;
  // ---- halo cells: population and first sorted index -----------------------------------------------------------
  for (int i = threadIdx.x; i < kTileHalo; i += kPairThreads) {
    int jx = ox - 1 + i % kTileH, jy = oy - 1 + (i / kTileH) % kTileH, jz = oz - 1 + i / (kTileH * kTileH);
    bool valid = true;
    // cells past the end of the grid belong to no block's interior; one wrap like Grid::pbc_cell in periodic dimensions
    if (jx < 0) { if (g.mx != 0.0f) jx += g.nx; else valid = false; }
    else if (jx >= g.nx) { if (g.mx != 0.0f && jx == g.nx) jx = 0; else valid = false; }
    if (jy < 0) { if (g.my != 0.0f) jy += g.ny; else valid = false; }
    else if (jy >= g.ny) { if (g.my != 0.0f && jy == g.ny) jy = 0; else valid = false; }
    if (jz < 0) { if (g.mz != 0.0f) jz += g.nz; else valid = false; }
    else if (jz >= g.nz) { if (g.mz != 0.0f && jz == g.nz) jz = 0; else valid = false; }
    int s = 0, cnt = 0;
    if (valid) {
      const uint32_t code = cellBin(g, jx, jy, jz);
      s = (int)__ldg(binStart + code);
      cnt = (int)__ldg(binStart + code + 1) - s;
    }
    sGlob[i] = s;
    sOff[i + 1] = cnt; // turned into a prefix below
  }
  if (threadIdx.x == 0) sOff[0] = 0;
  __syncthreads();
  if (warp == 0) { // inclusive scan of the 216 counts by one warp, 32 at a time
    int carry = 0;
    for (int c0 = 0; c0 < kTileHalo; c0 += 32) {
      const int i = c0 + lane;
      int v = i < kTileHalo ? sOff[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (i < kTileHalo) sOff[i + 1] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  const int total = sOff[kTileHalo];
  if (total > kTileCap) {
    // dense block: the per-cell algorithm, cell by cell, with this CTA's staging arrays
    for (int hcell = 0; hcell < kTileB * kTileB * kTileB; hcell++) {
      const int cx = ox + hcell % kTileB, cy = oy + (hcell / kTileB) % kTileB, cz = oz + hcell / (kTileB * kTileB);
      if (cx >= g.nx || cy >= g.ny || cz >= g.nz) continue; // block uniform
      dpdOneCellBlockwide(sortPos, groupIndex, binStart, g, cx, cy, cz, vel, par, force, globalIdx, ownerLo, ownerHi, accumulate,
                          noiseId, sPos, sVel, queue);
    }
    return;
  }
  // ---- stage the halo once: positions at the periodic image nearest the block centre, velocities, noise ids ------
  const float3 tc = make_float3(__fmaf_rn((float)ox + 0.5f * kTileB, g.csx, -g.hLx), __fmaf_rn((float)oy + 0.5f * kTileB, g.csy, -g.hLy),
                                __fmaf_rn((float)oz + 0.5f * kTileB, g.csz, -g.hLz));
  for (int i = threadIdx.x; i < kTileHalo; i += kPairThreads) {
    const int s = sGlob[i], o = sOff[i], cnt = sOff[i + 1] - o;
    for (int k = 0; k < cnt; k++) {
      float4 p = ldg4(sortPos + s + k);
      toHomeImage(p, g, tc);
      const int gi = groupIndex[s + k];
      const int id = globalIdx ? globalIdx[gi] : gi;
      sPos[o + k] = p;
      sVel[o + k] = make_float4(__ldg(vel + 3 * (size_t)id), __ldg(vel + 3 * (size_t)id + 1), __ldg(vel + 3 * (size_t)id + 2),
                                __int_as_float(noiseId ? __ldg(noiseId + id) : id));
    }
  }
  __syncthreads();
  // ---- every warp walks home cells of the block on its own ------------------------------------------------------
  unsigned short *idx = sIdx[warp];
  unsigned short *q = queue[warp];
  for (int hcell = warp; hcell < kTileB * kTileB * kTileB; hcell += kPairWarps) {
    const int lx = hcell % kTileB, ly = (hcell / kTileB) % kTileB, lz = hcell / (kTileB * kTileB);
    if (ox + lx >= g.nx || oy + ly >= g.ny || oz + lz >= g.nz) continue; // warp uniform
    const int home = (lz + 1) * kTileH * kTileH + (ly + 1) * kTileH + (lx + 1);
    const int hCount = sOff[home + 1] - sOff[home];
    if (hCount == 0) continue;
    // the 27 neighbour cells in the reference's visiting order (x offset fastest): lane l < 27 owns neighbour l
    int nslot = 0, ncnt = 0;
    if (lane < 27) {
      const int nb = home + (lane % 3 - 1) + (lane / 3 % 3 - 1) * kTileH + (lane / 9 - 1) * kTileH * kTileH;
      nslot = sOff[nb];
      ncnt = sOff[nb + 1] - nslot;
    }
    int inc = ncnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int noff = inc - ncnt;
    const int ntotal = __shfl_sync(0xffffffffu, inc, 31);
    const bool listed = ntotal <= kDpdCap;
    __syncwarp();
    if (listed)
      for (int k = 0; k < ncnt; k++) idx[noff + k] = (unsigned short)(nslot + k);
    __syncwarp();
    const int hGlob = sGlob[home];
    for (int h = 0; h < hCount; h++) {
      const int gih = groupIndex[hGlob + h];
      const int idi = globalIdx ? globalIdx[gih] : gih;
      if (idi < ownerLo || idi >= ownerHi) continue; // warp uniform
      const float4 pi = sPos[sOff[home] + h];
      const float4 vi = sVel[sOff[home] + h];
      const int nidi = __float_as_int(vi.w);
      float fx = 0.f, fy = 0.f, fz = 0.f;
      auto body = [&](int slot) {
        const float4 pj = sPos[slot];
        const float4 vj = sVel[slot];
        dpdPair(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z, vi.x - vj.x, vi.y - vj.y, vi.z - vj.z, nidi, __float_as_int(vj.w), par,
                fx, fy, fz);
      };
      if (listed) {
        int qn = 0; // warp uniform
        for (int t0 = 0; t0 < ntotal; t0 += 32) {
          const int t = t0 + lane;
          bool in = false;
          int slot = 0;
          if (t < ntotal) {
            slot = idx[t];
            const float4 pj = sPos[slot];
            const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
            const float rmod = __fsqrt_rn(__fmaf_rn(rz, rz, __fmaf_rn(ry, ry, rx * rx)));
            in = rmod != 0.0f && __frcp_rn(rmod) > par.invrcut;
          }
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) q[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
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
        // more candidates than the per-cell kernel stages: its cell-by-cell walk (same lanes per cell)
        for (int c = 0; c < 27; c++) {
          const int cnt = __shfl_sync(0xffffffffu, ncnt, c);
          const int st = __shfl_sync(0xffffffffu, nslot, c);
          for (int t = lane; t < cnt; t += 32) body(st + t);
        }
      }
      fx = warpSum(fx);
      fy = warpSum(fy);
      fz = warpSum(fz);
      if (lane == 0) {
        float4 f = accumulate ? force[idi] : make_float4(0.f, 0.f, 0.f, 0.f);
        f.x += fx; f.y += fy; f.z += fz;
        force[idi] = f;
      }
    }
    __syncwarp();
  }
}

} // namespace ub200

using namespace ub200;

int ub200::dpdSum(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut, uint32_t seed,
                  uint32_t step, int idStride, void *d_force, const int *d_globalIdx, int ownerLo, int ownerHi, int accumulate,
                  void *stream, const int *d_noiseId, const int *ownerHiDev) {
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
  // UB200_DPD_TILE: 1 = one CTA per 4x4x4 block of cells (dpdTileTraversal), 0 = one CTA per cell (dpdCellTraversal).
  // Default: the tile kernel at DPD-like densities (at most 4 particles per cell on average) on grids where every
  // periodic dimension has at least 8 cells; the forces are bit-identical either way.
  const char *tileEnv = getenv("UB200_DPD_TILE");
  const bool roomy = (g.mx == 0.0f || g.nx >= 8) && (g.my == 0.0f || g.ny >= 8) && (g.mz == 0.0f || g.nz >= 8) && g.nx > 1 && g.ny > 1 &&
                     g.nz > 1;
  const bool sparse = (double)cl->N <= 4.0 * (double)cl->ncells;
  const bool useTile = tileEnv ? tileEnv[0] == '1' : sparse;
  if (useTile && !pairMic && roomy) {
    const int tx = (g.nx + kTileB - 1) / kTileB, ty = (g.ny + kTileB - 1) / kTileB, tz = (g.nz + kTileB - 1) / kTileB;
    dpdTileTraversal<<<tx * ty * tz, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(),
                                                            cl->binStart.as<uint32_t>(), g, tx, ty, (const float *)d_vel, par,
                                                            (float4 *)d_force, d_globalIdx, ownerLo, ownerHi, accumulate, d_noiseId, ownerHiDev);
    UB200_LAUNCHED();
    return UB200_OK;
  }
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
                                                          d_noiseId, ownerHiDev);
  else
    dpdCellTraversal<false><<<grid, kPairThreads, 0, st>>>(cl->sortPos.as<float4>(), cl->groupIndex.as<int>(),
                                                           cl->binStart.as<uint32_t>(), g, cl->ncells,
                                                           (const float *)d_vel, par, (float4 *)d_force, d_globalIdx, ownerLo, ownerHi, accumulate,
                                                          d_noiseId, ownerHiDev);
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
