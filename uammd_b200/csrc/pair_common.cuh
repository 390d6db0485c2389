// Shared machinery of the cell-tile pair traversal kernels (LJ, DPD, RPY near field), sm_100a.
#pragma once
#include "common.cuh"

namespace ub200 {

constexpr int kPairThreads = 128;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kCandCap = 1024; // staged candidates per home cell (16 KB); denser neighbourhoods take the direct path

struct NeighbourCells {
  int start, count, off; // per lane (= neighbour cell slot): first sorted index, population, exclusive prefix
  int total, centre;     // warp uniform
};

// Lane l < 27 describes neighbour cell l of home cell (cx,cy,cz) in the reference's visiting order
// (x offset fastest; dims with one cell are not expanded; NeighbourContainer.cuh:95-115).
__device__ __forceinline__ NeighbourCells describeNeighbours(const GridF &g, int cx, int cy, int cz,
                                                             const uint32_t *__restrict__ binStart, int lane) {
  NeighbourCells nc;
  const int npx = g.nx > 1 ? 3 : 1, npy = g.ny > 1 ? 3 : 1, npz = g.nz > 1 ? 3 : 1;
  const int ncell = npx * npy * npz;
  nc.centre = (npx > 1) + npx * (npy > 1) + npx * npy * (npz > 1);
  nc.start = 0; nc.count = 0;
  if (lane < ncell) {
    int jx = cx + (npx > 1 ? lane % 3 - 1 : 0);
    int jy = cy + (npy > 1 ? (lane / npx) % 3 - 1 : 0);
    int jz = cz + (npz > 1 ? lane / (npx * npy) - 1 : 0);
    bool valid = true;
    // Grid::pbc_cell (utils/Grid.cuh:81-106): single wrap in periodic dims; non periodic dims keep the raw
    // coordinate, whose out-of-range cells can hold nothing within the cut-off -> skipped here.
    if (jx < 0) { if (g.mx != 0.0f) jx += g.nx; else valid = false; }
    else if (jx >= g.nx) { if (g.mx != 0.0f) jx -= g.nx; else valid = false; }
    if (jy < 0) { if (g.my != 0.0f) jy += g.ny; else valid = false; }
    else if (jy >= g.ny) { if (g.my != 0.0f) jy -= g.ny; else valid = false; }
    if (jz < 0) { if (g.mz != 0.0f) jz += g.nz; else valid = false; }
    else if (jz >= g.nz) { if (g.mz != 0.0f) jz -= g.nz; else valid = false; }
    if (valid) {
      const uint32_t code = cellBin(g, jx, jy, jz);
      const uint32_t s = __ldg(binStart + code), e = __ldg(binStart + code + 1);
      nc.start = (int)s;
      nc.count = (int)(e - s);
    }
  }
  int inc = nc.count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  nc.off = inc - nc.count;
  nc.total = __shfl_sync(0xffffffffu, inc, 31);
  return nc;
}

// centre of cell (cx,cy,cz) in box coordinates
__device__ __forceinline__ float3 cellCentre(const GridF &g, int cx, int cy, int cz) {
  return make_float3(__fmaf_rn((float)cx + 0.5f, g.csx, -g.hLx), __fmaf_rn((float)cy + 0.5f, g.csy, -g.hLy),
                     __fmaf_rn((float)cz + 0.5f, g.csz, -g.hLz));
}
// bring p to the periodic image nearest the home cell centre c
__device__ __forceinline__ void toHomeImage(float4 &p, const GridF &g, const float3 &c) {
  p.x = imageNear(p.x, c.x, g.Lx, g.mx);
  p.y = imageNear(p.y, c.y, g.Ly, g.my);
  p.z = imageNear(p.z, c.z, g.Lz, g.mz);
}

__device__ __forceinline__ float warpSum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

} // namespace ub200
