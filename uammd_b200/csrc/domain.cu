// Brick domain decomposition of the short-range pair path over the GPUs of one box ("3-D brick
// domain decomposition, owned particles + ghost shell, full-neighbour scheme => no reverse force communication").
// The reference is single-GPU; the decomposition is defined on the reference's own neighbour grid
// (CellList::createUpdateGrid, Interactor/NeighbourList/CellList.cuh:100-126; Grid::getCell, utils/Grid.cuh:49-71):
//   * rank (kx,ky,kz) of a px x py x pz rank grid owns the CELLS [floor(k n/p), floor((k+1) n/p)) per dimension, and
//     the particles whose cell - computed with exactly the arithmetic of the cell list build - lies in that range;
//   * rank r needs as ghosts the particles of every cell adjacent (27-neighbourhood, periodic wrap like
//     Grid::pbc_cell, utils/Grid.cuh:81-106) to one of its cells: the cell size is >= the cut-off, so the traversal
//     from an owned home cell only ever visits owned or ghost cells, whole.
// Because a cell is wholly owned or wholly ghost, a local array ordered [owned by id | ghosts by id] keeps the
// within-cell order of the single-GPU list, and the forces come out bit-identical to the single-GPU ones.
//
// brickClassify is pure integer work after the cell assignment: R 16 B, W 12 B per particle, HBM bound.
#include "common.cuh"

namespace ub200 {

struct BrickGrid {
  int px, py, pz;
};

// index k of the brick whose cell range [floor(k n/p), floor((k+1) n/p)) holds cell c
__host__ __device__ __forceinline__ int brickOfCell(int c, int n, int p) { return ((c + 1) * p - 1) / n; }

__global__ void __launch_bounds__(256)
brickClassify(const float4 *__restrict__ pos, int N, GridF g, BrickGrid b, int *__restrict__ cellOut,
              int *__restrict__ ownerOut, uint32_t *__restrict__ ghostMaskOut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 p = ldg4(pos + i);
  int cx = cellCoord(p.x, g.Lx, g.mx, g.hLx, g.ix, g.nx);
  int cy = cellCoord(p.y, g.Ly, g.my, g.hLy, g.iy, g.ny);
  int cz = cellCoord(p.z, g.Lz, g.mz, g.hLz, g.iz, g.nz);
  // outside a non periodic box: clamped like binParticles (the cell list build raises its error flag)
  cx = min(max(cx, 0), g.nx - 1);
  cy = min(max(cy, 0), g.ny - 1);
  cz = min(max(cz, 0), g.nz - 1);
  const int own = brickOfCell(cx, g.nx, b.px) + b.px * (brickOfCell(cy, g.ny, b.py) + b.py * brickOfCell(cz, g.nz, b.pz));
  uint32_t mask = 0;
  for (int oz = -1; oz <= 1; oz++) {
    int jz = cz + oz;
    if (jz < 0) { if (g.mz != 0.0f) jz += g.nz; else continue; }
    else if (jz >= g.nz) { if (g.mz != 0.0f) jz -= g.nz; else continue; }
    if (jz < 0 || jz >= g.nz) continue; // single-cell dimension
    const int kz = brickOfCell(jz, g.nz, b.pz);
    for (int oy = -1; oy <= 1; oy++) {
      int jy = cy + oy;
      if (jy < 0) { if (g.my != 0.0f) jy += g.ny; else continue; }
      else if (jy >= g.ny) { if (g.my != 0.0f) jy -= g.ny; else continue; }
      if (jy < 0 || jy >= g.ny) continue;
      const int ky = brickOfCell(jy, g.ny, b.py);
      for (int ox = -1; ox <= 1; ox++) {
        int jx = cx + ox;
        if (jx < 0) { if (g.mx != 0.0f) jx += g.nx; else continue; }
        else if (jx >= g.nx) { if (g.mx != 0.0f) jx -= g.nx; else continue; }
        if (jx < 0 || jx >= g.nx) continue;
        const int r = brickOfCell(jx, g.nx, b.px) + b.px * (ky + b.py * kz);
        mask |= 1u << r;
      }
    }
  }
  mask &= ~(1u << own);
  if (cellOut) cellOut[i] = cx + g.nx * (cy + g.ny * cz);
  ownerOut[i] = own;
  ghostMaskOut[i] = mask;
}

} // namespace ub200

using namespace ub200;

extern "C" int ub200_brick_classify_f32(const void *d_pos, int N, const float L[3], const int periodic[3],
                                        const int cellDim[3], const int rankGrid[3], int *d_cell, int *d_owner,
                                        uint32_t *d_ghostMask, void *stream) {
  if (!L || !periodic || !cellDim || !rankGrid || N < 0) return UB200_ERR_INVALID_ARGUMENT;
  long world = 1;
  for (int d = 0; d < 3; d++) {
    // every brick needs at least one cell; the ghost mask has one bit per rank
    if (rankGrid[d] < 1 || cellDim[d] < 1 || rankGrid[d] > cellDim[d]) return UB200_ERR_INVALID_ARGUMENT;
    world *= rankGrid[d];
  }
  if (world > 32) return UB200_ERR_UNSUPPORTED;
  if (N == 0) return UB200_OK;
  if (!d_pos || !d_owner || !d_ghostMask) return UB200_ERR_INVALID_ARGUMENT;
  const GridF g = makeGridF(L, periodic, cellDim);
  const BrickGrid b = {rankGrid[0], rankGrid[1], rankGrid[2]};
  brickClassify<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4 *)d_pos, N, g, b, d_cell, d_owner, d_ghostMask);
  UB200_LAUNCHED();
  return UB200_OK;
}
