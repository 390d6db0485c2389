// Index geometry of the column traversal (engine-private half-cell grid), shared by the CUDA kernels and by a
// host unit test (tests/test_colgeom.py compiles this header with g++ and checks it against a brute-force stencil).
//
// The neighbour search of the reference visits, for a particle in cell c of a grid with cells >= cutOff, the 27 cells
// around c (Interactor/NeighbourList/CellList/NeighbourContainer.cuh:95-138). The engine bins the particles a second
// time on a grid of HALF cells (edge >= cutOff/2) sorted x-fastest, so the same neighbourhood is covered by the
// 5 x 5 x 5 half cells around the particle's own: 125 (cutOff/2)^3 = 15.6 cutOff^3 instead of 27 cutOff^3.
// A *column* is a run of TZ half cells along z at fixed (x0, y0). Its halo is the 5 x 5 x (TZ + 4) block around it,
// staged plane by plane: plane p (z = z0 - 2 + p) holds the 5 rows y = y0 - 2 .. y0 + 2, and a row is the x-run
// x0 - 2 .. x0 + 2, which is contiguous in the sorted arrays (two pieces when it crosses the periodic boundary).
// With that order the neighbourhood of home cell hz is the contiguous range of planes hz .. hz + 4.
//
// Windows (multi-GPU bricks). A rank of the brick decomposition bins its owned particles and ghosts on a WINDOW of the
// global half-cell grid: local cell l of a windowed dimension is global cell (o + l) mod g, the local grid is not
// periodic in that dimension, and a neighbour reached across the global periodic boundary carries the image shift
// floor((o + l) / g) - floor((o + l_home) / g) - exactly the shift the single-GPU traversal applies, so that both
// evaluate bit-identical separations. A dimension that is not decomposed is "whole": o = 0, g = n, periodic or not.
#pragma once

#ifdef __CUDACC__
#define UB200_HD __host__ __device__ __forceinline__
#else
#define UB200_HD inline
#endif

namespace ub200 {

struct ColGrid {
  int nx, ny, nz;       // (local) half cells per dimension
  int px, py, pz;       // the local grid wraps in this dimension (whole periodic dimension)
  int ox, oy, oz;       // global index of local cell 0 (0 for a whole dimension; may be negative for a window)
  int gx, gy, gz;       // half cells of the global grid (= n for a whole dimension)
  int wx, wy, wz;       // windowed dimension: local cells 0, 1 and n - 2, n - 1 are the ghost layers of a brick
};

inline ColGrid makeWholeColGrid(const int dims[3], const int periodic[3]) {
  return ColGrid{dims[0], dims[1], dims[2], periodic[0], periodic[1], periodic[2], 0, 0, 0, dims[0], dims[1], dims[2], 0, 0, 0};
}

// One staged row: up to two x segments of consecutive cells. c0[s] = linear (local) index of the first cell, n[s] =
// number of cells (0 = absent), sx[s] = image shift of the segment in box lengths; sy, sz = image shift of the row;
// hs = segment holding the home cell x0 (meaningful for the dy = 0 row of a home plane).
struct ColRow {
  int c0[2], n[2], sx[2];
  int sy, sz, hs;
};

// floor(u / g) without a division for -g <= u < 2 g
UB200_HD int colFloorDiv(int u, int g) {
  if (u >= 0) {
    if (u < g) return 0;
    if (u < 2 * g) return 1;
    return u / g;
  }
  if (u >= -g) return -1;
  return -((-u + g - 1) / g);
}

// Cell v of one dimension (n local cells, wraps locally iff per, window origin o on a global grid of g cells): local
// index w and image shift floor((o + v) / g). Home cells that are computed always lie in image 0 (a whole dimension has
// its cells in [0, n); the owned cells of a window are cells of the global grid proper), so the shift is also the
// shift relative to the home cell. False when v lies outside a non-wrapping grid.
UB200_HD bool colNeighbour(int v, int n, int per, int o, int g, int &w, int &shift) {
  w = v;
  shift = 0;
  if (v < 0 || v >= n) {
    if (!per) return false;
    const int q = colFloorDiv(v, n);
    w = v - q * n;
  }
  shift = colFloorDiv(o + v, g);
  return true;
}

// Row r (0 <= r < 5 * nPlanes; plane p = r / 5, dy = r % 5 - 2) of the column with first home cell (x0, y0, z0).
UB200_HD ColRow columnRow(const ColGrid &g, int x0, int y0, int z0, int r) {
  ColRow row;
  row.c0[0] = row.c0[1] = 0;
  row.n[0] = row.n[1] = 0;
  row.sx[0] = row.sx[1] = 0;
  row.sy = row.sz = row.hs = 0;
  const int p = r / 5, dy = r - 5 * p - 2;
  int y, z;
  if (!colNeighbour(y0 + dy, g.ny, g.py, g.oy, g.gy, y, row.sy)) return row;
  if (!colNeighbour(z0 - 2 + p, g.nz, g.pz, g.oz, g.gz, z, row.sz)) return row;
  const int base = g.nx * (y + g.ny * z);
  const int xa = x0 - 2, xb = x0 + 2;
  if (g.px) {
    // whole periodic dimension (o = 0, g = n >= 5): at most one wrap
    if (xa < 0) {
      row.c0[0] = base + xa + g.nx; row.n[0] = -xa; row.sx[0] = -1;
      row.c0[1] = base;             row.n[1] = xb + 1; row.sx[1] = 0;
      row.hs = 1;
    } else if (xb >= g.nx) {
      row.c0[0] = base + xa; row.n[0] = g.nx - xa;      row.sx[0] = 0;
      row.c0[1] = base;      row.n[1] = xb - g.nx + 1;  row.sx[1] = 1;
    } else {
      row.c0[0] = base + xa; row.n[0] = 5;
    }
  } else {
    const int a = xa < 0 ? 0 : xa, b = xb >= g.nx ? g.nx - 1 : xb;
    const int qa = colFloorDiv(g.ox + a, g.gx), qb = colFloorDiv(g.ox + b, g.gx);
    if (qa == qb) {
      row.c0[0] = base + a; row.n[0] = b - a + 1; row.sx[0] = qa;
    } else {
      const int ls = qb * g.gx - g.ox; // first local cell of the upper image
      row.c0[0] = base + a;  row.n[0] = ls - a;     row.sx[0] = qa;
      row.c0[1] = base + ls; row.n[1] = b - ls + 1; row.sx[1] = qb;
      row.hs = x0 >= ls ? 1 : 0;
    }
  }
  return row;
}

// number of half cells per dimension for a cut-off: edge L / n >= (1 + 1e-5) cutOff / 2 (the margin absorbs the
// single precision rounding of the cell assignment)
inline int colCellsFor(double L, double cutOff) {
  const double n = 2.0 * L / (cutOff * 1.00001);
  if (!(n >= 1.0)) return 1;
  if (n > 2.0e9) return 2000000000;
  return (int)n;
}

} // namespace ub200
