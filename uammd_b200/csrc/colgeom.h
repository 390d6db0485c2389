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
#pragma once

#ifdef __CUDACC__
#define UB200_HD __host__ __device__ __forceinline__
#else
#define UB200_HD inline
#endif

namespace ub200 {

struct ColGrid {
  int nx, ny, nz;       // half cells per dimension
  int px, py, pz;       // periodic flags
};

// One staged row: up to two x segments of consecutive cells. c0[s] = linear index of the first cell, n[s] = number of
// cells (0 = absent), sx[s] = image shift of the segment in box lengths; sy, sz = image shift of the row.
struct ColRow {
  int c0[2], n[2], sx[2];
  int sy, sz;
};

// v -> wrapped index in [0, n) and the number of box lengths the unwrapped cell lies away (image shift).
// Returns false when v is outside a non periodic dimension.
UB200_HD bool colWrap(int v, int n, int periodic, int &w, int &shift) {
  shift = 0;
  w = v;
  if (v >= 0 && v < n) return true;
  if (!periodic) return false;
  // one box length away in all but degenerate cases (no division on the common path)
  if (v < 0) { w = v + n; shift = -1; } else { w = v - n; shift = 1; }
  if (w >= 0 && w < n) return true;
  int q = v / n;
  if (v - q * n < 0) q--; // floor division
  shift = q;
  w = v - q * n;
  return true;
}

// Row r (0 <= r < 5 * nPlanes; plane p = r / 5, dy = r % 5 - 2) of the column with first home cell (x0, y0, z0).
UB200_HD ColRow columnRow(const ColGrid &g, int x0, int y0, int z0, int r) {
  ColRow row;
  row.c0[0] = row.c0[1] = 0;
  row.n[0] = row.n[1] = 0;
  row.sx[0] = row.sx[1] = 0;
  row.sy = row.sz = 0;
  const int p = r / 5, dy = r - 5 * p - 2;
  int y, z;
  if (!colWrap(y0 + dy, g.ny, g.py, y, row.sy)) return row;
  if (!colWrap(z0 - 2 + p, g.nz, g.pz, z, row.sz)) return row;
  const int base = g.nx * (y + g.ny * z);
  const int xa = x0 - 2, xb = x0 + 2;
  if (g.px) {
    // callers guarantee nx >= 5 in a periodic dimension: at most one wrap
    if (xa < 0) {
      row.c0[0] = base + xa + g.nx; row.n[0] = -xa; row.sx[0] = -1;
      row.c0[1] = base;             row.n[1] = xb + 1; row.sx[1] = 0;
    } else if (xb >= g.nx) {
      row.c0[0] = base + xa; row.n[0] = g.nx - xa;      row.sx[0] = 0;
      row.c0[1] = base;      row.n[1] = xb - g.nx + 1;  row.sx[1] = 1;
    } else {
      row.c0[0] = base + xa; row.n[0] = 5;
    }
  } else {
    const int a = xa < 0 ? 0 : xa, b = xb >= g.nx ? g.nx - 1 : xb;
    row.c0[0] = base + a; row.n[0] = b - a + 1;
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
