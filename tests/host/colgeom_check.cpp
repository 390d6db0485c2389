// Host check of include/uammd_b200/colgeom.h: for every home half cell of every column of a set of grids, the cells (and
// image shifts) reached through planes hz .. hz+4 of the column's staged rows must be exactly the 5 x 5 x 5 stencil
// around the home cell (each periodic image once, nothing outside a non periodic box), in ascending (z, y, x) order.
#include "../../include/uammd_b200/colgeom.h"
#include <cstdio>
#include <cstdlib>
#include <tuple>
#include <vector>
using namespace ub200;

struct Img { int cell, sx, sy, sz; bool operator==(const Img &o) const { return cell == o.cell && sx == o.sx && sy == o.sy && sz == o.sz; } };

static int check(ColGrid g, int TZ) {
  int bad = 0;
  const int nzc = (g.nz + TZ - 1) / TZ;
  for (int zc = 0; zc < nzc; zc++)
    for (int y0 = 0; y0 < g.ny; y0++)
      for (int x0 = 0; x0 < g.nx; x0++) {
        const int z0 = zc * TZ;
        const int nHome = std::min(TZ, g.nz - z0);
        const int nRows = 5 * (nHome + 4);
        std::vector<std::vector<Img>> rows(nRows);
        for (int r = 0; r < nRows; r++) {
          const ColRow row = columnRow(g, x0, y0, z0, r);
          for (int s = 0; s < 2; s++)
            for (int k = 0; k < row.n[s]; k++) rows[r].push_back({row.c0[s] + k, row.sx[s], row.sy, row.sz});
        }
        for (int hz = 0; hz < nHome; hz++) {
          std::vector<Img> got, want;
          for (int r = 5 * hz; r < 5 * hz + 25; r++) got.insert(got.end(), rows[r].begin(), rows[r].end());
          for (int dz = -2; dz <= 2; dz++)
            for (int dy = -2; dy <= 2; dy++)
              for (int dx = -2; dx <= 2; dx++) {
                int c[3] = {x0 + dx, y0 + dy, z0 + hz + dz}, n[3] = {g.nx, g.ny, g.nz}, p[3] = {g.px, g.py, g.pz}, sh[3] = {0, 0, 0};
                bool ok = true;
                for (int d = 0; d < 3; d++) {
                  while (c[d] < 0) { if (!p[d]) { ok = false; break; } c[d] += n[d]; sh[d]--; }
                  while (ok && c[d] >= n[d]) { if (!p[d]) { ok = false; break; } c[d] -= n[d]; sh[d]++; }
                  if (!ok) break;
                }
                if (ok) want.push_back({c[0] + g.nx * (c[1] + g.ny * c[2]), sh[0], sh[1], sh[2]});
              }
          if (!(got == want)) {
            if (bad < 5) fprintf(stderr, "mismatch grid %dx%dx%d per %d%d%d col (%d,%d,%d) hz %d: got %zu want %zu\n", g.nx, g.ny, g.nz, g.px, g.py, g.pz, x0, y0, z0, hz, got.size(), want.size());
            bad++;
          }
        }
        // the home cell's own row: dy == 0 of plane hz + 2 must contain cell (x0, y0, z0 + hz) with zero shift
        for (int hz = 0; hz < nHome; hz++) {
          const ColRow row = columnRow(g, x0, y0, z0, 5 * (hz + 2) + 2);
          const int cc = x0 + g.nx * (y0 + g.ny * (z0 + hz));
          const int seg = row.sx[0] < 0 ? 1 : 0;
          if (seg != row.hs) { bad++; if (bad < 5) fprintf(stderr, "hs mismatch\n"); }
          if (!(cc >= row.c0[seg] && cc < row.c0[seg] + row.n[seg] && row.sx[seg] == 0 && row.sy == 0 && row.sz == 0)) {
            if (bad < 5) fprintf(stderr, "home cell not in segment %d: grid %dx%dx%d col (%d,%d,%d) hz %d\n", seg, g.nx, g.ny, g.nz, x0, y0, z0, hz);
            bad++;
          }
        }
      }
  return bad;
}

int main() {
  int bad = 0, grids = 0;
  const int dims[][3] = {{5, 5, 5}, {6, 5, 7}, {8, 9, 13}, {12, 7, 6}, {5, 11, 20}, {1, 5, 9}, {3, 2, 1}, {16, 16, 1}, {7, 1, 4}};
  for (auto &d : dims)
    for (int per = 0; per < 8; per++) {
      const int pp[3] = {per & 1, (per >> 1) & 1, (per >> 2) & 1};
      ColGrid g = makeWholeColGrid(d, pp);
      if ((g.px && g.nx < 5) || (g.py && g.ny < 5) || (g.pz && g.nz < 5)) continue; // callers never build such grids
      for (int TZ : {1, 4, 6, 8}) { bad += check(g, TZ); grids++; }
    }
  if (colCellsFor(107.7217, 2.5) != 86 || colCellsFor(10.0, 2.5) != 7 || colCellsFor(0.0, 2.5) != 1) { fprintf(stderr, "colCellsFor\n"); bad++; }
  printf("%d grids checked, %d mismatches\n", grids, bad);
  return bad ? 1 : 0;
}
