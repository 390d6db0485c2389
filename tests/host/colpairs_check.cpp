// Host emulation of the column traversal's data path (uammd_b200/csrc/lj_column.cu + colgeom.h): bin random particles on
// the half-cell grid with canonical coordinates, stage every column row by row with the image shifts, and check that
// the in-range pair set of every home particle equals the brute-force minimum-image pair set.
#include "../../include/uammd_b200/colgeom.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
using namespace ub200;

struct P { float x, y, z; int id; };

static void canonical(float r, float L, int per, int n, int &c, float &rf) {
  rf = r;
  if (per) rf = r + std::floor(r * (-1.0f / L) + 0.5f) * L;
  c = (int)((rf + 0.5f * L) * (1.0f / (L / n)));
  if (per) { if (c >= n) { c -= n; rf -= L; } else if (c < 0) { c += n; rf += L; } }
  c = std::min(std::max(c, 0), n - 1);
}

static int run(const float L[3], const int per[3], float rc, int N, int TZ, unsigned seed) {
  ColGrid g;
  int dims[3];
  for (int d = 0; d < 3; d++) dims[d] = colCellsFor(L[d], rc);
  g = makeWholeColGrid(dims, per);
  for (int d = 0; d < 3; d++) if (per[d] && dims[d] < 5) return 0;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> U(-0.5f, 0.5f);
  std::vector<P> raw(N);
  for (int i = 0; i < N; i++) {
    raw[i] = {U(rng) * L[0], U(rng) * L[1], U(rng) * L[2], i};
    if (i % 7 == 0) for (int d = 0; d < 3; d++) if (per[d]) (&raw[i].x)[d] += L[d] * (float)((int)(rng() % 5) - 2); // outside the primary box
  }
  if (per[0]) raw[0].x = 0.5f * L[0]; // the +L/2 corner case of Grid::getCell
  if (per[1]) raw[1].y = -0.5f * L[1];
  const int ncells = dims[0] * dims[1] * dims[2];
  std::vector<std::vector<P>> cells(ncells);
  for (int i = 0; i < N; i++) {
    int c[3]; float f[3];
    for (int d = 0; d < 3; d++) canonical((&raw[i].x)[d], L[d], per[d], dims[d], c[d], f[d]);
    cells[c[0] + dims[0] * (c[1] + dims[1] * c[2])].push_back({f[0], f[1], f[2], i});
  }
  // brute force
  std::vector<std::vector<int>> want(N), got(N);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      if (i == j) continue;
      double d2 = 0;
      for (int d = 0; d < 3; d++) {
        double dx = (double)(&raw[j].x)[d] - (double)(&raw[i].x)[d];
        if (per[d]) dx -= std::floor(dx / L[d] + 0.5) * L[d];
        d2 += dx * dx;
      }
      if (d2 < (double)rc * rc * (1 - 1e-5)) want[i].push_back(j); // exclude pairs within rounding of the cut-off
    }
  std::vector<std::vector<int>> maybe(N); // pairs within rounding of the cut-off may go either way
  int bad = 0;
  const int nzc = (g.nz + TZ - 1) / TZ;
  for (int zc = 0; zc < nzc; zc++)
    for (int y0 = 0; y0 < g.ny; y0++)
      for (int x0 = 0; x0 < g.nx; x0++) {
        const int z0 = zc * TZ, nHome = std::min(TZ, g.nz - z0), nRows = 5 * (nHome + 4);
        std::vector<P> slice;
        std::vector<int> planeOff(nHome + 5, 0);
        for (int r = 0; r < nRows; r++) {
          if (r % 5 == 0) planeOff[r / 5] = (int)slice.size();
          const ColRow row = columnRow(g, x0, y0, z0, r);
          for (int s = 0; s < 2; s++)
            for (int k = 0; k < row.n[s]; k++)
              for (const P &p : cells[row.c0[s] + k])
                slice.push_back({p.x + row.sx[s] * L[0], p.y + row.sy * L[1], p.z + row.sz * L[2], p.id});
        }
        planeOff[nHome + 4] = (int)slice.size();
        for (int hz = 0; hz < nHome; hz++)
          for (const P &h : cells[x0 + g.nx * (y0 + g.ny * (z0 + hz))])
            for (int t = planeOff[hz]; t < planeOff[hz + 5]; t++) {
              const P &c = slice[t];
              const float dx = c.x - h.x, dy = c.y - h.y, dz = c.z - h.z, r2 = dz * dz + (dy * dy + dx * dx);
              if (c.id != h.id && r2 < rc * rc) got[h.id].push_back(c.id);
              if (c.id == h.id && r2 != 0.0f) { bad++; if (bad < 5) fprintf(stderr, "self image seen at r2=%g\n", r2); }
            }
      }
  for (int i = 0; i < N; i++) {
    std::sort(got[i].begin(), got[i].end());
    std::sort(want[i].begin(), want[i].end());
    // every wanted neighbour present, no duplicates, extras only within rounding of the cut-off
    if (std::adjacent_find(got[i].begin(), got[i].end()) != got[i].end()) { bad++; if (bad < 5) fprintf(stderr, "duplicate neighbour of %d\n", i); }
    if (!std::includes(got[i].begin(), got[i].end(), want[i].begin(), want[i].end())) { bad++; if (bad < 5) fprintf(stderr, "missing neighbour of %d (%zu vs %zu)\n", i, got[i].size(), want[i].size()); }
    if (got[i].size() > want[i].size() + 2) { bad++; if (bad < 5) fprintf(stderr, "spurious neighbours of %d (%zu vs %zu)\n", i, got[i].size(), want[i].size()); }
  }
  return bad;
}

// Brick decomposition: rank (kx,ky,kz) owns the half cells [floor(k g/p), floor((k+1) g/p)) of every decomposed
// dimension and works on the window brick +- 2 cells; every owned particle must see exactly the brute-force pair set.
static int runBricks(const float L[3], const int per[3], const int rankGrid[3], float rc, int N, int TZ, unsigned seed) {
  int dims[3];
  for (int d = 0; d < 3; d++) dims[d] = colCellsFor(L[d], rc);
  for (int d = 0; d < 3; d++) if (per[d] && dims[d] < 5) return 0;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> U(-0.5f, 0.5f);
  std::vector<P> raw(N);
  for (int i = 0; i < N; i++) raw[i] = {U(rng) * L[0], U(rng) * L[1], U(rng) * L[2], i};
  if (per[0]) raw[0].x = 0.5f * L[0];
  std::vector<int> cell3(3 * N);
  std::vector<P> canon(N);
  for (int i = 0; i < N; i++) {
    float f[3];
    for (int d = 0; d < 3; d++) canonical((&raw[i].x)[d], L[d], per[d], dims[d], cell3[3 * i + d], f[d]);
    canon[i] = {f[0], f[1], f[2], i};
  }
  std::vector<std::vector<int>> want(N), got(N);
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      if (i == j) continue;
      double d2 = 0;
      for (int d = 0; d < 3; d++) {
        double dx = (double)(&raw[j].x)[d] - (double)(&raw[i].x)[d];
        if (per[d]) dx -= std::floor(dx / L[d] + 0.5) * L[d];
        d2 += dx * dx;
      }
      if (d2 < (double)rc * rc * (1 - 1e-5)) want[i].push_back(j);
    }
  int bad = 0;
  const int world = rankGrid[0] * rankGrid[1] * rankGrid[2];
  for (int rank = 0; rank < world; rank++) {
    const int k[3] = {rank % rankGrid[0], (rank / rankGrid[0]) % rankGrid[1], rank / (rankGrid[0] * rankGrid[1])};
    int lo[3], hi[3], o[3], w[3], lp[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = (k[d] * dims[d]) / rankGrid[d]; hi[d] = ((k[d] + 1) * dims[d]) / rankGrid[d];
      if (rankGrid[d] == 1) { o[d] = 0; w[d] = dims[d]; lp[d] = per[d]; }
      else { o[d] = lo[d] - 2; w[d] = hi[d] - lo[d] + 4; lp[d] = 0; if (w[d] > dims[d]) return 0; }
    }
    ColGrid g{w[0], w[1], w[2], lp[0], lp[1], lp[2], o[0], o[1], o[2], dims[0], dims[1], dims[2], rankGrid[0] > 1, rankGrid[1] > 1, rankGrid[2] > 1};
    std::vector<std::vector<P>> cells((size_t)w[0] * w[1] * w[2]);
    std::vector<char> owned(N, 0);
    for (int i = 0; i < N; i++) {
      int l[3]; bool in = true, own = true;
      for (int d = 0; d < 3; d++) {
        int c = cell3[3 * i + d];
        own = own && c >= lo[d] && c < hi[d];
        l[d] = c - o[d];
        if (l[d] < 0) l[d] += dims[d];
        if (l[d] >= dims[d]) l[d] -= dims[d];
        if (l[d] >= w[d]) { if (per[d] || rankGrid[d] == 1) in = false; else in = false; }
        if (!per[d] && rankGrid[d] > 1) { // non periodic decomposed dimension: no wrap into the window
          const int u = c - o[d];
          if (u < 0 || u >= w[d]) in = false; else l[d] = u;
        }
      }
      if (!in) continue;
      owned[i] = own;
      cells[l[0] + (size_t)w[0] * (l[1] + (size_t)w[1] * l[2])].push_back(canon[i]);
    }
    const int nzc = (g.nz + TZ - 1) / TZ;
    for (int zc = 0; zc < nzc; zc++)
      for (int y0 = 0; y0 < g.ny; y0++)
        for (int x0 = 0; x0 < g.nx; x0++) {
          const int z0 = zc * TZ, nHome = std::min(TZ, g.nz - z0), nRows = 5 * (nHome + 4);
          std::vector<P> slice;
          std::vector<int> planeOff(nHome + 5, 0);
          for (int r = 0; r < nRows; r++) {
            if (r % 5 == 0) planeOff[r / 5] = (int)slice.size();
            const ColRow row = columnRow(g, x0, y0, z0, r);
            for (int s = 0; s < 2; s++)
              for (int kk = 0; kk < row.n[s]; kk++)
                for (const P &p : cells[row.c0[s] + kk])
                  slice.push_back({p.x + row.sx[s] * L[0], p.y + row.sy * L[1], p.z + row.sz * L[2], p.id});
          }
          planeOff[nHome + 4] = (int)slice.size();
          for (int hz = 0; hz < nHome; hz++)
            for (const P &h : cells[x0 + (size_t)g.nx * (y0 + (size_t)g.ny * (z0 + hz))]) {
              if (!owned[h.id]) continue;
              for (int t = planeOff[hz]; t < planeOff[hz + 5]; t++) {
                const P &c = slice[t];
                const float dx = c.x - h.x, dy = c.y - h.y, dz = c.z - h.z, r2 = dz * dz + (dy * dy + dx * dx);
                if (c.id != h.id && r2 < rc * rc) got[h.id].push_back(c.id);
              }
            }
        }
  }
  for (int i = 0; i < N; i++) {
    std::sort(got[i].begin(), got[i].end());
    std::sort(want[i].begin(), want[i].end());
    if (std::adjacent_find(got[i].begin(), got[i].end()) != got[i].end()) { bad++; if (bad < 5) fprintf(stderr, "brick: duplicate neighbour of %d\n", i); }
    if (!std::includes(got[i].begin(), got[i].end(), want[i].begin(), want[i].end())) { bad++; if (bad < 5) fprintf(stderr, "brick: missing neighbour of %d (%zu vs %zu)\n", i, got[i].size(), want[i].size()); }
    if (got[i].size() > want[i].size() + 2) { bad++; if (bad < 5) fprintf(stderr, "brick: spurious neighbours of %d\n", i); }
  }
  return bad;
}

int main() {
  int bad = 0, n = 0;
  {
    const float Lb[3] = {20.f, 17.f, 24.f};
    const int grids[][3] = {{2, 1, 1}, {2, 2, 2}, {1, 3, 2}, {4, 1, 2}, {1, 1, 3}};
    for (auto &rg : grids)
      for (int pm : {7, 5, 0}) {
        const int per[3] = {pm & 1, (pm >> 1) & 1, (pm >> 2) & 1};
        bad += runBricks(Lb, per, rg, 2.5f, 3000, 6, 99 + pm);
        n++;
      }
  }
  const float boxes[][3] = {{12.f, 12.f, 12.f}, {7.f, 9.5f, 21.f}, {30.f, 7.f, 26.f}, {6.3f, 6.3f, 6.3f}, {16.f, 16.f, 2.f}};
  for (auto &b : boxes)
    for (int pm = 0; pm < 8; pm++) {
      const int per[3] = {pm & 1, (pm >> 1) & 1, (pm >> 2) & 1};
      for (int TZ : {4, 6}) {
        const double vol = (double)b[0] * b[1] * b[2];
        bad += run(b, per, 2.5f, (int)std::min(3000.0, vol * 0.8), TZ, 17 + pm);
        n++;
      }
    }
  printf("%d configurations, %d errors\n", n, bad);
  return bad ? 1 : 0;
}
