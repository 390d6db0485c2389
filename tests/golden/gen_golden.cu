/* Golden-vector generator: runs the UNMODIFIED reference's own __host__ __device__ code on the host of the
 * build container (no GPU needed) and writes small fixtures next to this file. Compile + run:
 *   nvcc -std=c++14 --expt-relaxed-constexpr -I/root/reference/src -I/root/reference/src/third_party \
 *        -Xcompiler -ffp-contract=off tests/golden/gen_golden.cu -o /tmp/gen_golden && /tmp/gen_golden tests/golden
 * Fixtures:
 *   saru.bin      : for 64 seed triples: 8 u32 draws, then f() x2 and gf(0.5,2.0) of a fresh generator
 *   morton.bin    : 4096 random cells (x,y,z < 1024) and their Sorter::MortonHash::hash
 *   getcell_f32.bin : 3 grids x 4096 points: Grid::getCell (float build) incl. points outside the box
 * (double precision fixtures come from gen_golden_f64.cu)
 */
#include "global/defines.h"
#include "utils/vector.cuh"
#include "utils/Box.cuh"
#include "utils/Grid.cuh"
#include "third_party/saruprng.cuh"
#include <cstdio>
#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include "utils/ParticleSorter.cuh"

using namespace uammd;

int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  {
    FILE *f = fopen((dir + "/saru.bin").c_str(), "wb");
    std::mt19937 gen(99);
    for (int t = 0; t < 64; t++) {
      uint32_t s[3] = {(uint32_t)gen(), (uint32_t)gen(), (uint32_t)gen()};
      if (t < 4) { s[0] = t; s[1] = 0; s[2] = t * 7; }
      fwrite(s, 4, 3, f);
      Saru a(s[0], s[1], s[2]);
      uint32_t u[8];
      for (int k = 0; k < 8; k++) u[k] = a.u32();
      fwrite(u, 4, 8, f);
      Saru b(s[0], s[1], s[2]);
      float fl[4];
      fl[0] = b.f(); fl[1] = b.f();
      Saru c(s[0], s[1], s[2]);
      float2 g = c.gf(0.5f, 2.0f);
      fl[2] = g.x; fl[3] = g.y;
      fwrite(fl, 4, 4, f);
    }
    fclose(f);
  }
  {
    FILE *f = fopen((dir + "/morton.bin").c_str(), "wb");
    std::mt19937 gen(5);
    Sorter::MortonHash mh(Grid(Box(1.0), make_int3(8, 8, 8)));
    for (int t = 0; t < 4096; t++) {
      int c[3] = {(int)(gen() % 1024), (int)(gen() % 1024), (int)(gen() % 1024)};
      uint32_t h = mh.hash(make_int3(c[0], c[1], c[2]));
      fwrite(c, 4, 3, f);
      fwrite(&h, 4, 1, f);
    }
    fclose(f);
  }
  {
    FILE *f = fopen((dir + "/getcell_f32.bin").c_str(), "wb");
    std::mt19937_64 gen(17);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    float Ls[3][3] = {{107.7217f, 107.7217f, 107.7217f}, {64.f, 32.f, 7.f}, {29.24f, 50.5f, 110.064f}};
    int cds[3][3] = {{43, 43, 43}, {64, 32, 7}, {11, 20, 110}};
    for (int gidx = 0; gidx < 3; gidx++) {
      Box box(make_real3(Ls[gidx][0], Ls[gidx][1], Ls[gidx][2]));
      Grid grid(box, make_int3(cds[gidx][0], cds[gidx][1], cds[gidx][2]));
      fwrite(Ls[gidx], 4, 3, f);
      fwrite(cds[gidx], 4, 3, f);
      for (int t = 0; t < 4096; t++) {
        // 3/4 inside the box, 1/4 up to 2.5 L outside; a few exactly on faces
        double scale = (t % 4 == 3) ? 2.5 : 0.5;
        float p[3];
        for (int d = 0; d < 3; d++) p[d] = (float)(U(gen) * scale * Ls[gidx][d]);
        if (t % 97 == 0) p[0] = -0.5f * Ls[gidx][0];
        if (t % 89 == 0) p[1] = 0.5f * Ls[gidx][1];
        if (t % 83 == 0) p[2] = Ls[gidx][2] / cds[gidx][2] * (float)(t % cds[gidx][2]) - 0.5f * Ls[gidx][2];
        int3 c = grid.getCell(make_real3(p[0], p[1], p[2]));
        int ci[3] = {c.x, c.y, c.z};
        fwrite(p, 4, 3, f);
        fwrite(ci, 4, 3, f);
      }
    }
    fclose(f);
  }
  return 0;
}
