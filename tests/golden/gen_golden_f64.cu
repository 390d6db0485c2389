/* Golden-vector generator, double precision build (-DDOUBLE_PRECISION): reference host code run in the build
 * container. Compile + run:
 *   nvcc -std=c++14 --expt-relaxed-constexpr -DDOUBLE_PRECISION -I/root/reference/src \
 *        -I/root/reference/src/third_party -Xcompiler -ffp-contract=off tests/golden/gen_golden_f64.cu \
 *        -o /tmp/gen_golden_f64 && /tmp/gen_golden_f64 tests/golden
 * Fixtures:
 *   peskin_f64.bin   : 801 radii r in [-2.2h, 2.2h] (h = 0.7): r, Peskin::threePoint::phi, fourPoint::phi
 *   gaussian_f64.bin : for 6 (h, tol) pairs: h, tol, support, rmax, hydrodynamic radius a, then phi at 65 radii
 *   getcell_f64.bin  : 2 grids x 2048 points: Grid::getCell in double
 */
#include "global/defines.h"
#include "utils/vector.cuh"
#include "utils/Box.cuh"
#include "utils/Grid.cuh"
#include "misc/IBM_kernels.cuh"
#include "Integrator/BDHI/FCM/FCM_kernels.cuh"
#include <cstdio>
#include <random>
#include <string>
using namespace uammd;
int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  {
    FILE *f = fopen((dir + "/peskin_f64.bin").c_str(), "wb");
    const double h = 0.7;
    IBM_kernels::Peskin::threePoint p3(h);
    IBM_kernels::Peskin::fourPoint p4(h);
    for (int i = 0; i <= 800; i++) {
      double r = (-2.2 + 4.4 * i / 800.0) * h;
      double v[3] = {r, p3.phi(r), p4.phi(r)};
      fwrite(v, 8, 3, f);
    }
    fclose(f);
  }
  {
    FILE *f = fopen((dir + "/gaussian_f64.bin").c_str(), "wb");
    const double hs[6] = {1.0, 0.5, 1.0, 0.83, 1.0, 2.0};
    const double tols[6] = {1e-3, 1e-5, 1e-8, 1e-6, 1e-2, 1e-4};
    for (int k = 0; k < 6; k++) {
      BDHI::FCM_ns::Kernels::Gaussian g(hs[k], tols[k]);
      double head[5] = {hs[k], tols[k], (double)g.support, g.rmax, g.fixHydrodynamicRadius(0, hs[k])};
      fwrite(head, 8, 5, f);
      for (int i = 0; i < 65; i++) {
        double r = g.rmax * 1.05 * i / 64.0;
        double v = g.phi(r, real3());
        fwrite(&v, 8, 1, f);
      }
    }
    fclose(f);
  }
  {
    FILE *f = fopen((dir + "/getcell_f64.bin").c_str(), "wb");
    std::mt19937_64 gen(23);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    double Ls[2][3] = {{128., 128., 128.}, {64., 32., 7.}};
    int cds[2][3] = {{128, 128, 128}, {64, 32, 7}};
    for (int gi = 0; gi < 2; gi++) {
      Grid grid(Box(make_real3(Ls[gi][0], Ls[gi][1], Ls[gi][2])), make_int3(cds[gi][0], cds[gi][1], cds[gi][2]));
      fwrite(Ls[gi], 8, 3, f);
      fwrite(cds[gi], 4, 3, f);
      for (int t = 0; t < 2048; t++) {
        double scale = (t % 4 == 3) ? 2.5 : 0.5;
        double p[3];
        for (int d = 0; d < 3; d++) p[d] = U(gen) * scale * Ls[gi][d];
        if (t % 97 == 0) p[0] = -0.5 * Ls[gi][0];
        if (t % 89 == 0) p[1] = 0.5 * Ls[gi][1];
        int3 c = grid.getCell(make_real3(p[0], p[1], p[2]));
        int ci[3] = {c.x, c.y, c.z};
        fwrite(p, 8, 3, f);
        fwrite(ci, 4, 3, f);
      }
    }
    fclose(f);
  }
  return 0;
}
