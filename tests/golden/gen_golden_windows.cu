/* Golden-vector generator for the IBM windows added in round 2 (double precision build): reference host code run in the
 * build container. Compile + run:
 *   nvcc -std=c++14 --expt-relaxed-constexpr -DDOUBLE_PRECISION -I/root/reference/src \
 *        -I/root/reference/src/third_party -Xcompiler -ffp-contract=off tests/golden/gen_golden_windows.cu \
 *        -o oracle/_ref/gen_golden_windows && oracle/_ref/gen_golden_windows tests/golden bm
 * and, on a machine with a GPU (sixPoint's constructor fills a device-side table; the binary travels with the snapshot):
 *   oracle/_ref/gen_golden_windows gpurun_out six      -> copy gpurun_out/windows_six_f64.bin to tests/golden/
 * Fixtures:
 *   windows_bm_f64.bin : 3 Barnett-Magland kernels (alpha, beta): alpha, beta, phi(0) = 1/norm, then phi at 129 radii in
 *                        [-1.05 alpha, 1.05 alpha]
 *   windows_six_f64.bin: sixPoint (h = 0.7): h, then r and phi(r) at 801 radii in [-3.3 h, 3.3 h]
 */
#include "global/defines.h"
#include "utils/vector.cuh"
#include "misc/IBM_kernels.cuh"
#include <cstdio>
#include <string>
using namespace uammd;
int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  std::string what = argc > 2 ? argv[2] : "bm";
  if (what == "bm") {
  FILE *f = fopen((dir + "/windows_bm_f64.bin").c_str(), "wb");
  const double alphas[3] = {2.0, 3.0 * 0.7, 1.5}, betas[3] = {1.8 * 4, 1.714 * 6, 5.0};
  for (int k = 0; k < 3; k++) {
    IBM_kernels::BarnettMagland bm(alphas[k], betas[k]);
    double head[3] = {alphas[k], betas[k], bm.phi(0.0)};
    fwrite(head, 8, 3, f);
    for (int i = 0; i <= 128; i++) {
      double r = (-1.05 + 2.1 * i / 128.0) * alphas[k];
      double v = bm.phi(r);
      fwrite(&v, 8, 1, f);
    }
  }
  fclose(f);
  } else {
  FILE *f = fopen((dir + "/windows_six_f64.bin").c_str(), "wb");
  const double h = 0.7;
  IBM_kernels::GaussianFlexible::sixPoint six(h);
  fwrite(&h, 8, 1, f);
  for (int i = 0; i <= 800; i++) {
    double r = (-3.3 + 6.6 * i / 800.0) * h;
    double v[2] = {r, six.phi(r)};
    fwrite(v, 8, 2, f);
  }
  fclose(f);
  }
  return 0;
}
