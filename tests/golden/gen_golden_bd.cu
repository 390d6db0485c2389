/* Golden-vector generator for BASELINE config 0 (BD::EulerMaruyama, README example): runs the UNMODIFIED
 * reference's host code in the build container. Compile + run:
 *   nvcc -std=c++14 --expt-relaxed-constexpr -I/root/reference/src -I/root/reference/src/third_party \
 *        -Xcompiler -ffp-contract=off tests/golden/gen_golden_bd.cu -o /tmp/gen_golden_bd && /tmp/gen_golden_bd tests/golden
 * Fixture xorshift_bd.bin:
 *   u64 seed; for each of 3 seeds {1234, 0xdeadbeef, default ctor}: 16 x next32 (u32), 48 x uniform(-0.5,0.5) (f64)
 *   then, for seed 1234 after 3e5 uniform draws (the README's 1e5 uniform3 calls): next32 x3 (the third is the
 *   integrator's Saru seed, BrownianDynamics.cu:14-16)
 *   then 64 particles: Saru(i, step=5, seed) -> gf(0, B) twice (float4) with the HOST libm (what the C oracle uses)
 */
#include "global/defines.h"
#include "utils/vector.cuh"
#include "utils/utils.h"
#include "third_party/saruprng.cuh"
#include <cstdio>
#include <string>
using namespace uammd;
int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  FILE *f = fopen((dir + "/xorshift_bd.bin").c_str(), "wb");
  for (int t = 0; t < 3; t++) {
    Xorshift128plus r;
    if (t == 0) r.setSeed(1234);
    if (t == 1) r.setSeed(0xdeadbeefULL);
    for (int k = 0; k < 16; k++) { uint32_t v = r.next32(); fwrite(&v, 4, 1, f); }
    for (int k = 0; k < 48; k++) { double v = r.uniform(-0.5, 0.5); fwrite(&v, 8, 1, f); }
  }
  Xorshift128plus r;
  r.setSeed(1234);
  for (int k = 0; k < 100000; k++) r.uniform3(-0.5, 0.5);
  uint32_t s3[3] = {r.next32(), r.next32(), r.next32()};
  fwrite(s3, 4, 3, f);
  const float B = (float)sqrt(2.0 * 1.0 * (1.0 / (6.0 * M_PI)) * 0.1);
  for (int i = 0; i < 64; i++) {
    Saru rng(i, 5, s3[2]);
    float2 a = rng.gf(0, B), b = rng.gf(0, B);
    float o[4] = {a.x, a.y, b.x, b.y};
    fwrite(o, 4, 4, f);
  }
  fclose(f);
  return 0;
}
