/* TEST INFRASTRUCTURE - reference harness for BASELINE config 0: BD::EulerMaruyama, ideal particles, fp64
 * (the README example, README.md:82-106). A tiny main() of OUR OWN that includes the UNMODIFIED reference
 * headers under /root/reference/src and drives BD::EulerMaruyama::forwardTime
 * (Integrator/BrownianDynamics.cu:148-173). Compiled by oracle/Makefile into oracle/_ref/ref_bd with
 * -DDOUBLE_PRECISION. Used by tests/ (-m gpu) as the bit-exactness oracle. Never linked by the product.
 *
 * usage: ref_bd N steps temperature viscosity radius dt sysseed outprefix [pos.bin [force.bin]]
 *   without pos.bin the initial positions are drawn like the README does: sys->rng().uniform3(-0.5, 0.5)
 *   after sys->rng().setSeed(sysseed). With force.bin a constant external force (double4[N]) acts.
 * writes outprefix.pos0.bin / .pos.bin (double4[N]) and prints {"seed": <Saru seed the integrator drew>, ...}
 */
#include "uammd.cuh"
#include "Integrator/BrownianDynamics.cuh"
#include "Interactor/ExternalForces.cuh"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace uammd;

struct ExposeSeed : public BD::EulerMaruyama {
  using BD::EulerMaruyama::EulerMaruyama;
  uint getSeed() const { return seed; }
};

struct ConstantForce {
  real4 *f;
  __device__ ForceEnergyVirial sum(Interactor::Computables comp, int id) {
    ForceEnergyVirial r;
    r.force = make_real3(f[id]);
    r.energy = 0;
    r.virial = 0;
    return r;
  }
  auto getArrays(ParticleData *pd) {
    auto id = pd->getId(access::location::gpu, access::mode::read);
    return std::make_tuple(id.raw());
  }
};

template <class T> static std::vector<T> readBin(const std::string &fn, size_t n) {
  std::vector<T> v(n);
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", fn.c_str()); exit(2); }
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read %s\n", fn.c_str()); exit(2); }
  fclose(f);
  return v;
}
template <class T> static void writeBin(const std::string &fn, const T *p, size_t n) {
  FILE *f = fopen(fn.c_str(), "wb");
  fwrite(p, sizeof(T), n, f);
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc < 9) return 1;
  const int N = atoi(argv[1]), steps = atoi(argv[2]);
  const real T = atof(argv[3]), vis = atof(argv[4]), a = atof(argv[5]), dt = atof(argv[6]);
  const uint64_t sysseed = strtoull(argv[7], nullptr, 10);
  const std::string out = argv[8];
  auto sys = std::make_shared<System>();
  sys->rng().setSeed(sysseed);
  auto pd = std::make_shared<ParticleData>(N, sys);
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::write);
    if (argc > 9) {
      auto hp = readBin<real4>(argv[9], N);
      std::copy(hp.begin(), hp.end(), pos.begin());
    } else {
      std::generate(pos.begin(), pos.end(), [&]() { return make_real4(sys->rng().uniform3(-0.5, 0.5), 0); });
    }
    writeBin(out + ".pos0.bin", pos.raw(), N);
  }
  BD::EulerMaruyama::Parameters par;
  par.temperature = T;
  par.viscosity = vis;
  par.hydrodynamicRadius = a;
  par.dt = dt;
  auto bd = std::make_shared<ExposeSeed>(pd, par);
  thrust::device_vector<real4> dforce;
  if (argc > 10) {
    auto hf = readBin<real4>(argv[10], N);
    dforce = hf;
    auto ext = std::make_shared<ExternalForces<ConstantForce>>(pd, std::make_shared<ConstantForce>(ConstantForce{dforce.data().get()}));
    bd->addInteractor(ext);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, 0);
  for (int i = 0; i < steps; i++) bd->forwardTime();
  cudaDeviceSynchronize();
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::read);
    writeBin(out + ".pos.bin", pos.raw(), N);
  }
  printf("{\"mode\":\"bd\",\"N\":%d,\"steps\":%d,\"seed\":%u,\"ms_per_step\":%.6f}\n", N, steps, bd->getSeed(), ms / steps);
  sys->finish();
  return 0;
}
