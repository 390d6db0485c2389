/* TEST INFRASTRUCTURE - reference harness for VerletNVT::GronbechJensen (the integrator that
 * generic_md and examples/misc/benchmark.cu instantiate). A tiny main() of OUR OWN that includes the UNMODIFIED
 * reference headers under /root/reference/src and drives VerletNVT::GronbechJensen::forwardTime
 * (Integrator/VerletNVT/GronbechJensen.cu:96-127), optionally with PairForces<Potential::LJ, VerletList> like benchmark.cu.
 * Compiled by oracle/Makefile into oracle/_ref/ref_nvt (single precision). Never linked by the product.
 *
 * usage: ref_nvt N L steps temperature friction dt sysseed lj(0|1) initVelocities(0|1) outprefix [pos.bin vel.bin [rcutmult warmup]]
 *   positions / velocities: float4[N] / float3[N] files, or (without them) a jittered lattice and zero velocities.
 *   rcutmult: VerletList::setCutOffMultiplier (benchmark.cu uses 1.2; default 1.08); warmup: untimed steps before `steps`.
 * writes outprefix.pos0.bin .vel0.bin (state after construction, i.e. after initVelocities) and .pos.bin .vel.bin
 * (after `steps` steps); prints {"seed": <Saru seed the integrator drew>, "ms_per_step": ...}
 */
#include "uammd.cuh"
#include "Integrator/VerletNVT.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "Interactor/NeighbourList/VerletList.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using namespace uammd;

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

/* The reference declares Basic(shared_ptr<ParticleGroup>, Parameters) (VerletNVT.cuh:92) but defines only the protected
   three-argument constructor (Basic.cu:31-52): a program that constructs VerletNVT::Basic does not link. A derived class
   reaches the protected constructor; everything that runs is the unmodified Basic. */
struct BasicExposed : VerletNVT::Basic {
  BasicExposed(std::shared_ptr<ParticleData> pd, VerletNVT::Basic::Parameters par)
      : VerletNVT::Basic(std::make_shared<ParticleGroup>(pd, "All"), par, "VerletNVT::Basic") {}
};

int main(int argc, char **argv) {
  if (argc < 11) return 1;
  const int N = atoi(argv[1]);
  const real L = atof(argv[2]);
  const int steps = atoi(argv[3]);
  const real T = atof(argv[4]), friction = atof(argv[5]), dt = atof(argv[6]);
  const uint64_t sysseed = strtoull(argv[7], nullptr, 10);
  const bool lj = atoi(argv[8]) != 0, initVel = atoi(argv[9]) != 0;
  const std::string out = argv[10];
  auto sys = std::make_shared<System>();
  sys->rng().setSeed(sysseed);
  auto pd = std::make_shared<ParticleData>(N, sys);
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::write);
    auto vel = pd->getVel(access::location::cpu, access::mode::write);
    if (argc > 12) {
      auto hp = readBin<real4>(argv[11], N);
      auto hv = readBin<real3>(argv[12], N);
      std::copy(hp.begin(), hp.end(), pos.begin());
      std::copy(hv.begin(), hv.end(), vel.begin());
    } else {
      std::mt19937_64 gen(2024);
      std::uniform_real_distribution<double> U(-0.5, 0.5);
      const int n = (int)std::ceil(std::cbrt((double)N));
      const double a = L / n;
      for (int i = 0; i < N; i++) {
        const int ix = i % n, iy = (i / n) % n, iz = i / (n * n);
        pos[i] = make_real4((ix + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), (iy + 0.5) * a - 0.5 * L + 0.2 * a * U(gen),
                            (iz + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), 0);
        vel[i] = make_real3(0);
      }
    }
  }
  // the integrator's constructor draws its Saru seed as the third next32() of the system generator (Basic.cu:36-38)
  auto rngCopy = sys->rng();
  rngCopy.next32(); rngCopy.next32();
  const uint seed = rngCopy.next32();
  const uint velSeed = rngCopy.next32(); // initVelocities draws the next one (Basic.cu:74-76)
  VerletNVT::GronbechJensen::Parameters par;
  par.temperature = T;
  par.dt = dt;
  par.friction = friction;
  par.initVelocities = initVel;
  // REF_NVT_SCHEME=basic drives VerletNVT::Basic (Integrator/VerletNVT/Basic.cu:87-172), the class GronbechJensen derives from
  const char *scheme = getenv("REF_NVT_SCHEME");
  const bool basic = scheme && std::string(scheme) == "basic";
  std::shared_ptr<VerletNVT::Basic> nvt;
  if (basic) nvt = std::make_shared<BasicExposed>(pd, par);
  else nvt = std::make_shared<VerletNVT::GronbechJensen>(pd, par);
  CudaSafeCall(cudaDeviceSynchronize());
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::read);
    auto vel = pd->getVel(access::location::cpu, access::mode::read);
    writeBin(out + ".pos0.bin", pos.raw(), N);
    writeBin(out + ".vel0.bin", vel.raw(), N);
  }
  if (lj) {
    Potential::LJ::InputPairParameters p;
    p.epsilon = 1.0; p.sigma = 1.0; p.cutOff = 2.5; p.shift = false;
    auto pot = std::make_shared<Potential::LJ>();
    pot->setPotParameters(0, 0, p);
    using PF = PairForces<Potential::LJ, VerletList>;
    PF::Parameters pp;
    pp.box = Box(make_real3(L));
    if (argc > 13) {
      auto nl = std::make_shared<VerletList>(pd);
      nl->setCutOffMultiplier(atof(argv[13]));
      pp.nl = nl;
    }
    nvt->addInteractor(std::make_shared<PF>(pd, pp, pot));
  }
  const int warmup = argc > 14 ? atoi(argv[14]) : 0;
  for (int i = 0; i < warmup; i++) nvt->forwardTime();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, 0);
  for (int i = 0; i < steps; i++) nvt->forwardTime();
  cudaDeviceSynchronize();
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::read);
    auto vel = pd->getVel(access::location::cpu, access::mode::read);
    writeBin(out + ".pos.bin", pos.raw(), N);
    writeBin(out + ".vel.bin", vel.raw(), N);
  }
  printf("{\"mode\":\"%s\",\"N\":%d,\"steps\":%d,\"seed\":%u,\"vel_seed\":%u,\"ms_per_step\":%.6f}\n", basic ? "nvt_basic" : "nvt_gj", N, steps, seed, velSeed,
         steps ? ms / steps : 0.f);
  sys->finish();
  return 0;
}
