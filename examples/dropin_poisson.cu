/* b200::Poisson next to the UNMODIFIED reference Poisson (Interactor/SpectralEwaldPoisson.cuh) on the same charges, double
 * precision like the reference's test build:
 *   1. test/Potentials/Poisson/TriplyPeriodic/test_poisson.cu:192-222 (SingleSimulationTest): three charges, L = 100,
 *      r = 2, tolerance 1e-7, gw = 0.001, split 0.2: force and field on charge 0 against the analytic answer (1e-3) - both.
 *   2. a neutral cloud of N random unit charges: forces, energies and computeFieldPotentialAtParticles() of the two
 *      implementations against each other (largest difference in units of the largest value).
 * Built by oracle/Makefile into oracle/_ref/dropin_poisson (links cuFFT for the reference side only); run by
 * tests/test_dropin_gpu.py.
 */
#include "uammd.cuh"
#include "Interactor/SpectralEwaldPoisson.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

static real theoreticalField(real r, real gw) {
  const double pi = M_PI;
  return -exp(-r * r / (4.0 * gw * gw)) / (4 * pi * sqrt(pi) * gw * r) - erf(r / (2.0 * gw)) / (4 * pi * r * r);
}

template <class P> static void fill(typename P::Parameters &par, real L, real tol, real gw, real split, real eps) {
  par.box = Box(L); par.epsilon = eps; par.gw = gw; par.tolerance = tol; par.split = split;
}

struct Result {
  std::vector<real4> force, fp;
  std::vector<real> energy;
};
template <class P> static Result run(std::shared_ptr<System> sys, const std::vector<real4> &pos, const std::vector<real> &q, real L,
                                     real tol, real gw, real split, real eps) {
  const int N = (int)pos.size();
  auto pd = std::make_shared<ParticleData>(N, sys);
  {
    auto p = pd->getPos(access::location::cpu, access::mode::write);
    auto c = pd->getCharge(access::location::cpu, access::mode::write);
    auto f = pd->getForce(access::location::cpu, access::mode::write);
    auto e = pd->getEnergy(access::location::cpu, access::mode::write);
    for (int i = 0; i < N; i++) { p[i] = pos[i]; c[i] = q[i]; f[i] = real4(); e[i] = 0; }
  }
  typename P::Parameters par;
  fill<P>(par, L, tol, gw, split, eps);
  auto poisson = std::make_shared<P>(pd, par);
  poisson->sum({.force = true, .energy = true, .virial = false}, 0);
  CudaSafeCall(cudaDeviceSynchronize());
  Result r;
  {
    auto f = pd->getForce(access::location::cpu, access::mode::read);
    auto e = pd->getEnergy(access::location::cpu, access::mode::read);
    r.force.assign(f.begin(), f.end());
    r.energy.assign(e.begin(), e.end());
  }
  thrust::device_vector<real4> fp = poisson->computeFieldPotentialAtParticles();
  r.fp.resize(N);
  thrust::copy(fp.begin(), fp.end(), r.fp.begin());
  return r;
}

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 2000;
  auto sys = std::make_shared<System>();
  // ---- 1. the reference's own known answer
  double katRef = 0, katOurs = 0, katFieldOurs = 0;
  {
    const real L = 100, r = 2, tol = 1e-7, gw = 0.001, split = 0.2;
    std::vector<real4> pos = {make_real4(-0.5 * r + 3.1, -7.7, 12.3, 0), make_real4(0.5 * r + 3.1, -7.7, 12.3, 0),
                              make_real4(0.5 * r + 3.1, -7.7, 12.3, 0)};
    std::vector<real> q = {1.0, -0.5, -0.5};
    const real want = theoreticalField(r, gw);
    Result a = run<Poisson>(sys, pos, q, L, tol, gw, split, 1.0);
    Result b = run<b200::Poisson>(sys, pos, q, L, tol, gw, split, 1.0);
    katRef = std::abs(1.0 - std::abs(a.force[0].x / want));
    katOurs = std::abs(1.0 - std::abs(b.force[0].x / want));
    katFieldOurs = std::abs(1.0 - std::abs(b.fp[0].x / want));
  }
  // ---- 2. a neutral cloud, both implementations
  double dForce = 0, dEnergy = 0, dField = 0, dPhi = 0, sF = 0, sE = 0, sFld = 0, sPhi = 0;
  {
    const real L = 40, tol = 1e-6, gw = 0.5, split = 0.6, eps = 1.3;
    std::mt19937_64 gen(7);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    std::vector<real4> pos(N);
    std::vector<real> q(N);
    double total = 0;
    for (int i = 0; i < N; i++) {
      pos[i] = make_real4(U(gen) * L, U(gen) * L, U(gen) * L, 0);
      q[i] = (i % 2) ? 1.0 : -1.0;
      total += q[i];
    }
    q[N - 1] -= total;
    Result a = run<Poisson>(sys, pos, q, L, tol, gw, split, eps);
    Result b = run<b200::Poisson>(sys, pos, q, L, tol, gw, split, eps);
    for (int i = 0; i < N; i++) {
      dForce = std::max({dForce, (double)std::abs(a.force[i].x - b.force[i].x), (double)std::abs(a.force[i].y - b.force[i].y),
                         (double)std::abs(a.force[i].z - b.force[i].z)});
      sF = std::max({sF, (double)std::abs(a.force[i].x), (double)std::abs(a.force[i].y), (double)std::abs(a.force[i].z)});
      dEnergy = std::max(dEnergy, (double)std::abs(a.energy[i] - b.energy[i]));
      sE = std::max(sE, (double)std::abs(a.energy[i]));
      dField = std::max({dField, (double)std::abs(a.fp[i].x - b.fp[i].x), (double)std::abs(a.fp[i].y - b.fp[i].y),
                         (double)std::abs(a.fp[i].z - b.fp[i].z)});
      sFld = std::max({sFld, (double)std::abs(a.fp[i].x), (double)std::abs(a.fp[i].y), (double)std::abs(a.fp[i].z)});
      dPhi = std::max(dPhi, (double)std::abs(a.fp[i].w - b.fp[i].w));
      sPhi = std::max(sPhi, (double)std::abs(a.fp[i].w));
    }
  }
  // ---- 3. timing of sum(force + energy) on a larger neutral cloud (argv[2] charges at number density 0.1, 0 skips it)
  const int Nt = argc > 2 ? atoi(argv[2]) : 0;
  double msRef = 0, msOurs = 0;
  if (Nt > 0) {
    const real L = std::cbrt(Nt / 0.1), tol = 1e-4, gw = 0.25, split = 1.0, eps = 1.0;
    std::mt19937_64 gen(11);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    auto pd = std::make_shared<ParticleData>(Nt, sys);
    {
      auto p = pd->getPos(access::location::cpu, access::mode::write);
      auto c = pd->getCharge(access::location::cpu, access::mode::write);
      for (int i = 0; i < Nt; i++) { p[i] = make_real4(U(gen) * L, U(gen) * L, U(gen) * L, 0); c[i] = (i % 2) ? 1.0 : -1.0; }
    }
    auto timeIt = [&](auto poisson) {
      for (int w = 0; w < 2; w++) poisson->sum({.force = true, .energy = true, .virial = false}, 0);
      CudaSafeCall(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, 0);
      for (int w = 0; w < 10; w++) poisson->sum({.force = true, .energy = true, .virial = false}, 0);
      cudaEventRecord(e1, 0);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      return (double)ms / 10;
    };
    Poisson::Parameters pa; fill<Poisson>(pa, L, tol, gw, split, eps);
    b200::Poisson::Parameters pb; fill<b200::Poisson>(pb, L, tol, gw, split, eps);
    msRef = timeIt(std::make_shared<Poisson>(pd, pa));
    msOurs = timeIt(std::make_shared<b200::Poisson>(pd, pb));
  }
  printf("{\"N\":%d,\"timing_N\":%d,\"sum_ms_reference\":%.4f,\"sum_ms_ours\":%.4f,\"kat_reference\":%.3e,\"kat_ours\":%.3e,\"kat_field_ours\":%.3e,\"force_vs_ref\":%.3e,\"energy_vs_ref\":%.3e,"
         "\"field_vs_ref\":%.3e,\"potential_vs_ref\":%.3e}\n",
         N, Nt, msRef, msOurs, katRef, katOurs, katFieldOurs, dForce / sF, dEnergy / sE, dField / sFld, dPhi / sPhi);
  sys->finish();
  return 0;
}
