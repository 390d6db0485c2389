/* Drop-in check of the all-pairs fallback (single precision build).
 *
 * PairForces switches from the neighbour list to NBody::transverse when the box is no larger than three cut-offs in
 * every dimension (Interactor/PairForces.cu:49-53,61-66). The same UAMMD program computes force, energy and virial of a
 * small LJ system with (A) the stock PairForces<Potential::LJ> and (C) b200::PairForcesLJ (ub200_lj_nbody_f32) and prints
 * the largest deviations in units of the largest reference value.
 * Built by oracle/Makefile into oracle/_ref/dropin_nbody; run by tests/test_dropin_gpu.py.   usage: dropin_nbody N L
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

struct Result {
  std::vector<real4> force;
  std::vector<real> energy, virial;
};

static Result evaluate(std::shared_ptr<ParticleData> pd, std::shared_ptr<Interactor> it) {
  {
    auto f = pd->getForce(access::gpu, access::write);
    thrust::fill(thrust::cuda::par, f.begin(), f.end(), real4());
    auto e = pd->getEnergy(access::gpu, access::write);
    thrust::fill(thrust::cuda::par, e.begin(), e.end(), real());
    auto v = pd->getVirial(access::gpu, access::write);
    thrust::fill(thrust::cuda::par, v.begin(), v.end(), real());
  }
  Interactor::Computables comp;
  comp.force = comp.energy = comp.virial = true;
  it->sum(comp, 0);
  CudaSafeCall(cudaDeviceSynchronize());
  Result r;
  auto f = pd->getForce(access::cpu, access::read);
  auto e = pd->getEnergy(access::cpu, access::read);
  auto v = pd->getVirial(access::cpu, access::read);
  r.force.assign(f.begin(), f.end());
  r.energy.assign(e.begin(), e.end());
  r.virial.assign(v.begin(), v.end());
  return r;
}

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 300;
  const real L = argc > 2 ? atof(argv[2]) : 7.0;
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  Box box(make_real3(L));
  {
    auto pos = pd->getPos(access::cpu, access::write);
    std::mt19937_64 gen(77);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    const int n = (int)std::ceil(std::cbrt((double)N));
    for (int i = 0; i < N; i++) {
      const int ix = i % n, iy = (i / n) % n, iz = i / (n * n);
      const double a = L / n;
      pos[i] = make_real4((ix + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), (iy + 0.5) * a - 0.5 * L + 0.2 * a * U(gen),
                          (iz + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), 0);
    }
  }
  Potential::LJ::InputPairParameters par;
  par.epsilon = 1.0; par.sigma = 1.0; par.cutOff = 2.5; par.shift = true;
  auto potA = std::make_shared<Potential::LJ>();
  potA->setPotParameters(0, 0, par);
  auto potC = std::make_shared<b200::LJ>();
  potC->setPotParameters(0, 0, par);
  using PFA = PairForces<Potential::LJ>;
  PFA::Parameters pa; pa.box = box;
  b200::PairForcesLJ::Parameters pc; pc.box = box;
  auto A = std::make_shared<PFA>(pd, pa, potA);
  auto C = std::make_shared<b200::PairForcesLJ>(pd, pc, potC);
  const Result a = evaluate(pd, A), c = evaluate(pd, C);
  double fmax = 0, emax = 0, vmax = 0, df = 0, de = 0, dv = 0;
  for (int i = 0; i < N; i++) {
    fmax = std::max({fmax, (double)std::abs(a.force[i].x), (double)std::abs(a.force[i].y), (double)std::abs(a.force[i].z)});
    emax = std::max(emax, (double)std::abs(a.energy[i]));
    vmax = std::max(vmax, (double)std::abs(a.virial[i]));
    df = std::max({df, (double)std::abs(a.force[i].x - c.force[i].x), (double)std::abs(a.force[i].y - c.force[i].y),
                   (double)std::abs(a.force[i].z - c.force[i].z)});
    de = std::max(de, (double)std::abs(a.energy[i] - c.energy[i]));
    dv = std::max(dv, (double)std::abs(a.virial[i] - c.virial[i]));
  }
  printf("{\"N\":%d,\"nbody\":%d,\"fmax\":%.6g,\"force_vs_ref\":%.6g,\"energy_vs_ref\":%.6g,\"virial_vs_ref\":%.6g}\n", N,
         (int)(L <= 3 * par.cutOff), fmax, df / fmax, de / emax, dv / vmax);
  sys->finish();
  return 0;
}
