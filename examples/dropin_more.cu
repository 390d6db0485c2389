/* Drop-in demonstration / integration test for the round-1 additions (single precision build):
 *   (1) Verlet list:  stock PairForces<Potential::LJ, VerletList>  vs  PairForces<Potential::LJ, b200::VerletList> (the
 *       reference's traversal kernel + functor over OUR list: identical bits expected) vs b200::PairForcesLJ over the
 *       b200::VerletList (fast path), plus a few VerletNVE steps;
 *   (2) BD::EulerMaruyama vs b200::BDEulerMaruyama: identical positions after 50 steps (same seeds, same Saru stream);
 *   (3) BDHI::EulerMaruyama<BDHI::PSE> vs BDHI::EulerMaruyama<b200::PSE> at T = 0: positions after 3 steps.
 * Built by oracle/Makefile into oracle/_ref/dropin_more (needs the reference tree + the LAPACK shim); run by
 * tests/test_dropin_gpu.py.   usage: dropin_more N L
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/ExternalForces.cuh"
#include "Integrator/VerletNVE.cuh"
#include "Integrator/BrownianDynamics.cuh"
#include "Integrator/BDHI/BDHI_EulerMaruyama.cuh"
#include "Integrator/BDHI/BDHI_PSE.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

static void fillLattice(std::shared_ptr<ParticleData> pd, int N, real L) {
  auto pos = pd->getPos(access::cpu, access::write);
  std::mt19937_64 gen(2024);
  std::uniform_real_distribution<double> U(-0.5, 0.5);
  const int n = (int)std::ceil(std::cbrt((double)N));
  for (int i = 0; i < N; i++) {
    const int ix = i % n, iy = (i / n) % n, iz = i / (n * n);
    const double a = L / n;
    pos[i] = make_real4((ix + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), (iy + 0.5) * a - 0.5 * L + 0.2 * a * U(gen),
                        (iz + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), 0);
  }
}
static std::vector<real4> forcesOf(std::shared_ptr<ParticleData> pd, std::shared_ptr<Interactor> it) {
  {
    auto f = pd->getForce(access::gpu, access::write);
    thrust::fill(thrust::cuda::par, f.begin(), f.end(), real4());
  }
  Interactor::Computables comp;
  comp.force = true;
  it->sum(comp, 0);
  CudaSafeCall(cudaDeviceSynchronize());
  auto f = pd->getForce(access::cpu, access::read);
  return std::vector<real4>(f.begin(), f.end());
}
static double maxDiff(std::shared_ptr<ParticleData> a, std::shared_ptr<ParticleData> b, int N) {
  auto p1 = a->getPos(access::cpu, access::read);
  auto p2 = b->getPos(access::cpu, access::read);
  double d = 0;
  for (int i = 0; i < N; i++)
    d = std::max({d, (double)std::abs(p1[i].x - p2[i].x), (double)std::abs(p1[i].y - p2[i].y), (double)std::abs(p1[i].z - p2[i].z)});
  return d;
}
struct Pull { // constant external force along x on every particle
  __device__ ForceEnergyVirial sum(Interactor::Computables comp, real4 pos) {
    ForceEnergyVirial r;
    r.force = make_real3(1, 0.5, -0.25); r.energy = 0; r.virial = 0;
    return r;
  }
  auto getArrays(ParticleData *pd) { return std::make_tuple(pd->getPos(access::gpu, access::read).raw()); }
};

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 50000;
  const real L = argc > 2 ? atof(argv[2]) : 40.0;
  auto sys = std::make_shared<System>();
  Box box(make_real3(L));
  // ---------------- (1) Verlet list ----------------
  auto pd = std::make_shared<ParticleData>(N, sys);
  fillLattice(pd, N, L);
  Potential::LJ::InputPairParameters par;
  par.epsilon = 1.0; par.sigma = 1.0; par.cutOff = 2.5; par.shift = false;
  auto potA = std::make_shared<Potential::LJ>(); potA->setPotParameters(0, 0, par);
  auto potC = std::make_shared<b200::LJ>(); potC->setPotParameters(0, 0, par);
  using PFA = PairForces<Potential::LJ, VerletList>;
  using PFB = PairForces<Potential::LJ, b200::VerletList>;
  PFA::Parameters pa; pa.box = box; pa.nl = std::make_shared<VerletList>(pd);
  PFB::Parameters pb; pb.box = box; pb.nl = std::make_shared<b200::VerletList>(pd);
  b200::PairForcesLJ::Parameters pc; pc.box = box; pc.verletList = std::make_shared<b200::VerletList>(pd);
  auto A = std::make_shared<PFA>(pd, pa, potA);
  auto B = std::make_shared<PFB>(pd, pb, potA);
  auto C = std::make_shared<b200::PairForcesLJ>(pd, pc, potC);
  auto fA = forcesOf(pd, A), fB = forcesOf(pd, B), fC = forcesOf(pd, C);
  double fmax = 0, dB = 0, dC = 0;
  for (int i = 0; i < N; i++) {
    fmax = std::max({fmax, (double)std::abs(fA[i].x), (double)std::abs(fA[i].y), (double)std::abs(fA[i].z)});
    dB = std::max({dB, (double)std::abs(fA[i].x - fB[i].x), (double)std::abs(fA[i].y - fB[i].y), (double)std::abs(fA[i].z - fB[i].z)});
    dC = std::max({dC, (double)std::abs(fA[i].x - fC[i].x), (double)std::abs(fA[i].y - fC[i].y), (double)std::abs(fA[i].z - fC[i].z)});
  }
  // ---------------- (2) BD::EulerMaruyama ----------------
  double bdDiff = 0, bdForceDiff = 0;
  for (int withForce = 0; withForce < 2; withForce++) {
    auto sysA = std::make_shared<System>(); sysA->rng().setSeed(1234);
    auto sysB = std::make_shared<System>(); sysB->rng().setSeed(1234);
    auto pA = std::make_shared<ParticleData>(N, sysA), pB = std::make_shared<ParticleData>(N, sysB);
    fillLattice(pA, N, L); fillLattice(pB, N, L);
    BD::EulerMaruyama::Parameters bp; bp.temperature = 1.0; bp.viscosity = 1.0; bp.hydrodynamicRadius = 1.0; bp.dt = 0.01;
    b200::BDEulerMaruyama::Parameters bq; bq.temperature = 1.0; bq.viscosity = 1.0; bq.hydrodynamicRadius = 1.0; bq.dt = 0.01;
    auto bdA = std::make_shared<BD::EulerMaruyama>(pA, bp);
    auto bdB = std::make_shared<b200::BDEulerMaruyama>(pB, bq);
    if (withForce) {
      bdA->addInteractor(std::make_shared<ExternalForces<Pull>>(pA, std::make_shared<Pull>()));
      bdB->addInteractor(std::make_shared<ExternalForces<Pull>>(pB, std::make_shared<Pull>()));
    }
    for (int s = 0; s < 50; s++) { bdA->forwardTime(); bdB->forwardTime(); }
    CudaSafeCall(cudaDeviceSynchronize());
    (withForce ? bdForceDiff : bdDiff) = maxDiff(pA, pB, N);
  }
  // ---------------- (3) BDHI::EulerMaruyama<PSE> ----------------
  double pseDiff = 0, pseMove = 0;
  {
    auto sysA = std::make_shared<System>(); sysA->rng().setSeed(99);
    auto sysB = std::make_shared<System>(); sysB->rng().setSeed(99);
    auto pA = std::make_shared<ParticleData>(N, sysA), pB = std::make_shared<ParticleData>(N, sysB), p0 = std::make_shared<ParticleData>(N, sysA);
    fillLattice(pA, N, L); fillLattice(pB, N, L); fillLattice(p0, N, L);
    using SA = BDHI::EulerMaruyama<BDHI::PSE>;
    using SB = BDHI::EulerMaruyama<b200::PSE>;
    SA::Parameters qa; qa.temperature = 0; qa.viscosity = 1.0; qa.hydrodynamicRadius = 0.5; qa.dt = 0.01; qa.box = box; qa.tolerance = 1e-4; qa.psi = 0.8;
    SB::Parameters qb; qb.temperature = 0; qb.viscosity = 1.0; qb.hydrodynamicRadius = 0.5; qb.dt = 0.01; qb.box = box; qb.tolerance = 1e-4; qb.psi = 0.8;
    auto ia = std::make_shared<SA>(pA, qa);
    auto ib = std::make_shared<SB>(pB, qb);
    ia->addInteractor(std::make_shared<ExternalForces<Pull>>(pA, std::make_shared<Pull>()));
    ib->addInteractor(std::make_shared<ExternalForces<Pull>>(pB, std::make_shared<Pull>()));
    for (int s = 0; s < 3; s++) { ia->forwardTime(); ib->forwardTime(); }
    CudaSafeCall(cudaDeviceSynchronize());
    pseDiff = maxDiff(pA, pB, N);
    pseMove = maxDiff(pA, p0, N);
  }
  printf("{\"N\":%d,\"fmax\":%.6g,\"verlet_generic_vs_ref\":%.6g,\"verlet_fast_vs_ref\":%.6g,\"bd_max_dpos\":%.6g,\"bd_force_max_dpos\":%.6g,"
         "\"pse_max_dpos\":%.6g,\"pse_displacement\":%.6g}\n",
         N, fmax, dB / fmax, dC / fmax, bdDiff, bdForceDiff, pseDiff, pseMove);
  sys->finish();
  return 0;
}
