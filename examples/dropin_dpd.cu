/* Drop-in demonstration / integration test for the DPD fluid (BASELINE config 4 arithmetic), single precision build.
 *
 * A UAMMD program (reference headers, ParticleData, VerletNVE, PairForces, CellList) with the DPD module swapped:
 *   (A) reference  PairForces<b200::DPDPotential, CellList>     the reference's own PairForces + CellList + traversal kernel
 *                                                                driving the reference's own DPD_impl::ForceTransverser
 *                                                                through the getTransverser the glue adds (the stock
 *                                                                Potential::DPD is silently skipped by PairForces)
 *   (B) fast       b200::PairForcesDPD                           our cell list + specialised DPD traversal
 * Both are given the same Saru seed and step, so the pairwise noise is the same stream; the forces may differ by the
 * summation order only. Then VerletNVE + (B) runs a DPD fluid at rho = 3 and the kinetic temperature is measured: the
 * thermostat (dissipative + random force, sigma^2 = 2 gamma kT / dt) must hold kT.
 * Built by oracle/Makefile into oracle/_ref/dropin_dpd; run by tests/test_dropin_gpu.py.
 * usage: dropin_dpd N steps
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Integrator/VerletNVE.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

/* (A): same potential class with a seed we can read: getForceTransverser keeps its seed in a function-local static
   (DPD.cuh:165), so the comparison builds the reference's ForceTransverser with OUR seed through its public constructor */
struct SeededDPD : public b200::DPDPotential {
  using b200::DPDPotential::DPDPotential;
  uint seed = 0;
  auto getTransverser(Interactor::Computables comp, Box box, std::shared_ptr<ParticleData> pd) {
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    auto vel = pd->getVel(access::location::gpu, access::mode::read);
    auto force = pd->getForce(access::location::gpu, access::mode::readwrite);
    step++;
    return ForceTransverser(pos.raw(), vel.raw(), force.raw(), seed, step, box, pd->getNumParticles(), rcut, gamma, sigma, A);
  }
};

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

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 81000;
  const int steps = argc > 2 ? atoi(argv[2]) : 2000;
  const real L = std::cbrt(N / 3.0), dt = 0.01, kT = 1.0;
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  Box box(make_real3(L));
  {
    auto pos = pd->getPos(access::cpu, access::write);
    auto vel = pd->getVel(access::cpu, access::write);
    std::mt19937_64 gen(21);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    std::normal_distribution<double> G(0, 1);
    for (int i = 0; i < N; i++) {
      pos[i] = make_real4(L * U(gen), L * U(gen), L * U(gen), 0);
      vel[i] = make_real3(G(gen), G(gen), G(gen));
    }
  }
  SeededDPD::Parameters par;
  par.cutOff = 1.0; par.A = 25.0; par.gamma.gamma = 4.5; par.temperature = kT; par.dt = dt;
  auto potA = std::make_shared<SeededDPD>(par);
  auto potB = std::make_shared<b200::DPDPotential>(par);
  using PFA = PairForces<SeededDPD, CellList>;
  PFA::Parameters pa; pa.box = box;
  b200::PairForcesDPD::Parameters pb; pb.box = box;
  auto A = std::make_shared<PFA>(pd, pa, potA);
  auto B = std::make_shared<b200::PairForcesDPD>(pd, pb, potB);
  potA->seed = B->getSeed();
  auto fA = forcesOf(pd, A), fB = forcesOf(pd, B); // both at step 1
  // (C) the reference's DPD transverser (a GENERAL transverser: getInfo brings velocities and ids) through the B200
  // generic-Transverser column traversal
  auto potC = std::make_shared<SeededDPD>(par);
  potC->seed = B->getSeed();
  using PFC = PairForces<SeededDPD, b200::ColumnList>;
  PFC::Parameters pcc; pcc.box = box;
  auto Cc = std::make_shared<PFC>(pd, pcc, potC);
  auto fC = forcesOf(pd, Cc);
  double dAC = 0;
  for (int i = 0; i < N; i++)
    dAC = std::max({dAC, (double)std::abs(fA[i].x - fC[i].x), (double)std::abs(fA[i].y - fC[i].y), (double)std::abs(fA[i].z - fC[i].z)});
  double fmax = 0, dAB = 0, fsum = 0;
  for (int i = 0; i < N; i++) {
    fmax = std::max({fmax, (double)std::abs(fA[i].x), (double)std::abs(fA[i].y), (double)std::abs(fA[i].z)});
    dAB = std::max({dAB, (double)std::abs(fA[i].x - fB[i].x), (double)std::abs(fA[i].y - fB[i].y), (double)std::abs(fA[i].z - fB[i].z)});
    fsum += std::abs(fA[i].x);
  }
  // thermostat: VerletNVE + b200::PairForcesDPD from the random cloud; kT over the second half of the run
  VerletNVE::Parameters vp; vp.dt = dt; vp.initVelocities = false;
  auto nve = std::make_shared<VerletNVE>(pd, vp);
  nve->addInteractor(B);
  double ktSum = 0; int ktCount = 0;
  for (int s = 0; s < steps; s++) {
    nve->forwardTime();
    if (s >= steps / 2 && s % 20 == 0) {
      auto vel = pd->getVel(access::cpu, access::read);
      double k = 0;
      for (int i = 0; i < N; i++) k += (double)vel[i].x * vel[i].x + (double)vel[i].y * vel[i].y + (double)vel[i].z * vel[i].z;
      ktSum += k / (3.0 * N); ktCount++;
    }
  }
  CudaSafeCall(cudaDeviceSynchronize());
  printf("{\"N\":%d,\"fmax\":%.6g,\"fast_vs_ref\":%.6g,\"column_generic_vs_ref\":%.6g,\"mean_abs_fx\":%.6g,\"steps\":%d,\"kT_measured\":%.6g,"
         "\"kT_target\":%.6g}\n", N, fmax, dAB / fmax, dAC / fmax, fsum / N, steps, ktCount ? ktSum / ktCount : 0.0, (double)kT);
  sys->finish();
  return 0;
}
