/* Drop-in demonstration / integration test for path 1 (single precision build).
 *
 * A UAMMD program (reference headers, reference ParticleData / VerletNVE) where only the pair-force module is
 * swapped:  (A) stock   PairForces<Potential::LJ, CellList>                 (the reference)
 *           (B) generic PairForces<Potential::LJ, b200::CellList>          (our list + the reference's own
 *                                                                           traversal kernel and user functor)
 *           (C) fast    b200::PairForcesLJ                                 (our list + our LJ traversal)
 *           (D) column  PairForces<Potential::LJ, b200::ColumnList>        (the reference's own LJ functor through the B200
 *                                                                           generic-Transverser column traversal)
 * It prints the largest force deviation of (B) and (C) from (A) in units of the largest force, checks that the
 * CellListData of (A) and (B) are bit identical, and runs VerletNVE with (C) for a few steps next to (A).
 * Built by oracle/Makefile (needs the reference tree) into oracle/_ref/dropin_lj; run by tests/test_dropin_gpu.py.
 * usage: dropin_lj N L
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Integrator/VerletNVE.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

template <class T> std::vector<T> toHost(const T *d, size_t n) {
  std::vector<T> h(n);
  CudaSafeCall(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
  return h;
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

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 100000;
  const real L = argc > 2 ? atof(argv[2]) : 50.0;
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  Box box(make_real3(L));
  {
    auto pos = pd->getPos(access::cpu, access::write);
    std::mt19937_64 gen(2024);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    // jittered simple cubic lattice: liquid-like, no overlaps
    const int n = (int)std::ceil(std::cbrt((double)N));
    for (int i = 0; i < N; i++) {
      const int ix = i % n, iy = (i / n) % n, iz = i / (n * n);
      const double a = L / n;
      pos[i] = make_real4((ix + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), (iy + 0.5) * a - 0.5 * L + 0.2 * a * U(gen),
                          (iz + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), 0);
    }
  }
  Potential::LJ::InputPairParameters par;
  par.epsilon = 1.0; par.sigma = 1.0; par.cutOff = 2.5; par.shift = false;
  auto potA = std::make_shared<Potential::LJ>();
  potA->setPotParameters(0, 0, par);
  auto potC = std::make_shared<b200::LJ>();
  potC->setPotParameters(0, 0, par);

  using PFA = PairForces<Potential::LJ, CellList>;
  using PFB = PairForces<Potential::LJ, b200::CellList>;
  PFA::Parameters pa; pa.box = box; pa.nl = std::make_shared<CellList>(pd);
  PFB::Parameters pb; pb.box = box; pb.nl = std::make_shared<b200::CellList>(pd);
  b200::PairForcesLJ::Parameters pc; pc.box = box;
  auto A = std::make_shared<PFA>(pd, pa, potA);
  auto B = std::make_shared<PFB>(pd, pb, potA);
  auto C = std::make_shared<b200::PairForcesLJ>(pd, pc, potC);
  using PFD = PairForces<Potential::LJ, b200::ColumnList>;
  PFD::Parameters pdd; pdd.box = box;
  auto D = std::make_shared<PFD>(pd, pdd, potA);

  auto fA = forcesOf(pd, A), fB = forcesOf(pd, B), fC = forcesOf(pd, C), fD = forcesOf(pd, D);
  double fmax = 0, dB = 0, dC = 0, dD = 0;
  for (int i = 0; i < N; i++) {
    fmax = std::max({fmax, (double)std::abs(fA[i].x), (double)std::abs(fA[i].y), (double)std::abs(fA[i].z)});
    dB = std::max({dB, (double)std::abs(fA[i].x - fB[i].x), (double)std::abs(fA[i].y - fB[i].y), (double)std::abs(fA[i].z - fB[i].z)});
    dC = std::max({dC, (double)std::abs(fA[i].x - fC[i].x), (double)std::abs(fA[i].y - fC[i].y), (double)std::abs(fA[i].z - fC[i].z)});
    dD = std::max({dD, (double)std::abs(fA[i].x - fD[i].x), (double)std::abs(fA[i].y - fD[i].y), (double)std::abs(fA[i].z - fD[i].z)});
  }
  // energy + virial through the reference's own functor (a five-word quantity) over the column traversal
  double dE = 0, dV = 0, emax = 0, vmax = 0;
  {
    auto ev = [&](std::shared_ptr<Interactor> it, std::vector<real> &e, std::vector<real> &v) {
      {
        auto en = pd->getEnergy(access::gpu, access::write); thrust::fill(thrust::cuda::par, en.begin(), en.end(), real());
        auto vi = pd->getVirial(access::gpu, access::write); thrust::fill(thrust::cuda::par, vi.begin(), vi.end(), real());
        auto f = pd->getForce(access::gpu, access::write); thrust::fill(thrust::cuda::par, f.begin(), f.end(), real4());
      }
      Interactor::Computables comp; comp.force = true; comp.energy = true; comp.virial = true;
      it->sum(comp, 0);
      CudaSafeCall(cudaDeviceSynchronize());
      auto en = pd->getEnergy(access::cpu, access::read); e.assign(en.begin(), en.end());
      auto vi = pd->getVirial(access::cpu, access::read); v.assign(vi.begin(), vi.end());
    };
    std::vector<real> eA, vA, eD, vD;
    ev(A, eA, vA); ev(D, eD, vD);
    for (int i = 0; i < N; i++) {
      emax = std::max(emax, (double)std::abs(eA[i])); vmax = std::max(vmax, (double)std::abs(vA[i]));
      dE = std::max(dE, (double)std::abs(eA[i] - eD[i])); dV = std::max(dV, (double)std::abs(vA[i] - vD[i]));
    }
  }
  // CellListData bit parity between the reference list and ours
  auto clA = pa.nl->getCellList();
  auto clB = pb.nl->getCellList();
  const int ncells = clA.grid.getNumberCells();
  auto giA = toHost(clA.groupIndex, N), giB = toHost(clB.groupIndex, N);
  auto spA = toHost(clA.sortPos, N), spB = toHost(clB.sortPos, N);
  auto csA = toHost(clA.cellStart, ncells), csB = toHost(clB.cellStart, ncells);
  auto ceA = toHost(clA.cellEnd, ncells), ceB = toHost(clB.cellEnd, ncells);
  long mismatches = 0;
  for (int i = 0; i < N; i++) mismatches += (giA[i] != giB[i]) + (memcmp(&spA[i], &spB[i], sizeof(real4)) != 0);
  for (int c = 0; c < ncells; c++) {
    const bool eA = csA[c] < clA.VALID_CELL, eB = csB[c] < clB.VALID_CELL;
    if (eA != eB) mismatches++;
    else if (!eA) mismatches += (csA[c] - clA.VALID_CELL != csB[c] - clB.VALID_CELL) + (ceA[c] != ceB[c]);
  }
  // a few NVE steps: reference integrator driving (A) on one ParticleData, (C) on a copy
  auto pd2 = std::make_shared<ParticleData>(N, sys);
  {
    auto p1 = pd->getPos(access::cpu, access::read);
    auto p2 = pd2->getPos(access::cpu, access::write);
    std::copy(p1.begin(), p1.end(), p2.begin());
    auto v1 = pd->getVel(access::cpu, access::write);
    auto v2 = pd2->getVel(access::cpu, access::write);
    std::mt19937 gen(7);
    std::normal_distribution<float> G(0, 1);
    for (int i = 0; i < N; i++) { v1[i] = make_real3(G(gen), G(gen), G(gen)); v2[i] = v1[i]; }
  }
  VerletNVE::Parameters vp; vp.dt = 0.002; vp.initVelocities = false;
  auto nveA = std::make_shared<VerletNVE>(pd, vp);
  PFA::Parameters pa2; pa2.box = box;
  nveA->addInteractor(std::make_shared<PFA>(pd, pa2, potA));
  auto nveC = std::make_shared<VerletNVE>(pd2, vp);
  nveC->addInteractor(std::make_shared<b200::PairForcesLJ>(pd2, pc, potC));
  for (int s = 0; s < 20; s++) { nveA->forwardTime(); nveC->forwardTime(); }
  CudaSafeCall(cudaDeviceSynchronize());
  double dpos = 0;
  {
    auto p1 = pd->getPos(access::cpu, access::read);
    auto p2 = pd2->getPos(access::cpu, access::read);
    for (int i = 0; i < N; i++)
      dpos = std::max({dpos, (double)std::abs(p1[i].x - p2[i].x), (double)std::abs(p1[i].y - p2[i].y), (double)std::abs(p1[i].z - p2[i].z)});
  }
  printf("{\"N\":%d,\"fmax\":%.6g,\"generic_vs_ref\":%.6g,\"fast_vs_ref\":%.6g,\"column_generic_vs_ref\":%.6g,\"column_energy_vs_ref\":%.6g,"
         "\"column_virial_vs_ref\":%.6g,\"celllist_mismatches\":%ld,\"nve20_max_dpos\":%.6g}\n",
         N, fmax, dB / fmax, dC / fmax, dD / fmax, dE / emax, dV / vmax, mismatches, dpos);
  sys->finish();
  return 0;
}
