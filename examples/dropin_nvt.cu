/* Drop-in check of the configuration examples/misc/benchmark.cu runs (single precision build):
 * VerletNVT::GronbechJensen + PairForces<Potential::LJ, VerletList>.
 *   (A) stock   VerletNVT::GronbechJensen + PairForces<Potential::LJ, VerletList>            (the reference)
 *   (B) ours    b200::VerletNVTGronbechJensen, no interactor vs the reference integrator alone (bit identical expected)
 *   (B') ours   b200::VerletNVTBasic vs the reference's VerletNVT::Basic alone                (bit identical expected)
 *   (C) ours    b200::VerletNVTGronbechJensen + b200::PairForcesLJ over b200::VerletList      next to (A)
 * Both systems start from the same System seed, so the integrators draw the same Saru seeds and initial velocities.
 * Built by oracle/Makefile into oracle/_ref/dropin_nvt; run by tests/test_dropin_gpu.py.  usage: dropin_nvt N L steps
 */
#include "uammd.cuh"
#include "Integrator/VerletNVT.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "Interactor/NeighbourList/VerletList.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;

static void lattice(std::shared_ptr<ParticleData> pd, int N, real L) {
  auto pos = pd->getPos(access::cpu, access::write);
  std::mt19937_64 gen(2024);
  std::uniform_real_distribution<double> U(-0.5, 0.5);
  const int n = (int)std::ceil(std::cbrt((double)N));
  const double a = L / n;
  for (int i = 0; i < N; i++) {
    const int ix = i % n, iy = (i / n) % n, iz = i / (n * n);
    pos[i] = make_real4((ix + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), (iy + 0.5) * a - 0.5 * L + 0.2 * a * U(gen),
                        (iz + 0.5) * a - 0.5 * L + 0.2 * a * U(gen), 0);
  }
}

struct Diff {
  double dpos = 0, dvel = 0;
  long words = 0;
};
static Diff compare(std::shared_ptr<ParticleData> a, std::shared_ptr<ParticleData> b, int N) {
  Diff d;
  auto p1 = a->getPos(access::cpu, access::read);
  auto p2 = b->getPos(access::cpu, access::read);
  auto v1 = a->getVel(access::cpu, access::read);
  auto v2 = b->getVel(access::cpu, access::read);
  for (int i = 0; i < N; i++) {
    d.dpos = std::max({d.dpos, (double)std::abs(p1[i].x - p2[i].x), (double)std::abs(p1[i].y - p2[i].y), (double)std::abs(p1[i].z - p2[i].z)});
    d.dvel = std::max({d.dvel, (double)std::abs(v1[i].x - v2[i].x), (double)std::abs(v1[i].y - v2[i].y), (double)std::abs(v1[i].z - v2[i].z)});
    d.words += (memcmp(&p1[i], &p2[i], sizeof(real4)) != 0) + (memcmp(&v1[i], &v2[i], sizeof(real3)) != 0);
  }
  return d;
}

/* VerletNVT::Basic's public constructor is declared (VerletNVT.cuh:92) but never defined in the reference; a derived class
   reaches the protected one. What runs is the unmodified Basic. */
struct BasicExposed : VerletNVT::Basic {
  BasicExposed(std::shared_ptr<ParticleData> pd, VerletNVT::Basic::Parameters par)
      : VerletNVT::Basic(std::make_shared<ParticleGroup>(pd, "All"), par, "VerletNVT::Basic") {}
};

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 32768;
  const real L = argc > 2 ? atof(argv[2]) : 38.0;
  const int steps = argc > 3 ? atoi(argv[3]) : 20;
  Box box(make_real3(L));
  Potential::LJ::InputPairParameters par;
  par.epsilon = 1.0; par.sigma = 1.0; par.cutOff = 2.5; par.shift = false;
  VerletNVT::GronbechJensen::Parameters np;
  np.temperature = 1.0; np.dt = 0.002; np.friction = 1.0; np.initVelocities = true;
  b200::VerletNVTGronbechJensen::Parameters bp;
  bp.temperature = np.temperature; bp.dt = np.dt; bp.friction = np.friction; bp.initVelocities = true;

  // ---- integrator alone: bit identity -------------------------------------------------------------------------
  Diff ideal;
  {
    auto sysA = std::make_shared<System>(); sysA->rng().setSeed(4242);
    auto sysB = std::make_shared<System>(); sysB->rng().setSeed(4242);
    auto pdA = std::make_shared<ParticleData>(N, sysA), pdB = std::make_shared<ParticleData>(N, sysB);
    lattice(pdA, N, L); lattice(pdB, N, L);
    auto A = std::make_shared<VerletNVT::GronbechJensen>(pdA, np);
    auto B = std::make_shared<b200::VerletNVTGronbechJensen>(pdB, bp);
    for (int s = 0; s < steps; s++) { A->forwardTime(); B->forwardTime(); }
    CudaSafeCall(cudaDeviceSynchronize());
    ideal = compare(pdA, pdB, N);
  }
  // ---- VerletNVT::Basic alone: bit identity -------------------------------------------------------------------
  Diff basic;
  {
    auto sysA = std::make_shared<System>(); sysA->rng().setSeed(99);
    auto sysB = std::make_shared<System>(); sysB->rng().setSeed(99);
    auto pdA = std::make_shared<ParticleData>(N, sysA), pdB = std::make_shared<ParticleData>(N, sysB);
    lattice(pdA, N, L); lattice(pdB, N, L);
    auto A = std::make_shared<BasicExposed>(pdA, np);
    auto B = std::make_shared<b200::VerletNVTBasic>(pdB, bp);
    for (int s = 0; s < steps; s++) { A->forwardTime(); B->forwardTime(); }
    CudaSafeCall(cudaDeviceSynchronize());
    basic = compare(pdA, pdB, N);
  }
  // ---- with the LJ interactor over the Verlet list (benchmark.cu) ---------------------------------------------
  Diff lj;
  {
    auto sysA = std::make_shared<System>(); sysA->rng().setSeed(777);
    auto sysC = std::make_shared<System>(); sysC->rng().setSeed(777);
    auto pdA = std::make_shared<ParticleData>(N, sysA), pdC = std::make_shared<ParticleData>(N, sysC);
    lattice(pdA, N, L); lattice(pdC, N, L);
    auto potA = std::make_shared<Potential::LJ>(); potA->setPotParameters(0, 0, par);
    auto potC = std::make_shared<b200::LJ>(); potC->setPotParameters(0, 0, par);
    using PFA = PairForces<Potential::LJ, VerletList>;
    PFA::Parameters pa; pa.box = box;
    b200::PairForcesLJ::Parameters pc; pc.box = box; pc.verletList = std::make_shared<b200::VerletList>(pdC);
    auto A = std::make_shared<VerletNVT::GronbechJensen>(pdA, np);
    A->addInteractor(std::make_shared<PFA>(pdA, pa, potA));
    auto C = std::make_shared<b200::VerletNVTGronbechJensen>(pdC, bp);
    C->addInteractor(std::make_shared<b200::PairForcesLJ>(pdC, pc, potC));
    for (int s = 0; s < steps; s++) { A->forwardTime(); C->forwardTime(); }
    CudaSafeCall(cudaDeviceSynchronize());
    lj = compare(pdA, pdC, N);
  }
  printf("{\"N\":%d,\"steps\":%d,\"ideal_mismatch_words\":%ld,\"ideal_max_dpos\":%.6g,\"ideal_max_dvel\":%.6g,"
         "\"basic_mismatch_words\":%ld,\"basic_max_dvel\":%.6g,\"lj_max_dpos\":%.6g,\"lj_max_dvel\":%.6g}\n",
         N, steps, ideal.words, ideal.dpos, ideal.dvel, basic.words, basic.dvel, lj.dpos, lj.dvel);
  return 0;
}
