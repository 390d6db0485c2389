/* Drop-in check of the double precision pair path (a -DDOUBLE_PRECISION build of UAMMD: real = double).
 *
 * The same UAMMD program computes force, energy and virial of an LJ liquid with (A) the stock PairForces<Potential::LJ>
 * (CellList in double, Radial<LJFunctor>::Transverser) and (C) an Interactor that forwards Interactor::sum to
 * ub200_lj_sum_f64 (uammd_b200/csrc/pair_lj_f64.cu), and prints the largest deviations in units of the largest reference
 * value. Both evaluate the same pair terms in double; only the order of the sums differs.
 * Built by oracle/Makefile into oracle/_ref/dropin_lj64; run by tests/test_dropin_gpu.py.   usage: dropin_lj64 N [shift]
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "uammd_b200.h"
#include <random>
#include <stdexcept>
#include <vector>
using namespace uammd;

static_assert(sizeof(real) == 8, "compile with -DDOUBLE_PRECISION");

/* Interactor concept (Interactor/Interactor.cuh:56-119) over the C ABI: what b200::PairForcesLJ is for single precision */
class PairForcesLJ64 : public Interactor {
  ub200_lj64 *handle = nullptr;
  Box box;
  std::vector<double> table; // LJFunctor::PairParameters rows {cutOff2, sigma2, epsilonDivSigma2, shift}
  int ntypes = 0;
  double cutOff = 0;

public:
  PairForcesLJ64(std::shared_ptr<ParticleData> pd, Box box) : Interactor(pd, "b200::PairForcesLJ64"), box(box) {
    if (ub200_lj64_create(&handle)) throw std::runtime_error("ub200_lj64_create");
  }
  ~PairForcesLJ64() { ub200_lj64_destroy(handle); }
  /* one particle type; more types fill an ntypes x ntypes table the same way */
  void setPotParameters(Potential::LJ::InputPairParameters in) {
    const auto p = Potential::LJFunctor::processPairParameters(in); // the reference's own host arithmetic
    table = {(double)p.cutOff2, (double)p.sigma2, (double)p.epsilonDivSigma2, (double)p.shift};
    ntypes = 1;
    cutOff = in.cutOff;
  }
  void updateBox(Box newBox) override { box = newBox; }
  void sum(Computables comp, cudaStream_t st = 0) override {
    auto force = comp.force ? pd->getForce(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    auto energy = comp.energy ? pd->getEnergy(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    auto virial = comp.virial ? pd->getVirial(access::location::gpu, access::mode::readwrite).raw() : nullptr;
    auto pos = pd->getPos(access::location::gpu, access::mode::read);
    const double L[3] = {box.boxSize.x, box.boxSize.y, box.boxSize.z};
    const int periodic[3] = {box.isPeriodicX(), box.isPeriodicY(), box.isPeriodicZ()};
    const int rc = ub200_lj_sum_f64(handle, pos.raw(), pd->getNumParticles(), L, periodic, cutOff, table.data(), ntypes, force, energy,
                                    virial, (void *)st);
    if (rc) throw std::runtime_error(std::string("ub200_lj_sum_f64: ") + ub200_error_string(rc));
  }
};

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
  const int N = argc > 1 ? atoi(argv[1]) : 4000;
  const bool shift = argc > 2 ? atoi(argv[2]) != 0 : true;
  const real L = std::cbrt(N / 0.8);
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
  par.epsilon = 1.0; par.sigma = 1.0; par.cutOff = 2.5; par.shift = shift;
  auto potA = std::make_shared<Potential::LJ>();
  potA->setPotParameters(0, 0, par);
  using PFA = PairForces<Potential::LJ>;
  PFA::Parameters pa; pa.box = box;
  auto A = std::make_shared<PFA>(pd, pa, potA);
  auto C = std::make_shared<PairForcesLJ64>(pd, box);
  C->setPotParameters(par);
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
  printf("{\"N\":%d,\"L\":%.6g,\"fmax\":%.6g,\"force_vs_ref\":%.6g,\"energy_vs_ref\":%.6g,\"virial_vs_ref\":%.6g}\n", N, (double)L, fmax,
         df / fmax, de / emax, dv / vmax);
  sys->finish();
  return 0;
}
