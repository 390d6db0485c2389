/* TEST INFRASTRUCTURE - reference harness for the DPD transverser (BASELINE config 4 arithmetic).
 *
 * At this commit the reference's Potential::DPD only exposes the pre-v2 getForceTransverser(box, pd); PairForces looks for
 * getTransverser(comp, box, pd) and silently falls back to a null transverser, so the stock
 * PairForces<Potential::DPD> computes nothing. The ARITHMETIC - DPD_impl::ForceTransverser::{getInfo,compute,set},
 * Interactor/Potential/DPD.cuh:92-159 - is intact. This main() of OUR OWN includes the UNMODIFIED reference headers and adds
 * the ten-line adaptor that forwards getTransverser to the reference's own getForceTransverser, so that the reference's
 * PairForces + CellList + transverseWithNeighbourContainer drive the reference's ForceTransverser unchanged.
 * Compiled by oracle/Makefile into oracle/_ref/ref_dpd (single precision). Never linked by the product.
 *
 * usage: ref_dpd N L rcut A gamma temperature dt sysseed calls pos.bin vel.bin outprefix
 *   calls: number of force evaluations (the reference increments `step` before each one, DPD.cuh:165); the forces of the
 *   LAST evaluation are written to outprefix.force.bin (float4[N]). Prints {"seed": low 32 bits of the Saru seed, "step": ...}.
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/NeighbourList/CellList.cuh"
#include "Interactor/Potential/DPD.cuh"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace uammd;

struct DPDWithTransverser : public Potential::DPD {
  using Potential::DPD::DPD;
  auto getTransverser(Interactor::Computables comp, Box box, std::shared_ptr<ParticleData> pd) {
    return this->getForceTransverser(box, pd);
  }
  int currentStep() const { return this->step; }
};

// the reference's own detection must now see a transverser (it does not for the stock Potential::DPD: SURVEY F3)
static_assert(Potential::has_getTransverser<DPDWithTransverser>::value, "adaptor not detected by PairForces");
static_assert(!Potential::has_getTransverser<Potential::DPD>::value, "the stock Potential::DPD gained a getTransverser");

template <class T> static std::vector<T> readBin(const std::string &fn, size_t n) {
  std::vector<T> v(n);
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", fn.c_str()); exit(2); }
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read %s\n", fn.c_str()); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char **argv) {
  if (argc < 13) return 1;
  const int N = atoi(argv[1]);
  const real L = atof(argv[2]);
  DPDWithTransverser::Parameters par;
  par.cutOff = atof(argv[3]);
  par.A = atof(argv[4]);
  par.gamma.gamma = atof(argv[5]);
  par.temperature = atof(argv[6]);
  par.dt = atof(argv[7]);
  const uint64_t sysseed = strtoull(argv[8], nullptr, 10);
  const int calls = atoi(argv[9]);
  const std::string out = argv[12];
  auto sys = std::make_shared<System>();
  sys->rng().setSeed(sysseed);
  auto pd = std::make_shared<ParticleData>(N, sys);
  {
    auto pos = pd->getPos(access::location::cpu, access::mode::write);
    auto vel = pd->getVel(access::location::cpu, access::mode::write);
    auto hp = readBin<real4>(argv[10], N);
    auto hv = readBin<real3>(argv[11], N);
    std::copy(hp.begin(), hp.end(), pos.begin());
    std::copy(hv.begin(), hv.end(), vel.begin());
  }
  auto pot = std::make_shared<DPDWithTransverser>(par);
  using PF = PairForces<DPDWithTransverser, CellList>;
  PF::Parameters pp;
  pp.box = Box(make_real3(L));
  auto pf = std::make_shared<PF>(pd, pp, pot);
  Interactor::Computables comp;
  comp.force = true;
  // getForceTransverser draws its seed as `static auto seed = sys->rng().next()` at the first evaluation (DPD.cuh:164);
  // Saru takes it as an unsigned int, i.e. its low 32 bits. Everything else is constructed by now, so the next draw is it.
  auto rngCopy = sys->rng();
  const unsigned int seed32 = (unsigned int)rngCopy.next();
  for (int c = 0; c < calls; c++) {
    {
      auto f = pd->getForce(access::location::gpu, access::mode::write);
      thrust::fill(thrust::cuda::par, f.begin(), f.end(), real4());
    }
    pf->sum(comp, 0);
    CudaSafeCall(cudaDeviceSynchronize());
  }
  {
    auto f = pd->getForce(access::location::cpu, access::mode::read);
    FILE *fo = fopen((out + ".force.bin").c_str(), "wb");
    fwrite(f.raw(), sizeof(real4), N, fo);
    fclose(fo);
  }
  printf("{\"mode\":\"dpd\",\"N\":%d,\"seed\":%u,\"step\":%d}\n", N, seed32, pot->currentStep());
  sys->finish();
  return 0;
}
