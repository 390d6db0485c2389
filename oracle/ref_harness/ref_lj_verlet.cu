/* TEST INFRASTRUCTURE - reference harness for PairForces<Potential::LJ, VerletList> (the neighbour list generic_md
 * instantiates). A tiny main() of OUR OWN over the UNMODIFIED reference headers
 * (Interactor/NeighbourList/VerletList.cuh:83-201, VerletList/VerletListBase.cuh:73-199, BasicList/BasicListBase.cuh:42-215).
 * Compiled by oracle/Makefile into oracle/_ref/ref_lj_verlet. Never linked by the product.
 *
 * usage:
 *   ref_lj_verlet forces N Lx Ly Lz rc sigma eps pos.bin outprefix
 *       dumps the Verlet list of the first update (numberNeighbours int[N], neighbourList int[maxk*N] in the
 *       reference's [k*N + i] layout truncated to maxk = max numberNeighbours, groupIndex int[N]) and the LJ forces
 *   ref_lj_verlet md N Lx Ly Lz rc sigma eps dt warmup steps flush pos.bin vel.bin outprefix|-
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/NeighbourList/VerletList.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "Integrator/VerletNVE.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace uammd;
using PF = PairForces<Potential::LJ, VerletList>;

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

int main(int argc, char *argv[]) {
  if (argc < 11) return 1;
  std::string mode = argv[1];
  int a = 2;
  const int N = atoi(argv[a++]);
  real Lx = atof(argv[a++]), Ly = atof(argv[a++]), Lz = atof(argv[a++]);
  real rc = atof(argv[a++]), sigma = atof(argv[a++]), eps = atof(argv[a++]);
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  Box box(make_real3(Lx, Ly, Lz));
  auto pot = std::make_shared<Potential::LJ>();
  Potential::LJ::InputPairParameters ip;
  ip.epsilon = eps; ip.shift = false; ip.sigma = sigma; ip.cutOff = rc;
  pot->setPotParameters(0, 0, ip);
  auto nl = std::make_shared<VerletList>(pd);
  PF::Parameters params;
  params.box = box;
  params.nl = nl;
  auto pf = std::make_shared<PF>(pd, params, pot);
  if (mode == "forces") {
    std::string posf = argv[a++], out = argv[a++];
    {
      auto h = readBin<float4>(posf, N);
      auto pos = pd->getPos(access::location::cpu, access::mode::write);
      for (int i = 0; i < N; i++) pos[i] = make_real4(h[i].x, h[i].y, h[i].z, h[i].w);
      auto f = pd->getForce(access::location::cpu, access::mode::write);
      std::fill(f.begin(), f.end(), real4());
    }
    Interactor::Computables comp;
    comp.force = true;
    pf->sum(comp, 0);
    CudaSafeCall(cudaDeviceSynchronize());
    auto vl = nl->getVerletList();
    std::vector<int> nn(N), gi(N);
    CudaSafeCall(cudaMemcpy(nn.data(), vl.numberNeighbours, N * sizeof(int), cudaMemcpyDeviceToHost));
    CudaSafeCall(cudaMemcpy(gi.data(), vl.groupIndex, N * sizeof(int), cudaMemcpyDeviceToHost));
    const int maxk = *std::max_element(nn.begin(), nn.end());
    std::vector<int> list((size_t)maxk * N);
    CudaSafeCall(cudaMemcpy(list.data(), vl.neighbourList, list.size() * sizeof(int), cudaMemcpyDeviceToHost));
    writeBin(out + ".nn.bin", nn.data(), N);
    writeBin(out + ".index.bin", gi.data(), N);
    writeBin(out + ".list.bin", list.data(), list.size());
    auto f = pd->getForce(access::location::cpu, access::mode::read);
    writeBin(out + ".force.bin", f.raw(), N);
    printf("{\"mode\":\"forces\",\"N\":%d,\"maxk\":%d,\"stride\":%d}\n", N, maxk, (int)vl.particleStride[0]);
  } else {
    real dt = atof(argv[a++]);
    int warm = atoi(argv[a++]), steps = atoi(argv[a++]), flush = atoi(argv[a++]);
    std::string posf = argv[a++], velf = argv[a++], out = argv[a++];
    {
      auto h = readBin<float4>(posf, N);
      auto hv = readBin<float>(velf, 3 * (size_t)N);
      auto pos = pd->getPos(access::location::cpu, access::mode::write);
      auto vel = pd->getVel(access::location::cpu, access::mode::write);
      for (int i = 0; i < N; i++) {
        pos[i] = make_real4(h[i].x, h[i].y, h[i].z, h[i].w);
        vel[i] = make_real3(hv[3 * i], hv[3 * i + 1], hv[3 * i + 2]);
      }
    }
    VerletNVE::Parameters par;
    par.dt = dt;
    par.initVelocities = false;
    auto verlet = std::make_shared<VerletNVE>(pd, par);
    verlet->addInteractor(pf);
    for (int i = 0; i < warm; i++) verlet->forwardTime();
    CudaSafeCall(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    char *scrub = nullptr;
    const size_t scrubBytes = 256ull << 20;
    if (flush) CudaSafeCall(cudaMalloc(&scrub, scrubBytes));
    double total = 0;
    int rebuilds = 0;
    for (int i = 0; i < steps; i++) {
      if (flush) CudaSafeCall(cudaMemsetAsync(scrub, i & 0xff, scrubBytes, 0));
      cudaEventRecord(e0, 0);
      verlet->forwardTime();
      cudaEventRecord(e1, 0);
      CudaSafeCall(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      total += ms;
      if (nl->getNumberOfStepsSinceLastUpdate() == 0) rebuilds++;
    }
    printf("{\"mode\":\"md_summary\",\"N\":%d,\"steps\":%d,\"ms_per_step_mean\":%.6f,\"rebuilds\":%d}\n", N, steps, total / steps, rebuilds);
    if (out != "-") {
      auto p = pd->getPos(access::location::cpu, access::mode::read);
      writeBin(out + ".pos_final.bin", p.raw(), N);
    }
  }
  sys->finish();
  return 0;
}
