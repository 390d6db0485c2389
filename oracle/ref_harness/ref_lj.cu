/* TEST INFRASTRUCTURE - reference harness for path 1 (LJ pair forces over CellList + VerletNVE).
 *
 * This is a tiny main() of OUR OWN that includes the UNMODIFIED reference headers where they lie
 * under /root/reference/src (never copied into this repo) and drives the reference's stock code
 * path: PairForces<Potential::LJ, CellList> (src/Interactor/PairForces.cu:43-78) and VerletNVE
 * (src/Integrator/VerletNVE.cu:174-188). It is compiled by oracle/Makefile into oracle/_ref/ref_lj
 * (git-ignored; travels to the GPU box with the snapshot). It is used
 *   - by tests/ (-m gpu) as the parity oracle: dumps sortPos/groupIndex/cellStart/cellEnd/force
 *   - by bench.py --impl reference as the reference arm (the reference has no CPU implementation,
 *     its only implementation is this CUDA path).
 * Nothing in the product path links or executes it.
 *
 * usage:
 *   ref_lj forces N Lx Ly Lz rc sigma eps shift pos.bin outprefix
 *   ref_lj md     N Lx Ly Lz rc sigma eps dt warmup steps reps flush pos.bin vel.bin outprefix
 *                 (flush=1: a 256 MiB write evicts L2 before every step and steps are timed one by one,
 *                  the same protocol bench.py applies to the new engine)
 * pos.bin: float4[N] (x,y,z,type); vel.bin: float3[N]
 */
#include "uammd.cuh"
#include "Interactor/PairForces.cuh"
#include "Interactor/NeighbourList/CellList.cuh"
#include "Interactor/Potential/Potential.cuh"
#include "Integrator/VerletNVE.cuh"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>

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
  if (!f) { fprintf(stderr, "cannot open %s\n", fn.c_str()); exit(2); }
  fwrite(p, sizeof(T), n, f);
  fclose(f);
}

using PF = PairForces<Potential::LJ, CellList>;

static std::shared_ptr<PF> makePairForces(std::shared_ptr<ParticleData> pd, std::shared_ptr<CellList> nl,
                                          Box box, real rc, real sigma, real eps, bool shift) {
  auto pot = std::make_shared<Potential::LJ>();
  Potential::LJ::InputPairParameters par;
  par.epsilon = eps;
  par.shift = shift;
  par.sigma = sigma;
  par.cutOff = rc;
  pot->setPotParameters(0, 0, par);
  PF::Parameters params;
  params.box = box;
  params.nl = nl;
  return std::make_shared<PF>(pd, params, pot);
}

int main(int argc, char *argv[]) {
  if (argc < 2) return 1;
  std::string mode = argv[1];
  int a = 2;
  int N = atoi(argv[a++]);
  real Lx = atof(argv[a++]), Ly = atof(argv[a++]), Lz = atof(argv[a++]);
  real rc = atof(argv[a++]), sigma = atof(argv[a++]), eps = atof(argv[a++]);
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  Box box(make_real3(Lx, Ly, Lz));
  if (mode == "forces") {
    bool shift = atoi(argv[a++]);
    std::string posf = argv[a++], out = argv[a++];
    {
      auto h = readBin<float4>(posf, N);
      auto pos = pd->getPos(access::location::cpu, access::mode::write);
      for (int i = 0; i < N; i++) pos[i] = make_real4(h[i].x, h[i].y, h[i].z, h[i].w);
    }
    auto nl = std::make_shared<CellList>(pd);
    auto pf = makePairForces(pd, nl, box, rc, sigma, eps, shift);
    {
      auto f = pd->getForce(access::location::gpu, access::mode::write);
      thrust::fill(thrust::cuda::par, f.begin(), f.end(), real4());
      auto e = pd->getEnergy(access::location::gpu, access::mode::write);
      thrust::fill(thrust::cuda::par, e.begin(), e.end(), real());
      auto v = pd->getVirial(access::location::gpu, access::mode::write);
      thrust::fill(thrust::cuda::par, v.begin(), v.end(), real());
    }
    Interactor::Computables comp;
    comp.force = true; comp.energy = true; comp.virial = true;
    pf->sum(comp, 0);
    CudaSafeCall(cudaDeviceSynchronize());
    auto cl = nl->getCellList();
    int ncells = cl.grid.getNumberCells();
    std::vector<real4> sp(N);
    std::vector<int> gi(N), ce(ncells);
    std::vector<uint> cs(ncells);
    CudaSafeCall(cudaMemcpy(sp.data(), cl.sortPos, N * sizeof(real4), cudaMemcpyDeviceToHost));
    CudaSafeCall(cudaMemcpy(gi.data(), cl.groupIndex, N * sizeof(int), cudaMemcpyDeviceToHost));
    CudaSafeCall(cudaMemcpy(cs.data(), cl.cellStart, ncells * sizeof(uint), cudaMemcpyDeviceToHost));
    CudaSafeCall(cudaMemcpy(ce.data(), cl.cellEnd, ncells * sizeof(int), cudaMemcpyDeviceToHost));
    // Normalise the VALID_CELL epoch trick (CellListBase.cuh:210-230): empty -> start=end=-1
    std::vector<int> cs_n(ncells), ce_n(ncells);
    for (int c = 0; c < ncells; c++) {
      bool empty = cs[c] < cl.VALID_CELL;
      cs_n[c] = empty ? -1 : int(cs[c] - cl.VALID_CELL);
      ce_n[c] = empty ? -1 : ce[c];
    }
    int3 cd = cl.grid.cellDim;
    int meta[4] = {cd.x, cd.y, cd.z, ncells};
    writeBin(out + ".celldim.bin", meta, 4);
    writeBin(out + ".sortpos.bin", sp.data(), N);
    writeBin(out + ".index.bin", gi.data(), N);
    writeBin(out + ".cellstart.bin", cs_n.data(), ncells);
    writeBin(out + ".cellend.bin", ce_n.data(), ncells);
    {
      auto f = pd->getForce(access::location::cpu, access::mode::read);
      writeBin(out + ".force.bin", f.raw(), N);
      auto e = pd->getEnergy(access::location::cpu, access::mode::read);
      writeBin(out + ".energy.bin", e.raw(), N);
      auto v = pd->getVirial(access::location::cpu, access::mode::read);
      writeBin(out + ".virial.bin", v.raw(), N);
    }
    printf("{\"mode\":\"forces\",\"N\":%d,\"celldim\":[%d,%d,%d]}\n", N, cd.x, cd.y, cd.z);
  } else if (mode == "md") {
    real dt = atof(argv[a++]);
    int warm = atoi(argv[a++]), steps = atoi(argv[a++]), reps = atoi(argv[a++]), flush = atoi(argv[a++]);
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
    auto nl = std::make_shared<CellList>(pd);
    verlet->addInteractor(makePairForces(pd, nl, box, rc, sigma, eps, false));
    for (int i = 0; i < warm; i++) verlet->forwardTime();
    CudaSafeCall(cudaDeviceSynchronize());
    if (out != "-") {
      auto p = pd->getPos(access::location::cpu, access::mode::read);
      writeBin(out + ".pos_warm.bin", p.raw(), N);
      auto v = pd->getVel(access::location::cpu, access::mode::read);
      writeBin(out + ".vel_warm.bin", v.raw(), N);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30, total = 0;
    for (int r = 0; r < reps; r++) {
      CudaSafeCall(cudaDeviceSynchronize());
      // VerletNVE enqueues on its own private stream (VerletNVE.cu:56); legacy-stream events order with it.
      float ms = 0;
      if (!flush) {
        cudaEventRecord(e0, 0);
        for (int i = 0; i < steps; i++) verlet->forwardTime();
        cudaEventRecord(e1, 0);
        CudaSafeCall(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
      } else {
        static char *scrub = nullptr;
        const size_t scrubBytes = 256ull << 20;
        if (!scrub) CudaSafeCall(cudaMalloc(&scrub, scrubBytes));
        for (int i = 0; i < steps; i++) {
          CudaSafeCall(cudaMemsetAsync(scrub, i & 0xff, scrubBytes, 0));
          cudaEventRecord(e0, 0);
          verlet->forwardTime();
          cudaEventRecord(e1, 0);
          CudaSafeCall(cudaEventSynchronize(e1));
          float m1; cudaEventElapsedTime(&m1, e0, e1);
          ms += m1;
        }
      }
      best = std::min(best, (double)ms); total += ms;
      printf("{\"mode\":\"md\",\"rep\":%d,\"N\":%d,\"steps\":%d,\"ms\":%.4f,\"ms_per_step\":%.6f,\"steps_per_s\":%.3f}\n",
             r, N, steps, ms, ms / steps, 1000.0 * steps / ms);
    }
    printf("{\"mode\":\"md_summary\",\"N\":%d,\"steps\":%d,\"reps\":%d,\"ms_per_step_mean\":%.6f,\"ms_per_step_best\":%.6f}\n",
           N, steps, reps, total / reps / steps, best / steps);
    if (out != "-") {
      auto p = pd->getPos(access::location::cpu, access::mode::read);
      writeBin(out + ".pos_final.bin", p.raw(), N);
    }
  } else {
    fprintf(stderr, "unknown mode\n");
    return 1;
  }
  sys->finish();
  return 0;
}
