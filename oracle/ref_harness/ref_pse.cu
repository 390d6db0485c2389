/* TEST INFRASTRUCTURE - reference harness for BDHI::PSE (Positively Split Ewald RPY), BASELINE config 3.
 *
 * A tiny main() of OUR OWN that includes the UNMODIFIED reference headers under /root/reference/src and drives
 * BDHI::PSE (Integrator/BDHI/BDHI_PSE.cuh:82-176). The reference's Lanczos needs <cblas.h>/<lapacke.h>, absent in
 * this container: oracle/ref_harness/shim/ provides the two functions it calls (gemv, steqr). Compiled by
 * oracle/Makefile into oracle/_ref/ref_pse (-DDOUBLE_PRECISION) and oracle/_ref/ref_pse_f32. Used by tests/ (-m gpu)
 * as the parity oracle and by bench.py's PSE reference leg. Never linked by the product.
 *
 * usage:
 *   ref_pse mdot N L viscosity a tolerance psi shear temperature dt sysseed pos.bin force.bin outprefix
 *       writes outprefix.far.bin  (computeMFFarField: Mw F + far noise, real3[N])
 *              outprefix.near.bin (computeMFNearField: Mr F)
 *              outprefix.bdw.bin  (computeBdW: near-field noise through Lanczos; zeros when temperature == 0)
 *   ref_pse time N L viscosity a tolerance psi shear temperature dt sysseed warmup steps flush pos.bin force.bin
 *       times the body of BDHI::EulerMaruyama<PSE>::forwardTime: computeMF + computeBdW + position update.
 * pos.bin / force.bin: real4[N] in the build's precision.
 */
#include "uammd.cuh"
#include "Integrator/BDHI/BDHI_PSE.cuh"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace uammd;

template <class T> static std::vector<T> readBin(const std::string &fn, size_t n) {
  std::vector<T> v(n);
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", fn.c_str()); exit(2); }
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read %s\n", fn.c_str()); exit(2); }
  fclose(f);
  return v;
}
static void dump(const std::string &fn, thrust::device_vector<real3> &v) {
  std::vector<real3> h(v.size());
  thrust::copy(v.begin(), v.end(), h.begin());
  FILE *f = fopen(fn.c_str(), "wb");
  fwrite(h.data(), sizeof(real3), h.size(), f);
  fclose(f);
}

struct UpdatePos {
  real dt, sq;
  __device__ real4 operator()(thrust::tuple<real4, real3, real3> t) const {
    real4 p = thrust::get<0>(t);
    real3 m = thrust::get<1>(t), b = thrust::get<2>(t);
    return make_real4(p.x + m.x * dt + sq * b.x, p.y + m.y * dt + sq * b.y, p.z + m.z * dt + sq * b.z, p.w);
  }
};

int main(int argc, char **argv) {
  if (argc < 14) return 1;
  const std::string mode = argv[1];
  int a = 2;
  const int N = atoi(argv[a++]);
  const real L = atof(argv[a++]);
  BDHI::PSE::Parameters par;
  par.viscosity = atof(argv[a++]);
  par.hydrodynamicRadius = atof(argv[a++]);
  par.tolerance = atof(argv[a++]);
  par.psi = atof(argv[a++]);
  par.shearStrain = atof(argv[a++]);
  par.temperature = atof(argv[a++]);
  par.dt = atof(argv[a++]);
  par.box = Box(make_real3(L));
  const uint64_t sysseed = strtoull(argv[a++], nullptr, 10);
  auto sys = std::make_shared<System>();
  sys->rng().setSeed(sysseed);
  auto pd = std::make_shared<ParticleData>(N, sys);
  int warm = 0, steps = 0, flush = 0;
  if (mode == "time") { warm = atoi(argv[a++]); steps = atoi(argv[a++]); flush = atoi(argv[a++]); }
  {
    auto hp = readBin<real4>(argv[a++], N), hf = readBin<real4>(argv[a++], N);
    auto pos = pd->getPos(access::location::cpu, access::mode::write);
    auto force = pd->getForce(access::location::cpu, access::mode::write);
    std::copy(hp.begin(), hp.end(), pos.begin());
    std::copy(hf.begin(), hf.end(), force.begin());
  }
  auto pse = std::make_shared<BDHI::PSE>(pd, par);
  thrust::device_vector<real3> MF(N), BdW(N);
  if (mode == "mdot") {
    const std::string out = argv[a++];
    thrust::fill(MF.begin(), MF.end(), real3());
    pse->computeMFFarField(MF.data().get(), 0);
    CudaSafeCall(cudaDeviceSynchronize());
    dump(out + ".far.bin", MF);
    thrust::fill(MF.begin(), MF.end(), real3());
    pse->computeMFNearField(MF.data().get(), 0);
    CudaSafeCall(cudaDeviceSynchronize());
    dump(out + ".near.bin", MF);
    thrust::fill(BdW.begin(), BdW.end(), real3());
    pse->computeBdW(BdW.data().get(), 0);
    CudaSafeCall(cudaDeviceSynchronize());
    dump(out + ".bdw.bin", BdW);
    printf("{\"mode\":\"mdot\",\"N\":%d,\"M0\":%.17g}\n", N, (double)pse->getSelfMobility());
  } else {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    char *scrub = nullptr;
    const size_t scrubBytes = 256ull << 20;
    if (flush) CudaSafeCall(cudaMalloc(&scrub, scrubBytes));
    double total = 0;
    const real sq = sqrt(2 * par.temperature * par.dt);
    for (int i = 0; i < warm + steps; i++) {
      if (flush) CudaSafeCall(cudaMemsetAsync(scrub, i & 0xff, scrubBytes, 0));
      cudaEventRecord(e0, 0);
      {
        pse->computeMF(MF.data().get(), 0);
        if (par.temperature > 0) pse->computeBdW(BdW.data().get(), 0);
        auto pos = pd->getPos(access::location::gpu, access::mode::readwrite);
        thrust::device_ptr<real4> pp(pos.raw());
        auto zip = thrust::make_zip_iterator(thrust::make_tuple(pp, MF.begin(), BdW.begin()));
        thrust::transform(thrust::cuda::par, zip, zip + N, pp, UpdatePos{par.dt, par.temperature > 0 ? sq : real(0)});
      }
      cudaEventRecord(e1, 0);
      CudaSafeCall(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (i >= warm) total += ms;
    }
    printf("{\"mode\":\"time\",\"N\":%d,\"steps\":%d,\"ms_per_step\":%.6f,\"steps_per_s\":%.3f}\n", N, steps, total / steps,
           1000.0 * steps / total);
  }
  sys->finish();
  return 0;
}
