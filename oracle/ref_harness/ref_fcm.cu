/* TEST INFRASTRUCTURE - reference harness for path 2 (BDHI::FCM), double precision build.
 *
 * A tiny main() of OUR OWN that includes the UNMODIFIED reference headers from /root/reference/src and drives
 * FCM_impl<Kernel, GaussianTorque>::computeHydrodynamicDisplacements (Integrator/BDHI/FCM/FCM_impl.cuh:652-693)
 * with Kernel = Peskin::threePoint (the BASELINE config; the stock BDHI::FCM hard-codes the Gaussian kernel,
 * SURVEY F4), Peskin::fourPoint or Gaussian. Compiled by oracle/Makefile into oracle/_ref/ref_fcm. Used by
 * tests/ (-m gpu) as the parity oracle and by bench.py's FCM reference leg. Never linked by the product.
 *
 * usage:
 *   ref_fcm mdot  KERNEL N L n viscosity tolerance temperature prefactor seed pos.bin force.bin out.bin
 *   ref_fcm mdott KERNEL N L n viscosity tolerance temperature prefactor seed pos.bin force.bin torque.bin lin.bin ang.bin
 *                 (rotational FCM: KernelTorque = GaussianTorque built like detail::initializeKernelTorque, BDHI_FCM.cuh:69-80)
 *   ref_fcm time  KERNEL N L n viscosity tolerance temperature dt warmup steps flush pos.bin force.bin
 * KERNEL: peskin3 | peskin4 | gaussian ; pos.bin / force.bin: double4[N] ; out.bin: double3[N]
 * "time" runs the body of BDHI::EulerMaruyama<FCM>::forwardTime (BDHI_EulerMaruyama.cu:125-166) with fixed
 * external forces: computeHydrodynamicDisplacements(T, 1/sqrt(dt)) + copy into MF (BDHI_FCM.cuh:131-142) +
 * position update x += MF dt (integrateGPUD :82-113).
 */
#include "uammd.cuh"
#include "Integrator/BDHI/FCM/FCM_impl.cuh"
#include "Integrator/BDHI/FCM/FCM_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace uammd;
using BDHI::FCM_impl;
using KernelTorque = BDHI::FCM_ns::Kernels::GaussianTorque;

template <class T> static std::vector<T> readBin(const std::string &fn, size_t n) {
  std::vector<T> v(n);
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", fn.c_str()); exit(2); }
  if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read %s\n", fn.c_str()); exit(2); }
  fclose(f);
  return v;
}

struct UpdatePos {
  real dt;
  __device__ real4 operator()(thrust::tuple<real4, real3> t) const {
    real4 p = thrust::get<0>(t);
    real3 m = thrust::get<1>(t);
    return make_real4(p.x + m.x * dt, p.y + m.y * dt, p.z + m.z * dt, p.w);
  }
};

template <class Kernel> int run(int argc, char **argv) {
  using FCM = FCM_impl<Kernel, KernelTorque>;
  std::string mode = argv[1];
  int a = 3;
  const int N = atoi(argv[a++]);
  const real L = atof(argv[a++]);
  const int n = atoi(argv[a++]);
  const real viscosity = atof(argv[a++]);
  const real tolerance = atof(argv[a++]);
  const real temperature = atof(argv[a++]);
  typename FCM::Parameters par;
  par.viscosity = viscosity;
  par.tolerance = tolerance;
  par.box = Box(make_real3(L));
  par.cells = make_int3(n, n, n);
  const real h = L / n;
  par.kernel = std::make_shared<Kernel>(h, tolerance);
  par.kernelTorque = std::make_shared<KernelTorque>(real(1.0), h, real(1e-3)); // unused (no torques)
  par.hydrodynamicRadius = par.kernel->fixHydrodynamicRadius(h, h);
  if (mode == "mdott") {
    const real prefactor = atof(argv[a++]);
    par.seed = (uint)atoll(argv[a++]);
    std::string posf = argv[a++], forcef = argv[a++], torquef = argv[a++], linf = argv[a++], angf = argv[a++];
    const real width = par.hydrodynamicRadius / (pow(6 * sqrt(M_PI), 1 / 3.));
    par.kernelTorque = std::make_shared<KernelTorque>(width, h, tolerance);
    auto fcm = std::make_shared<FCM>(par);
    auto hp = readBin<real4>(posf, N), hf = readBin<real4>(forcef, N), ht = readBin<real4>(torquef, N);
    thrust::device_vector<real4> pos(hp), force(hf), torque(ht);
    auto disp = fcm->computeHydrodynamicDisplacements(pos.data().get(), force.data().get(), torque.data().get(), N, temperature,
                                                      prefactor, 0);
    CudaSafeCall(cudaDeviceSynchronize());
    std::vector<real3> lin(N), ang(N);
    CudaSafeCall(cudaMemcpy(lin.data(), disp.first.data().get(), N * sizeof(real3), cudaMemcpyDeviceToHost));
    CudaSafeCall(cudaMemcpy(ang.data(), disp.second.data().get(), N * sizeof(real3), cudaMemcpyDeviceToHost));
    FILE *f = fopen(linf.c_str(), "wb"); fwrite(lin.data(), sizeof(real3), N, f); fclose(f);
    f = fopen(angf.c_str(), "wb"); fwrite(ang.data(), sizeof(real3), N, f); fclose(f);
    printf("{\"mode\":\"mdott\",\"N\":%d,\"n\":%d,\"support\":%d,\"supportTorque\":%d,\"a\":%.17g,\"widthTorque\":%.17g}\n", N, n,
           (int)par.kernel->support, (int)par.kernelTorque->support, (double)par.hydrodynamicRadius, (double)width);
  } else if (mode == "mdot") {
    const real prefactor = atof(argv[a++]);
    par.seed = (uint)atoll(argv[a++]);
    std::string posf = argv[a++], forcef = argv[a++], outf = argv[a++];
    auto fcm = std::make_shared<FCM>(par);
    auto hp = readBin<real4>(posf, N), hf = readBin<real4>(forcef, N);
    thrust::device_vector<real4> pos(hp), force(hf);
    auto disp = fcm->computeHydrodynamicDisplacements(pos.data().get(), force.data().get(), nullptr, N, temperature,
                                                      prefactor, 0);
    CudaSafeCall(cudaDeviceSynchronize());
    std::vector<real3> out(N);
    CudaSafeCall(cudaMemcpy(out.data(), disp.first.data().get(), N * sizeof(real3), cudaMemcpyDeviceToHost));
    FILE *f = fopen(outf.c_str(), "wb");
    fwrite(out.data(), sizeof(real3), N, f);
    fclose(f);
    printf("{\"mode\":\"mdot\",\"N\":%d,\"n\":%d,\"support\":%d,\"a\":%.17g}\n", N, n, (int)par.kernel->support,
           (double)par.hydrodynamicRadius);
  } else {
    const real dt = atof(argv[a++]);
    const int warm = atoi(argv[a++]), steps = atoi(argv[a++]), flush = atoi(argv[a++]);
    std::string posf = argv[a++], forcef = argv[a++];
    par.seed = 1234;
    auto fcm = std::make_shared<FCM>(par);
    auto hp = readBin<real4>(posf, N), hf = readBin<real4>(forcef, N);
    thrust::device_vector<real4> pos(hp), force(hf);
    thrust::device_vector<real3> MF(N);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    char *scrub = nullptr;
    const size_t scrubBytes = 256ull << 20;
    if (flush) CudaSafeCall(cudaMalloc(&scrub, scrubBytes));
    double total = 0;
    const real prefactor = real(1.0) / sqrt(dt);
    for (int i = 0; i < warm + steps; i++) {
      if (flush) CudaSafeCall(cudaMemsetAsync(scrub, i & 0xff, scrubBytes, 0));
      cudaEventRecord(e0, 0);
      {
        auto disp = fcm->computeHydrodynamicDisplacements(pos.data().get(), force.data().get(), nullptr, N,
                                                          temperature, prefactor, 0);
        thrust::copy(thrust::cuda::par, disp.first.begin(), disp.first.end(), MF.begin());
        auto zip = thrust::make_zip_iterator(thrust::make_tuple(pos.begin(), MF.begin()));
        thrust::transform(thrust::cuda::par, zip, zip + N, pos.begin(), UpdatePos{dt});
      }
      cudaEventRecord(e1, 0);
      CudaSafeCall(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (i >= warm) total += ms;
    }
    printf("{\"mode\":\"time\",\"N\":%d,\"n\":%d,\"steps\":%d,\"ms_per_step\":%.6f,\"steps_per_s\":%.3f}\n", N, n, steps,
           total / steps, 1000.0 * steps / total);
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 4) return 1;
  auto sys = std::make_shared<System>();
  std::string k = argv[2];
  int rc;
  if (k == "peskin3") rc = run<BDHI::FCM_ns::Kernels::Peskin::threePoint>(argc, argv);
  else if (k == "peskin4") rc = run<BDHI::FCM_ns::Kernels::Peskin::fourPoint>(argc, argv);
  else rc = run<BDHI::FCM_ns::Kernels::Gaussian>(argc, argv);
  sys->finish();
  return rc;
}
