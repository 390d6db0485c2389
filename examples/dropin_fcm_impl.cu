/* The reference's own tests of FCM_impl and IBM re-hosted on the glue classes (double precision build, no GoogleTest in
 * this image: plain main with the same constants).
 *
 *  1. test/BDHI/FCM/fcm_test.cu:85-144  FCM_impl, SelfMobilityIsCorrectUpToTolerance: eta = 1.12321, a = 1.012312,
 *     L = 96 h ceil(a / h), Gaussian kernel at tolerance 1e-8, 20 Saru(1234) positions, unit force along x, y, z:
 *     every component of the displacement within the tolerance of the Hasimoto self mobility.
 *     The same loop runs through the reference's FCM_impl for comparison (largest difference reported).
 *  2. test/misc/ibm/test_ibm_regular.cu:113-136,240-274  Peskin 3-point spread and gather of one particle on an 8^3 grid
 *     against the manual triple loops of the test (1e-10), and the adjointness <S f, u> = <f, J u> (:156-214).
 * Built by oracle/Makefile into oracle/_ref/dropin_fcm_impl; run by tests/test_dropin_gpu.py.
 */
#include "uammd.cuh"
#include "Integrator/BDHI/FCM/FCM_impl.cuh"
#include "misc/IBM.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
#include <vector>
using namespace uammd;
using Kernel = BDHI::FCM_ns::Kernels::Gaussian;
using KernelTorque = BDHI::FCM_ns::Kernels::GaussianTorque;

static real selfMobility(real a, real eta, real L) {
  long double rh = a, x = rh / L, x3 = x * x * x;
  const long double c = 2.83729747948061947666591710460773907l, b = 0.19457l;
  const long double a6pref = 16.0l * M_PIl * M_PIl / 45.0l + 630.0L * b * b;
  return 1.0l / (6.0l * M_PIl * eta * rh) * (1.0l - c * x + (4.0l / 3.0l) * M_PIl * x3 - a6pref * x3 * x3);
}

template <class FCMType> static void selfMobilityRun(typename FCMType::Parameters par, real a, real3 L, std::vector<real3> &out) {
  auto fcm = std::make_shared<FCMType>(par);
  thrust::device_vector<real4> pos(1), force(1);
  Saru rng(1234);
  for (int j = 0; j < 20; j++) {
    const real3 p = make_real3(rng.f(-0.5, 0.5), rng.f(-0.5, 0.5), rng.f(-0.5, 0.5)) * L;
    pos[0] = make_real4(p);
    for (int i = 0; i < 3; i++) {
      force[0] = make_real4(i == 0, i == 1, i == 2, 0);
      auto disp = fcm->computeHydrodynamicDisplacements(thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(force.data()),
                                                        nullptr, 1, 0, 0, 0);
      CudaSafeCall(cudaDeviceSynchronize());
      const real3 dx = disp.first[0];
      out.push_back(dx);
    }
  }
}

int main(int argc, char **argv) {
  const bool withReference = argc < 2 || atoi(argv[1]) != 0;
  auto sys = std::make_shared<System>();
  // ---- 1. FCM_impl self mobility
  const real tol = 1e-8, a = 1.012312, eta = 1.12321;
  const real h = Kernel::adviseGridSize(a, tol);
  const real3 L = make_real3(96 * h * ceil(a / h));
  const int3 cells = make_int3(L / h);
  auto mk = [&](auto &par) {
    par.viscosity = eta; par.tolerance = tol; par.dt = 1; par.cells = cells; par.box = Box(L); par.hydrodynamicRadius = a;
    const real hh = std::min({L.x / cells.x, L.y / cells.y, L.z / cells.z});
    par.kernel = std::make_shared<Kernel>(hh, tol);
    par.kernelTorque = std::make_shared<KernelTorque>(a / (pow(6 * sqrt(M_PI), 1 / 3.)), hh, tol);
    par.seed = 1234;
  };
  std::vector<real3> ours, ref;
  {
    b200::FCM_impl<Kernel, KernelTorque>::Parameters par;
    mk(par);
    selfMobilityRun<b200::FCM_impl<Kernel, KernelTorque>>(par, a, L, ours);
  }
  if (withReference) {
    BDHI::FCM_impl<Kernel, KernelTorque>::Parameters par;
    mk(par);
    selfMobilityRun<BDHI::FCM_impl<Kernel, KernelTorque>>(par, a, L, ref);
  }
  const real m0 = selfMobility(a, eta, L.x);
  double worst = 0, vsRef = 0;
  for (size_t k = 0; k < ours.size(); k++) {
    const int i = k % 3;
    const real3 want = make_real3(i == 0 ? m0 : 0, i == 1 ? m0 : 0, i == 2 ? m0 : 0);
    worst = std::max({worst, (double)std::abs(ours[k].x - want.x), (double)std::abs(ours[k].y - want.y), (double)std::abs(ours[k].z - want.z)});
    if (withReference)
      vsRef = std::max({vsRef, (double)std::abs(ours[k].x - ref[k].x), (double)std::abs(ours[k].y - ref[k].y), (double)std::abs(ours[k].z - ref[k].z)});
  }
  // ---- 2. IBM: Peskin 3-point spread / gather of one particle against the manual loops of the reference's test
  using Peskin = IBM_kernels::Peskin::threePoint;
  const int n = 8;
  const real Lb = 16.0, hb = Lb / n;
  Grid grid(Box(make_real3(Lb)), make_int3(n));
  auto kern = std::make_shared<Peskin>(hb);
  b200::IBM<Peskin> ibm(kern, grid);
  IBM<Peskin> refIbm(kern, grid);
  double spreadErr = 0, gatherErr = 0, spreadVsRef = 0, adjoint = 0;
  {
    const real3 p = make_real3(0.3, -1.1, 2.6);
    thrust::device_vector<real3> pos(1, p);
    thrust::device_vector<real> q(1, real(1.0));
    thrust::device_vector<real> field(n * n * n, real(0)), fieldRef(n * n * n, real(0));
    ibm.spread(pos.begin(), q.begin(), thrust::raw_pointer_cast(field.data()), 1);
    auto fr = thrust::raw_pointer_cast(fieldRef.data());
    refIbm.spread(thrust::raw_pointer_cast(pos.data()), thrust::raw_pointer_cast(q.data()), fr, 1);
    CudaSafeCall(cudaDeviceSynchronize());
    thrust::host_vector<real> hf = field, hr = fieldRef;
    std::vector<real> expect(n * n * n);
    Peskin host(hb);
    auto mic = [&](real d) { return d - Lb * std::floor(d / Lb + 0.5); };
    for (int iz = 0; iz < n; iz++)
      for (int iy = 0; iy < n; iy++)
        for (int ix = 0; ix < n; ix++) {
          const real x = -Lb / 2 + (ix + 0.5) * hb, y = -Lb / 2 + (iy + 0.5) * hb, z = -Lb / 2 + (iz + 0.5) * hb;
          expect[ix + n * (iy + n * iz)] = host.phi(mic(x - p.x)) * host.phi(mic(y - p.y)) * host.phi(mic(z - p.z));
        }
    for (int c = 0; c < n * n * n; c++) {
      spreadErr = std::max(spreadErr, (double)std::abs(hf[c] - expect[c]));
      spreadVsRef = std::max(spreadVsRef, (double)std::abs(hf[c] - hr[c]));
    }
    // gather of a smooth field: u(x) = sin(2 pi x / L) sampled at the cell centres, against the manual sum
    thrust::host_vector<real> hu(n * n * n);
    double manual = 0;
    for (int iz = 0; iz < n; iz++)
      for (int iy = 0; iy < n; iy++)
        for (int ix = 0; ix < n; ix++) {
          const real x = -Lb / 2 + (ix + 0.5) * hb;
          hu[ix + n * (iy + n * iz)] = std::sin(2 * M_PI * x / Lb) + 0.1 * iy - 0.05 * iz;
          manual += hu[ix + n * (iy + n * iz)] * expect[ix + n * (iy + n * iz)] * hb * hb * hb;
        }
    thrust::device_vector<real> u = hu;
    thrust::device_vector<real> Jq(1, real(0));
    ibm.gather(pos.begin(), thrust::raw_pointer_cast(Jq.data()), u.begin(), 1);
    CudaSafeCall(cudaDeviceSynchronize());
    const real got = Jq[0];
    gatherErr = std::abs(got - manual);
    // adjointness: <S q, u> dV = q J u
    double lhs = 0;
    for (int c = 0; c < n * n * n; c++) lhs += (double)hf[c] * hu[c] * hb * hb * hb;
    adjoint = std::abs(lhs - got);
  }
  printf("{\"cells\":%d,\"self_mobility\":%.12g,\"worst_vs_hasimoto\":%.3g,\"tolerance\":%.3g,\"max_vs_reference_fcm_impl\":%.3g,"
         "\"ibm_spread_vs_manual\":%.3g,\"ibm_spread_vs_reference\":%.3g,\"ibm_gather_vs_manual\":%.3g,\"ibm_adjointness\":%.3g}\n",
         cells.x, (double)m0, worst, (double)tol, vsRef, spreadErr, spreadVsRef, gatherErr, adjoint);
  sys->finish();
  return 0;
}
