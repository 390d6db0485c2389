/* Drop-in demonstration / integration test for path 2 (double precision build, -DDOUBLE_PRECISION).
 * The reference's BDHI::EulerMaruyama<Method> integrator instantiated with Method = b200::FCM<Peskin::threePoint>
 * next to the reference FCM_impl<Peskin::threePoint, GaussianTorque>: MF of the first step is compared, then both
 * take deterministic (T = 0) steps under a constant force and the final positions are compared.
 * Built by oracle/Makefile into oracle/_ref/dropin_fcm; run by tests/test_dropin_gpu.py.  usage: dropin_fcm N n
 */
#include "uammd.cuh"
#include "Integrator/BDHI/BDHI_EulerMaruyama.cuh"
#include "Integrator/BDHI/FCM/FCM_impl.cuh"
#include "uammd_b200/uammd_b200.cuh"
#include <random>
using namespace uammd;
using Kernel = BDHI::FCM_ns::Kernels::Peskin::threePoint;
using RefFCM = BDHI::FCM_impl<Kernel, BDHI::FCM_ns::Kernels::GaussianTorque>;

// constant external force interactor
struct ConstantForce : public Interactor {
  thrust::device_vector<real4> f;
  ConstantForce(std::shared_ptr<ParticleData> pd, const std::vector<real4> &h) : Interactor(pd, "ConstantForce"), f(h) {}
  struct Add { __device__ real4 operator()(real4 a, real4 b) const { return make_real4(a.x + b.x, a.y + b.y, a.z + b.z, a.w); } };
  void sum(Computables comp, cudaStream_t st) override {
    auto force = pd->getForce(access::gpu, access::readwrite);
    thrust::transform(thrust::cuda::par.on(st), force.begin(), force.end(), f.begin(), force.begin(), Add());
  }
};

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 20000;
  const int n = argc > 2 ? atoi(argv[2]) : 64;
  const real L = n;
  auto sys = std::make_shared<System>();
  auto pd = std::make_shared<ParticleData>(N, sys);
  std::vector<real4> hf(N);
  {
    auto pos = pd->getPos(access::cpu, access::write);
    std::mt19937_64 gen(11);
    std::uniform_real_distribution<double> U(-0.5, 0.5);
    std::normal_distribution<double> G(0, 1);
    for (int i = 0; i < N; i++) {
      pos[i] = make_real4(U(gen) * L, U(gen) * L, U(gen) * L, 0);
      hf[i] = make_real4(G(gen), G(gen), G(gen), 0);
    }
  }
  using Method = b200::FCM<Kernel>;
  Method::Parameters par;
  par.temperature = 0; par.viscosity = 1.0; par.dt = 0.01; par.box = Box(make_real3(L));
  par.cells = make_int3(n, n, n); par.tolerance = 1e-3; par.seed = 1234;
  auto bdhi = std::make_shared<BDHI::EulerMaruyama<Method>>(pd, par);
  bdhi->addInteractor(std::make_shared<ConstantForce>(pd, hf));

  // reference MF on the same initial state
  RefFCM::Parameters rp;
  rp.temperature = 0; rp.viscosity = 1.0; rp.dt = 0.01; rp.box = par.box; rp.cells = par.cells; rp.tolerance = 1e-3; rp.seed = 1234;
  rp.kernel = std::make_shared<Kernel>(L / n, rp.tolerance);
  rp.kernelTorque = std::make_shared<BDHI::FCM_ns::Kernels::GaussianTorque>(real(1.0), L / n, real(1e-3));
  rp.hydrodynamicRadius = L / n;
  auto ref = std::make_shared<RefFCM>(rp);
  thrust::device_vector<real4> dforce(hf);
  thrust::device_vector<real4> rpos(N);
  {
    auto pos = pd->getPos(access::gpu, access::read);
    thrust::copy(thrust::cuda::par, pos.begin(), pos.end(), rpos.begin());
  }
  const int steps = 5;
  for (int s = 0; s < steps; s++) {
    bdhi->forwardTime();
    auto disp = ref->computeHydrodynamicDisplacements(rpos.data().get(), dforce.data().get(), nullptr, N, 0, 0, 0);
    thrust::host_vector<real3> m = disp.first;
    thrust::host_vector<real4> p = rpos;
    for (int i = 0; i < N; i++) p[i] = make_real4(p[i].x + m[i].x * par.dt, p[i].y + m[i].y * par.dt, p[i].z + m[i].z * par.dt, p[i].w);
    rpos = p;
  }
  CudaSafeCall(cudaDeviceSynchronize());
  double dmax = 0, moved = 0;
  {
    auto pos = pd->getPos(access::cpu, access::read);
    thrust::host_vector<real4> p = rpos;
    for (int i = 0; i < N; i++)
      dmax = std::max({dmax, std::abs((double)pos[i].x - p[i].x), std::abs((double)pos[i].y - p[i].y), std::abs((double)pos[i].z - p[i].z)});
  }
  printf("{\"N\":%d,\"n\":%d,\"steps\":%d,\"max_dpos_vs_reference\":%.6g,\"a\":%.6g,\"M0\":%.10g}\n", N, n, steps, dmax,
         (double)bdhi->getHydrodynamicRadius(), (double)bdhi->getSelfMobility());
  sys->finish();
  return 0;
}
