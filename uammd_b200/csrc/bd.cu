// BD::EulerMaruyama position update for sm_100a (BASELINE config 0: ideal Brownian particles, the README example).
// Replaces EulerMaruyama_ns::integrateGPU (Integrator/BrownianDynamics.cu:117-145): one pass over pos (+ force),
// noise from Saru(particle index, step, seed) drawn in registers - HBM bound, 2 x sizeof(real4) bytes per particle
// (+ sizeof(real4) when forces act).
#include "common.cuh"
#include "saru.cuh"

namespace ub200 {

template <class T> struct Shear3 { T k[9]; };

// Expression shapes follow the reference line by line (R += dt*(KR + M*F); R += dW) so that nvcc's FMA
// contraction makes the same rounding decisions: the result is BIT-IDENTICAL to the reference's (tests/test_bd_gpu.py).
template <class T4>
__global__ void __launch_bounds__(128)
bdEulerMaruyama(T4 *__restrict__ pos, const int *__restrict__ groupIdx, const T4 *__restrict__ force, Shear3<decltype(T4::x)> K,
                decltype(T4::x) selfMobility, const decltype(T4::x) *__restrict__ radius, decltype(T4::x) dt, int is2D,
                decltype(T4::x) temperature, int N, uint32_t stepNum, uint32_t seed) {
  using T = decltype(T4::x);
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (uint32_t)N) return;
  const int i = groupIdx ? groupIdx[id] : (int)id;
  const T4 p = pos[i];
  T Rx = p.x, Ry = p.y, Rz = p.z;
  T Fx = T(0), Fy = T(0), Fz = T(0);
  if (force) { const T4 f = force[i]; Fx = f.x; Fy = f.y; Fz = f.z; }
  const T KRx = K.k[0] * Rx + K.k[1] * Ry + K.k[2] * Rz;
  const T KRy = K.k[3] * Rx + K.k[4] * Ry + K.k[5] * Rz;
  const T KRz = K.k[6] * Rx + K.k[7] * Ry + K.k[8] * Rz;
  const T M = selfMobility * (radius ? (T(1.0) / radius[i]) : T(1.0));
  Rx += dt * (KRx + M * Fx);
  Ry += dt * (KRy + M * Fy);
  Rz += dt * (KRz + M * Fz);
  if (temperature > T(0)) {
    Saru rng((uint32_t)i, stepNum, seed);
    const T B = sqrt(T(2.0) * temperature * M * dt);
    const float2 g0 = rng.gauss2((float)B);
    const float2 g1 = rng.gauss2((float)B);
    Rx = Rx + (T)g0.x; Ry = Ry + (T)g0.y; Rz = Rz + (T)g1.x; // R += dW: a plain add of the already rounded noise
  }
  T4 out = p;
  out.x = Rx; out.y = Ry;
  if (!is2D) out.z = Rz;
  pos[i] = out;
}

} // namespace ub200

using namespace ub200;

extern "C" int ub200_bd_euler_maruyama_step(int precisionBytes, void *d_pos, const int *d_groupIdx, const void *d_force,
                                            const double *K9, double selfMobility, const void *d_radius, double dt,
                                            int is2D, double temperature, int N, uint32_t stepNum, uint32_t seed,
                                            void *stream) {
  if (!d_pos || N <= 0 || (precisionBytes != 4 && precisionBytes != 8)) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 127) / 128;
  if (precisionBytes == 4) {
    Shear3<float> K;
    for (int q = 0; q < 9; q++) K.k[q] = K9 ? (float)K9[q] : 0.f;
    bdEulerMaruyama<float4><<<nb, 128, 0, st>>>((float4 *)d_pos, d_groupIdx, (const float4 *)d_force, K, (float)selfMobility,
                                                (const float *)d_radius, (float)dt, is2D, (float)temperature, N, stepNum, seed);
  } else {
    Shear3<double> K;
    for (int q = 0; q < 9; q++) K.k[q] = K9 ? K9[q] : 0.0;
    bdEulerMaruyama<double4><<<nb, 128, 0, st>>>((double4 *)d_pos, d_groupIdx, (const double4 *)d_force, K, selfMobility,
                                                 (const double *)d_radius, dt, is2D, temperature, N, stepNum, seed);
  }
  UB200_LAUNCHED();
  return UB200_OK;
}
