// Langevin (NVT) velocity Verlet of Gronbech-Jensen & Farago, Mol. Phys. 111 (2013) 983, sm_100a (single precision).
// Replaces VerletNVT::GronbechJensen_ns::integrateGPU<step> (Integrator/VerletNVT/GronbechJensen.cu:30-66) - the
// integrator that generic_md and examples/misc/benchmark.cu drive - and
// VerletNVT::Basic_ns::initialVelocities (Integrator/VerletNVT/Basic.cu:12-29).
//   step 1: x += b dt v + b dt/(2m) (dt f + beta),  v = a v + a dt/(2m) f + b/m beta,  f = 0
//   step 2: v += dt/(2m) f
// with b = 1/(1 + friction dt/2), a = (1 - friction dt/2) b and beta ~ N(0, 2 kT m friction dt) from
// Saru(index in group, step, seed). One pass over pos (RW 32 B), vel (RW 24 B) and force (R 16 B + W 16 B in step 1):
// HBM bound, 88 B / 40 B per particle.
// The roundings (which products are fused, the order of the scalar prefactors) are spelled out exactly as nvcc contracts
// the reference kernel for sm_100a (read from its PTX), so that positions and velocities match the reference bit for bit.
#include "common.cuh"
#include "saru.cuh"
#include <cfloat>

namespace ub200 {

template <int STEP>
__global__ void __launch_bounds__(128)
gjIntegrate(float4 *__restrict__ pos, float *__restrict__ vel, float4 *__restrict__ force, const float *__restrict__ mass,
            float defaultMass, const int *__restrict__ groupIdx, int N, float dt, float friction, int is2D,
            float noiseAmplitude, uint32_t stepNum, uint32_t seed) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  const int i = groupIdx ? groupIdx[id] : id;
  const float invMass = __frcp_rn(defaultMass > 0.0f ? defaultMass : mass[i]);
  float *v = vel + 3 * (size_t)i;
  if (STEP == 1) {
    Saru rng((uint32_t)id, stepNum, seed);
    const float amp = __fmul_rn(noiseAmplitude, rsqrtf(invMass)); // sqrt(2 kT m friction dt)
    const float2 n01 = rng.gauss2(amp);
    const float nx = n01.x, ny = n01.y;
    const float nz = is2D ? 0.0f : rng.gauss2(amp).x;
    const float g = __fmul_rn(__fmul_rn(dt, friction), 0.5f);
    const float b = __frcp_rn(__fadd_rn(g, 1.0f));
    const float a = __fmul_rn(__fsub_rn(1.0f, g), b);
    const float4 p = pos[i];
    const float4 f = force[i];
    const float vx = v[0], vy = v[1], vz = v[2];
    const float bdt = __fmul_rn(dt, b);
    const float c2 = __fmul_rn(b, __fmul_rn(dt, __fmul_rn(invMass, 0.5f)));
    float4 q = p;
    q.x = __fmaf_rn(c2, __fmaf_rn(dt, f.x, nx), __fmaf_rn(bdt, vx, p.x));
    q.y = __fmaf_rn(c2, __fmaf_rn(dt, f.y, ny), __fmaf_rn(bdt, vy, p.y));
    q.z = __fmaf_rn(c2, __fmaf_rn(dt, f.z, nz), __fmaf_rn(bdt, vz, p.z));
    pos[i] = q;
    const float c3 = __fmul_rn(a, __fmul_rn(__fmul_rn(dt, 0.5f), invMass));
    const float c4 = __fmul_rn(b, invMass);
    v[0] = __fmaf_rn(c4, nx, __fmaf_rn(a, vx, __fmul_rn(c3, f.x)));
    v[1] = __fmaf_rn(c4, ny, __fmaf_rn(a, vy, __fmul_rn(c3, f.y)));
    v[2] = is2D ? 0.0f : __fmaf_rn(c4, nz, __fmaf_rn(a, vz, __fmul_rn(c3, f.z)));
    force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const float c = __fmul_rn(__fmul_rn(dt, 0.5f), invMass);
    const float4 f = force[i];
    v[0] = __fmaf_rn(f.x, c, v[0]);
    v[1] = __fmaf_rn(f.y, c, v[1]);
    v[2] = is2D ? 0.0f : __fmaf_rn(c, f.z, v[2]);
  }
}

// VerletNVT::Basic_ns::integrateGPU<step> (Integrator/VerletNVT/Basic.cu:87-117): the plain Langevin velocity Verlet,
//   both half steps: v += (f/m - friction v) dt/2 + noise,  noise ~ N(0, noiseAmplitude^2 / (2 m)) from
//   Saru(index in group + N (step - 1), stepNum, seed) - a fresh draw in EACH half step;  step 1 then: x += v dt, f = 0.
// Roundings as nvcc contracts the reference kernel for sm_100a (read from its PTX): two products and a difference, one
// fma with the noise as addend, an addition; the drift is one fma.
template <int STEP>
__global__ void __launch_bounds__(128)
basicIntegrate(float4 *__restrict__ pos, float *__restrict__ vel, float4 *__restrict__ force, const float *__restrict__ mass,
               float defaultMass, const int *__restrict__ groupIdx, int N, float dt, float friction, int is2D,
               float noiseAmplitude, uint32_t stepNum, uint32_t seed) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  const int i = groupIdx ? groupIdx[id] : id;
  const float invMass = __frcp_rn(defaultMass > 0.0f ? defaultMass : mass[i]);
  Saru rng((uint32_t)(id + N * (STEP - 1)), stepNum, seed);
  const float amp = __fmul_rn(noiseAmplitude, __fsqrt_rn(__fmul_rn(invMass, 0.5f)));
  const float2 n01 = rng.gauss2(amp);
  const float nz = rng.gauss2(amp).x;
  float *v = vel + 3 * (size_t)i;
  const float4 f = force[i];
  const float hdt = __fmul_rn(dt, 0.5f);
  const float v0 = v[0], v1 = v[1], v2 = v[2];
  const float vx = __fadd_rn(v0, __fmaf_rn(hdt, __fsub_rn(__fmul_rn(invMass, f.x), __fmul_rn(friction, v0)), n01.x));
  const float vy = __fadd_rn(v1, __fmaf_rn(hdt, __fsub_rn(__fmul_rn(invMass, f.y), __fmul_rn(friction, v1)), n01.y));
  float vz = __fadd_rn(__fmaf_rn(hdt, __fsub_rn(__fmul_rn(invMass, f.z), __fmul_rn(friction, v2)), nz), v2);
  if (is2D) vz = 0.0f;
  v[0] = vx; v[1] = vy; v[2] = vz;
  if (STEP == 1) {
    float4 p = pos[i];
    p.x = __fmaf_rn(dt, vx, p.x);
    p.y = __fmaf_rn(dt, vy, p.y);
    p.z = __fmaf_rn(dt, vz, p.z);
    pos[i] = p;
    force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Basic_ns::initialVelocities (Basic.cu:12-29): Saru(id, seed) two-seed constructor (saruprng.cuh:236-251), gd() =
// Box-Muller on float uniforms with float log/sin/cos/sqrt and a double product (saruprng.cuh:130-143). The reference
// ignores the mass here (mass_i = 1) and indexes the group twice (vel[index[index[id]]]); both are kept.
__global__ void __launch_bounds__(128)
nvtInitialVelocities(float *__restrict__ vel, const int *__restrict__ groupIdx, float vamp, int is2D, int N, uint32_t seed) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  uint32_t s1 = (uint32_t)id, s2 = seed;
  s2 += s1 << 16;
  s1 += s2 << 11;
  s2 += (uint32_t)(((int32_t)s1) >> 7);
  s1 ^= (uint32_t)(((int32_t)s2) >> 3);
  s2 *= 0xA5366B4Du;
  s2 ^= s2 >> 10;
  s2 ^= (uint32_t)(((int32_t)s2) >> 19);
  s1 += s2 ^ 0x6d2d4e11u;
  Saru rng(0u, 0u, 0u);
  rng.lcg = 0x79dedea3u * (s1 ^ (uint32_t)(((int32_t)s1) >> 14));
  rng.weyl = (rng.lcg + s2) ^ (uint32_t)(((int32_t)rng.lcg) >> 8);
  rng.lcg = rng.lcg + (rng.weyl * (rng.weyl ^ 0xdddf97f5u));
  rng.weyl = 0xABCB96F7u + (rng.weyl >> 1);
  const double std = (double)vamp;
  double g[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int k = 0; k < (is2D ? 1 : 2); k++) {
    double u0;
    do { u0 = (double)rng.f(); } while (u0 <= DBL_MIN);
    const double u1 = (double)rng.f();
    const double r = (double)sqrtf(-2.0f * logf((float)u0));
    const float theta = (float)(6.283185307179586 * u1);
    g[k][0] = __dmul_rn(__dmul_rn(r, (double)sinf(theta)), std);
    g[k][1] = __dmul_rn(__dmul_rn(r, (double)cosf(theta)), std);
  }
  const int i = groupIdx ? groupIdx[id] : id;
  const int index = groupIdx ? groupIdx[i] : i;
  vel[3 * (size_t)index + 0] = (float)g[0][0];
  vel[3 * (size_t)index + 1] = (float)g[0][1];
  vel[3 * (size_t)index + 2] = (float)g[1][0];
}

} // namespace ub200

using namespace ub200;

extern "C" int ub200_nvt_gj_half_step_f32(void *d_pos, void *d_vel, void *d_force, const float *d_mass, float defaultMass,
                                          const int *d_groupIdx, int N, float dt, float friction, int is2D,
                                          float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step, void *stream) {
  if (!d_pos || !d_vel || !d_force || N <= 0 || (step != 1 && step != 2)) return UB200_ERR_INVALID_ARGUMENT;
  if (!(defaultMass > 0.0f) && !d_mass) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 127) / 128;
  if (step == 1)
    gjIntegrate<1><<<nb, 128, 0, st>>>((float4 *)d_pos, (float *)d_vel, (float4 *)d_force, d_mass, defaultMass, d_groupIdx, N,
                                       dt, friction, is2D, noiseAmplitude, stepNum, seed);
  else
    gjIntegrate<2><<<nb, 128, 0, st>>>((float4 *)d_pos, (float *)d_vel, (float4 *)d_force, d_mass, defaultMass, d_groupIdx, N,
                                       dt, friction, is2D, noiseAmplitude, stepNum, seed);
  UB200_LAUNCHED();
  return UB200_OK;
}

extern "C" int ub200_nvt_basic_half_step_f32(void *d_pos, void *d_vel, void *d_force, const float *d_mass, float defaultMass,
                                             const int *d_groupIdx, int N, float dt, float friction, int is2D,
                                             float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step, void *stream) {
  if (!d_pos || !d_vel || !d_force || N <= 0 || (step != 1 && step != 2)) return UB200_ERR_INVALID_ARGUMENT;
  if (!(defaultMass > 0.0f) && !d_mass) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 127) / 128;
  if (step == 1)
    basicIntegrate<1><<<nb, 128, 0, st>>>((float4 *)d_pos, (float *)d_vel, (float4 *)d_force, d_mass, defaultMass, d_groupIdx,
                                          N, dt, friction, is2D, noiseAmplitude, stepNum, seed);
  else
    basicIntegrate<2><<<nb, 128, 0, st>>>((float4 *)d_pos, (float *)d_vel, (float4 *)d_force, d_mass, defaultMass, d_groupIdx,
                                          N, dt, friction, is2D, noiseAmplitude, stepNum, seed);
  UB200_LAUNCHED();
  return UB200_OK;
}

extern "C" int ub200_nvt_initial_velocities_f32(void *d_vel, const int *d_groupIdx, int N, float velAmplitude, int is2D,
                                                uint32_t seed, void *stream) {
  if (!d_vel || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  nvtInitialVelocities<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>((float *)d_vel, d_groupIdx, velAmplitude, is2D, N, seed);
  UB200_LAUNCHED();
  return UB200_OK;
}
