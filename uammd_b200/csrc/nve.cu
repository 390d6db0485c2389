// Velocity Verlet (NVE) half steps and the fused LJ molecular dynamics engine, sm_100a.
// Replaces VerletNVE_ns::integrateGPU<step> / VerletNVE::forwardTime (Integrator/VerletNVE.cu:64-85,174-188).
#include "lj_engine.cuh"

namespace ub200 {

// v += (F/m) dt/2 ; step 1 also x += v dt. Same operation order as the reference (force/m is (1/m)*force,
// utils/vector.cuh:191-193; the trailing multiply-adds are contracted by nvcc there, spelled out here).
template <int STEP>
__global__ void __launch_bounds__(256)
nveHalfStep(float4 *__restrict__ pos, float *__restrict__ vel, const float4 *__restrict__ force,
            const float *__restrict__ mass, float defaultMass, const int *__restrict__ groupIdx, int N, float dt,
            int is2D) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= N) return;
  const int i = groupIdx ? groupIdx[id] : id;
  const float invm = 1.0f / (mass ? mass[i] : defaultMass);
  const float4 f = force[i];
  float vx = vel[3 * (size_t)i + 0], vy = vel[3 * (size_t)i + 1], vz = vel[3 * (size_t)i + 2];
  vx = __fmaf_rn(__fmul_rn(__fmul_rn(invm, f.x), dt), 0.5f, vx);
  vy = __fmaf_rn(__fmul_rn(__fmul_rn(invm, f.y), dt), 0.5f, vy);
  vz = __fmaf_rn(__fmul_rn(__fmul_rn(invm, f.z), dt), 0.5f, vz);
  if (is2D) vz = 0.0f;
  vel[3 * (size_t)i + 0] = vx;
  vel[3 * (size_t)i + 1] = vy;
  vel[3 * (size_t)i + 2] = vz;
  if (STEP == 1) {
    float4 p = pos[i];
    p.x = __fmaf_rn(vx, dt, p.x);
    p.y = __fmaf_rn(vy, dt, p.y);
    p.z = __fmaf_rn(vz, dt, p.z);
    pos[i] = p;
  }
}

// second kick of step n fused with the first kick + drift of step n+1 (same roundings as two separate passes)
__global__ void __launch_bounds__(256)
nveKickKickDrift(float4 *__restrict__ pos, float *__restrict__ vel, const float4 *__restrict__ force, int N,
                 float dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 f = force[i];
  float vx = vel[3 * (size_t)i + 0], vy = vel[3 * (size_t)i + 1], vz = vel[3 * (size_t)i + 2];
  const float hx = __fmul_rn(f.x, dt), hy = __fmul_rn(f.y, dt), hz = __fmul_rn(f.z, dt);
  vx = __fmaf_rn(hx, 0.5f, vx); vy = __fmaf_rn(hy, 0.5f, vy); vz = __fmaf_rn(hz, 0.5f, vz);
  vx = __fmaf_rn(hx, 0.5f, vx); vy = __fmaf_rn(hy, 0.5f, vy); vz = __fmaf_rn(hz, 0.5f, vz);
  vel[3 * (size_t)i + 0] = vx;
  vel[3 * (size_t)i + 1] = vy;
  vel[3 * (size_t)i + 2] = vz;
  float4 p = pos[i];
  p.x = __fmaf_rn(vx, dt, p.x);
  p.y = __fmaf_rn(vy, dt, p.y);
  p.z = __fmaf_rn(vz, dt, p.z);
  pos[i] = p;
}

// sum of m v^2 / 2 (unit mass) in double: per-block partial sums, then one block adds them up (deterministic order)
__global__ void __launch_bounds__(256) kineticPartial(const float *__restrict__ vel, int N, double *__restrict__ partial) {
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float x = vel[3 * (size_t)i], y = vel[3 * (size_t)i + 1], z = vel[3 * (size_t)i + 2];
    s += (double)x * x + (double)y * y + (double)z * z;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) kineticFinal(const double *__restrict__ partial, int n, double *__restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = 0.5 * sh[0];
}

} // namespace ub200

using namespace ub200;

struct ub200_md {
  ub200_ljengine *eng = nullptr; // PairForces<LJ, CellList>::sum: private half-cell list + column traversal (lj_column.cu)
  DevBuf dpos, dvel, dforce; // device state for the host-buffer entry point
  DevBuf keScratch;                  // kinetic energy: per-block partial sums + the result
  cudaStream_t copyStream = nullptr; // host-buffer entry point: transfers overlapped with the force evaluations
  cudaEvent_t evPosUp = nullptr, evVelUp = nullptr, evDrift = nullptr, evPosDown = nullptr;
};

extern "C" {

int ub200_nve_half_step_f32(void *d_pos, void *d_vel, const void *d_force, const float *d_mass, float defaultMass,
                            const int *d_groupIdx, int N, float dt, int is2D, int step, void *stream) {
  if (!d_pos || !d_vel || !d_force || N <= 0 || (step != 1 && step != 2)) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 255) / 256;
  if (step == 1)
    nveHalfStep<1><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, d_mass, defaultMass,
                                       d_groupIdx, N, dt, is2D);
  else
    nveHalfStep<2><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, d_mass, defaultMass,
                                       d_groupIdx, N, dt, is2D);
  UB200_LAUNCHED();
  return UB200_OK;
}

int ub200_nve_kick_kick_drift_f32(void *d_pos, void *d_vel, const void *d_force, int N, float dt, void *stream) {
  if (!d_pos || !d_vel || !d_force || N <= 0) return UB200_ERR_INVALID_ARGUMENT;
  nveKickKickDrift<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float4 *)d_pos, (float *)d_vel,
                                                                    (const float4 *)d_force, N, dt);
  UB200_LAUNCHED();
  return UB200_OK;
}

int ub200_md_create(ub200_md **out) {
  if (!out) return UB200_ERR_INVALID_ARGUMENT;
  ub200_md *md = new (std::nothrow) ub200_md();
  if (!md) return UB200_ERR_ALLOC;
  int rc = ub200_ljengine_create(&md->eng);
  if (rc) { delete md; return rc; }
  *out = md;
  return UB200_OK;
}

int ub200_md_destroy(ub200_md *md) {
  if (!md) return UB200_OK;
  ub200_ljengine_destroy(md->eng);
  md->dpos.release(); md->dvel.release(); md->dforce.release(); md->keScratch.release();
  if (md->copyStream) {
    cudaStreamDestroy(md->copyStream);
    cudaEventDestroy(md->evPosUp); cudaEventDestroy(md->evVelUp); cudaEventDestroy(md->evDrift); cudaEventDestroy(md->evPosDown);
  }
  delete md;
  return UB200_OK;
}

ub200_ljengine *ub200_md_engine(ub200_md *md) { return md ? md->eng : nullptr; }

int ub200_md_kinetic_energy_f32(ub200_md *md, const void *d_vel, int N, double *h_out, void *stream) {
  if (!md || !d_vel || N <= 0 || !h_out) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = 2 * kNumSMs;
  if (const int e = md->keScratch.reserve(sizeof(double) * (nb + 1))) return e;
  double *part = md->keScratch.as<double>();
  kineticPartial<<<nb, 256, 0, st>>>((const float *)d_vel, N, part + 1);
  UB200_LAUNCHED();
  kineticFinal<<<1, 256, 0, st>>>(part + 1, nb, part);
  UB200_LAUNCHED();
  UB200_CUDA(cudaMemcpyAsync(h_out, part, sizeof(double), cudaMemcpyDeviceToHost, st));
  return UB200_OK;
}

static int mdForces(ub200_md *md, void *d_pos, void *d_force, int N, const float L[3], float rc, const float *params,
                    int ntypes, cudaStream_t st) {
  const int periodic[3] = {1, 1, 1};
  (void)rc; // the neighbour search radius is the largest pair cut-off of the table, as Radial::getCutOff returns it
  // sole interactor: forces are written, not accumulated (replaces resetForces + sum)
  return ljEngineSum(md->eng, (const float4 *)d_pos, nullptr, N, L, periodic, params, ntypes, (float4 *)d_force, nullptr,
                     nullptr, nullptr, false, 0, 0x7fffffff, st);
}

int ub200_md_lj_nve_prepare_f32(ub200_md *md, void *d_pos, void *d_force, int N, const float L[3], float rc,
                                const float *params, int ntypes, void *stream) {
  if (!md || !d_pos || !d_force || N <= 0 || !L || !params) return UB200_ERR_INVALID_ARGUMENT;
  return mdForces(md, d_pos, d_force, N, L, rc, params, ntypes, (cudaStream_t)stream);
}

int ub200_md_lj_nve_run_f32(ub200_md *md, void *d_pos, void *d_vel, void *d_force, int N, const float L[3], float rc,
                            const float *params, int ntypes, float dt, int nsteps, void *stream) {
  if (!md || !d_pos || !d_vel || !d_force || N <= 0 || !L || !params || nsteps < 0) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 255) / 256;
  for (int s = 0; s < nsteps; s++) {
    int e;
    if (s == 0) {
      nveHalfStep<1><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, nullptr, 1.0f,
                                         nullptr, N, dt, 0);
      UB200_LAUNCHED();
    }
    if ((e = mdForces(md, d_pos, d_force, N, L, rc, params, ntypes, st))) return e;
    if (s == nsteps - 1) {
      nveHalfStep<2><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, nullptr, 1.0f,
                                         nullptr, N, dt, 0);
    } else {
      nveKickKickDrift<<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, N, dt);
    }
    UB200_LAUNCHED();
  }
  return UB200_OK;
}

// Same loop over PairForces<LJ, VerletList>: drift check (host-synchronous, like the reference) + sortPos refresh every
// step, list rebuild only when a particle left its skin.
int ub200_md_lj_nve_verlet_run_f32(ub200_md *md, ub200_verletlist *vl, void *d_pos, void *d_vel, void *d_force, int N,
                                   const float L[3], float rc, const float *params, int ntypes, float dt, int nsteps,
                                   int forcesAreCurrent, void *stream) {
  if (!md || !vl || !d_pos || !d_vel || !d_force || N <= 0 || !L || !params || nsteps < 0) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (N + 255) / 256;
  const int periodic[3] = {1, 1, 1};
  auto forces = [&]() -> int {
    int e = ub200_verletlist_update_f32(vl, d_pos, nullptr, N, L, periodic, rc, 0, nullptr, stream);
    if (e) return e;
    // sole interactor: the forces are written, which replaces VerletNVE::resetForces (VerletNVE.cu:152-158) + sum
    return ljVerletSum(vl, params, ntypes, (float4 *)d_force, nullptr, nullptr, nullptr, false, st);
  };
  int e;
  if (!forcesAreCurrent && (e = forces())) return e;
  for (int s = 0; s < nsteps; s++) {
    if (s == 0) {
      nveHalfStep<1><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, nullptr, 1.0f, nullptr, N, dt, 0);
      UB200_LAUNCHED();
    }
    if ((e = forces())) return e;
    if (s == nsteps - 1)
      nveHalfStep<2><<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, nullptr, 1.0f, nullptr, N, dt, 0);
    else
      nveKickKickDrift<<<nb, 256, 0, st>>>((float4 *)d_pos, (float *)d_vel, (const float4 *)d_force, N, dt);
    UB200_LAUNCHED();
  }
  return UB200_OK;
}

int ub200_md_lj_nve_run_host_f32(ub200_md *md, float *h_pos4, float *h_vel3, float *h_force4, int N, const float L[3],
                                 float rc, const float *params, int ntypes, float dt, int nsteps, void *stream) {
  if (!md || !h_pos4 || !h_vel3 || N <= 0 || nsteps < 1) return UB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int e;
  if ((e = md->dpos.reserve(sizeof(float4) * (size_t)N))) return e;
  if ((e = md->dvel.reserve(sizeof(float) * 3 * (size_t)N))) return e;
  if ((e = md->dforce.reserve(sizeof(float4) * (size_t)N))) return e;
  if (!md->copyStream) {
    UB200_CUDA(cudaStreamCreateWithFlags(&md->copyStream, cudaStreamNonBlocking));
    UB200_CUDA(cudaEventCreateWithFlags(&md->evPosUp, cudaEventDisableTiming));
    UB200_CUDA(cudaEventCreateWithFlags(&md->evVelUp, cudaEventDisableTiming));
    UB200_CUDA(cudaEventCreateWithFlags(&md->evDrift, cudaEventDisableTiming));
    UB200_CUDA(cudaEventCreateWithFlags(&md->evPosDown, cudaEventDisableTiming));
  }
  cudaStream_t cs = md->copyStream;
  const int nb = (N + 255) / 256;
  // positions up on the compute stream; velocities up on the copy stream while F(t) is evaluated (it needs positions only)
  UB200_CUDA(cudaMemcpyAsync(md->dpos.p, h_pos4, sizeof(float4) * (size_t)N, cudaMemcpyHostToDevice, st));
  UB200_CUDA(cudaEventRecord(md->evPosUp, st));
  UB200_CUDA(cudaStreamWaitEvent(cs, md->evPosUp, 0)); // also orders the copy stream after the caller's earlier work
  UB200_CUDA(cudaMemcpyAsync(md->dvel.p, h_vel3, sizeof(float) * 3 * (size_t)N, cudaMemcpyHostToDevice, cs));
  UB200_CUDA(cudaEventRecord(md->evVelUp, cs));
  if ((e = ub200_md_lj_nve_prepare_f32(md, md->dpos.p, md->dforce.p, N, L, rc, params, ntypes, stream))) return e;
  UB200_CUDA(cudaStreamWaitEvent(st, md->evVelUp, 0));
  if (nsteps > 1 && (e = ub200_md_lj_nve_run_f32(md, md->dpos.p, md->dvel.p, md->dforce.p, N, L, rc, params, ntypes, dt, nsteps - 1, stream)))
    return e;
  // last step by hand: after its drift the positions are final and go down while F(t+dt) and the last kick run
  nveHalfStep<1><<<nb, 256, 0, st>>>(md->dpos.as<float4>(), md->dvel.as<float>(), md->dforce.as<float4>(), nullptr, 1.0f, nullptr, N, dt, 0);
  UB200_LAUNCHED();
  UB200_CUDA(cudaEventRecord(md->evDrift, st));
  UB200_CUDA(cudaStreamWaitEvent(cs, md->evDrift, 0));
  UB200_CUDA(cudaMemcpyAsync(h_pos4, md->dpos.p, sizeof(float4) * (size_t)N, cudaMemcpyDeviceToHost, cs));
  UB200_CUDA(cudaEventRecord(md->evPosDown, cs));
  if ((e = mdForces(md, md->dpos.p, md->dforce.p, N, L, rc, params, ntypes, st))) return e;
  nveHalfStep<2><<<nb, 256, 0, st>>>(md->dpos.as<float4>(), md->dvel.as<float>(), md->dforce.as<float4>(), nullptr, 1.0f, nullptr, N, dt, 0);
  UB200_LAUNCHED();
  UB200_CUDA(cudaMemcpyAsync(h_vel3, md->dvel.p, sizeof(float) * 3 * (size_t)N, cudaMemcpyDeviceToHost, st));
  if (h_force4)
    UB200_CUDA(cudaMemcpyAsync(h_force4, md->dforce.p, sizeof(float4) * (size_t)N, cudaMemcpyDeviceToHost, st));
  UB200_CUDA(cudaStreamWaitEvent(st, md->evPosDown, 0));
  UB200_CUDA(cudaStreamSynchronize(st));
  return UB200_OK;
}
}
