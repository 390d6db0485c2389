// Head-to-head of the hand-written 3-D real FFT (ub200_fft3d_*) with cuFFT on the plans the reference makes
// (FCM_impl.cuh:179-211, PSE/FarField.cuh:555-603: three interleaved batches, stride 3, in place, padded rows):
// forward + inverse of a real3 grid, L2 flushed between repetitions, device-timed. Plain CUDA host program:
//   nvcc -O2 scripts/fft_vs_cufft.cu -Iinclude -Luammd_b200 -luammd_b200 -lcufft -o scripts/_bin/fft_vs_cufft
#include "uammd_b200.h"
#include <cuda_runtime.h>
#include <cufft.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <class Fn> static float timed(Fn fn, int reps, char *scrub) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<float> ts;
  for (int r = 0; r < reps + 2; r++) {
    cudaMemsetAsync(scrub, r, 256u << 20, 0);
    cudaEventRecord(a, 0);
    fn();
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (r >= 2) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

static void run(int n, bool dbl, int reps, char *scrub) {
  const int nxPad = 2 * (n / 2 + 1);
  const size_t elems = (size_t)n * n * nxPad * 3, bytes = elems * (dbl ? 8 : 4);
  void *grid;
  cudaMalloc(&grid, bytes);
  cudaMemset(grid, 0, bytes);
  // ---- cuFFT, the reference's plans
  cufftHandle fwd, inv;
  cufftCreate(&fwd); cufftCreate(&inv);
  int dims[3] = {n, n, n}, inembed[3] = {n, n, nxPad}, oembed[3] = {n, n, n / 2 + 1};
  size_t ws1 = 0, ws2 = 0;
  cufftMakePlanMany(fwd, 3, dims, inembed, 3, 1, oembed, 3, 1, dbl ? CUFFT_D2Z : CUFFT_R2C, 3, &ws1);
  cufftMakePlanMany(inv, 3, dims, oembed, 3, 1, inembed, 3, 1, dbl ? CUFFT_Z2D : CUFFT_C2R, 3, &ws2);
  auto cufwd = [&]() {
    if (dbl) cufftExecD2Z(fwd, (cufftDoubleReal *)grid, (cufftDoubleComplex *)grid);
    else cufftExecR2C(fwd, (cufftReal *)grid, (cufftComplex *)grid);
  };
  auto cuinv = [&]() {
    if (dbl) cufftExecZ2D(inv, (cufftDoubleComplex *)grid, (cufftDoubleReal *)grid);
    else cufftExecC2R(inv, (cufftComplex *)grid, (cufftReal *)grid);
  };
  const float cf = timed(cufwd, reps, scrub), ci = timed(cuinv, reps, scrub);
  // ---- ours
  ub200_fft3d *plan = nullptr;
  const int rc = ub200_fft3d_create(&plan, dbl ? 8 : 4, n, n, n);
  float of = -1, oi = -1;
  if (!rc) {
    of = timed([&]() { ub200_fft3d_exec(plan, grid, -1, nullptr); }, reps, scrub);
    oi = timed([&]() { ub200_fft3d_exec(plan, grid, +1, nullptr); }, reps, scrub);
    ub200_fft3d_destroy(plan);
  }
  const double gb = 2.0 * bytes / 1e9; // one read + one write of the grid: the floor of an out-of-cache transform pass
  printf("{\"n\":%d,\"precision\":\"%s\",\"grid_MB\":%.1f,\"cufft_fwd_ms\":%.4f,\"cufft_inv_ms\":%.4f,\"ours_fwd_ms\":%.4f,"
         "\"ours_inv_ms\":%.4f,\"cufft_workspace_MB\":%.1f,\"one_pass_floor_ms_at_6.5TBs\":%.4f}\n",
         n, dbl ? "f64" : "f32", bytes / 1e6, cf, ci, of, oi, std::max(ws1, ws2) / 1e6, gb / 6545.3 * 1e3);
  cufftDestroy(fwd); cufftDestroy(inv);
  cudaFree(grid);
}

int main(int argc, char **argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 20;
  char *scrub;
  cudaMalloc(&scrub, 256u << 20);
  run(128, true, reps, scrub);   // BASELINE config 2: FCM 128^3 fp64
  run(128, false, reps, scrub);
  run(256, false, reps, scrub);  // BASELINE config 3: PSE 256^3 fp32
  run(256, true, reps, scrub);
  run(250, true, reps, scrub);   // 2 5^3: the grid the Poisson parameter resolution picks for 2e5 charges (generic radix path)
  run(216, true, reps, scrub);   // 2^3 3^3
  run(154, true, reps, scrub);   // 2 7 11
  return 0;
}
