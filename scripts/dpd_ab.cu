// A/B of the tile variant of the DPD traversal (UB200_DPD_TILE=1) against the default per-cell kernel through the C ABI:
// same forces bit for bit? how fast? Plain CUDA host program (no Python); built by `make -C uammd_b200/csrc tools`.
// usage: dpd_ab [N] (default 4 000 000, rho = 3, rc = 1: BASELINE config 4 shape)
#include "uammd_b200.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 4000000;
  const float L = (float)cbrt(N / 3.0);
  std::vector<float> pos(4 * (size_t)N), vel(3 * (size_t)N);
  std::mt19937 gen(21);
  std::uniform_real_distribution<float> U(-0.5f, 0.5f);
  std::normal_distribution<float> G(0.f, 1.f);
  for (size_t i = 0; i < (size_t)N; i++) {
    for (int d = 0; d < 3; d++) { pos[4 * i + d] = U(gen) * L; vel[3 * i + d] = G(gen); }
    pos[4 * i + 3] = 0.f;
  }
  float *d_pos, *d_vel, *d_f[2];
  cudaMalloc(&d_pos, 16 * (size_t)N); cudaMalloc(&d_vel, 12 * (size_t)N);
  cudaMalloc(&d_f[0], 16 * (size_t)N); cudaMalloc(&d_f[1], 16 * (size_t)N);
  cudaMemcpy(d_pos, pos.data(), 16 * (size_t)N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_vel, vel.data(), 12 * (size_t)N, cudaMemcpyHostToDevice);
  const float Lv[3] = {L, L, L};
  const int per[3] = {1, 1, 1};
  int cd[3];
  ub200_neighbour_celldim_f32(Lv, 1.0f, cd);
  ub200_celllist *cl;
  ub200_celllist_create(&cl);
  int rc = ub200_celllist_build_f32(cl, d_pos, nullptr, N, Lv, per, cd, nullptr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2] = {0, 0};
  const float sigma = sqrtf(2.0f * 1.0f) / sqrtf(0.01f);
  for (int v = 0; v < 2; v++) {
    setenv("UB200_DPD_TILE", v ? "1" : "0", 1);
    for (int it = 0; it < 3; it++) { // 2 warm-ups, the third pass is timed
      cudaMemset(d_f[v], 0, 16 * (size_t)N);
      cudaEventRecord(e0);
      rc |= ub200_dpd_sum_f32(cl, d_vel, 25.0f, 4.5f, sigma, 1.0f, 99u, 7u, N, d_f[v], nullptr, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
    }
  }
  std::vector<float> f0(4 * (size_t)N), f1(4 * (size_t)N);
  cudaMemcpy(f0.data(), d_f[0], 16 * (size_t)N, cudaMemcpyDeviceToHost);
  cudaMemcpy(f1.data(), d_f[1], 16 * (size_t)N, cudaMemcpyDeviceToHost);
  size_t diff = 0;
  double fmax = 0, dmax = 0;
  for (size_t i = 0; i < f0.size(); i++) {
    diff += memcmp(&f0[i], &f1[i], 4) != 0;
    fmax = std::max(fmax, (double)std::fabs(f0[i]));
    dmax = std::max(dmax, (double)std::fabs(f0[i] - f1[i]));
  }
  printf("{\"N\":%d,\"cells\":[%d,%d,%d],\"rc\":%d,\"cuda\":\"%s\",\"ms_default\":%.4f,\"ms_tile\":%.4f,\"differing_words\":%zu,"
         "\"max_abs_diff\":%.3g,\"fmax\":%.4g}\n",
         N, cd[0], cd[1], cd[2], rc, cudaGetErrorString(cudaGetLastError()), ms[0], ms[1], diff, dmax, fmax);
  return 0;
}
