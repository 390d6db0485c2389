// A/B of the packed-fp32 (FFMA2) LJ traversal against the default one through the C ABI: same forces bit for bit? how fast?
// Plain CUDA host program (no Python): nvcc -O2 scripts/lj_ab.cu -Iinclude -Luammd_b200 -luammd_b200 -o scripts/_bin/lj_ab
#include "uammd_b200.h"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 63; // FCC cells per side: N = 4 n^3 (63 -> 1 000 188)
  const int N = 4 * n * n * n;
  const float L = (float)cbrt(N / 0.8);
  std::vector<float> pos(4 * (size_t)N);
  std::mt19937 gen(5);
  std::normal_distribution<float> jit(0.f, 0.05f);
  const float a = L / n, basis[4][3] = {{0, 0, 0}, {.5f, .5f, 0}, {.5f, 0, .5f}, {0, .5f, .5f}};
  size_t k = 0;
  for (int x = 0; x < n; x++) for (int y = 0; y < n; y++) for (int z = 0; z < n; z++) for (int b = 0; b < 4; b++) {
    pos[4 * k + 0] = (x + basis[b][0] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 1] = (y + basis[b][1] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 2] = (z + basis[b][2] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 3] = 0.f;
    k++;
  }
  float *d_pos, *d_f0, *d_f1;
  cudaMalloc(&d_pos, 16 * (size_t)N); cudaMalloc(&d_f0, 16 * (size_t)N); cudaMalloc(&d_f1, 16 * (size_t)N);
  cudaMemcpy(d_pos, pos.data(), 16 * (size_t)N, cudaMemcpyHostToDevice);
  const float Lv[3] = {L, L, L};
  const int per[3] = {1, 1, 1};
  int cd[3];
  ub200_neighbour_celldim_f32(Lv, 2.5f, cd);
  ub200_celllist *cl;
  ub200_celllist_create(&cl);
  int rc = ub200_celllist_build_f32(cl, d_pos, nullptr, N, Lv, per, cd, nullptr);
  const float par[4] = {6.25f, 1.f, 1.f, 0.f};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms[2] = {0, 0};
  float *out[2] = {d_f0, d_f1};
  for (int v = 0; v < 2; v++) {
    setenv("UB200_LJ_PACKED", v ? "1" : "0", 1);
    for (int it = 0; it < 3; it++) { // 2 warm-ups, the third pass is timed
      cudaMemset(out[v], 0, 16 * (size_t)N);
      cudaEventRecord(e0);
      rc |= ub200_lj_sum_f32(cl, par, 1, out[v], nullptr, nullptr, nullptr, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
    }
  }
  std::vector<float> f0(4 * (size_t)N), f1(4 * (size_t)N);
  cudaMemcpy(f0.data(), d_f0, 16 * (size_t)N, cudaMemcpyDeviceToHost);
  cudaMemcpy(f1.data(), d_f1, 16 * (size_t)N, cudaMemcpyDeviceToHost);
  size_t diff = 0;
  double fmax = 0, dmax = 0;
  for (size_t i = 0; i < f0.size(); i++) {
    diff += memcmp(&f0[i], &f1[i], 4) != 0;
    fmax = std::max(fmax, (double)std::fabs(f0[i]));
    dmax = std::max(dmax, (double)std::fabs(f0[i] - f1[i]));
  }
  printf("{\"N\":%d,\"rc\":%d,\"cuda\":\"%s\",\"ms_default\":%.4f,\"ms_packed\":%.4f,\"differing_words\":%zu,\"max_abs_diff\":%.3g,\"fmax\":%.4g}\n",
         N, rc, cudaGetErrorString(cudaGetLastError()), ms[0], ms[1], diff, dmax, fmax);
  return 0;
}
