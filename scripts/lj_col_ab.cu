// A/B of the LJ traversals through the C ABI on a jittered FCC liquid (N = 4 n^3): the column traversal over the engine's
// half-cell list with TMA staging / with per-lane row copies (UB200_LJ_STAGE) against the cell traversal over the
// reference-layout list. Plain CUDA host program (no Python) so that ncu captures start fast:
//   nvcc -O2 scripts/lj_col_ab.cu -Iinclude -Luammd_b200 -luammd_b200 -o scripts/_bin/lj_col_ab
#include "uammd_b200.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 63; // FCC cells per side: N = 4 n^3 (63 -> 1 000 188)
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const int N = 4 * n * n * n;
  const float L = (float)cbrt(N / 0.8);
  std::vector<float> pos(4 * (size_t)N);
  std::mt19937 gen(5);
  std::normal_distribution<float> jit(0.f, 0.08f);
  const float a = L / n, basis[4][3] = {{0, 0, 0}, {.5f, .5f, 0}, {.5f, 0, .5f}, {0, .5f, .5f}};
  size_t k = 0;
  for (int x = 0; x < n; x++) for (int y = 0; y < n; y++) for (int z = 0; z < n; z++) for (int b = 0; b < 4; b++) {
    pos[4 * k + 0] = (x + basis[b][0] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 1] = (y + basis[b][1] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 2] = (z + basis[b][2] + 0.25f) * a - 0.5f * L + jit(gen);
    pos[4 * k + 3] = 0.f;
    k++;
  }
  float *d_pos, *d_f[3];
  cudaMalloc(&d_pos, 16 * (size_t)N);
  for (auto &p : d_f) cudaMalloc(&p, 16 * (size_t)N);
  cudaMemcpy(d_pos, pos.data(), 16 * (size_t)N, cudaMemcpyHostToDevice);
  char *scrub;
  cudaMalloc(&scrub, 256u << 20);
  const float Lv[3] = {L, L, L};
  const int per[3] = {1, 1, 1};
  const float par[4] = {6.25f, 1.f, 1.f, 0.f};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  float ms[3] = {0, 0, 0}, msSum[2] = {0, 0};
  // 0: cell traversal over the reference-layout list
  int cd[3];
  ub200_neighbour_celldim_f32(Lv, 2.5f, cd);
  ub200_celllist *cl;
  ub200_celllist_create(&cl);
  rc |= ub200_celllist_build_f32(cl, d_pos, nullptr, N, Lv, per, cd, nullptr);
  std::vector<float> t;
  for (int it = 0; it < reps; it++) {
    cudaMemsetAsync(scrub, it, 256u << 20);
    cudaMemsetAsync(d_f[0], 0, 16 * (size_t)N);
    cudaEventRecord(e0);
    rc |= ub200_lj_sum_f32(cl, par, 1, d_f[0], nullptr, nullptr, nullptr, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float m; cudaEventElapsedTime(&m, e0, e1); t.push_back(m);
  }
  std::sort(t.begin(), t.end()); ms[0] = t[t.size() / 2];
  // 1, 2: column traversal, TMA / LDG staging
  int path[2] = {-1, -1}, flag[2] = {0, 0}, cells[3] = {0, 0, 0};
  for (int v = 0; v < 2; v++) {
    setenv("UB200_LJ_STAGE", v ? "ldg" : "tma", 1);
    ub200_ljengine *e;
    ub200_ljengine_create(&e);
    rc |= ub200_ljengine_sum_f32(e, d_pos, nullptr, N, Lv, per, par, 1, d_f[1 + v], nullptr, nullptr, nullptr, 0, 0, 0x7fffffff, nullptr);
    path[v] = ub200_ljengine_last_path(e);
    ub200_ljengine_grid(e, cells);
    t.clear();
    for (int it = 0; it < reps; it++) {
      cudaMemsetAsync(scrub, it, 256u << 20);
      cudaEventRecord(e0);
      rc |= ub200_ljengine_traverse_f32(e, d_f[1 + v], 0, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float m; cudaEventElapsedTime(&m, e0, e1); t.push_back(m);
    }
    std::sort(t.begin(), t.end()); ms[1 + v] = t[t.size() / 2];
    t.clear();
    for (int it = 0; it < reps; it++) {
      cudaMemsetAsync(scrub, it, 256u << 20);
      cudaEventRecord(e0);
      rc |= ub200_ljengine_sum_f32(e, d_pos, nullptr, N, Lv, per, par, 1, d_f[1 + v], nullptr, nullptr, nullptr, 0, 0, 0x7fffffff, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float m; cudaEventElapsedTime(&m, e0, e1); t.push_back(m);
    }
    std::sort(t.begin(), t.end()); msSum[v] = t[t.size() / 2];
    ub200_ljengine_error_flag(e, nullptr, &flag[v]);
    ub200_ljengine_destroy(e);
  }
  std::vector<float> f0(4 * (size_t)N), f1(4 * (size_t)N), f2(4 * (size_t)N);
  cudaMemcpy(f0.data(), d_f[0], 16 * (size_t)N, cudaMemcpyDeviceToHost);
  cudaMemcpy(f1.data(), d_f[1], 16 * (size_t)N, cudaMemcpyDeviceToHost);
  cudaMemcpy(f2.data(), d_f[2], 16 * (size_t)N, cudaMemcpyDeviceToHost);
  size_t diff12 = 0;
  double fmax = 0, d01 = 0;
  for (size_t i = 0; i < f0.size(); i++) {
    diff12 += memcmp(&f1[i], &f2[i], 4) != 0;
    fmax = std::max(fmax, (double)std::fabs(f0[i]));
    d01 = std::max(d01, (double)std::fabs(f0[i] - f1[i]));
  }
  printf("{\"N\":%d,\"rc\":%d,\"cuda\":\"%s\",\"half_cells\":[%d,%d,%d],\"path\":[%d,%d],\"error_flag\":[%d,%d],\"ms_cell_traversal\":%.4f,"
         "\"ms_column_tma\":%.4f,\"ms_column_ldg\":%.4f,\"ms_build_plus_column_tma\":%.4f,\"ms_build_plus_column_ldg\":%.4f,"
         "\"tma_vs_ldg_differing_words\":%zu,\"column_vs_cell_max_abs_diff\":%.3g,\"fmax\":%.4g}\n",
         N, rc, cudaGetErrorString(cudaGetLastError()), cells[0], cells[1], cells[2], path[0], path[1], flag[0], flag[1], ms[0], ms[1],
         ms[2], msSum[0], msSum[1], diff12, d01, fmax);
  return 0;
}
