/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH (see oracle.h).
 * CPU restatement of path 1: cell assignment, Morton-ordered cell list, LJ and DPD neighbour traversal,
 * velocity Verlet. Citations are file:line relative to /root/reference/src.
 *
 * Floating point: compiled with -ffp-contract=off; where nvcc's default FMA contraction shapes the
 * reference's device arithmetic we call fmaf() explicitly and say so.
 */
#include "oracle.h"
#include "oracle_saru.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------- Box / Grid ---------- */
/* utils/Box.cuh:16-36 (minusInvBoxSize, zero for L==0 / inf / non periodic), utils/Grid.cuh:35-48 */
void orc_grid_init_f(orc_grid_f *g, const float L[3], const int periodic[3], const int cellDim[3]) {
  for (int d = 0; d < 3; d++) {
    g->L[d] = L[d];
    g->minusInvL[d] = -1.0f / L[d];
    if (L[d] == 0.0f || isinf(L[d]) || !periodic[d]) g->minusInvL[d] = 0.0f;
    g->cellDim[d] = cellDim[d];
    if (d == 2 && g->cellDim[d] == 0) g->cellDim[d] = 1;
    g->cellSize[d] = L[d] / (float)g->cellDim[d];
    g->invCellSize[d] = 1.0f / g->cellSize[d];
  }
  if (L[2] == 0.0f) g->invCellSize[2] = 0.0f;
}

void orc_grid_init_d(orc_grid_d *g, const double L[3], const int periodic[3], const int cellDim[3]) {
  for (int d = 0; d < 3; d++) {
    g->L[d] = L[d];
    g->minusInvL[d] = -1.0 / L[d];
    if (L[d] == 0.0 || isinf(L[d]) || !periodic[d]) g->minusInvL[d] = 0.0;
    g->cellDim[d] = cellDim[d];
    if (d == 2 && g->cellDim[d] == 0) g->cellDim[d] = 1;
    g->cellSize[d] = L[d] / (double)g->cellDim[d];
    g->invCellSize[d] = 1.0 / g->cellSize[d];
  }
  if (L[2] == 0.0) g->invCellSize[2] = 0.0;
}

/* Interactor/NeighbourList/CellList.cuh:100-126: cellDim = int(L/rc), dims with <= 3 cells collapse to 1.
   (Infinite boxes - 64 cells of rc, non periodic - are resolved by the caller into a finite L.) */
void orc_neighbour_celldim_f(const float L[3], float rc, int cellDim[3]) {
  for (int d = 0; d < 3; d++) {
    int c = (int)(L[d] / rc);
    if (c <= 3) c = 1;
    cellDim[d] = c;
  }
}

/* Box::apply_pbc utils/Box.cuh:51-58. nvcc contracts r*minusInvL+0.5 and r+offset*L into FMAs. */
static inline float pbc1_f(float r, float L, float minusInvL) {
  if (minusInvL == 0.0f) return r;
  float offset = floorf(fmaf(r, minusInvL, 0.5f));
  return fmaf(offset, L, r);
}
static inline double pbc1_d(double r, double L, double minusInvL) {
  if (minusInvL == 0.0) return r;
  double offset = floor(fma(r, minusInvL, 0.5));
  return fma(offset, L, r);
}

/* Grid::getCell utils/Grid.cuh:49-71 */
void orc_get_cell_f(const orc_grid_f *g, const float *p, int cell[3]) {
  for (int d = 0; d < 3; d++) {
    float r = pbc1_f(p[d], g->L[d], g->minusInvL[d]);
    int c = (int)((r + 0.5f * g->L[d]) * g->invCellSize[d]);
    if (c == g->cellDim[d]) c = 0;
    cell[d] = c;
  }
}
void orc_get_cell_d(const orc_grid_d *g, const double *p, int cell[3]) {
  for (int d = 0; d < 3; d++) {
    double r = pbc1_d(p[d], g->L[d], g->minusInvL[d]);
    int c = (int)((r + 0.5 * g->L[d]) * g->invCellSize[d]);
    if (c == g->cellDim[d]) c = 0;
    cell[d] = c;
  }
}

/* Sorter::MortonHash utils/ParticleSorter.cuh:51-76: 10 bits per dimension interleaved, x lowest */
static inline uint32_t spread10(uint32_t v) {
  uint32_t x = v & 0x3ffu;
  x = (x | (x << 16)) & 0x30000ffu;
  x = (x | (x << 8)) & 0x300f00fu;
  x = (x | (x << 4)) & 0x30c30c3u;
  x = (x | (x << 2)) & 0x9249249u;
  return x;
}
uint32_t orc_morton_hash(int cx, int cy, int cz) {
  return spread10((uint32_t)cx) | (spread10((uint32_t)cy) << 1) | (spread10((uint32_t)cz) << 2);
}

/* ---------- cell list ---------- */
/* CellListBase::update Interactor/NeighbourList/CellList/CellListBase.cuh:124-140:
   hash (ParticleSorter.cuh:102-111) -> stable ascending sort of (hash,index) on the low bits
   (cub::DeviceRadixSort::SortPairs, ParticleSorter.cuh:303-321; keys < 2^end_bit so it equals a full
   stable sort) -> sortPos gather (:179-187) -> fillCellList (CellListBase.cuh:68-95). */
int orc_celllist_build_f(const orc_grid_f *g, const float *pos4, int N, float *sortPos4, int *index,
                         int *cellStart, int *cellEnd) {
  const int ncells = g->cellDim[0] * g->cellDim[1] * g->cellDim[2];
  uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)N * 2);
  int *idx = (int *)malloc(sizeof(int) * (size_t)N * 2);
  uint32_t *key2 = key + N;
  int *idx2 = idx + N;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    int c[3];
    orc_get_cell_f(g, pos4 + 4 * (size_t)i, c);
    key[i] = orc_morton_hash(c[0], c[1], c[2]);
    idx[i] = i;
  }
  /* stable LSD radix sort, 3 passes of 11 bits (keys are 30 bit) */
  for (int pass = 0; pass < 3; pass++) {
    const int shift = 11 * pass;
    size_t count[2049];
    memset(count, 0, sizeof(count));
    for (int i = 0; i < N; i++) count[((key[i] >> shift) & 2047u) + 1]++;
    for (int b = 0; b < 2048; b++) count[b + 1] += count[b];
    for (int i = 0; i < N; i++) {
      size_t d = count[(key[i] >> shift) & 2047u]++;
      key2[d] = key[i];
      idx2[d] = idx[i];
    }
    uint32_t *tk = key; key = key2; key2 = tk;
    int *ti = idx; idx = idx2; idx2 = ti;
  }
  for (int c = 0; c < ncells; c++) { cellStart[c] = -1; cellEnd[c] = -1; }
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < N; k++) {
    index[k] = idx[k];
    memcpy(sortPos4 + 4 * (size_t)k, pos4 + 4 * (size_t)idx[k], 4 * sizeof(float));
  }
  int prev = 0;
  for (int k = 0; k < N; k++) {
    int c[3];
    orc_get_cell_f(g, sortPos4 + 4 * (size_t)k, c);
    long icell = c[0] + (long)g->cellDim[0] * (c[1] + (long)g->cellDim[1] * c[2]);
    if (c[0] < 0 || c[1] < 0 || c[2] < 0 || icell >= ncells || icell < 0) { err = 1; break; }
    if (k == 0 || icell != prev) {
      cellStart[icell] = k;
      if (k > 0) cellEnd[prev] = k;
    }
    if (k == N - 1) cellEnd[icell] = N;
    prev = (int)icell;
  }
  /* after the three swaps key/idx point at the second halves; free the original bases */
  uint32_t *kbase = key < key2 ? key : key2;
  int *ibase = idx < idx2 ? idx : idx2;
  free(kbase);
  free(ibase);
  return err;
}

/* ---------- neighbour walk ---------- */
/* NeighbourIterator Interactor/NeighbourList/CellList/NeighbourContainer.cuh:95-138: up to 27 cells,
   x offset fastest, dims with one cell are not expanded, single +-ncells wrap (Grid.cuh:81-106),
   wrapped cells skipped in non periodic dims. Returns number of cells written to out[]. */
static int neighbour_cells(const orc_grid_f *g, const int celli[3], int out[27]) {
  const int *n = g->cellDim;
  const int np[3] = {n[0] > 1 ? 3 : 1, n[1] > 1 ? 3 : 1, n[2] > 1 ? 3 : 1};
  const int total = np[0] * np[1] * np[2];
  int m = 0;
  for (int c = 0; c < total; c++) {
    int cj[3] = {celli[0], celli[1], celli[2]};
    if (np[0] > 1) cj[0] += c % 3 - 1;
    if (np[1] > 1) cj[1] += (c / np[0]) % 3 - 1;
    if (np[2] > 1) cj[2] += c / (np[0] * np[1]) - 1;
    int skip = 0;
    for (int d = 0; d < 3; d++) {
      const int periodic = g->minusInvL[d] != 0.0f;
      const int nc = periodic ? n[d] : 0;
      if (cj[d] <= -1) cj[d] += nc;
      else if (cj[d] >= nc) cj[d] -= nc;
      /* non periodic: the reference leaves cj unwrapped and reads a far-away (or out of range) cell whose
         particles can never be within rc without MIC (>= 2 cells away since dims <= 3 collapse); skip it. */
      if (!periodic && (cj[d] < 0 || cj[d] >= n[d])) skip = 1;
    }
    if (!skip) out[m++] = cj[0] + n[0] * (cj[1] + n[1] * cj[2]);
  }
  return m;
}

/* ---------- LJ ---------- */
/* LJFunctor::force Interactor/Potential/Potential.cuh:37-46 (returns |F|/r) */
static inline float lj_force_f(float r2, const float *p) { /* p = {cutOff2, sigma2, epsDivSigma2, shift} */
  if (r2 >= p[0]) return 0.0f;
  const float invr2 = p[1] / r2;
  const float invr6 = invr2 * invr2 * invr2;
  return p[2] * fmaf(-48.0f, invr6, 24.0f) * invr6 * invr2;
}
/* LJFunctor::energy Potential.cuh:48-65 */
static inline float lj_energy_f(float r2, const float *p) {
  if (r2 >= p[0]) return 0.0f;
  const float invr2 = p[1] / r2;
  const float invr6 = invr2 * invr2 * invr2;
  const float E = p[2] * p[1] * 4.0f * invr6 * (invr6 - 1.0f) - p[3];
  return 0.5f * E;
}

/* transverseWithNeighbourContainer Interactor/NeighbourList/common.cuh:10-34 driving
   Radial<LJFunctor>::Transverser::compute/set Interactor/Potential/RadialPotential.cuh:107-127.
   One sequential fp32 accumulator per particle in the reference's visiting order (self included: r2==0 -> 0). */
void orc_lj_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                const int *cellEnd, int N, const float *params4, int ntypes, float *force4, float *energy,
                float *virial) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int id = 0; id < N; id++) {
    const float *pi = sortPos4 + 4 * (size_t)id;
    int celli[3], cells[27];
    orc_get_cell_f(g, pi, celli);
    const int ncl = neighbour_cells(g, celli, cells);
    float F[3] = {0, 0, 0}, E = 0, V = 0;
    for (int c = 0; c < ncl; c++) {
      const int cs = cellStart[cells[c]];
      if (cs < 0) continue;
      const int ce = cellEnd[cells[c]];
      for (int j = cs; j < ce; j++) {
        const float *pj = sortPos4 + 4 * (size_t)j;
        float r12[3];
        for (int d = 0; d < 3; d++) r12[d] = pbc1_f(pj[d] - pi[d], g->L[d], g->minusInvL[d]);
        const float *p = params4 + 4 * ((int)pi[3] * ntypes + (int)pj[3]);
        const float r2 = fmaf(r12[2], r12[2], fmaf(r12[1], r12[1], r12[0] * r12[0]));
        if (r2 == 0.0f) continue;
        const float fm = lj_force_f(r2, p);
        const float f[3] = {fm * r12[0], fm * r12[1], fm * r12[2]};
        if (energy) E += lj_energy_f(r2, p);
        if (virial) V += fmaf(f[2], r12[2], fmaf(f[1], r12[1], f[0] * r12[0]));
        for (int d = 0; d < 3; d++) F[d] = fmaf(fm, r12[d], F[d]);
      }
    }
    const int ori = index[id];
    if (force4) for (int d = 0; d < 3; d++) force4[4 * (size_t)ori + d] += F[d];
    if (energy) energy[ori] += E;
    if (virial) virial[ori] += V;
  }
}

void orc_lj_f64(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                const int *cellEnd, int N, const float *params4, int ntypes, double *force3, double *energy,
                double *virial, double *abssum, double *sens, double band, double *edge) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int id = 0; id < N; id++) {
    const float *pi = sortPos4 + 4 * (size_t)id;
    int celli[3], cells[27];
    orc_get_cell_f(g, pi, celli);
    const int ncl = neighbour_cells(g, celli, cells);
    double F[3] = {0, 0, 0}, E = 0, V = 0, A = 0, S = 0, G = 0;
    for (int c = 0; c < ncl; c++) {
      const int cs = cellStart[cells[c]];
      if (cs < 0) continue;
      const int ce = cellEnd[cells[c]];
      for (int j = cs; j < ce; j++) {
        const float *pj = sortPos4 + 4 * (size_t)j;
        double r12[3], r2 = 0;
        for (int d = 0; d < 3; d++) {
          double r = (double)pj[d] - (double)pi[d];
          if (g->minusInvL[d] != 0.0f) r -= floor(r / (double)g->L[d] + 0.5) * (double)g->L[d];
          r12[d] = r;
          r2 += r * r;
        }
        if (r2 == 0.0) continue;
        const float *p = params4 + 4 * ((int)pi[3] * ntypes + (int)pj[3]);
        const double invr2 = (double)p[1] / r2, invr6 = invr2 * invr2 * invr2;
        const double fm = (double)p[2] * (-48.0 * invr6 + 24.0) * invr6 * invr2;
        /* the unshifted force jumps by |f(rc)| at the cut-off: a pair within the fp32 uncertainty band of rc
           may legitimately be counted on either side */
        if (fabs(r2 - (double)p[0]) <= band) G += fabs(fm) * sqrt(r2);
        if (r2 >= (double)p[0]) continue;
        for (int d = 0; d < 3; d++) F[d] += fm * r12[d];
        A += fabs(fm) * sqrt(r2);
        /* |d|F|/dr| = 24 eps/sigma^2 * |7 u^4 - 26 u^7| (u = sigma^2/r^2): sensitivity of the pair force to r */
        S += 24.0 * (double)p[2] * fabs(7.0 * invr6 * invr2 - 26.0 * invr6 * invr6 * invr2);
        E += 0.5 * ((double)p[2] * (double)p[1] * 4.0 * invr6 * (invr6 - 1.0) - (double)p[3]);
        V += fm * r2;
      }
    }
    const int ori = index[id];
    if (force3) for (int d = 0; d < 3; d++) force3[3 * (size_t)ori + d] += F[d];
    if (energy) energy[ori] += E;
    if (virial) virial[ori] += V;
    if (abssum) abssum[ori] += A;
    if (sens) sens[ori] += S;
    if (edge) edge[ori] += G;
  }
}

/* ---------- DPD ---------- */
/* DPD_impl::ForceTransverser::compute Interactor/Potential/DPD.cuh:121-158; getInfo(pi) = {vel[pi], pi}
   with pi the GLOBAL (array) index (:154, common.cuh:17,29-31). ij = i + N*j wraps in int32 like the reference. */
void orc_dpd_ids_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                     const int *cellEnd, int N, const float *vel3, float A, float gamma, float sigma, float rcut,
                     uint32_t seed, uint32_t step, int use_double_acc, float *force4, double *force3d,
                     const int *noiseId, int idStride);
void orc_dpd_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                 const int *cellEnd, int N, const float *vel3, float A, float gamma, float sigma, float rcut,
                 uint32_t seed, uint32_t step, int use_double_acc, float *force4, double *force3d) {
  orc_dpd_ids_f32(g, sortPos4, index, cellStart, cellEnd, N, vel3, A, gamma, sigma, rcut, seed, step, use_double_acc,
                  force4, force3d, 0, N);
}
/* same with the Saru key taken from noiseId[array index] and an explicit stride (brick decomposition of
   uammd_b200/domain.py: local arrays, noise keyed on global ids; new functionality, no reference counterpart) */
void orc_dpd_ids_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                     const int *cellEnd, int N, const float *vel3, float A, float gamma, float sigma, float rcut,
                     uint32_t seed, uint32_t step, int use_double_acc, float *force4, double *force3d,
                     const int *noiseId, int idStride) {
  const float invrcut = 1.0f / rcut; /* host double 1.0/rcut narrowed to real (DPD.cuh:110) */
#pragma omp parallel for schedule(dynamic, 256)
  for (int id = 0; id < N; id++) {
    const float *pi = sortPos4 + 4 * (size_t)id;
    const int gi = index[id];
    int celli[3], cells[27];
    orc_get_cell_f(g, pi, celli);
    const int ncl = neighbour_cells(g, celli, cells);
    float F[3] = {0, 0, 0};
    double Fd3[3] = {0, 0, 0};
    for (int c = 0; c < ncl; c++) {
      const int cs = cellStart[cells[c]];
      if (cs < 0) continue;
      const int ce = cellEnd[cells[c]];
      for (int j = cs; j < ce; j++) {
        const float *pj = sortPos4 + 4 * (size_t)j;
        const int gj = index[j];
        float rij[3], vij[3];
        for (int d = 0; d < 3; d++) {
          rij[d] = pbc1_f(pi[d] - pj[d], g->L[d], g->minusInvL[d]);
          vij[d] = vel3[3 * (size_t)gi + d] - vel3[3 * (size_t)gj + d];
        }
        int i = noiseId ? noiseId[gi] : gi, jj = noiseId ? noiseId[gj] : gj;
        if (i > jj) { int t = i; i = jj; jj = t; }
        const uint32_t ij = (uint32_t)i + (uint32_t)idStride * (uint32_t)jj; /* int32 wrap == uint32 wrap */
        const float r2 = fmaf(rij[2], rij[2], fmaf(rij[1], rij[1], rij[0] * rij[0]));
        const float rmod = sqrtf(r2);
        if (rmod == 0.0f) continue;
        const float invrmod = 1.0f / rmod;
        if (invrmod <= invrcut) continue;
        orc_saru rng = orc_saru_seed3(ij, seed, step);
        const float wrf = fmaf(-rmod, invrcut, 1.0f); /* 1 - rmod*invrcut, contracted by nvcc */
        const float Fc = A * wrf * invrmod;
        const float wd = wrf * wrf;
        const float rv = fmaf(rij[2], vij[2], fmaf(rij[1], vij[1], rij[0] * vij[0]));
        const float Fd = -gamma * wd * invrmod * invrmod * rv;
        float gpair[2];
        orc_saru_gf(&rng, 0.0f, sigma * sqrtf(gamma) * wrf * invrmod, gpair);
        const float Fr = gpair[0];
        const float ftot = Fc + Fd + Fr;
        for (int d = 0; d < 3; d++) {
          F[d] = fmaf(ftot, rij[d], F[d]);
          Fd3[d] += (double)ftot * (double)rij[d];
        }
      }
    }
    if (force4 && !use_double_acc) for (int d = 0; d < 3; d++) force4[4 * (size_t)gi + d] += F[d];
    if (force3d) for (int d = 0; d < 3; d++) force3d[3 * (size_t)gi + d] += Fd3[d];
  }
}

void orc_saru3_u32(uint32_t s1, uint32_t s2, uint32_t s3, int n, uint32_t *out) {
  orc_saru r = orc_saru_seed3(s1, s2, s3);
  for (int i = 0; i < n; i++) out[i] = orc_saru_u32(&r);
}
void orc_saru3_gf(uint32_t s1, uint32_t s2, uint32_t s3, float mean, float std, float out[2]) {
  orc_saru r = orc_saru_seed3(s1, s2, s3);
  orc_saru_gf(&r, mean, std, out);
}

/* ---------- velocity Verlet ---------- */
/* VerletNVE_ns::integrateGPU<step> Integrator/VerletNVE.cu:64-85; force/m is (1/m)*force (vector.cuh:191-193) */
void orc_nve_half_f32(float *pos4, float *vel3, const float *force4, int N, float dt, float mass, int step) {
  const float invm = 1.0f / mass;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    for (int d = 0; d < 3; d++) {
      const float a = invm * force4[4 * (size_t)i + d];
      vel3[3 * (size_t)i + d] = fmaf(a * dt, 0.5f, vel3[3 * (size_t)i + d]);
    }
    if (step == 1)
      for (int d = 0; d < 3; d++) pos4[4 * (size_t)i + d] = fmaf(vel3[3 * (size_t)i + d], dt, pos4[4 * (size_t)i + d]);
  }
}

typedef struct {
  int N, ncells;
  float *sortPos;
  int *index, *cellStart, *cellEnd;
} md_scratch;

void *orc_md_scratch_new(int N, int ncells) {
  md_scratch *s = (md_scratch *)malloc(sizeof(md_scratch));
  s->N = N; s->ncells = ncells;
  s->sortPos = (float *)malloc(sizeof(float) * 4 * (size_t)N);
  s->index = (int *)malloc(sizeof(int) * (size_t)N);
  s->cellStart = (int *)malloc(sizeof(int) * (size_t)ncells);
  s->cellEnd = (int *)malloc(sizeof(int) * (size_t)ncells);
  return s;
}
void orc_md_scratch_free(void *p) {
  md_scratch *s = (md_scratch *)p;
  free(s->sortPos); free(s->index); free(s->cellStart); free(s->cellEnd); free(s);
}

/* VerletNVE::forwardTime Integrator/VerletNVE.cu:174-188 (force[] must hold F(t) on entry, as after
   firstStepPreparation :160-171): kick+drift, zero forces, rebuild list + LJ sum, kick. */
int orc_md_step_f32(const float L[3], float rc, const float *params4, float dt, int N, float *pos4, float *vel3,
                    float *force4, void *scratch) {
  md_scratch *s = (md_scratch *)scratch;
  const int periodic[3] = {1, 1, 1};
  int cellDim[3];
  orc_neighbour_celldim_f(L, rc, cellDim);
  orc_grid_f g;
  orc_grid_init_f(&g, L, periodic, cellDim);
  if (cellDim[0] * cellDim[1] * cellDim[2] > s->ncells) return 2;
  orc_nve_half_f32(pos4, vel3, force4, N, dt, 1.0f, 1);
  memset(force4, 0, sizeof(float) * 4 * (size_t)N);
  if (orc_celllist_build_f(&g, pos4, N, s->sortPos, s->index, s->cellStart, s->cellEnd)) return 1;
  orc_lj_f32(&g, s->sortPos, s->index, s->cellStart, s->cellEnd, N, params4, 1, force4, NULL, NULL);
  orc_nve_half_f32(pos4, vel3, force4, N, dt, 1.0f, 2);
  return 0;
}

/* ---------- BD::EulerMaruyama (BASELINE config 0) ---------- */
/* EulerMaruyama_ns::integrateGPU Integrator/BrownianDynamics.cu:117-145, double precision build:
 * M = selfMobility (/radius[i]); R += dt (K R + M F); Saru(i, step, seed): B = sqrt(2 T M dt),
 * dW = (gf(0,B).x, gf(0,B).y, gf'(0,B).x) with gf the FLOAT Box-Muller (saruprng.cuh:115-128). fma() spells out
 * the contractions nvcc applies to the reference kernel. The host libm logf/sinf/cosf differ from the device's in
 * the last ulp, so against the GPU this oracle agrees to ~1e-6 of the noise amplitude; bit parity is pinned by the
 * compiled reference (oracle/_ref/ref_bd). */
void orc_bd_euler_maruyama_f64(double *pos4, const double *force4, const double *K9, double selfMobility,
                               const double *radius, double dt, int is2D, double temperature, int N, uint32_t step,
                               uint32_t seed) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    double R[3] = {pos4[4 * (size_t)i], pos4[4 * (size_t)i + 1], pos4[4 * (size_t)i + 2]};
    double F[3] = {0, 0, 0}, KR[3] = {0, 0, 0};
    if (force4) for (int d = 0; d < 3; d++) F[d] = force4[4 * (size_t)i + d];
    if (K9) for (int d = 0; d < 3; d++) KR[d] = fma(K9[3 * d + 2], R[2], fma(K9[3 * d + 1], R[1], K9[3 * d] * R[0]));
    const double M = selfMobility * (radius ? 1.0 / radius[i] : 1.0);
    double out[3];
    for (int d = 0; d < 3; d++) out[d] = fma(dt, fma(M, F[d], KR[d]), R[d]);
    if (temperature > 0) {
      orc_saru r = orc_saru_seed3((uint32_t)i, step, seed);
      const float B = (float)sqrt(2.0 * temperature * M * dt);
      float g0[2], g1[2];
      orc_saru_gf(&r, 0.0f, B, g0);
      orc_saru_gf(&r, 0.0f, B, g1);
      out[0] += (double)g0[0]; out[1] += (double)g0[1]; out[2] += (double)g1[0];
    }
    pos4[4 * (size_t)i] = out[0];
    pos4[4 * (size_t)i + 1] = out[1];
    if (!is2D) pos4[4 * (size_t)i + 2] = out[2];
  }
}

/* ---------- brick domain decomposition (uammd_b200/csrc/domain.cu; new functionality, no reference counterpart) ----------
   Defined on the reference's neighbour grid: owner = brick holding the particle's cell (Grid::getCell, utils/Grid.cuh:49-71),
   ghost mask = ranks owning a cell of the 27-neighbourhood (wrap like Grid::pbc_cell, utils/Grid.cuh:81-106). */
static int brick_of_cell(int c, int n, int p) {
  int k = 0; /* largest k with floor(k n / p) <= c */
  while (k + 1 < p && ((k + 1) * n) / p <= c) k++;
  return k;
}
void orc_brick_classify_f(const orc_grid_f *g, const float *pos4, int N, const int rankGrid[3], int *cell, int *owner,
                          uint32_t *ghostMask) {
  const int *n = g->cellDim;
  for (int i = 0; i < N; i++) {
    int c[3];
    orc_get_cell_f(g, pos4 + 4 * (size_t)i, c);
    for (int d = 0; d < 3; d++) c[d] = c[d] < 0 ? 0 : (c[d] >= n[d] ? n[d] - 1 : c[d]);
    const int own = brick_of_cell(c[0], n[0], rankGrid[0]) +
                    rankGrid[0] * (brick_of_cell(c[1], n[1], rankGrid[1]) + rankGrid[1] * brick_of_cell(c[2], n[2], rankGrid[2]));
    uint32_t mask = 0;
    for (int o = 0; o < 27; o++) {
      int j[3] = {c[0] + o % 3 - 1, c[1] + (o / 3) % 3 - 1, c[2] + o / 9 - 1};
      int skip = 0;
      for (int d = 0; d < 3; d++) {
        const int periodic = g->minusInvL[d] != 0.0f;
        if (j[d] < 0) { if (periodic) j[d] += n[d]; else skip = 1; }
        else if (j[d] >= n[d]) { if (periodic) j[d] -= n[d]; else skip = 1; }
      }
      if (skip) continue;
      const int r = brick_of_cell(j[0], n[0], rankGrid[0]) +
                    rankGrid[0] * (brick_of_cell(j[1], n[1], rankGrid[1]) + rankGrid[1] * brick_of_cell(j[2], n[2], rankGrid[2]));
      mask |= 1u << r;
    }
    if (cell) cell[i] = c[0] + n[0] * (c[1] + n[1] * c[2]);
    owner[i] = own;
    ghostMask[i] = mask & ~(1u << own);
  }
}

/* ---------- VerletNVT::GronbechJensen ---------- */
/* GronbechJensen_ns::integrateGPU<step> Integrator/VerletNVT/GronbechJensen.cu:30-66, single precision build:
 *   step 1: Saru(id, stepNum, seed); beta = (gf.x, gf.y, gf'.x) with std = noiseAmplitude * rsqrt(1/m);
 *           b = 1/(1 + friction dt/2), a = (1 - friction dt/2) b;
 *           p += b dt v + (1/2)(1/m) dt b (dt f + beta);  v = a v + dt (1/2)(1/m) a f + b (1/m) beta;  f = 0
 *   step 2: v += dt (1/2)(1/m) f
 * fmaf() spells out the contractions nvcc applies to the reference kernel for sm_100a (read from its PTX). The host
 * libm logf/sinf/cosf and the exact 1/sqrt used here differ from the device's in the last ulp, so against the GPU
 * this oracle agrees to ~1e-6 of the noise amplitude; bit parity is pinned by the compiled reference (oracle/_ref/ref_nvt). */
void orc_nvt_gj_half_f32(float *pos4, float *vel3, float *force4, const float *mass, float defaultMass, int N, float dt,
                         float friction, int is2D, float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    const float invMass = 1.0f / (defaultMass > 0.0f ? defaultMass : mass[i]);
    float *p = pos4 + 4 * (size_t)i, *v = vel3 + 3 * (size_t)i, *f = force4 + 4 * (size_t)i;
    if (step == 1) {
      orc_saru rng = orc_saru_seed3((uint32_t)i, stepNum, seed);
      const float amp = noiseAmplitude * (1.0f / sqrtf(invMass));
      float n[3] = {0, 0, 0}, g2[2];
      orc_saru_gf(&rng, 0.0f, amp, g2);
      n[0] = g2[0]; n[1] = g2[1];
      if (!is2D) { orc_saru_gf(&rng, 0.0f, amp, g2); n[2] = g2[0]; }
      const float g = (dt * friction) * 0.5f;
      const float b = 1.0f / (g + 1.0f);
      const float a = (1.0f - g) * b;
      const float bdt = dt * b;
      const float c2 = b * (dt * (invMass * 0.5f));
      const float c3 = a * ((dt * 0.5f) * invMass);
      const float c4 = b * invMass;
      for (int d = 0; d < 3; d++) {
        const float pd = fmaf(c2, fmaf(dt, f[d], n[d]), fmaf(bdt, v[d], p[d]));
        const float vd = fmaf(c4, n[d], fmaf(a, v[d], c3 * f[d]));
        p[d] = pd;
        v[d] = vd;
      }
      f[0] = f[1] = f[2] = f[3] = 0.0f;
    } else {
      const float c = (dt * 0.5f) * invMass;
      for (int d = 0; d < 3; d++) v[d] = fmaf(f[d], c, v[d]);
    }
    if (is2D) v[2] = 0.0f;
  }
}

/* VerletNVT::Basic_ns::integrateGPU<step> Integrator/VerletNVT/Basic.cu:87-117, single precision build: in BOTH half
 * steps v += (f/m - friction v) dt/2 + noise with noise = gf(0, noiseAmplitude sqrt(1/(2m))) drawn from
 * Saru(id + N (step - 1), stepNum, seed); step 1 then drifts (x += v dt) and zeroes the force. fmaf() marks the
 * contractions nvcc applies to the reference kernel (read from its PTX). */
void orc_nvt_basic_half_f32(float *pos4, float *vel3, float *force4, const float *mass, float defaultMass, int N, int Ngroup,
                            float dt, float friction, int is2D, float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    const float invMass = 1.0f / (defaultMass > 0.0f ? defaultMass : mass[i]);
    float *p = pos4 + 4 * (size_t)i, *v = vel3 + 3 * (size_t)i, *f = force4 + 4 * (size_t)i;
    orc_saru rng = orc_saru_seed3((uint32_t)(i + Ngroup * (step - 1)), stepNum, seed); /* Ngroup: particles of the group (N may be a prefix of it) */
    const float amp = noiseAmplitude * sqrtf(invMass * 0.5f);
    float n[3], g2[2];
    orc_saru_gf(&rng, 0.0f, amp, g2);
    n[0] = g2[0]; n[1] = g2[1];
    orc_saru_gf(&rng, 0.0f, amp, g2);
    n[2] = g2[0];
    const float hdt = dt * 0.5f;
    for (int d = 0; d < 3; d++) {
      const float t = invMass * f[d], u = friction * v[d];
      v[d] = v[d] + fmaf(hdt, t - u, n[d]);
    }
    if (is2D) v[2] = 0.0f;
    if (step == 1) {
      for (int d = 0; d < 3; d++) p[d] = fmaf(dt, v[d], p[d]);
      f[0] = f[1] = f[2] = f[3] = 0.0f;
    }
  }
}

/* Basic_ns::initialVelocities Integrator/VerletNVT/Basic.cu:12-29 (mass ignored: mass_i = 1; no group here) */
void orc_nvt_initial_velocities_f32(float *vel3, int N, float vamp, int is2D, uint32_t seed) {
  for (int i = 0; i < N; i++) {
    orc_saru rng = orc_saru_seed2((uint32_t)i, seed);
    double g0[2], g1[2] = {0.0, 0.0};
    orc_saru_gd(&rng, 0.0, (double)(vamp / 1.0f), g0);
    if (!is2D) orc_saru_gd(&rng, 0.0, (double)(vamp / 1.0f), g1);
    vel3[3 * (size_t)i] = (float)g0[0];
    vel3[3 * (size_t)i + 1] = (float)g0[1];
    vel3[3 * (size_t)i + 2] = (float)g1[0];
  }
}
