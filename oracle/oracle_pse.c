/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH (see oracle.h).
 * CPU restatement (double precision) of BDHI::PSE: the far-field spectral operator and Fourier noise, the closed-form
 * near-field RPY coefficients, the tabulated lookup and a direct O(N^2) near-field mat-vec with the sheared minimum
 * image. Citations are file:line relative to /root/reference/src.
 */
#include "oracle.h"
#include "oracle_saru.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int foldk(int i, int n) { return i - n * (i >= (n / 2 + 1)); }

/* pse_ns::detail::greensFunction Integrator/BDHI/PSE/FarField.cuh:85-119; k = unsheared wave vector */
double orc_pse_greens_d(const double k[3], double shear, double rh, double viscosity, double split, double eta, double ntot) {
  const double k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2];
  if (k2 == 0) return 0.0;
  const double kE[3] = {k[0], k[1] - shear * k[0], k[2]}; /* shearWaveVector PSE/utils.cuh:36-39 */
  const double KE2 = kE[0] * kE[0] + kE[1] * kE[1] + kE[2] * kE[2];
  const double kmod = sqrt(KE2), invk2 = 1.0 / KE2, sink = sin(kmod * rh);
  const double kEw = KE2 / (4.0 * split * split), kNu = k2 / (4.0 * split * split);
  const double tau = eta * kNu - kEw;
  const double hashimoto = (1.0 + kEw) * exp(tau) / KE2;
  double B = sink * sink * invk2 * hashimoto / (viscosity * rh * rh);
  B /= ntot;
  return B;
}

/* projectFourier FarField.cuh:53-73 on one (re or im) triple */
static void project(const double k[3], const double f[3], double out[3]) {
  const double invk2 = 1.0 / (k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
  const double kf = (k[0] * f[0] + k[1] * f[1] + k[2] * f[2]) * invk2;
  for (int d = 0; d < 3; d++) out[d] = f[d] - k[d] * kf;
}

static void wave_vectors(const orc_grid_d *g, const int cell[3], double shear, double k[3], double kE[3]) {
  const int n[3] = {g->cellDim[0], g->cellDim[1], g->cellDim[2]};
  for (int d = 0; d < 3; d++) k[d] = (2.0 * M_PI / g->L[d]) * foldk(cell[d], n[d]);
  kE[0] = k[0]; kE[1] = k[1] - shear * k[0]; kE[2] = k[2];
}

/* forceFourier2Vel FarField.cuh:134-153. ghat: [(nx/2+1)*ny*nz][3 components][re,im] */
void orc_pse_force2vel_d(const orc_grid_d *g, double shear, double rh, double viscosity, double split, double eta,
                         double *ghat) {
  const int nx = g->cellDim[0], ny = g->cellDim[1], nz = g->cellDim[2], nkx = nx / 2 + 1;
  const double ntot = (double)(nx * ny * nz);
  for (int iz = 0; iz < nz; iz++)
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < nkx; ix++) {
        const size_t id = (size_t)ix + (size_t)nkx * ((size_t)iy + (size_t)ny * iz);
        double *v = ghat + 6 * id;
        if (id == 0) { memset(v, 0, 6 * sizeof(double)); continue; }
        const int cell[3] = {ix, iy, iz};
        double k[3], kE[3];
        wave_vectors(g, cell, shear, k, kE);
        const double B = orc_pse_greens_d(k, shear, rh, viscosity, split, eta, ntot);
        for (int c = 0; c < 2; c++) {
          const double f[3] = {B * v[0 + c], B * v[2 + c], B * v[4 + c]};
          double o[3];
          project(kE, f, o);
          v[0 + c] = o[0]; v[2 + c] = o[1]; v[4 + c] = o[2];
        }
      }
}

static int pse_is_nyquist(const int c[3], const int n[3]) { /* FarField.cuh:183-219 */
  const int xq = (c[0] == n[0] - c[0]) && (n[0] % 2 == 0);
  const int yq = (c[1] == n[1] - c[1]) && (n[1] % 2 == 0);
  const int zq = (c[2] == n[2] - c[2]) && (n[2] % 2 == 0);
  return (xq && c[1] == 0 && c[2] == 0) || (xq && yq && c[2] == 0) || (c[0] == 0 && yq && c[2] == 0) ||
         (xq && c[1] == 0 && zq) || (c[0] == 0 && c[1] == 0 && zq) || (c[0] == 0 && yq && zq) || (xq && yq && zq);
}

static void pse_noise_term(const orc_grid_d *g, double shear, double rh, double viscosity, double split, double eta,
                           const int cell[3], const double nzv[6], int conj, double *dst) {
  double k[3], kE[3];
  wave_vectors(g, cell, shear, k, kE);
  const double ntot = (double)(g->cellDim[0] * g->cellDim[1] * g->cellDim[2]);
  const double Bsq = sqrt(orc_pse_greens_d(k, shear, rh, viscosity, split, eta, ntot));
  for (int c = 0; c < 2; c++) {
    const double sgn = (c == 1 && conj) ? -1.0 : 1.0;
    const double f[3] = {sgn * nzv[0 + c], sgn * nzv[2 + c], sgn * nzv[4 + c]};
    double o[3];
    project(kE, f, o);
    for (int d = 0; d < 3; d++) dst[2 * d + c] += Bsq * o[d];
  }
}

/* fourierBrownianNoise FarField.cuh:235-308, node by node in index order (race-free sum of the reference's two
   non-atomic "+="). noisePrefactor = prefactor*sqrt(2T/dV) (addBrownianNoise :467-492). */
void orc_pse_add_noise_d(const orc_grid_d *g, double shear, double rh, double viscosity, double split, double eta,
                         double noisePrefactor, uint32_t seed1, uint32_t seed2, double *ghat) {
  const int n[3] = {g->cellDim[0], g->cellDim[1], g->cellDim[2]};
  const int nkx = n[0] / 2 + 1;
  for (int iz = 0; iz < n[2]; iz++)
    for (int iy = 0; iy < n[1]; iy++)
      for (int ix = 0; ix < nkx; ix++) {
        const uint32_t id = (uint32_t)ix + (uint32_t)nkx * ((uint32_t)iy + (uint32_t)n[1] * iz);
        const int cell[3] = {ix, iy, iz};
        if (id == 0 || (ix == 0 && iy == 0 && 2 * iz >= n[2] + 1) || (ix == 0 && 2 * iy >= n[1] + 1)) continue;
        orc_saru rng = orc_saru_seed3(id, seed1, seed2);
        const float sc = (float)(0.707106781186547 * noisePrefactor);
        double nzv[6];
        for (int c = 0; c < 3; c++) {
          float pr[2];
          orc_saru_gf(&rng, 0.0f, sc, pr);
          nzv[2 * c] = pr[0];
          nzv[2 * c + 1] = pr[1];
        }
        const int nyq = pse_is_nyquist(cell, n);
        if (nyq)
          for (int c = 0; c < 3; c++) { nzv[2 * c] *= 1.41421356237310; nzv[2 * c + 1] = 0.0; }
        pse_noise_term(g, shear, rh, viscosity, split, eta, cell, nzv, 0, ghat + 6 * (size_t)id);
        if (nyq) continue;
        if (ix == n[0] - ix || ix == 0) {
          const int cc[3] = {ix, (iy > 0) * (n[1] - iy), (iz > 0) * (n[2] - iz)};
          const size_t idc = (size_t)cc[0] + (size_t)nkx * ((size_t)cc[1] + (size_t)n[1] * cc[2]);
          pse_noise_term(g, shear, rh, viscosity, split, eta, cc, nzv, 1, ghat + 6 * idc);
        }
      }
}

/* RPYPSE_near::FandG + params2FG Integrator/BDHI/PSE/RPY_PSE.cuh:45-128 (not yet divided by 6 pi eta a) */
static double fg_combine(double r, double rh, double psi, const double c[8]) {
  const double psisq = psi * psi, a2mr = 2 * rh - r, a2pr = 2 * rh + r, rsq = r * r;
  return c[0] + c[1] * exp(-psisq * a2pr * a2pr) + c[2] * exp(-a2mr * a2mr * psisq) + c[3] * exp(-psisq * rsq) +
         c[4] * erfc(a2mr * psi) + c[5] * erfc(-a2mr * psi) + c[6] * erfc(a2pr * psi) + c[7] * erfc(r * psi);
}
void orc_rpy_near_fg(double r, double rh, double psi, double rcut, double out[2]) {
  out[0] = out[1] = 0.0;
  if (r >= rcut) return;
  if (r <= 0.0) {
    const double pi = M_PI;
    out[0] = (1.0 / (4 * sqrt(pi) * psi * rh)) *
             (1 - exp(-4 * rh * rh * psi * psi) + 4 * sqrt(pi) * rh * psi * erfc(2 * rh * psi));
    return;
  }
  const double r2 = r * r, a2mr = 2 * rh - r, a2pr = 2 * rh + r, rh2 = rh * rh, rh4 = rh2 * rh2;
  const double psi2 = psi * psi, psi3 = psi2 * psi, psi4 = psi2 * psi2, r3 = r2 * r, r4 = r3 * r, sp = sqrt(M_PI);
  double f[8], gq[8];
  if (r > 2 * rh) {
    f[0] = (64.0 * rh4 * psi4 + 96.0 * rh2 * r2 * psi4 - 128.0 * rh * r3 * psi4 + 36.0 * r4 * psi4 - 3.0) / (128.0 * rh * r3 * psi4);
    f[4] = (3.0 - 4.0 * psi4 * a2mr * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2)) / (256.0 * rh * r3 * psi4);
    f[5] = 0;
    gq[0] = (-64.0 * rh4 * psi4 + 96.0 * rh2 * r2 * psi4 - 64.0 * rh * r3 * psi4 + 12.0 * r4 * psi4 + 3.0) / (64.0 * rh * r3 * psi4);
    gq[4] = (4.0 * psi4 * a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r) - 3.0) / (128.0 * rh * r3 * psi4);
    gq[5] = 0;
  } else {
    f[0] = (-16.0 * rh4 - 24.0 * rh2 * r2 + 32.0 * rh * r3 - 9.0 * r4) / (32.0 * rh * r3);
    f[4] = 0;
    f[5] = (4.0 * psi4 * a2mr * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2) - 3.0) / (256.0 * rh * r3 * psi4);
    gq[0] = a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r) / (16.0 * rh * r3);
    gq[4] = 0;
    gq[5] = (3.0 - 4.0 * psi4 * a2mr * a2mr * a2mr * (2.0 * rh + 3.0 * r)) / (128.0 * rh * r3 * psi4);
  }
  f[1] = (-2.0 * psi2 * a2pr * (4.0 * rh2 - 4.0 * rh * r + 9.0 * r2) + 2.0 * rh - 3.0 * r) / (128.0 * rh * r3 * psi3 * sp);
  f[2] = (2.0 * psi2 * a2mr * (4.0 * rh2 + 4.0 * rh * r + 9.0 * r2) - 2.0 * rh - 3.0 * r) / (128.0 * rh * r3 * psi3 * sp);
  f[3] = 3.0 * (6.0 * r2 * psi2 + 1.0) / (64.0 * sp * rh * r2 * psi3);
  f[6] = (4.0 * psi4 * a2pr * a2pr * (4.0 * rh2 - 4.0 * rh * r + 9.0 * r2) - 3.0) / (256.0 * rh * r3 * psi4);
  f[7] = 3.0 * (1.0 - 12.0 * r4 * psi4) / (128.0 * rh * r3 * psi4);
  gq[1] = (2.0 * psi2 * a2pr * a2pr * (2.0 * rh - 3.0 * r) - 2.0 * rh + 3.0 * r) / (64.0 * sp * rh * r3 * psi3);
  gq[2] = (-2.0 * psi2 * a2mr * a2mr * (2.0 * rh + 3.0 * r) + 2.0 * rh + 3.0 * r) / (64.0 * sp * rh * r3 * psi3);
  gq[3] = (3.0 * (2.0 * r2 * psi2 - 1.0)) / (32.0 * sp * rh * r2 * psi3);
  gq[6] = (3.0 - 4.0 * psi4 * (2.0 * rh - 3.0 * r) * a2pr * a2pr * a2pr) / (128.0 * rh * r3 * psi4);
  gq[7] = -3.0 * (4.0 * r4 * psi4 + 1.0) / (64.0 * rh * r3 * psi4);
  out[0] = fg_combine(r, rh, psi, f);
  out[1] = fg_combine(r, rh, psi, gq);
}

/* TabulatedFunction constructor (misc/TabulatedFunction.cuh:103-117) for the near-field table: nPoints entries */
void orc_pse_near_table_d(int nPoints, double rh, double psi, double normalization, double rcut, double *table2) {
  const int Ntable = nPoints - 1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i <= Ntable; i++) {
    const double x = (i / (double)Ntable) * rcut;
    double fg[2];
    orc_rpy_near_fg(x, rh, psi, rcut, fg);
    table2[2 * i] = fg[0] / normalization;
    table2[2 * i + 1] = fg[1] / normalization;
  }
}

/* TabulatedFunction::operator() + LinearInterpolation (:63-75,148-158) */
static void table_get(const double *table2, int nPoints, double rcut, double rs, double out[2]) {
  const int Ntable = nPoints - 1;
  const double interval = 1.0 / rcut, dr = 1.0 / (double)Ntable;
  const double r = rs * interval;
  out[0] = out[1] = 0.0;
  if (rs >= rcut) return;
  if (r <= 0.0) { out[0] = table2[0]; out[1] = table2[1]; return; }
  const int i = (int)(r * Ntable);
  const double r0 = i * dr, t = (r - r0) * (double)Ntable;
  for (int c = 0; c < 2; c++) out[c] = fma(t, table2[2 * (i + 1) + c], fma(-t, table2[2 * i + c], table2[2 * i + c]));
}

/* RPYNearTransverser::compute over ALL pairs (NearField.cuh:131-182); out3[i] += sum_j M_ij v_j. useTable = 0
   evaluates F, G in closed form instead (measures the tabulation error). */
void orc_pse_near_mdot_d(int N, const double *pos4, const double *v, int vStride, const double L[3], double shear,
                         double rh, double psi, double normalization, double rcut, const double *table2, int nPoints,
                         int useTable, double *out3) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < N; i++) {
    double acc[3] = {0, 0, 0};
    for (int j = 0; j < N; j++) {
      double r[3] = {pos4[4 * (size_t)j] - pos4[4 * (size_t)i], pos4[4 * (size_t)j + 1] - pos4[4 * (size_t)i + 1],
                     pos4[4 * (size_t)j + 2] - pos4[4 * (size_t)i + 2]};
      r[0] += shear * r[1];
      const double s1 = round(r[1] / L[1]);
      r[0] -= shear * L[1] * s1;
      r[1] -= L[1] * s1;
      r[2] -= L[2] * round(r[2] / L[2]);
      r[0] -= L[0] * round(r[0] / L[0]);
      const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      if (r2 >= rcut * rcut) continue;
      double fg[2];
      if (useTable) table_get(table2, nPoints, rcut, sqrt(r2), fg);
      else { orc_rpy_near_fg(sqrt(r2), rh, psi, rcut, fg); fg[0] /= normalization; fg[1] /= normalization; }
      const double *vj = v + (size_t)j * vStride;
      if (r2 == 0.0) { for (int d = 0; d < 3; d++) acc[d] += fg[0] * vj[d]; continue; }
      const double gmfv = (fg[1] - fg[0]) * (r[0] * vj[0] + r[1] * vj[1] + r[2] * vj[2]) / r2;
      for (int d = 0; d < 3; d++) acc[d] += fg[0] * vj[d] + gmfv * r[d];
    }
    for (int d = 0; d < 3; d++) out3[3 * (size_t)i + d] += acc[d];
  }
}
