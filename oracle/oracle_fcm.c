/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH (see oracle.h).
 * CPU restatement of path 2 in double precision: IBM window functions, spreading, interpolation, the FCM
 * Fourier-space Stokes operator and a naive separable DFT for small grids (large grids are transformed
 * with numpy.fft inside tests/). Citations are file:line relative to /root/reference/src.
 */
#include "oracle.h"
#include "oracle_saru.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------- window functions ---------- */
/* IBM_kernels::Peskin::threePoint::phi misc/IBM_kernels.cuh:118-137 */
static double peskin3(double rr, double invh) {
  const double r = fabs(rr) * invh;
  if (r < 0.5) return invh * (1.0 / 3.0) * (1.0 + sqrt(1.0 + (-3.0) * r * r));
  if (r < 1.5) {
    const double omr = 1.0 - r;
    return invh * (1.0 / 6.0) * (5.0 - 3.0 * r - sqrt(1.0 + (-3.0) * omr * omr));
  }
  return 0.0;
}
/* IBM_kernels::Peskin::fourPoint::phi misc/IBM_kernels.cuh:140-157 */
static double peskin4(double rr, double invh) {
  const double r = fabs(rr) * invh;
  if (r < 1.0) return invh * 0.125 * (3.0 - 2.0 * r + sqrt(1.0 + 4.0 * r * (1.0 - r)));
  if (r < 2.0) return invh * 0.125 * (5.0 - 2.0 * r - sqrt(-7.0 + 12.0 * r - 4.0 * r * r));
  return 0.0;
}
/* IBM_kernels::BM misc/IBM_kernels.cuh:83-90 */
static double bm(double zz, double alpha, double beta) {
  const double z = zz / alpha;
  const double dz2 = 1.0 - z * z;
  return dz2 < 0.0 ? 0.0 : exp(beta * (sqrt(dz2) - 1.0));
}
/* detail::kahanSum / detail::integrate misc/IBM_kernels.cuh:44-78: composite Simpson, the three partial sums compensated */
static double kahan_add(double *sum, double *c, double x) {
  const double y = x - *c, t = *sum + y;
  *c = (t - *sum) - y;
  *sum = t;
  return t;
}
double orc_ibm_bm_norm(double alpha, double beta) {
  const int n = 20000; /* even */
  const double dx = alpha / n;
  double s = 0.0, c = 0.0;
  kahan_add(&s, &c, bm(0.0, alpha, beta));
  for (int i = 1; i < n; i++) kahan_add(&s, &c, (i % 2 ? 4.0 : 2.0) * bm(i * dx, alpha, beta));
  kahan_add(&s, &c, bm(alpha, alpha, beta));
  return 2.0 * (dx / 3.0 * s);
}
/* GaussianFlexible::sixPoint::phi_impl misc/IBM_kernels.cuh:168-216, r = |distance| / h */
static double six_point(double r) {
  if (r >= 3.0) return 0.0;
  const double K = 0.714075092976608;
  const double R = r - ceil(r) + 1.0, R2 = R * R, R3 = R2 * R;
  const double alpha = 28.0;
  const double beta = 9.0 / 4.0 - 1.5 * (K + R2) + (22.0 / 3 - 7.0 * K) * R - 7.0 / 3.0 * R3;
  const double gamma = 0.25 * (0.5 * (161.0 / 36 - 59.0 / 6 * K + 5 * K * K) * R2 + 1.0 / 3 * (-109.0 / 24 + 5 * K) * R2 * R2 +
                               5.0 / 18 * R3 * R3);
  const double discr = beta * beta - 4.0 * alpha * gamma;
  const int sgn = (1.5 - K) > 0 ? 1 : -1;
  const double pre = 1.0 / (2 * alpha) * (-beta + sgn * sqrt(discr));
  if (r <= 0) {
    const double rp1 = r + 1.0;
    return 2.0 * pre + 0.25 + 1.0 / 6 * (4 - 3 * K) * rp1 - 1.0 / 6 * rp1 * rp1 * rp1;
  } else if (r <= 1) {
    return 2.0 * pre + 5.0 / 8 - 0.25 * (K + r * r);
  } else if (r <= 2) {
    const double rm1 = r - 1.0;
    return -3.0 * pre + 0.25 - 1.0 / 6.0 * (4 - 3 * K) * rm1 + 1.0 / 6 * rm1 * rm1 * rm1;
  }
  const double rm2 = r - 2.0;
  return pre - 1.0 / 16 + 1.0 / 8 * (K + rm2 * rm2) - 1.0 / 12 * (3 * K - 1) * rm2 - 1.0 / 12 * rm2 * rm2 * rm2;
}
/* FCM_ns::Kernels::Gaussian::phi Integrator/BDHI/FCM/FCM_kernels.cuh:54-56 over IBM_kernels::Gaussian
   misc/IBM_kernels.cuh:28-40 */
double orc_ibm_phi(const orc_ibm_kernel *k, double r) {
  switch (k->kind) {
  case ORC_KERNEL_PESKIN3: return peskin3(r, 1.0 / k->h);
  case ORC_KERNEL_PESKIN4: return peskin4(r, 1.0 / k->h);
  case ORC_KERNEL_BARNETT_MAGLAND: return bm(r, k->rmax, k->tau) * k->prefactor;
  case ORC_KERNEL_SIXPOINT: return six_point(fabs(r) / k->h) / k->h;
  default: return r >= k->rmax ? 0.0 : k->prefactor * exp(k->tau * r * r);
  }
}

static inline double pbc1(double r, double L, double minusInvL) {
  if (minusInvL == 0.0) return r;
  return r + floor(r * minusInvL + 0.5) * L;
}
/* Grid::distanceToCellCenter utils/Grid.cuh:124-131 (one coordinate) */
static inline double dist_to_center(const orc_grid_d *g, int d, double p, int cell) {
  return pbc1(p + g->L[d] * 0.5 - g->cellSize[d] * ((double)cell + 0.5), g->L[d], g->minusInvL[d]);
}
/* Grid::pbc_cell_coord utils/Grid.cuh:90-106 */
static inline int pbc_cell(const orc_grid_d *g, int d, int c) {
  const int nc = g->minusInvL[d] != 0.0 ? g->cellDim[d] : 0;
  if (c <= -1) c += nc;
  else if (c >= nc) c -= nc;
  return c;
}

/* per particle stencil: IBM_ns::detail::computeSupportShift misc/IBM.cu:11-31 and fillSharedWeights :33-66 */
typedef struct {
  int cell[3];
  int P[3];
  double w[3][32];
} stencil;

static void make_stencil(const orc_grid_d *g, const orc_ibm_kernel *k, const double *p, stencil *s) {
  orc_get_cell_d(g, p, s->cell);
  const int sup = k->support;
  for (int d = 0; d < 3; d++) {
    int P = sup / 2;
    const double dl = fabs(dist_to_center(g, d, p[d], s->cell[d] - P));
    if (g->cellSize[d] > 0 && dl > sup * g->cellSize[d] / 2.0) P -= 1;
    s->P[d] = P;
    for (int i = 0; i < sup; i++) {
      const int cj = pbc_cell(g, d, s->cell[d] + i - P);
      s->w[d][i] = 0.0;
      if (cj >= 0) s->w[d][i] = orc_ibm_phi(k, dist_to_center(g, d, p[d], cj));
    }
  }
}

/* IBM_ns::particles2GridD misc/IBM.cu:83-147 with DefaultWeightCompute misc/IBM.cuh:88-97
   (value*phiX*phiY*phiZ) and LinearIndex3D(nxPad, ny, nz) :65-78 */
void orc_ibm_spread_d(const orc_grid_d *g, const orc_ibm_kernel *k, const double *pos4, const double *val3,
                      int N, int nxPad, double *grid3) {
  const int sup = k->support;
  for (int n = 0; n < N; n++) {
    stencil s;
    make_stencil(g, k, pos4 + 4 * (size_t)n, &s);
    for (int kk = 0; kk < sup; kk++)
      for (int jj = 0; jj < sup; jj++)
        for (int ii = 0; ii < sup; ii++) {
          const int cx = pbc_cell(g, 0, s.cell[0] + ii - s.P[0]);
          const int cy = pbc_cell(g, 1, s.cell[1] + jj - s.P[1]);
          const int cz = pbc_cell(g, 2, s.cell[2] + kk - s.P[2]);
          if (cx < 0 || cy < 0 || cz < 0) continue;
          if (cx >= g->cellDim[0] || cy >= g->cellDim[1] || cz >= g->cellDim[2]) continue;
          const size_t c = (size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g->cellDim[1] * cz);
          for (int d = 0; d < 3; d++) grid3[3 * c + d] += val3[3 * (size_t)n + d] * s.w[0][ii] * s.w[1][jj] * s.w[2][kk];
        }
  }
}

/* IBM_ns::grid2ParticlesDTPP misc/IBM.cu:168-235; quadrature weight = cell volume (IBM.cuh:80-86) */
void orc_ibm_gather_d(const orc_grid_d *g, const orc_ibm_kernel *k, const double *pos4, int N, int nxPad,
                      const double *grid3, double *out3) {
  const int sup = k->support;
  double dV = g->cellSize[0] * g->cellSize[1];
  if (g->cellDim[2] > 1) dV *= g->cellSize[2];
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) {
    stencil s;
    make_stencil(g, k, pos4 + 4 * (size_t)n, &s);
    double acc[3] = {0, 0, 0};
    for (int kk = 0; kk < sup; kk++)
      for (int jj = 0; jj < sup; jj++)
        for (int ii = 0; ii < sup; ii++) {
          const int cx = pbc_cell(g, 0, s.cell[0] + ii - s.P[0]);
          const int cy = pbc_cell(g, 1, s.cell[1] + jj - s.P[1]);
          const int cz = pbc_cell(g, 2, s.cell[2] + kk - s.P[2]);
          if (cx < 0 || cy < 0 || cz < 0) continue;
          if (cx >= g->cellDim[0] || cy >= g->cellDim[1] || cz >= g->cellDim[2]) continue;
          const size_t c = (size_t)cx + (size_t)nxPad * ((size_t)cy + (size_t)g->cellDim[1] * cz);
          for (int d = 0; d < 3; d++) acc[d] += dV * (grid3[3 * c + d] * s.w[0][ii] * s.w[1][jj] * s.w[2][kk]);
        }
    for (int d = 0; d < 3; d++) out3[3 * (size_t)n + d] += acc[d];
  }
}

/* ---------- Fourier space ---------- */
/* fcm_detail::indexToWaveNumber Integrator/BDHI/FCM/utils.cuh:27-35 */
static inline int fold(int i, int n) { return i - n * (i >= (n / 2 + 1)); }

/* fcm_detail::forceFourier2Vel Integrator/BDHI/FCM/FCM_impl.cuh:375-397 with getGradientFourier utils.cuh:41-51
   and projectFourier :70-100. ghat: [(nx/2+1)*ny*nz][3 components][re,im] */
void orc_fcm_force2vel_d(const orc_grid_d *g, double viscosity, double *ghat) {
  const int nx = g->cellDim[0], ny = g->cellDim[1], nz = g->cellDim[2];
  const int nkx = nx / 2 + 1;
  const double norm = (double)(nx * ny * nz);
  for (int iz = 0; iz < nz; iz++)
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < nkx; ix++) {
        const size_t id = (size_t)ix + (size_t)nkx * ((size_t)iy + (size_t)ny * iz);
        double *v = ghat + 6 * id;
        if (id == 0) { memset(v, 0, 6 * sizeof(double)); continue; }
        const int ik[3] = {fold(ix, nx), fold(iy, ny), fold(iz, nz)};
        const int nk[3] = {nx, ny, nz};
        double k[3], dk[3], k2 = 0;
        for (int d = 0; d < 3; d++) {
          k[d] = (2.0 * M_PI / g->L[d]) * ik[d];
          dk[d] = (ik[d] == nk[d] - ik[d]) ? 0.0 : k[d];
          k2 += k[d] * k[d];
        }
        const double B = 1.0 / (viscosity * k2), invk2 = 1.0 / k2;
        for (int c = 0; c < 2; c++) { /* re, im */
          const double f[3] = {v[0 + c], v[2 + c], v[4 + c]};
          const double fdk = f[0] * (dk[0] * invk2) + f[1] * (dk[1] * invk2) + f[2] * (dk[2] * invk2);
          for (int d = 0; d < 3; d++) v[2 * d + c] = (f[d] - dk[d] * fdk) * (B / norm);
        }
      }
}

/* naive separable DFT (small grids): forward sign -, unnormalised, like cufftExecD2Z with the
   batched-3 interleaved plan of FCM_impl.cuh:179-234 */
void orc_dft3_r2c_d(int nx, int ny, int nz, int nxPad, const double *grid3, double *ghat) {
  const int nkx = nx / 2 + 1;
  const size_t ntot = (size_t)nx * ny * nz;
  double *a = (double *)calloc(ntot * 2, sizeof(double));
  double *b = (double *)calloc(ntot * 2, sizeof(double));
  for (int comp = 0; comp < 3; comp++) {
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int kx = 0; kx < nx; kx++) {
          double re = 0, im = 0;
          for (int x = 0; x < nx; x++) {
            const double v = grid3[3 * ((size_t)x + (size_t)nxPad * (y + (size_t)ny * z)) + comp];
            const double ang = -2.0 * M_PI * (double)((long)kx * x % nx) / nx;
            re += v * cos(ang); im += v * sin(ang);
          }
          const size_t o = 2 * ((size_t)kx + (size_t)nx * (y + (size_t)ny * z));
          a[o] = re; a[o + 1] = im;
        }
    for (int z = 0; z < nz; z++)
      for (int ky = 0; ky < ny; ky++)
        for (int kx = 0; kx < nx; kx++) {
          double re = 0, im = 0;
          for (int y = 0; y < ny; y++) {
            const size_t i = 2 * ((size_t)kx + (size_t)nx * (y + (size_t)ny * z));
            const double ang = -2.0 * M_PI * (double)((long)ky * y % ny) / ny;
            const double c = cos(ang), s = sin(ang);
            re += a[i] * c - a[i + 1] * s; im += a[i] * s + a[i + 1] * c;
          }
          const size_t o = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * z));
          b[o] = re; b[o + 1] = im;
        }
    for (int kz = 0; kz < nz; kz++)
      for (int ky = 0; ky < ny; ky++)
        for (int kx = 0; kx < nkx; kx++) {
          double re = 0, im = 0;
          for (int z = 0; z < nz; z++) {
            const size_t i = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * z));
            const double ang = -2.0 * M_PI * (double)((long)kz * z % nz) / nz;
            const double c = cos(ang), s = sin(ang);
            re += b[i] * c - b[i + 1] * s; im += b[i] * s + b[i + 1] * c;
          }
          const size_t o = 6 * ((size_t)kx + (size_t)nkx * (ky + (size_t)ny * kz)) + 2 * comp;
          ghat[o] = re; ghat[o + 1] = im;
        }
  }
  free(a); free(b);
}

void orc_dft3_c2r_d(int nx, int ny, int nz, int nxPad, const double *ghat, double *grid3) {
  const int nkx = nx / 2 + 1;
  const size_t ntot = (size_t)nx * ny * nz;
  double *a = (double *)calloc(ntot * 2, sizeof(double));
  double *b = (double *)calloc(ntot * 2, sizeof(double));
  for (int comp = 0; comp < 3; comp++) {
    /* rebuild the full spectrum from the stored half, like a C2R transform implicitly does
       (cuFFT ignores the imaginary parts that Hermitian symmetry forces to zero; here the stored half is
       simply mirrored, which agrees whenever the input IS Hermitian) */
    for (int kz = 0; kz < nz; kz++)
      for (int ky = 0; ky < ny; ky++)
        for (int kx = 0; kx < nx; kx++) {
          const size_t o = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * kz));
          if (kx < nkx) {
            const size_t i = 6 * ((size_t)kx + (size_t)nkx * (ky + (size_t)ny * kz)) + 2 * comp;
            a[o] = ghat[i]; a[o + 1] = ghat[i + 1];
          } else {
            const int cx = nx - kx, cy = (ny - ky) % ny, cz = (nz - kz) % nz;
            const size_t i = 6 * ((size_t)cx + (size_t)nkx * (cy + (size_t)ny * cz)) + 2 * comp;
            a[o] = ghat[i]; a[o + 1] = -ghat[i + 1];
          }
        }
    for (int z = 0; z < nz; z++)
      for (int ky = 0; ky < ny; ky++)
        for (int kx = 0; kx < nx; kx++) {
          double re = 0, im = 0;
          for (int kz = 0; kz < nz; kz++) {
            const size_t i = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * kz));
            const double ang = 2.0 * M_PI * (double)((long)kz * z % nz) / nz;
            const double c = cos(ang), s = sin(ang);
            re += a[i] * c - a[i + 1] * s; im += a[i] * s + a[i + 1] * c;
          }
          const size_t o = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * z));
          b[o] = re; b[o + 1] = im;
        }
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int kx = 0; kx < nx; kx++) {
          double re = 0, im = 0;
          for (int ky = 0; ky < ny; ky++) {
            const size_t i = 2 * ((size_t)kx + (size_t)nx * (ky + (size_t)ny * z));
            const double ang = 2.0 * M_PI * (double)((long)ky * y % ny) / ny;
            const double c = cos(ang), s = sin(ang);
            re += b[i] * c - b[i + 1] * s; im += b[i] * s + b[i + 1] * c;
          }
          const size_t o = 2 * ((size_t)kx + (size_t)nx * (y + (size_t)ny * z));
          a[o] = re; a[o + 1] = im;
        }
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          double re = 0;
          for (int kx = 0; kx < nx; kx++) {
            const size_t i = 2 * ((size_t)kx + (size_t)nx * (y + (size_t)ny * z));
            const double ang = 2.0 * M_PI * (double)((long)kx * x % nx) / nx;
            re += a[i] * cos(ang) - a[i + 1] * sin(ang);
          }
          grid3[3 * ((size_t)x + (size_t)nxPad * (y + (size_t)ny * z)) + comp] = re;
        }
  }
  free(a); free(b);
}

/* ---------- Brownian noise in Fourier space ---------- */
/* fcm_detail::isNyquistWaveNumber Integrator/BDHI/FCM/utils.cuh:132-168 */
static int is_nyquist(const int c[3], const int n[3]) {
  const int xq = (c[0] == n[0] - c[0]) && (n[0] % 2 == 0);
  const int yq = (c[1] == n[1] - c[1]) && (n[1] % 2 == 0);
  const int zq = (c[2] == n[2] - c[2]) && (n[2] % 2 == 0);
  return (xq && c[1] == 0 && c[2] == 0) || (xq && yq && c[2] == 0) || (c[0] == 0 && yq && c[2] == 0) ||
         (xq && c[1] == 0 && zq) || (c[0] == 0 && c[1] == 0 && zq) || (c[0] == 0 && yq && zq) || (xq && yq && zq);
}

static void noise_term(const orc_grid_d *g, double viscosity, const int cell[3], const double nz[6], int conj,
                       double *dst) {
  const int n[3] = {g->cellDim[0], g->cellDim[1], g->cellDim[2]};
  double k[3], dk[3], k2 = 0;
  for (int d = 0; d < 3; d++) {
    const int ik = fold(cell[d], n[d]);
    k[d] = (2.0 * M_PI / g->L[d]) * ik;
    dk[d] = (ik == n[d] - ik) ? 0.0 : k[d];
    k2 += k[d] * k[d];
  }
  const double Bsq = sqrt(1.0 / (k2 * viscosity)), invk2 = 1.0 / k2;
  for (int c = 0; c < 2; c++) {
    const double sgn = (c == 1 && conj) ? -1.0 : 1.0;
    const double f[3] = {sgn * nz[0 + c] * Bsq, sgn * nz[2 + c] * Bsq, sgn * nz[4 + c] * Bsq};
    const double fdk = f[0] * (dk[0] * invk2) + f[1] * (dk[1] * invk2) + f[2] * (dk[2] * invk2);
    for (int d = 0; d < 3; d++) dst[2 * d + c] += f[d] - dk[d] * fdk;
  }
}

/* fcm_detail::fourierBrownianNoise Integrator/BDHI/FCM/FCM_impl.cuh:437-512, executed node by node in index order
   (the reference kernel's two non-atomic "+=" per node race on the kx = nx/2 plane; this is the race-free sum).
   noisePrefactor = prefactor*sqrt(2T/(dV Nxyz)) as computed by addBrownianNoise :514-542. */
void orc_fcm_add_noise_d(const orc_grid_d *g, double viscosity, double noisePrefactor, uint32_t seed1, uint32_t seed2,
                         double *ghat) {
  const int n[3] = {g->cellDim[0], g->cellDim[1], g->cellDim[2]};
  const int nkx = n[0] / 2 + 1;
  for (int iz = 0; iz < n[2]; iz++)
    for (int iy = 0; iy < n[1]; iy++)
      for (int ix = 0; ix < nkx; ix++) {
        const uint32_t id = (uint32_t)ix + (uint32_t)nkx * ((uint32_t)iy + (uint32_t)n[1] * iz);
        const int cell[3] = {ix, iy, iz};
        if (id == 0 || (ix == 0 && iy == 0 && 2 * iz >= n[2] + 1) || (ix == 0 && 2 * iy >= n[1] + 1)) continue;
        /* generateNoise (FCM/utils.cuh:115-130): Saru(id, seed1, seed2), three float Box-Muller pairs */
        orc_saru rng = orc_saru_seed3(id, seed1, seed2);
        const float sc = (float)(0.707106781186547 * noisePrefactor);
        double nzv[6];
        for (int c = 0; c < 3; c++) {
          float pr[2];
          orc_saru_gf(&rng, 0.0f, sc, pr);
          nzv[2 * c] = pr[0];
          nzv[2 * c + 1] = pr[1];
        }
        const int nyq = is_nyquist(cell, n);
        if (nyq)
          for (int c = 0; c < 3; c++) { nzv[2 * c] *= 1.41421356237310; nzv[2 * c + 1] = 0.0; }
        noise_term(g, viscosity, cell, nzv, 0, ghat + 6 * (size_t)id);
        if (nyq) continue;
        if (ix == n[0] - ix || ix == 0) {
          const int cc[3] = {ix, (iy > 0) * (n[1] - iy), (iz > 0) * (n[2] - iz)};
          const size_t idc = (size_t)cc[0] + (size_t)nkx * ((size_t)cc[1] + (size_t)n[1] * cc[2]);
          noise_term(g, viscosity, cc, nzv, 1, ghat + 6 * idc);
        }
      }
}
