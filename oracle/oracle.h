/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement ("oracle") of the UAMMD hot paths that uammd_b200 replaces. Plain C, no CUDA.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under uammd_b200/ links, imports or executes it.
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it follows.
 * Pinning status (see DESIGN.md "Oracle pinning"):
 *   - path 1 (cell list, LJ, DPD, NVE): the reference's tests hold NO golden vectors for this path
 *     (SURVEY.md F6); the oracle is pinned instead against the compiled, unmodified reference run
 *     on the GPU box (oracle/_ref/ref_lj, ref_dpd) by tests/test_ref_parity_gpu.py, and against the
 *     two-particle KAT "F = -/+24 at r = sigma" (examples/uammd_as_a_library/wrapper.py:32).
 *   - path 2 (IBM spread/gather, FCM): pinned against the reference's own KATs restated in
 *     tests/ (Hasimoto self mobility test/BDHI/FCM/fcm_test.cu:85-144; Peskin spread/gather vs
 *     manual loops test/misc/ibm/test_ibm_regular.cu:113-136,240-274) and against oracle/_ref/ref_fcm.
 */
#ifndef UAMMD_B200_ORACLE_H
#define UAMMD_B200_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Box + Grid in single precision. utils/Box.cuh:16-36, utils/Grid.cuh:21-48 */
typedef struct {
  float L[3];
  float minusInvL[3]; /* 0 for a non periodic (or zero/infinite) dimension */
  int cellDim[3];
  float cellSize[3];
  float invCellSize[3];
} orc_grid_f;

typedef struct {
  double L[3];
  double minusInvL[3];
  int cellDim[3];
  double cellSize[3];
  double invCellSize[3];
} orc_grid_d;

void orc_grid_init_f(orc_grid_f *g, const float L[3], const int periodic[3], const int cellDim[3]);
void orc_grid_init_d(orc_grid_d *g, const double L[3], const int periodic[3], const int cellDim[3]);
/* CellList::createUpdateGrid  Interactor/NeighbourList/CellList.cuh:100-126 */
void orc_neighbour_celldim_f(const float L[3], float rc, int cellDim[3]);
void orc_get_cell_f(const orc_grid_f *g, const float *p, int cell[3]);
void orc_get_cell_d(const orc_grid_d *g, const double *p, int cell[3]);
uint32_t orc_morton_hash(int cx, int cy, int cz);

/* cell list build: sortPos float4[N], index int[N], cellStart/cellEnd int[ncells] (-1 when empty) */
int orc_celllist_build_f(const orc_grid_f *g, const float *pos4, int N, float *sortPos4, int *index,
                         int *cellStart, int *cellEnd);

/* LJ pair forces, reference summation order, fp32 arithmetic with fmaf accumulation */
void orc_lj_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                const int *cellEnd, int N, const float *params4, int ntypes, float *force4, float *energy,
                float *virial);
/* same pair set, all arithmetic in fp64 from the fp32 inputs ("truth"); abssum = sum_j |f_ij| per particle,
   sens = sum_j |d f_ij / d r_ij| (how far a perturbation of the separations moves the force),
   edge = sum of |f_ij| over pairs with |r2 - rc2| <= band (pairs that fp32 rounding can move across the cut-off) */
void orc_lj_f64(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                const int *cellEnd, int N, const float *params4, int ntypes, double *force3, double *energy,
                double *virial, double *abssum, double *sens, double band, double *edge);

/* DPD pair forces (Interactor/Potential/DPD.cuh:121-158) */
void orc_dpd_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                 const int *cellEnd, int N, const float *vel3, float A, float gamma, float sigma, float rcut,
                 uint32_t seed, uint32_t step, int use_double_acc, float *force4, double *force3d);
void orc_dpd_ids_f32(const orc_grid_f *g, const float *sortPos4, const int *index, const int *cellStart,
                     const int *cellEnd, int N, const float *vel3, float A, float gamma, float sigma, float rcut,
                     uint32_t seed, uint32_t step, int use_double_acc, float *force4, double *force3d,
                     const int *noiseId, int idStride);
/* brick domain decomposition (uammd_b200/csrc/domain.cu): owner rank and ghost-destination mask per particle */
void orc_brick_classify_f(const orc_grid_f *g, const float *pos4, int N, const int rankGrid[3], int *cell, int *owner,
                          uint32_t *ghostMask);
/* Saru known-answer helpers */
void orc_saru3_u32(uint32_t s1, uint32_t s2, uint32_t s3, int n, uint32_t *out);
void orc_saru3_gf(uint32_t s1, uint32_t s2, uint32_t s3, float mean, float std, float out[2]);

/* velocity Verlet half kicks (Integrator/VerletNVE.cu:64-85) */
void orc_nve_half_f32(float *pos4, float *vel3, const float *force4, int N, float dt, float mass, int step);
/* One full VerletNVE::forwardTime (VerletNVE.cu:174-188) with a PairForces<LJ,CellList>: returns 0 on success */
int orc_md_step_f32(const float L[3], float rc, const float *params4, float dt, int N, float *pos4, float *vel3,
                    float *force4, void *scratch);
void *orc_md_scratch_new(int N, int ncells);
void orc_md_scratch_free(void *);

/* VerletNVT::GronbechJensen half steps (Integrator/VerletNVT/GronbechJensen.cu:30-66) and Basic_ns::initialVelocities
   (Integrator/VerletNVT/Basic.cu:12-29), single precision */
void orc_nvt_gj_half_f32(float *pos4, float *vel3, float *force4, const float *mass, float defaultMass, int N, float dt,
                         float friction, int is2D, float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step);
/* VerletNVT::Basic_ns::integrateGPU<step> Integrator/VerletNVT/Basic.cu:87-117 */
void orc_nvt_basic_half_f32(float *pos4, float *vel3, float *force4, const float *mass, float defaultMass, int N, int Ngroup,
                            float dt, float friction, int is2D, float noiseAmplitude, uint32_t stepNum, uint32_t seed, int step);
void orc_nvt_initial_velocities_f32(float *vel3, int N, float vamp, int is2D, uint32_t seed);

/* ---------------- path 2: IBM + FCM (fp64) ---------------- */
/* kernel ids */
#define ORC_KERNEL_PESKIN3 0
#define ORC_KERNEL_PESKIN4 1
#define ORC_KERNEL_GAUSSIAN 2
#define ORC_KERNEL_BARNETT_MAGLAND 3 /* prefactor = 1/norm, tau = beta, rmax = alpha */
#define ORC_KERNEL_SIXPOINT 4
typedef struct {
  int kind;
  int support;
  double h;       /* Peskin: cell size */
  double prefactor, tau, rmax; /* Gaussian: FCM_kernels.cuh:22-58 */
} orc_ibm_kernel;

double orc_ibm_phi(const orc_ibm_kernel *k, double r);
/* BarnettMagland::computeNorm misc/IBM_kernels.cuh:93-97 (Simpson over [0, alpha], 20000 intervals, Kahan sums) */
double orc_ibm_bm_norm(double alpha, double beta);
/* IBM spread: misc/IBM.cu:83-147. grid is real3 AoS with row pitch nxPad (nxPad = 2(nx/2+1) for FCM) */
void orc_ibm_spread_d(const orc_grid_d *g, const orc_ibm_kernel *k, const double *pos4, const double *val3,
                      int N, int nxPad, double *grid3);
/* IBM gather: misc/IBM.cu:168-235 with DefaultQuadratureWeights = cell volume */
void orc_ibm_gather_d(const orc_grid_d *g, const orc_ibm_kernel *k, const double *pos4, int N, int nxPad,
                      const double *grid3, double *out3);
/* FCM spectral step FCM_impl.cuh:375-397 on complex3 AoS [(nx/2+1)*ny*nz][3][2] */
void orc_fcm_force2vel_d(const orc_grid_d *g, double viscosity, double *ghat);
/* FCM Brownian noise added to the Fourier grid, FCM_impl.cuh:437-542 */
void orc_fcm_add_noise_d(const orc_grid_d *g, double viscosity, double noisePrefactor, uint32_t seed1, uint32_t seed2,
                         double *ghat);
/* naive-but-exact separable 3D real-to-complex / complex-to-real DFT (unnormalised, cuFFT sign convention)
   on the interleaved-3 layout. For small grids only (O(n^4)). */
void orc_dft3_r2c_d(int nx, int ny, int nz, int nxPad, const double *grid3, double *ghat);
void orc_dft3_c2r_d(int nx, int ny, int nz, int nxPad, const double *ghat, double *grid3);

/* ---------------- BDHI::PSE (fp64) ---------------- */
double orc_pse_greens_d(const double k[3], double shear, double rh, double viscosity, double split, double eta, double ntot);
void orc_pse_force2vel_d(const orc_grid_d *g, double shear, double rh, double viscosity, double split, double eta,
                         double *ghat);
void orc_pse_add_noise_d(const orc_grid_d *g, double shear, double rh, double viscosity, double split, double eta,
                         double noisePrefactor, uint32_t seed1, uint32_t seed2, double *ghat);
void orc_rpy_near_fg(double r, double rh, double psi, double rcut, double out[2]);
void orc_pse_near_table_d(int nPoints, double rh, double psi, double normalization, double rcut, double *table2);
void orc_pse_near_mdot_d(int N, const double *pos4, const double *v, int vStride, const double L[3], double shear,
                         double rh, double psi, double normalization, double rcut, const double *table2, int nPoints,
                         int useTable, double *out3);

#ifdef __cplusplus
}
#endif
#endif
