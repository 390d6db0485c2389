/* TEST INFRASTRUCTURE - minimal stand-in for <lapacke.h>: only LAPACKE_{s,d}steqr, which the reference Lanczos
 * calls with compz = 'I' (LanczosAlgorithm.cu:56). Implemented in lapack_shim.cpp. */
#ifndef UB200_SHIM_LAPACKE_H
#define UB200_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_ssteqr(int layout, char compz, lapack_int n, float *d, float *e, float *z, lapack_int ldz);
lapack_int LAPACKE_dsteqr(int layout, char compz, lapack_int n, double *d, double *e, double *z, lapack_int ldz);
#ifdef __cplusplus
}
#endif
#endif
