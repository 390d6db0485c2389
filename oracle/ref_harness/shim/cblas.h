/* TEST INFRASTRUCTURE - minimal stand-in for <cblas.h> so that the UNMODIFIED reference Lanczos
 * (misc/lapack_and_blas_defines.h:6-7, LanczosAlgorithm.cu:77) compiles in a container without CBLAS.
 * Only gemv is used by the reference. Implemented in lapack_shim.cpp. */
#ifndef UB200_SHIM_CBLAS_H
#define UB200_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int M, int N, float alpha, const float *A, int lda,
                 const float *X, int incX, float beta, float *Y, int incY);
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int M, int N, double alpha, const double *A, int lda,
                 const double *X, int incX, double beta, double *Y, int incY);
#ifdef __cplusplus
}
#endif
#endif
