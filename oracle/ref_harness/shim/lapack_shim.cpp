/* TEST INFRASTRUCTURE - implementations behind shim/cblas.h and shim/lapacke.h (see those headers).
 * steqr('I'): eigenvalues (ascending, like LAPACK) and eigenvectors of a symmetric tridiagonal matrix by cyclic
 * Jacobi rotations on the dense matrix in long double - deliberately a different algorithm from the product's
 * QL solver (uammd_b200/csrc/pse.cu), so that the two check each other. */
#include "cblas.h"
#include "lapacke.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

template <class T> static int steqr(int layout, char compz, int n, T *d, T *e, T *z, int ldz) {
  if (compz != 'I' && compz != 'i') return -2;
  std::vector<long double> A((size_t)n * n, 0.0L), V((size_t)n * n, 0.0L);
  for (int i = 0; i < n; i++) {
    A[(size_t)i * n + i] = d[i];
    V[(size_t)i * n + i] = 1.0L;
    if (i + 1 < n) A[(size_t)i * n + i + 1] = A[(size_t)(i + 1) * n + i] = e[i];
  }
  for (int sweep = 0; sweep < 100; sweep++) {
    long double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
    if (off < 1e-60L) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const long double apq = A[(size_t)p * n + q];
        if (fabsl(apq) < 1e-4000L) continue;
        const long double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0L * apq);
        const long double t = (theta >= 0 ? 1.0L : -1.0L) / (fabsl(theta) + sqrtl(theta * theta + 1.0L));
        const long double c = 1.0L / sqrtl(t * t + 1.0L), s = t * c;
        for (int k = 0; k < n; k++) {
          const long double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const long double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const long double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  for (int j = 0; j < n; j++) {
    const int src = order[j];
    d[j] = (T)A[(size_t)src * n + src];
    for (int r = 0; r < n; r++) { // eigenvector j, component r
      if (layout == LAPACK_COL_MAJOR) z[(size_t)j * ldz + r] = (T)V[(size_t)r * n + src];
      else z[(size_t)r * ldz + j] = (T)V[(size_t)r * n + src];
    }
  }
  return 0;
}

template <class T>
static void gemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int M, int N, T alpha, const T *A, int lda, const T *X, int incX,
                 T beta, T *Y, int incY) {
  const bool rowOfA = (order == CblasColMajor) == (trans == CblasNoTrans); // y_r = sum_c op(A)(r,c) x_c
  const int rows = trans == CblasNoTrans ? M : N, cols = trans == CblasNoTrans ? N : M;
  std::vector<long double> acc(rows, 0.0L);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      const T a = rowOfA ? A[(size_t)c * lda + r] : A[(size_t)r * lda + c];
      acc[r] += (long double)a * X[(size_t)c * incX];
    }
  for (int r = 0; r < rows; r++) Y[(size_t)r * incY] = (T)(alpha * acc[r] + (beta == T(0) ? 0.0L : (long double)beta * Y[(size_t)r * incY]));
}

extern "C" {
lapack_int LAPACKE_ssteqr(int layout, char compz, lapack_int n, float *d, float *e, float *z, lapack_int ldz) { return steqr(layout, compz, n, d, e, z, ldz); }
lapack_int LAPACKE_dsteqr(int layout, char compz, lapack_int n, double *d, double *e, double *z, lapack_int ldz) { return steqr(layout, compz, n, d, e, z, ldz); }
void cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int M, int N, float alpha, const float *A, int lda, const float *X, int incX, float beta, float *Y, int incY) { gemv(order, trans, M, N, alpha, A, lda, X, incX, beta, Y, incY); }
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int M, int N, double alpha, const double *A, int lda, const double *X, int incX, double beta, double *Y, int incY) { gemv(order, trans, M, N, alpha, A, lda, X, incX, beta, Y, incY); }
}
