/* oracle/shim/cblas.h -- TEST INFRASTRUCTURE ONLY.
 * The three CBLAS entry points /root/reference/src/helpers/RawMatrix.cpp:133-201 uses, mapped to the scipy-wheel
 * OpenBLAS (prefix scipy_). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
double scipy_cblas_dnrm2(int n, const double* x, int incx);
void scipy_cblas_dscal(int n, double alpha, double* x, int incx);
void scipy_cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, double alpha,
                       const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
#ifdef __cplusplus
}
#endif
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_dscal scipy_cblas_dscal
#define cblas_dgemv scipy_cblas_dgemv
