/* oracle/shim/lapack/fortran.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for LAPACK++'s <lapack/fortran.h> as used by
 * /root/reference/include/hcorepp/helpers/LapackWrappers.hpp:20-25 and src/helpers/RawMatrix.cpp:111-116.
 * Maps LAPACK_x macros onto the scipy-wheel OpenBLAS Fortran symbols (prefix scipy_, LP64). */
#pragma once
#include <complex>
typedef int lapack_int;
typedef std::complex<float> lapack_complex_float;
typedef std::complex<double> lapack_complex_double;

#ifdef __cplusplus
extern "C" {
#endif
void scipy_dlatms_(const int*, const int*, const char*, int*, const char*, double*, const int*, const double*,
                   const double*, const int*, const int*, const char*, double*, const int*, double*, int*);
void scipy_slatms_(const int*, const int*, const char*, int*, const char*, float*, const int*, const float*,
                   const float*, const int*, const int*, const char*, float*, const int*, float*, int*);
void scipy_zlatms_(const int*, const int*, const char*, int*, const char*, double*, const int*, const double*,
                   const double*, const int*, const int*, const char*, lapack_complex_double*, const int*,
                   lapack_complex_double*, int*);
void scipy_clatms_(const int*, const int*, const char*, int*, const char*, float*, const int*, const float*,
                   const float*, const int*, const int*, const char*, lapack_complex_float*, const int*,
                   lapack_complex_float*, int*);
void scipy_dlacpy_(const char*, const int*, const int*, const double*, const int*, double*, const int*);
void scipy_slacpy_(const char*, const int*, const int*, const float*, const int*, float*, const int*);
void scipy_zlacpy_(const char*, const int*, const int*, const lapack_complex_double*, const int*,
                   lapack_complex_double*, const int*);
void scipy_clacpy_(const char*, const int*, const int*, const lapack_complex_float*, const int*,
                   lapack_complex_float*, const int*);
void scipy_dgesvd_(const char*, const char*, const int*, const int*, double*, const int*, double*, double*,
                   const int*, double*, const int*, double*, const int*, int*);
void scipy_sgesvd_(const char*, const char*, const int*, const int*, float*, const int*, float*, float*,
                   const int*, float*, const int*, float*, const int*, int*);
/* complex gesvd has an extra rwork argument that the reference wrapper never passes; the complex branches are
 * discarded `if constexpr` arms (only float/double are instantiated) but must still parse, hence variadic. */
void scipy_zgesvd_(...);
void scipy_cgesvd_(...);
double scipy_dlange_(const char*, const int*, const int*, const double*, const int*, double*);
float scipy_slange_(const char*, const int*, const int*, const float*, const int*, float*);
double scipy_zlange_(const char*, const int*, const int*, const lapack_complex_double*, const int*, double*);
float scipy_clange_(const char*, const int*, const int*, const lapack_complex_float*, const int*, float*);
#ifdef __cplusplus
}
#endif

#define LAPACK_dlatms scipy_dlatms_
#define LAPACK_slatms scipy_slatms_
#define LAPACK_zlatms scipy_zlatms_
#define LAPACK_clatms scipy_clatms_
#define LAPACK_dlacpy scipy_dlacpy_
#define LAPACK_slacpy scipy_slacpy_
#define LAPACK_zlacpy scipy_zlacpy_
#define LAPACK_clacpy scipy_clacpy_
#define LAPACK_dgesvd scipy_dgesvd_
#define LAPACK_sgesvd scipy_sgesvd_
#define LAPACK_zgesvd scipy_zgesvd_
#define LAPACK_cgesvd scipy_cgesvd_
#define LAPACK_dlange scipy_dlange_
#define LAPACK_slange scipy_slange_
#define LAPACK_zlange scipy_zlange_
#define LAPACK_clange scipy_clange_

/* RawMatrix.cpp:114-116 calls the bare symbol dlange_ with a trailing hidden-length argument. */
#ifdef __cplusplus
extern "C" {
#endif
static inline double dlange_(const char* norm, const lapack_int* m, const lapack_int* n, const double* a,
                             const lapack_int* lda, double* work, ...) {
    return scipy_dlange_(norm, m, n, a, lda, work);
}
#ifdef __cplusplus
}
#endif
