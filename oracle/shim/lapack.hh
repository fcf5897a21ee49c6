// oracle/shim/lapack.hh -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal stand-in for the LAPACK++ v2023.01.00 C++ interface (un-vendored dependency of the reference, pinned in
// /root/reference/cmake/ImportLapackPP.cmake:23-29).  Only what /root/reference/src/kernels/omp/kernels.cpp uses:
// geqrf, ungqr, unmqr, gesvd, gesdd, lacpy, laset, potrf + enums Job / MatrixType and the Op/Uplo/Side aliases.
// Workspace queries are done internally, as LAPACK++ does.  Arithmetic = scipy-wheel OpenBLAS LAPACK (LP64).
#pragma once
#include "blas.hh"
#include <vector>

extern "C" {
#define HCB_DECL2(RET, name, ...) RET scipy_d##name##_(__VA_ARGS__);
void scipy_dgeqrf_(const int*, const int*, double*, const int*, double*, double*, const int*, int*);
void scipy_sgeqrf_(const int*, const int*, float*, const int*, float*, float*, const int*, int*);
void scipy_dorgqr_(const int*, const int*, const int*, double*, const int*, const double*, double*, const int*, int*);
void scipy_sorgqr_(const int*, const int*, const int*, float*, const int*, const float*, float*, const int*, int*);
void scipy_dormqr_(const char*, const char*, const int*, const int*, const int*, const double*, const int*,
                   const double*, double*, const int*, double*, const int*, int*);
void scipy_sormqr_(const char*, const char*, const int*, const int*, const int*, const float*, const int*,
                   const float*, float*, const int*, float*, const int*, int*);
void scipy_dgesvd_(const char*, const char*, const int*, const int*, double*, const int*, double*, double*,
                   const int*, double*, const int*, double*, const int*, int*);
void scipy_sgesvd_(const char*, const char*, const int*, const int*, float*, const int*, float*, float*,
                   const int*, float*, const int*, float*, const int*, int*);
void scipy_dgesdd_(const char*, const int*, const int*, double*, const int*, double*, double*, const int*, double*,
                   const int*, double*, const int*, int*, int*);
void scipy_sgesdd_(const char*, const int*, const int*, float*, const int*, float*, float*, const int*, float*,
                   const int*, float*, const int*, int*, int*);
void scipy_dlacpy_(const char*, const int*, const int*, const double*, const int*, double*, const int*);
void scipy_slacpy_(const char*, const int*, const int*, const float*, const int*, float*, const int*);
void scipy_dlaset_(const char*, const int*, const int*, const double*, const double*, double*, const int*);
void scipy_slaset_(const char*, const int*, const int*, const float*, const float*, float*, const int*);
void scipy_dpotrf_(const char*, const int*, double*, const int*, int*);
void scipy_spotrf_(const char*, const int*, float*, const int*, int*);
#undef HCB_DECL2
}

namespace lapack {

using Op = blas::Op;
using Uplo = blas::Uplo;
using Side = blas::Side;

enum class Job : char {
    NoVec = 'N', Vec = 'V', UpdateVec = 'U', AllVec = 'A', SomeVec = 'S', OverwriteVec = 'O',
    CompactVec = 'P', SomeVecTol = 'C', VecJacobi = 'J', Workspace = 'W'
};
enum class MatrixType : char {
    General = 'G', Lower = 'L', Upper = 'U', Hessenberg = 'H', LowerBand = 'B', UpperBand = 'Q', Band = 'Z'
};

#define HCB_REAL_DISPATCH(T, dfn, sfn, ...)                                    \
    do { if constexpr (std::is_same<T, double>::value) dfn(__VA_ARGS__);       \
         else sfn(__VA_ARGS__); } while (0)

template<typename T>
int64_t geqrf(int64_t m, int64_t n, T* A, int64_t lda, T* tau) {
    int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, info = 0, lwork = -1;
    T q;
    HCB_REAL_DISPATCH(T, scipy_dgeqrf_, scipy_sgeqrf_, &m_, &n_, A, &lda_, tau, &q, &lwork, &info);
    lwork = std::max(1, (int) q);
    std::vector<T> work(lwork);
    HCB_REAL_DISPATCH(T, scipy_dgeqrf_, scipy_sgeqrf_, &m_, &n_, A, &lda_, tau, work.data(), &lwork, &info);
    return info;
}

template<typename T>
int64_t ungqr(int64_t m, int64_t n, int64_t k, T* A, int64_t lda, const T* tau) {
    int m_ = (int) m, n_ = (int) n, k_ = (int) k, lda_ = (int) lda, info = 0, lwork = -1;
    T q;
    HCB_REAL_DISPATCH(T, scipy_dorgqr_, scipy_sorgqr_, &m_, &n_, &k_, A, &lda_, tau, &q, &lwork, &info);
    lwork = std::max(1, (int) q);
    std::vector<T> work(lwork);
    HCB_REAL_DISPATCH(T, scipy_dorgqr_, scipy_sorgqr_, &m_, &n_, &k_, A, &lda_, tau, work.data(), &lwork, &info);
    return info;
}

template<typename T>
int64_t unmqr(Side side, Op trans, int64_t m, int64_t n, int64_t k, const T* A, int64_t lda, const T* tau, T* C,
              int64_t ldc) {
    char s = (char) side, t = (trans == Op::NoTrans) ? 'N' : 'T';
    int m_ = (int) m, n_ = (int) n, k_ = (int) k, lda_ = (int) lda, ldc_ = (int) ldc, info = 0, lwork = -1;
    T q;
    HCB_REAL_DISPATCH(T, scipy_dormqr_, scipy_sormqr_, &s, &t, &m_, &n_, &k_, A, &lda_, tau, C, &ldc_, &q, &lwork,
                      &info);
    lwork = std::max(1, (int) q);
    std::vector<T> work(lwork);
    HCB_REAL_DISPATCH(T, scipy_dormqr_, scipy_sormqr_, &s, &t, &m_, &n_, &k_, A, &lda_, tau, C, &ldc_, work.data(),
                      &lwork, &info);
    return info;
}

template<typename T>
int64_t gesvd(Job jobu, Job jobvt, int64_t m, int64_t n, T* A, int64_t lda, T* S, T* U, int64_t ldu, T* VT,
              int64_t ldvt) {
    char ju = (char) jobu, jv = (char) jobvt;
    int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, ldu_ = (int) ldu, ldvt_ = (int) ldvt, info = 0, lwork = -1;
    T q;
    HCB_REAL_DISPATCH(T, scipy_dgesvd_, scipy_sgesvd_, &ju, &jv, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_, &q,
                      &lwork, &info);
    lwork = std::max(1, (int) q);
    std::vector<T> work(lwork);
    HCB_REAL_DISPATCH(T, scipy_dgesvd_, scipy_sgesvd_, &ju, &jv, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_,
                      work.data(), &lwork, &info);
    return info;
}

template<typename T>
int64_t gesdd(Job jobz, int64_t m, int64_t n, T* A, int64_t lda, T* S, T* U, int64_t ldu, T* VT, int64_t ldvt) {
    char jz = (char) jobz;
    int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, ldu_ = (int) ldu, ldvt_ = (int) ldvt, info = 0, lwork = -1;
    std::vector<int> iwork(8 * std::max<int64_t>(1, std::min(m, n)));
    T q;
    HCB_REAL_DISPATCH(T, scipy_dgesdd_, scipy_sgesdd_, &jz, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_, &q, &lwork,
                      iwork.data(), &info);
    lwork = std::max(1, (int) q);
    std::vector<T> work(lwork);
    HCB_REAL_DISPATCH(T, scipy_dgesdd_, scipy_sgesdd_, &jz, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_,
                      work.data(), &lwork, iwork.data(), &info);
    return info;
}

template<typename T>
void lacpy(MatrixType type, int64_t m, int64_t n, const T* A, int64_t lda, T* B, int64_t ldb) {
    char t = (char) type;
    int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, ldb_ = (int) ldb;
    HCB_REAL_DISPATCH(T, scipy_dlacpy_, scipy_slacpy_, &t, &m_, &n_, A, &lda_, B, &ldb_);
}

template<typename T>
void laset(MatrixType type, int64_t m, int64_t n, T offdiag, T diag, T* A, int64_t lda) {
    char t = (char) type;
    int m_ = (int) m, n_ = (int) n, lda_ = (int) lda;
    HCB_REAL_DISPATCH(T, scipy_dlaset_, scipy_slaset_, &t, &m_, &n_, &offdiag, &diag, A, &lda_);
}

template<typename T>
int64_t potrf(Uplo uplo, int64_t n, T* A, int64_t lda) {
    char u = (char) uplo;
    int n_ = (int) n, lda_ = (int) lda, info = 0;
    HCB_REAL_DISPATCH(T, scipy_dpotrf_, scipy_spotrf_, &u, &n_, A, &lda_, &info);
    return info;
}

#undef HCB_REAL_DISPATCH
}  // namespace lapack
