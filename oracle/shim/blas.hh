// oracle/shim/blas.hh -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal stand-in for the BLAS++ v2023.01.00 C++ interface (un-vendored dependency of the reference,
// pinned in /root/reference/cmake/ImportBlasPP.cmake:25-31) so that the reference's CPU path can be
// compiled UNMODIFIED from /root/reference into oracle/_ref/.  Only the surface the reference uses is
// provided (grep over src/ include/ examples/: gemm, trmm, trsm, syrk, scal, conj, real_type,
// is_complex, Gflop, Queue, op2str and the five char enums).  All arithmetic is forwarded to the
// Fortran BLAS inside the scipy wheel's OpenBLAS (symbols prefixed scipy_, LP64).
#pragma once
#ifdef HCB_SHIM_CUDA_QUEUE
#include <cuda_runtime.h>
#endif
#include <complex>
#include <cstdint>
#include <cstddef>
#include <algorithm>
#include <stdexcept>
#include <type_traits>
// BLAS++ pulls these in transitively and the reference relies on that (e.g. ContextManager.hpp uses std::vector).
#include <vector>
#include <string>
#include <limits>
#include <cmath>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#define HCB_F77(name) scipy_##name##_

extern "C" {
void HCB_F77(dgemm)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                    const int*, const double*, const int*, const double*, double*, const int*);
void HCB_F77(sgemm)(const char*, const char*, const int*, const int*, const int*, const float*, const float*,
                    const int*, const float*, const int*, const float*, float*, const int*);
void HCB_F77(dtrmm)(const char*, const char*, const char*, const char*, const int*, const int*, const double*,
                    const double*, const int*, double*, const int*);
void HCB_F77(strmm)(const char*, const char*, const char*, const char*, const int*, const int*, const float*,
                    const float*, const int*, float*, const int*);
void HCB_F77(dtrsm)(const char*, const char*, const char*, const char*, const int*, const int*, const double*,
                    const double*, const int*, double*, const int*);
void HCB_F77(strsm)(const char*, const char*, const char*, const char*, const int*, const int*, const float*,
                    const float*, const int*, float*, const int*);
void HCB_F77(dsyrk)(const char*, const char*, const int*, const int*, const double*, const double*, const int*,
                    const double*, double*, const int*);
void HCB_F77(ssyrk)(const char*, const char*, const int*, const int*, const float*, const float*, const int*,
                    const float*, float*, const int*);
void HCB_F77(dscal)(const int*, const double*, double*, const int*);
void HCB_F77(sscal)(const int*, const float*, float*, const int*);
}

namespace blas {

enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
enum class Uplo : char { Upper = 'U', Lower = 'L', General = 'G' };
enum class Diag : char { NonUnit = 'N', Unit = 'U' };
enum class Side : char { Left = 'L', Right = 'R' };

inline const char* op2str(Op op) {
    return op == Op::NoTrans ? "notrans" : (op == Op::Trans ? "trans" : "conj");
}

template<typename T> struct real_type_traits { using type = T; };
template<typename T> struct real_type_traits<std::complex<T>> { using type = T; };
template<typename T> using real_type = typename real_type_traits<T>::type;

template<typename T> struct is_complex : std::false_type {};
template<typename T> struct is_complex<std::complex<T>> : std::true_type {};

inline float conj(float x) { return x; }
inline double conj(double x) { return x; }

#ifdef HCB_SHIM_CUDA_QUEUE
// drop-in build of the reference's -DUSE_CUDA operator layer (tests/dropin): the queue is what BLAS++ makes it there,
// a CUDA stream with sync(); no cuBLAS handle behind it (nothing on the new path uses one)
class Queue {
public:
    Queue() : Queue(0, 0) {}
    Queue(int device, int64_t) {
        cudaSetDevice(device);
        cudaStreamCreateWithFlags(&mStream, cudaStreamNonBlocking);
    }
    Queue(const Queue &) = delete;
    ~Queue() { if (mStream) cudaStreamDestroy(mStream); }
    cudaStream_t stream() const { return mStream; }
    void sync() { cudaStreamSynchronize(mStream); }
private:
    cudaStream_t mStream = nullptr;
};
#else
class Queue {
public:
    Queue() = default;
    Queue(int, int64_t) {}
    void sync() {}
};
#endif

namespace detail {
    inline char flip_uplo(Uplo u) { return u == Uplo::Upper ? 'L' : (u == Uplo::Lower ? 'U' : 'G'); }
    inline char flip_side(Side s) { return s == Side::Left ? 'R' : 'L'; }
    inline char opc(Op o) { return o == Op::ConjTrans ? 'T' : (char) o; }  // real types only
    inline char flip_op(Op o) { return o == Op::NoTrans ? 'T' : 'N'; }
}

inline void gemm(Layout layout, Op ta, Op tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    char a = detail::opc(ta), b = detail::opc(tb);
    int m_ = (int) m, n_ = (int) n, k_ = (int) k, lda_ = (int) lda, ldb_ = (int) ldb, ldc_ = (int) ldc;
    if (layout == Layout::ColMajor) {
        HCB_F77(dgemm)(&a, &b, &m_, &n_, &k_, &alpha, A, &lda_, B, &ldb_, &beta, C, &ldc_);
    } else {  // C^T = op(B)^T op(A)^T in column-major terms
        HCB_F77(dgemm)(&b, &a, &n_, &m_, &k_, &alpha, B, &ldb_, A, &lda_, &beta, C, &ldc_);
    }
}

inline void gemm(Layout layout, Op ta, Op tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc) {
    char a = detail::opc(ta), b = detail::opc(tb);
    int m_ = (int) m, n_ = (int) n, k_ = (int) k, lda_ = (int) lda, ldb_ = (int) ldb, ldc_ = (int) ldc;
    if (layout == Layout::ColMajor) {
        HCB_F77(sgemm)(&a, &b, &m_, &n_, &k_, &alpha, A, &lda_, B, &ldb_, &beta, C, &ldc_);
    } else {
        HCB_F77(sgemm)(&b, &a, &n_, &m_, &k_, &alpha, B, &ldb_, A, &lda_, &beta, C, &ldc_);
    }
}

#define HCB_TRXM(NAME, FD, FS)                                                                                   \
    inline void NAME(Layout layout, Side side, Uplo uplo, Op trans, Diag diag, int64_t m, int64_t n,             \
                     double alpha, const double* A, int64_t lda, double* B, int64_t ldb) {                       \
        char s = (char) side, u = (char) uplo, t = detail::opc(trans), d = (char) diag;                          \
        int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, ldb_ = (int) ldb;                                      \
        if (layout == Layout::RowMajor) {                                                                        \
            s = detail::flip_side(side); u = detail::flip_uplo(uplo); std::swap(m_, n_);                         \
        }                                                                                                        \
        HCB_F77(FD)(&s, &u, &t, &d, &m_, &n_, &alpha, A, &lda_, B, &ldb_);                                       \
    }                                                                                                            \
    inline void NAME(Layout layout, Side side, Uplo uplo, Op trans, Diag diag, int64_t m, int64_t n,             \
                     float alpha, const float* A, int64_t lda, float* B, int64_t ldb) {                          \
        char s = (char) side, u = (char) uplo, t = detail::opc(trans), d = (char) diag;                          \
        int m_ = (int) m, n_ = (int) n, lda_ = (int) lda, ldb_ = (int) ldb;                                      \
        if (layout == Layout::RowMajor) {                                                                        \
            s = detail::flip_side(side); u = detail::flip_uplo(uplo); std::swap(m_, n_);                         \
        }                                                                                                        \
        HCB_F77(FS)(&s, &u, &t, &d, &m_, &n_, &alpha, A, &lda_, B, &ldb_);                                       \
    }
HCB_TRXM(trmm, dtrmm, strmm)
HCB_TRXM(trsm, dtrsm, strsm)
#undef HCB_TRXM

template<typename T>
inline void syrk(Layout layout, Uplo uplo, Op trans, int64_t n, int64_t k, T alpha, const T* A, int64_t lda,
                 T beta, T* C, int64_t ldc) {
    char u = (char) uplo, t = detail::opc(trans);
    int n_ = (int) n, k_ = (int) k, lda_ = (int) lda, ldc_ = (int) ldc;
    if (layout == Layout::RowMajor) { u = detail::flip_uplo(uplo); t = detail::flip_op(trans); }
    if constexpr (std::is_same<T, double>::value) HCB_F77(dsyrk)(&u, &t, &n_, &k_, &alpha, A, &lda_, &beta, C, &ldc_);
    else HCB_F77(ssyrk)(&u, &t, &n_, &k_, &alpha, A, &lda_, &beta, C, &ldc_);
}

inline void scal(int64_t n, double alpha, double* x, int64_t incx) {
    int n_ = (int) n, inc = (int) incx; HCB_F77(dscal)(&n_, &alpha, x, &inc);
}
inline void scal(int64_t n, float alpha, float* x, int64_t incx) {
    int n_ = (int) n, inc = (int) incx; HCB_F77(sscal)(&n_, &alpha, x, &inc);
}

template<typename T>
struct Gflop {
    static double gemm(double m, double n, double k) { return 2.0 * m * n * k * 1e-9; }
    static double potrf(double n) { return (n * n * n / 3.0 + n * n / 2.0 + n / 6.0) * 1e-9; }
    static double trsm(Side, double m, double n) { return m * m * n * 1e-9; }
    static double syrk(double n, double k) { return k * n * (n + 1) * 1e-9; }
};

}  // namespace blas
