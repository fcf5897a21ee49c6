// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin C ABI over the UNMODIFIED reference CPU path (compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libhcorepp_ref.so).  Nothing under hcorepp_b200/ may link or load this library; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it (as checker / CPU baseline).
//
// Every entry point calls straight into the reference's own classes:
//   tiles      -> hcorepp::operators::{DenseTile,CompressedTile}   (include/hcorepp/operators/concrete/*.hpp)
//   gemm       -> hcorepp::api::HCore<T>::Gemm                     (src/api/HCore.cpp:22-344)
//   matmul     -> the tile loop of examples/matrix_multiplication/omp_main.cpp:112-126
//   kernels    -> hcorepp::kernels::HCoreKernels<T>                (src/kernels/omp/kernels.cpp)
//   generators -> helpers::generators::{Latms,TileLatms}Generator, matrixhelpers::generate_dense_matrix
#include <hcorepp/api/HCore.hpp>
#include <hcorepp/operators/concrete/Dense.hpp>
#include <hcorepp/operators/concrete/Compressed.hpp>
#include <hcorepp/kernels/kernels.hpp>
#include <hcorepp/kernels/ContextManager.hpp>
#include <hcorepp/data-units/memory-handlers/MemoryHandler.hpp>
#include <hcorepp/helpers/MatrixHelpers.hpp>
#include <hcorepp/helpers/generators/concrete/LatmsGenerator.hpp>
#include <hcorepp/helpers/generators/concrete/TileLatmsGenerator.hpp>

#include <omp.h>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <vector>

using namespace hcorepp;
using namespace hcorepp::operators;
using hcorepp::kernels::RunContext;

extern "C" void scipy_openblas_set_num_threads(int);

namespace {

RunContext &ctx() { return kernels::ContextManager::GetInstance().GetContext(); }

// ContextManager::GetInstance() is not thread-safe (src/kernels/ContextManager.cpp:13-20): create it at load time.
__attribute__((constructor)) void hcref_init() { (void) ctx(); }

CompressionParameters make_params(double acc, int use_trmm, int use_ungqr, int trunc, int64_t fixed_rank, int svd) {
    return CompressionParameters(acc, use_trmm != 0, use_ungqr != 0, trunc != 0, (size_t) fixed_rank,
                                 svd == 0 ? common::CompressionType::LAPACK_GESVD
                                          : common::CompressionType::LAPACK_GESDD);
}

template<typename T>
void *tile_dense(int64_t m, int64_t n, const T *data, int64_t ld) {
    return new DenseTile<T>(m, n, const_cast<T *>(data), ld, blas::Layout::ColMajor, ctx());
}

template<typename T>
void *tile_uv(int64_t m, int64_t n, const T *U, const T *V, int64_t rank) {
    return new CompressedTile<T>(m, n, const_cast<T *>(U), const_cast<T *>(V), m, rank, blas::Layout::ColMajor,
                                 ctx());
}

// A tile in exactly the state the compressing constructor leaves it in (capacity max_rank, current rank `rank`,
// V stored with ld = rank at offset m*max_rank; Compressed.cpp:83-87,132-141) without paying for an SVD: build it at
// full capacity through the (m, n, UV, ld, rank) constructor and shrink the rank with the public ReadjustTileRank().
template<typename T>
void *tile_uv_cap(int64_t m, int64_t n, const T *U, const T *V, int64_t rank, int64_t max_rank) {
    std::vector<T> uv((size_t) (m + n) * max_rank, T(0));
    std::memcpy(uv.data(), U, sizeof(T) * m * rank);
    std::memcpy(uv.data() + (size_t) m * max_rank, V, sizeof(T) * rank * n);
    auto *t = new CompressedTile<T>(m, n, uv.data(), m, max_rank, blas::Layout::ColMajor, ctx());
    t->ReadjustTileRank(rank, ctx());
    return t;
}

template<typename T>
void *tile_compress(int64_t m, int64_t n, const T *data, int64_t ld, const CompressionParameters &p) {
    return new CompressedTile<T>(m, n, const_cast<T *>(data), ld, p, blas::Layout::ColMajor, ctx());
}

template<typename T>
void tile_info(void *h, int64_t *out) {
    auto *t = static_cast<Tile<T> *>(h);
    out[0] = (int64_t) t->GetNumOfRows();
    out[1] = (int64_t) t->GetNumOfCols();
    out[2] = t->isDense() ? 0 : (int64_t) t->GetTileRank();
    out[3] = t->isDense() ? 1 : 0;
    if (t->isCompressed()) {
        auto *c = static_cast<CompressedTile<T> *>(h);
        // capacity actually reserved after the U block: V starts at data + m*maxRank (Compressed.cpp:180-185)
        out[4] = (int64_t) ((c->GetVMatrix() - c->GetUMatrix()) / (ptrdiff_t) c->GetNumOfRows());
        out[5] = (int64_t) c->GetVLeadingDim();
    } else {
        out[4] = 0;
        out[5] = (int64_t) t->GetDataHolder().get().GetLeadingDim();
    }
}

template<typename T>
void tile_read(void *h, T *u_or_dense, T *v) {
    auto *t = static_cast<Tile<T> *>(h);
    if (t->isDense()) {
        auto &dh = t->GetDataHolder().get();
        size_t m = t->GetNumOfRows(), n = t->GetNumOfCols(), ld = dh.GetLeadingDim();
        for (size_t j = 0; j < n; ++j) std::memcpy(u_or_dense + j * m, dh.GetData() + j * ld, m * sizeof(T));
    } else {
        auto *c = static_cast<CompressedTile<T> *>(h);
        size_t m = c->GetNumOfRows(), n = c->GetNumOfCols(), rk = c->GetTileRank();
        std::memcpy(u_or_dense, c->GetUMatrix(), m * rk * sizeof(T));
        std::memcpy(v, c->GetVMatrix(), rk * n * sizeof(T));
    }
}

template<typename T>
int tile_gemm(T alpha, void *a, int opa, void *b, int opb, T beta, void *c, const CompressionParameters &p,
              int64_t *flops_out) {
    try {
        dataunits::MemoryUnit<T> unit(ctx());
        size_t flops = 0;
        api::HCore<T>::Gemm(alpha, *static_cast<Tile<T> *>(a), opa ? blas::Op::Trans : blas::Op::NoTrans,
                            *static_cast<Tile<T> *>(b), opb ? blas::Op::Trans : blas::Op::NoTrans, beta,
                            *static_cast<Tile<T> *>(c), ctx(), flops, unit, p);
        if (flops_out) *flops_out = (int64_t) flops;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "hcref gemm: %s\n", e.what());
        return 1;
    }
}

// TLR Cholesky pieces (HCore.cpp:482-647).  The reference has no driver and no enabled tests for
// them, so these wrappers are the only pin for that row.  uplo: 'L' / 'U'; side 'L' / 'R'; diag 'N' / 'U'.
template<typename T>
int tile_potrf(void *a, int uplo) {
    try {
        dataunits::MemoryUnit<T> unit(ctx());
        size_t flops = 0;
        api::HCore<T>::Potrf(*static_cast<Tile<T> *>(a), (blas::Uplo) uplo, ctx(), flops, unit);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "hcref potrf: %s\n", e.what());
        return 1;
    }
}
template<typename T>
int tile_trsm(int side, int uplo, int trans, int diag, T alpha, void *a, void *b) {
    try {
        dataunits::MemoryUnit<T> unit(ctx());
        size_t flops = 0;
        api::HCore<T>::Trsm((blas::Side) side, (blas::Uplo) uplo, trans ? blas::Op::Trans : blas::Op::NoTrans, (blas::Diag) diag,
                            alpha, *static_cast<Tile<T> *>(a), *static_cast<Tile<T> *>(b), ctx(), flops, unit);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "hcref trsm: %s\n", e.what());
        return 1;
    }
}
template<typename T>
int tile_syrk(T alpha, void *a, int opa, int uplo, T beta, void *c) {
    try {
        dataunits::MemoryUnit<T> unit(ctx());
        size_t flops = 0;
        api::HCore<T>::Syrk(alpha, *static_cast<Tile<T> *>(a), opa ? blas::Op::Trans : blas::Op::NoTrans, (blas::Uplo) uplo, beta,
                            *static_cast<Tile<T> *>(c), ctx(), flops, unit);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "hcref syrk: %s\n", e.what());
        return 1;
    }
}

// The tile loop of examples/matrix_multiplication/omp_main.cpp:112-126: C(j,i) += A(j,k) B(k,i) for k = 0..kt-1,
// OMP-parallel over the independent C tiles, one MemoryUnit per thread. Tile arrays are column-major grids:
// A[j + k*mt], B[k + i*kt], C[j + i*mt].
template<typename T>
int tile_matmul(int64_t mt, int64_t nt, int64_t kt, void **A, void **B, void **C, T alpha, T beta,
                const CompressionParameters &p, int nthreads, double *seconds, int64_t *flops_out) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    scipy_openblas_set_num_threads(1);
    size_t flops = 0;
    int failed = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for collapse(2) num_threads(nthreads) reduction(+:flops) reduction(|:failed) schedule(dynamic, 1)
    for (int64_t i = 0; i < nt; i++) {
        for (int64_t j = 0; j < mt; j++) {
            try {
                dataunits::MemoryUnit<T> unit(ctx());
                auto *c_tile = static_cast<Tile<T> *>(C[j + i * mt]);
                for (int64_t k = 0; k < kt; k++) {
                    auto *a_tile = static_cast<Tile<T> *>(A[j + k * mt]);
                    auto *b_tile = static_cast<Tile<T> *>(B[k + i * kt]);
                    api::HCore<T>::Gemm(alpha, *a_tile, blas::Op::NoTrans, *b_tile, blas::Op::NoTrans, beta, *c_tile,
                                        ctx(), flops, unit, p);
                }
            } catch (const std::exception &) {
                failed |= 1;
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (flops_out) *flops_out = (int64_t) flops;
    return failed;
}

template<typename T>
void latms_law(int64_t m, int64_t n, int64_t tile_size, int64_t *seed, T *out, int64_t ld, int reps) {
    // One generator object, `reps` consecutive draws into out (each m*ld... caller strides): the examples draw A then
    // B from the same generator so the seed state carries over (omp_main.cpp:232-242).
    helpers::generators::Generator<T> *g;
    if (tile_size > 0) g = new helpers::generators::TileLatmsGenerator<T>(seed, 0, 1, (size_t) tile_size);
    else g = new helpers::generators::LatmsGenerator<T>(seed, 0, 1);
    for (int r = 0; r < reps; ++r) g->GenerateValues(m, n, ld, out + (size_t) r * ld * n);
    delete g;
}

}  // namespace

#define HCREF_API(P, T)                                                                                              \
    extern "C" void *hcref_##P##tile_dense(int64_t m, int64_t n, const T *d, int64_t ld) {                           \
        return tile_dense<T>(m, n, d, ld);                                                                           \
    }                                                                                                                \
    extern "C" void *hcref_##P##tile_uv(int64_t m, int64_t n, const T *U, const T *V, int64_t rank) {                \
        return tile_uv<T>(m, n, U, V, rank);                                                                         \
    }                                                                                                                \
    extern "C" void *hcref_##P##tile_uv_cap(int64_t m, int64_t n, const T *U, const T *V, int64_t rank,              \
                                            int64_t max_rank) {                                                      \
        return tile_uv_cap<T>(m, n, U, V, rank, max_rank);                                                           \
    }                                                                                                                \
    extern "C" void *hcref_##P##tile_compress(int64_t m, int64_t n, const T *d, int64_t ld, double acc,              \
                                              int use_trmm, int use_ungqr, int trunc, int64_t fixed_rank,            \
                                              int svd) {                                                             \
        return tile_compress<T>(m, n, d, ld, make_params(acc, use_trmm, use_ungqr, trunc, fixed_rank, svd));         \
    }                                                                                                                \
    extern "C" void hcref_##P##tile_info(void *h, int64_t *out6) { tile_info<T>(h, out6); }                          \
    extern "C" void hcref_##P##tile_read(void *h, T *u, T *v) { tile_read<T>(h, u, v); }                             \
    extern "C" void hcref_##P##tile_free(void *h) { delete static_cast<Tile<T> *>(h); }                              \
    extern "C" int hcref_##P##gemm(T alpha, void *a, int opa, void *b, int opb, T beta, void *c, double acc,         \
                                   int use_trmm, int use_ungqr, int trunc, int64_t fixed_rank, int svd,              \
                                   int64_t *flops) {                                                                 \
        return tile_gemm<T>(alpha, a, opa, b, opb, beta, c,                                                          \
                            make_params(acc, use_trmm, use_ungqr, trunc, fixed_rank, svd), flops);                   \
    }                                                                                                                \
    extern "C" int hcref_##P##potrf(void *a, int uplo) { return tile_potrf<T>(a, uplo); }                            \
    extern "C" int hcref_##P##trsm(int side, int uplo, int trans, int diag, T alpha, void *a, void *b) {             \
        return tile_trsm<T>(side, uplo, trans, diag, alpha, a, b);                                                   \
    }                                                                                                                \
    extern "C" int hcref_##P##syrk(T alpha, void *a, int opa, int uplo, T beta, void *c) {                           \
        return tile_syrk<T>(alpha, a, opa, uplo, beta, c);                                                           \
    }                                                                                                                \
    extern "C" int hcref_##P##matmul(int64_t mt, int64_t nt, int64_t kt, void **A, void **B, void **C, T alpha,      \
                                     T beta, double acc, int use_trmm, int use_ungqr, int trunc,                     \
                                     int64_t fixed_rank, int svd, int nthreads, double *seconds, int64_t *flops) {   \
        return tile_matmul<T>(mt, nt, kt, A, B, C, alpha, beta,                                                      \
                              make_params(acc, use_trmm, use_ungqr, trunc, fixed_rank, svd), nthreads, seconds,      \
                              flops);                                                                                \
    }                                                                                                                \
    extern "C" void hcref_##P##latms_law(int64_t m, int64_t n, int64_t tile_size, int64_t *seed4, T *out,            \
                                         int64_t ld, int reps) {                                                     \
        latms_law<T>(m, n, tile_size, seed4, out, ld, reps);                                                         \
    }                                                                                                                \
    extern "C" void hcref_##P##generate_dense(int64_t m, int64_t n, T *A, int64_t lda, int64_t *seed4) {             \
        helpers::matrixhelpers::generate_dense_matrix<T>(m, n, A, lda, seed4, 0, 1);                                 \
    }                                                                                                                \
    extern "C" int64_t hcref_##P##compress_dense(int64_t m, int64_t n, const T *A, int64_t lda, double acc,          \
                                                 T *UV_out) {                                                        \
        T *uv = nullptr;                                                                                             \
        int64_t rk = 0;                                                                                              \
        helpers::matrixhelpers::compress_dense_matrix<T>(m, n, A, lda, &uv, rk, (T) acc);                            \
        if (UV_out) std::memcpy(UV_out, uv, (size_t) (lda + n) * rk * sizeof(T));                                    \
        std::free(uv);                                                                                               \
        return rk;                                                                                                   \
    }                                                                                                                \
    /* ---- kernel table (src/kernels/omp/kernels.cpp), 1:1 ---- */                                                  \
    extern "C" void hcref_##P##k_gemm(int ta, int tb, int64_t m, int64_t n, int64_t k, T alpha, const T *A,          \
                                      int64_t lda, const T *B, int64_t ldb, T beta, T *C, int64_t ldc) {             \
        kernels::HCoreKernels<T>::Gemm(blas::Layout::ColMajor, ta ? blas::Op::Trans : blas::Op::NoTrans,             \
                                       tb ? blas::Op::Trans : blas::Op::NoTrans, m, n, k, alpha, A, lda, B, ldb,     \
                                       beta, C, ldc, ctx());                                                         \
    }                                                                                                                \
    extern "C" void hcref_##P##k_multiply_by_alpha(T *arr, int64_t rows, int64_t cols, int64_t m, int64_t rank,      \
                                                   T alpha) {                                                        \
        kernels::HCoreKernels<T>::MultiplyByAlpha(arr, rows, cols, m, rank, alpha, ctx());                           \
    }                                                                                                                \
    extern "C" void hcref_##P##k_process_v(int64_t n, int64_t crank, int ungqr, int64_t vm, T beta, T *cv,           \
                                           int64_t ldcv, T *V, int64_t arank, const T *b, int cholesky) {            \
        kernels::HCoreKernels<T>::ProcessVpointer(n, crank, ungqr != 0, vm, beta, cv, ldcv, V, arank, b, ctx(),      \
                                                  cholesky != 0);                                                    \
    }                                                                                                                \
    extern "C" int64_t hcref_##P##k_new_rank(int trunc, T *sigma, int64_t size_s, T acc) {                           \
        size_t rk = 0;                                                                                               \
        kernels::HCoreKernels<T>::CalculateNewRank(rk, trunc != 0, sigma, size_s, acc, ctx());                       \
        return (int64_t) rk;                                                                                         \
    }                                                                                                                \
    extern "C" void hcref_##P##k_uvptr(int64_t rank, int64_t vm, T *uvptr, const T *vnew) {                          \
        kernels::HCoreKernels<T>::CalculateUVptr(rank, vm, uvptr, vnew, ctx());                                      \
    }                                                                                                                \
    extern "C" void hcref_##P##k_vtnew(int64_t rk, int ungqr, int64_t min_vm_vn, T *sigma, T *vtnew,                 \
                                       int64_t size_s, int64_t vm) {                                                 \
        kernels::HCoreKernels<T>::CalculateVTnew(rk, ungqr != 0, min_vm_vn, sigma, vtnew, size_s, vm, ctx());        \
    }                                                                                                                \
    extern "C" void hcref_##P##k_fill_identity(int64_t n, T *a) {                                                    \
        kernels::HCoreKernels<T>::FillIdentityMatrix(n, a, ctx());                                                   \
    }                                                                                                                \
    extern "C" void hcref_##P##k_lacpy(int type, int64_t m, int64_t n, T *A, int64_t lda, T *B, int64_t ldb) {       \
        kernels::HCoreKernels<T>::LaCpy((common::MatrixType) type, m, n, A, lda, B, ldb, ctx());                     \
    }                                                                                                                \
    extern "C" void hcref_##P##k_laset(int type, int64_t m, int64_t n, T off, T diag, T *A, int64_t lda) {           \
        kernels::HCoreKernels<T>::Laset((common::MatrixType) type, m, n, off, diag, A, lda, ctx());                  \
    }                                                                                                                \
    extern "C" void hcref_##P##k_geqrf(int64_t m, int64_t n, T *A, int64_t lda, T *tau) {                            \
        kernels::HCoreKernels<T>::Geqrf(m, n, A, lda, tau, nullptr, 0, 0, ctx());                                    \
    }                                                                                                                \
    extern "C" void hcref_##P##k_ungqr(int64_t m, int64_t n, int64_t k, T *A, int64_t lda, T *tau) {                 \
        kernels::HCoreKernels<T>::ungqr(m, n, k, A, lda, tau, nullptr, 0, ctx());                                    \
    }                                                                                                                \
    extern "C" void hcref_##P##k_unmqr(int side, int trans, int64_t m, int64_t n, int64_t k, const T *A,             \
                                       int64_t lda, const T *tau, T *C, int64_t ldc) {                               \
        kernels::HCoreKernels<T>::Unmqr((common::SideMode) side, (common::BlasOperation) trans, m, n, k, A, lda,     \
                                        tau, C, ldc, nullptr, 0, ctx());                                             \
    }                                                                                                                \
    extern "C" void hcref_##P##k_svd(int64_t m, int64_t n, T *A, int64_t lda, T *S, T *U, int64_t ldu, T *VT,        \
                                     int64_t ldvt, int svd) {                                                        \
        kernels::HCoreKernels<T>::SVD(common::Job::SomeVec, common::Job::SomeVec, m, n, A, lda, S, U, ldu, VT, ldvt, \
                                      svd == 0 ? common::CompressionType::LAPACK_GESVD                               \
                                               : common::CompressionType::LAPACK_GESDD,                              \
                                      nullptr, 0, 0, ctx());                                                         \
    }                                                                                                                \
    extern "C" void hcref_##P##k_trmm(int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha,        \
                                      const T *A, int64_t lda, T *B, int64_t ldb) {                                  \
        kernels::HCoreKernels<T>::Trmm(blas::Layout::ColMajor, (blas::Side) side, (blas::Uplo) uplo,                 \
                                       (blas::Op) trans, (blas::Diag) diag, m, n, alpha, A, lda, B, ldb, ctx());     \
    }

HCREF_API(d, double)
HCREF_API(s, float)

extern "C" void hcref_set_blas_threads(int n) { scipy_openblas_set_num_threads(n); }
extern "C" int hcref_max_threads() { return omp_get_max_threads(); }
extern "C" const char *hcref_version() { return "hcorepp reference CPU path (unmodified sources) + oracle/shim"; }
