/*
 * hcore_b200.h -- C ABI of libhcore_b200.so: the B200-native (sm_100a) replacement for the reference's device
 * backend on the TLR-GEMM hot path.
 *
 * What it replaces (file:line relative to the ecrc/hcorepp reference tree):
 *   - the kernel table   hcorepp::kernels::HCoreKernels<T>      include/hcorepp/kernels/kernels.hpp:27-129,
 *     as implemented for CUDA by src/kernels/cuda/kernels.cpp + src/kernels/cuda/CudaKernels.cu
 *     (cuBLAS via BLAS++ queue, cuSOLVER, element-wise SIMT kernels);
 *   - hcorepp::memory::{AllocateArray,DestroyArray,Memcpy,Memset}  include/hcorepp/kernels/cuda/memory.hpp:15-43;
 *   - hcorepp::kernels::RunContext (CUDA)                          include/hcorepp/kernels/cuda/RunContext.hpp:15-53;
 *   - and, as ONE fused batched call, the whole per-tile flow
 *       HCore<T>::Gemm (src/api/HCore.cpp:22-344) -> CompressedTile<T>::Gemm (src/operators/concrete/Compressed.cpp:208-694)
 *     for many (A,B,C) tile triples per launch with ranks kept on the device (no per-tile host sync, which the
 *     reference pays in CudaKernels.cu:656-697), plus the compressing constructor (Compressed.cpp:75-146).
 *
 * Conventions: plain C, no exceptions, every function returns 0 on success or an HCB_E* code
 * (hcb_last_error() gives the text).  All matrices are column-major.  Pointers named d_* are DEVICE pointers.
 * Everything is enqueued on the context's stream; nothing synchronises unless documented.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with HCB_ENODEVICE.
 */
#ifndef HCORE_B200_H
#define HCORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCB_OK 0
#define HCB_EINVAL 1    /* bad argument */
#define HCB_ECUDA 2     /* CUDA runtime error (text in hcb_last_error) */
#define HCB_ENODEVICE 3 /* no CUDA device: there is no CPU fallback */
#define HCB_ENOMEM 4
#define HCB_EUNSUPPORTED 5
#define HCB_EBOUND 6    /* hcb_ctx_sync: a rank exceeded its rank_bound in a fused call (that tile was left untouched) */

typedef struct hcb_ctx hcb_ctx;

/* ---- context / memory : RunContext + hcorepp::memory (cuda/RunContext.hpp:15-53, cuda/memory.hpp:15-43) -------- */
int hcb_ctx_create(int device, hcb_ctx **out);                       /* owns a new non-blocking stream */
int hcb_ctx_create_on_stream(int device, void *cuda_stream, hcb_ctx **out); /* borrows the caller's stream */
int hcb_ctx_destroy(hcb_ctx *ctx);
int hcb_ctx_sync(hcb_ctx *ctx);                                      /* RunContext::Sync() */
void *hcb_ctx_stream(hcb_ctx *ctx);                                  /* RunContext::GetStream() */
int hcb_ctx_device(hcb_ctx *ctx);
int hcb_ctx_sm_count(hcb_ctx *ctx);
/* grow-only scratch arena (replaces MemoryUnit pool, pool/MemoryHandler.hpp:21-157); never memset per call */
int hcb_ctx_reserve_workspace(hcb_ctx *ctx, size_t bytes);
size_t hcb_ctx_workspace_bytes(hcb_ctx *ctx);
/* Counters of the adaptive fast paths of the fused recompression since the context was created (or last reset); the call
 * synchronises.  out8[0] new-column panels factored by CholeskyQR2, [1] / [2] panels that fell back to the Householder path in
 * its first / second pass, [3] block Gram-Schmidt second passes skipped by the twice-is-enough test, [4] second passes run, [5] negligible new
 * columns (norm below 1e-13 of the largest new column) deflated to zero by the CholeskyQR2 path, [6] r x r graded triangular
 * factors of the incremental V side computed by Cholesky, [7] tiles that fell back to the Householder R-only QR for it. */
int hcb_ctx_stats(hcb_ctx *ctx, uint64_t *out8, int reset);
int hcb_malloc(hcb_ctx *ctx, size_t bytes, void **d_out);            /* memory::AllocateArray */
int hcb_free(hcb_ctx *ctx, void *d_ptr);                             /* memory::DestroyArray */
/* kind: 0 H2D, 1 D2D, 2 D2H, 3 H2H, 4 automatic (memory::MemoryTransfer, memory.hpp:22-44); async on the stream */
int hcb_memcpy(hcb_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);
int hcb_memset(hcb_ctx *ctx, void *d_dst, int value, size_t bytes);  /* memory::Memset */
/* Per-phase device timing of the fused batched path, measured with CUDA events recorded on the context's stream
 * (the stream the kernels are launched on). Phases: 0 setup, 1 contraction GEMMs, 2 stack assembly, 3 panel QR,
 * 4 core build + LQ preconditioning, 5 apply-Q rebuild, 6 finalize, 7 Jacobi SVD, 8 V*Sigma GEMM + truncation.
 * hcb_ctx_phase_times synchronises, adds the elapsed milliseconds / launch counts since the last call into
 * ms[HCB_N_PHASES] / launches[HCB_N_PHASES] and clears the records. */
#define HCB_N_PHASES 9
int hcb_ctx_phase_timing(hcb_ctx *ctx, int enable);
int hcb_ctx_phase_times(hcb_ctx *ctx, double *ms, uint64_t *launches);
const char *hcb_phase_name(int phase);
const char *hcb_last_error(void);
const char *hcb_version(void);
/* number of kernel launches issued by this library since the last reset (bench.py's gpu_launches claim) */
uint64_t hcb_launch_count(void);
void hcb_launch_count_reset(void);

/* ---- compression parameters : operators::CompressionParameters (CompressionParameters.hpp:44-46) -------------- */
typedef struct hcb_compress_params {
    double accuracy;      /* default 1e-4 */
    int32_t use_trmm;     /* accepted; the fused path always forms RU*RV^T (same product) */
    int32_t use_ungqr;    /* accepted; Q is always applied implicitly (same product) */
    int32_t truncated_svd;/* 1: threshold relative to sigma_0 (omp/kernels.cpp:87-95) */
    int64_t fixed_rank;   /* 0 = truncate by accuracy */
    int32_t svd_type;     /* 0 GESVD / 1 GESDD in the reference; one-sided Jacobi here either way */
    int32_t reserved;
} hcb_compress_params;

/* ---- tiles : operators::TileMetadata + data buffer (Tile.hpp:30-52) ------------------------------------------ */
#define HCB_TILE_DENSE 0
#define HCB_TILE_COMPRESSED 1
typedef struct hcb_tile {
    int32_t type;       /* HCB_TILE_DENSE | HCB_TILE_COMPRESSED */
    int32_t m, n;       /* rows, cols */
    int32_t ld;         /* dense: leading dimension; compressed: ignored (ldU = m, ldV = rank, Compressed.cpp:29-34) */
    int32_t max_rank;   /* compressed: capacity; V starts at d_data + m*max_rank (Compressed.cpp:180-185) */
    int32_t rank_bound; /* host-known upper bound on *d_rank used to size grids/scratch (0 -> max_rank) */
    int32_t *d_rank;    /* compressed: DEVICE int32 holding the current rank (the device owns rank truth) */
    void *d_data;       /* DEVICE buffer: dense m x n (ld) or [U (m x max_rank) | V (max_rank x n)] */
    /* compressed, optional (NULL = unknown): DEVICE int32 of state bits the library keeps next to the rank.
     * bit 0 (HCB_STATE_ORTHO_U): "U has orthonormal columns" -- set by the library at the end of every recompression /
     * compression (Compressed.cpp:558-560 guarantees it: U = Q_U * Unew), and what lets the next update orthogonalise
     * only the NEW columns against U instead of re-factoring the whole stack.  Whoever writes the factors through
     * d_data directly must clear it (the host mirrors do so in every constructor / PackTile / load).
     * bits 8..: number of consecutive incremental updates (the library re-factors the whole stack every 16th time). */
    int32_t *d_state;
    /* compressed C tiles: per-tile fixed rank for the replay drivers (par_fixed_rank_streams_main.cpp:465-477,540-541;
     * Compressed.cpp:510-515): > 0 overrides hcb_compress_params.fixed_rank for this tile, 0 = use the batch value */
    int32_t fixed_rank;
    int32_t reserved;
} hcb_tile;
#define HCB_STATE_ORTHO_U 1
#define HCB_STATE_ORTHO_V 2  /* the rows of V are mutually orthogonal (V = diag(sigma) W^T, Compressed.cpp:598-622) */

/* ---- (2) fine-grained kernel table : one symbol per HCoreKernels<T> entry (kernels.hpp:27-129) --------------- */
/* trans: 0 NoTrans, 1 Trans.  type for lacpy/laset: 'G','U','L' (common::MatrixType).  side 'L'/'R'. */
#define HCB_DECLARE_KERNEL_TABLE(P, T)                                                                                  \
    int hcb_##P##gemm(hcb_ctx *, int transA, int transB, int64_t m, int64_t n, int64_t k, T alpha, const T *dA,           \
                      int64_t lda, const T *dB, int64_t ldb, T beta, T *dC, int64_t ldc);  /* kernels.hpp:27-29 */      \
    int hcb_##P##multiply_by_alpha(hcb_ctx *, T *dArr, int64_t rows, int64_t cols, int64_t m, int64_t rank, T alpha);   \
    int hcb_##P##process_v(hcb_ctx *, int64_t n, int64_t crank, int ungqr, int64_t vm, T beta, const T *dCV,            \
                           int64_t ldcv, T *dV, int64_t arank, const T *dB, int cholesky);  /* kernels.hpp:36-38 */     \
    int hcb_##P##new_rank(hcb_ctx *, int truncated, const T *dSigma, int64_t size_s, T accuracy,                       \
                          int64_t *host_rank); /* SYNCS (host result, like CudaKernels.cu:656-697) */                  \
    int hcb_##P##new_rank_device(hcb_ctx *, int truncated, const T *dSigma, int64_t size_s, T accuracy,                \
                                 int32_t *d_rank); /* async variant: rank stays on the device */                       \
    int hcb_##P##uvptr(hcb_ctx *, int64_t rank, int64_t vm, T *dUV, const T *dVnew);       /* CalculateUVptr */        \
    int hcb_##P##vtnew(hcb_ctx *, int64_t rk, int ungqr, int64_t min_vm_vn, const T *dSigma, T *dVT, int64_t size_s,   \
                       int64_t vm);                                                        /* CalculateVTnew */        \
    int hcb_##P##uvptr_conj(hcb_ctx *, int64_t rank, int64_t vm, T *dUV);                  /* no-op for real T */      \
    int hcb_##P##fill_identity(hcb_ctx *, int64_t n, T *dA);                               /* FillIdentityMatrix */    \
    int hcb_##P##lacpy(hcb_ctx *, int type, int64_t m, int64_t n, const T *dA, int64_t lda, T *dB, int64_t ldb);        \
    int hcb_##P##laset(hcb_ctx *, int type, int64_t m, int64_t n, T offdiag, T diag, T *dA, int64_t lda);               \
    int hcb_##P##geqrf(hcb_ctx *, int64_t m, int64_t n, T *dA, int64_t lda, T *dTau);      /* LAPACK layout */         \
    int hcb_##P##ungqr(hcb_ctx *, int64_t m, int64_t n, int64_t k, T *dA, int64_t lda, const T *dTau);                  \
    int hcb_##P##unmqr(hcb_ctx *, int side, int trans, int64_t m, int64_t n, int64_t k, const T *dA, int64_t lda,       \
                       const T *dTau, T *dC, int64_t ldc);                                                             \
    int hcb_##P##svd(hcb_ctx *, int64_t m, int64_t n, T *dA, int64_t lda, T *dS, T *dU, int64_t ldu, T *dVT,            \
                     int64_t ldvt); /* SomeVec/SomeVec; dA destroyed */                                                \
    int hcb_##P##trmm(hcb_ctx *, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *dA,   \
                      int64_t lda, T *dB, int64_t ldb);                                                                \
    /* ---- TLR Cholesky pieces of the kernel table (kernels.hpp potrf / trsm / syrk / FillMatrixTriangle /          */  \
    /* Symmetrize; src/kernels/omp/kernels.cpp:234-303).  uplo 'L'/'U', side 'L'/'R', trans 0/'N' or 1/'T', diag      */  \
    /* 'N'/'U'.  potrf: in place, the other triangle is left as it was; d_info (device int32, may be NULL) receives  */  \
    /* LAPACK's info (0, or 1-based index of the first non-positive pivot).                                          */  \
    int hcb_##P##potrf(hcb_ctx *, int uplo, int64_t n, T *dA, int64_t lda, int32_t *d_info);                            \
    int hcb_##P##trsm(hcb_ctx *, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *dA,   \
                      int64_t lda, T *dB, int64_t ldb);                                                                \
    int hcb_##P##syrk(hcb_ctx *, int uplo, int trans, int64_t n, int64_t k, T alpha, const T *dA, int64_t lda, T beta,   \
                      T *dC, int64_t ldc);                                                                             \
    int hcb_##P##fill_triangle(hcb_ctx *, int uplo, int64_t n, T *dA, int64_t lda, T value); /* strict triangle */      \
    int hcb_##P##symmetrize(hcb_ctx *, int uplo, int64_t n, T *dA, int64_t lda);           /* copy uplo onto the other */ \
    /* HCoreKernels<T>::transpose (omp/kernels.cpp:305-333): Out (cols x rows, ldo) = A (rows x cols, lda)^T           */  \
    int hcb_##P##transpose(hcb_ctx *, int64_t rows, int64_t cols, const T *dA, int64_t lda, T *dOut, int64_t ldo);      \
    /* Batched tile forms for the Cholesky driver (HOST arrays of descriptors / device pointers):                    */  \
    /*   tlr_trsm_batched: X[t] := X[t] * L[t]^-T for compressed X (acts on the V factor only; HCore<T>::Trsm,        */  \
    /*                     HCore.cpp:624-647, in this library's V = rank x n convention)                             */  \
    /*   tlr_syrk_batched: C[t] := beta*C[t] + alpha*A[t]*A[t]^T, A compressed, C dense m x m (HCore<T>::Syrk with a  */  \
    /*                     compressed operand, HCore.cpp:484-575; both triangles are updated)                        */  \
    /*   tlr_potrf:        right-looking tile Cholesky of an SPD matrix: nt dense nb x nb diagonal tiles (diag[k], ldd)*/  \
    /*                     + compressed tiles below the diagonal (low: nt x nt column-major grid, entries i > j):     */  \
    /*                     potrf / trsm panel / syrk / ONE batched recompressing GEMM (opB = Trans) per step.  The    */  \
    /*                     reference has no such driver.  d_potrf_info: device int32[nt] or NULL;      */  \
    /*                     d_info: device int32[nt*nt] or NULL, flags of the recompressing updates, sticky over the  */  \
    /*                     steps, indexed by the tile's position in its step's batch (diagnostics).                  */  \
    int hcb_##P##tlr_trsm_batched(hcb_ctx *, int64_t n_tiles, const hcb_tile *X, const T *const *dL, const int64_t *ldl); \
    int hcb_##P##tlr_syrk_batched(hcb_ctx *, int64_t n_tiles, const hcb_tile *A, T *const *dC, const int64_t *ldc,      \
                                  T alpha, T beta);                                                                    \
    int hcb_##P##tlr_potrf(hcb_ctx *, int64_t nt, int64_t nb, T *const *diag, int64_t ldd, const hcb_tile *low,         \
                           const hcb_compress_params *p, int32_t *d_info, int32_t *d_potrf_info);                      \
    /* ---- (3) fused batched fast path -------------------------------------------------------------------------- */  \
    /* C[t] = alpha*op(A[t])*op(B[t]) + beta*C[t] for t < n_tiles, all eight Dense/Compressed operand mixes of      */  \
    /* HCore<T>::Gemm (HCore.cpp:36-313; the batch must be mix-homogeneous), recompression included, ranks on the   */  \
    /* device.  A,B,C are HOST arrays of descriptors.  d_info (device int32[n_tiles], may be NULL): low byte flags  */  \
    /* 0 ok | 1 Jacobi not converged | 2 rank clipped to max_rank | 4 a rank exceeded its rank_bound (tile left     */  \
    /* untouched -- also reported by the next hcb_ctx_sync as HCB_EBOUND); bits 8..15 = Jacobi sweeps used.  The    */  \
    /* call zeroes d_info first; flags are OR-ed, the sweep count is a maximum.  hcb_?tlr_matmul zeroes it once and   */  \
    /* keeps it sticky over its k loop.  Asynchronous.  The triples of one call may have DIFFERENT Dense/Compressed  */  \
    /* mixes: they are partitioned by mix and every group runs as one fused call (d_info in the caller's order).      */  \
    int hcb_##P##tlr_gemm_batched(hcb_ctx *, int64_t n_tiles, const hcb_tile *A, int opA, const hcb_tile *B, int opB,   \
                                  const hcb_tile *C, T alpha, T beta, const hcb_compress_params *p, int32_t *d_info);   \
    /* Compressing constructor, batched (Compressed.cpp:75-146): dense tile t (m x n, ld) -> out[t] (U, V, *d_rank) */  \
    /* fp64 tiles >= 288 on both sides and <= 2048 rows: sketched (range finder + small SVD), tiles whose spectrum   */  \
    /* is too flat for the sketch are detected on the device and redone with the full SVD (one stream sync per call). */  \
    /* d_info (device int32[n_tiles], may be NULL): zeroed, then |= 1 Jacobi not converged, |= 2 rank clipped at    */  \
    /* max_rank (silent in the reference, Compressed.cpp:117-119); bits 8..15 = Jacobi sweeps.                       */  \
    int hcb_##P##compress_batched(hcb_ctx *, int64_t n_tiles, const T *const *dense_ptrs_host, int64_t ld,              \
                                  const hcb_tile *out, const hcb_compress_params *p, int32_t *d_info);                 \
    /* Multi-tile driver (examples/matrix_multiplication/omp_main.cpp:112-126) on one GPU:                           */  \
    /* for k: C(j,i) += A(j,k)*B(k,i) for all owned (j,i) in ONE batched call per k.  Grids are column-major         */  \
    /* arrays of descriptors: A[j + k*mt], B[k + i*kt], C[j + i*mt]; owned == NULL means all C tiles, else a list of */  \
    /* n_owned linear C indices (2D block-cyclic ownership is decided by the caller). Only k in [k_begin, k_end)   */  \
    /* is accumulated (0, kt for the whole product).                                                                */  \
    int hcb_##P##tlr_matmul(hcb_ctx *, int64_t mt, int64_t nt, int64_t kt, const hcb_tile *A, const hcb_tile *B,        \
                            const hcb_tile *C, const int64_t *owned, int64_t n_owned, int64_t k_begin, int64_t k_end,   \
                            T alpha, T beta, const hcb_compress_params *p, int32_t *d_info);                                            \
    /* Local step of the multi-GPU driver (one process per GPU, C tiles 2D block-cyclic over a P x Q grid, SURVEY   */  \
    /* 8e / BASELINE configs[3]): C(j,i) += alpha*Apan[j]*Bpan[i] + (beta-1)*C for all j < mt, i < nt of THIS rank,   */  \
    /* where Apan / Bpan are the k-th row-panel of A / column-panel of B after the NCCL broadcast (the host layer     */  \
    /* hcorepp_b200/distributed.py moves them on a side stream, double buffered).  d_info as above, sticky over k:  */  \
    /* first != 0 zeroes it before the step.                                                                        */  \
    int hcb_##P##tlr_matmul_panel_step(hcb_ctx *, int64_t mt, int64_t nt, const hcb_tile *Apan, const hcb_tile *Bpan,   \
                                       const hcb_tile *C, T alpha, T beta, const hcb_compress_params *p,               \
                                       int32_t *d_info, int first);                                                    \
    /* scratch bytes the fused path needs for a batch of n_tiles (m x n) tiles with rank bound r = kc+ka            */  \
    /* (replaces HCore<T>::CalculateMemoryPoolSize, HCore.cpp:417-480)                                              */  \
    size_t hcb_##P##tlr_gemm_workspace(int64_t n_tiles, int64_t m, int64_t n, int64_t k, int64_t r_bound);

HCB_DECLARE_KERNEL_TABLE(d, double)
HCB_DECLARE_KERNEL_TABLE(s, float)

#ifdef __cplusplus
}
#endif
#endif /* HCORE_B200_H */
