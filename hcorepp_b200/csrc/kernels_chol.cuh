// kernels_chol.cuh -- the dense pieces of the tile-low-rank Cholesky (BASELINE.json configs[4]):
//   HCoreKernels<T>::potrf   include/hcorepp/kernels/kernels.hpp:103-105 (src/kernels/omp/kernels.cpp:234-243: lapack::potrf)
//   HCoreKernels<T>::trsm    kernels.hpp (omp/kernels.cpp:264-271: blas::trsm)
//   HCoreKernels<T>::syrk    kernels.hpp (omp/kernels.cpp:273-281: blas::syrk)
//   FillMatrixTriangle / Symmetrize (omp/kernels.cpp:245-262, 283-303)
// The reference's CUDA backend calls cuSOLVER potrf / cuBLAS trsm, syrk per tile; here the factorisation of a diagonal
// tile is a left-looking blocked Cholesky -- per 32-column block one DMMA GEMM (k_gemm_dmma) for the update and ONE fused
// kernel that factors the 32 x 32 diagonal block (redundantly per CTA: 11k flops) and solves its share of the rows below
// -- and the solve on the compressed tiles of a block column (V := V L^-T, rank x n) is the same scheme batched over
// the tiles with device-resident descriptors.
#pragma once
#include "common.cuh"

namespace hcb {

constexpr int CH_NB = 32;

// Diagonal block + panel below it.  A (n x n, lda) lower: block column [j0, j0 + jb) has been updated with the columns
// to its left.  Every CTA factors the jb x jb diagonal block in shared memory (CTA 0 writes it back) and solves
// X L_jj^T = A[rows, j0:j0+jb] for its 256 rows below the block.  grid.x = 1 + ceil((n - j0 - jb) / 256)
template<typename T>
__global__ void __launch_bounds__(256) k_potrf_panel(T *__restrict__ A, int n, int lda, int j0, int jb, int *__restrict__ info) {
    __shared__ T L[CH_NB][CH_NB + 1];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 256) {
        const int i = idx % CH_NB, j = idx / CH_NB;
        L[i][j] = (i < jb && j < jb && i >= j) ? A[(size_t) (j0 + i) + (size_t) (j0 + j) * lda] : T(0);
    }
    __syncthreads();
    // unblocked right-looking Cholesky of the jb x jb block (one warp would do; the 256 threads share the rank-1 updates)
    for (int j = 0; j < jb; ++j) {
        if (tid == 0) {
            const T d = L[j][j];
            if (!(d > T(0))) { s_bad = j0 + j + 1; L[j][j] = T(1); }
            else L[j][j] = t_sqrt(d);
        }
        __syncthreads();
        const T djj = L[j][j];
        if (tid > j && tid < jb) L[tid][j] /= djj;
        __syncthreads();
        for (int idx = tid; idx < jb * jb; idx += 256) {
            const int i = idx % jb, c = idx / jb;
            if (c > j && i >= c) L[i][c] -= L[i][j] * L[c][j];
        }
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        for (int idx = tid; idx < jb * jb; idx += 256) {
            const int i = idx % jb, j = idx / jb;
            if (i >= j) A[(size_t) (j0 + i) + (size_t) (j0 + j) * lda] = L[i][j];
        }
        if (tid == 0 && s_bad && info) atomicCAS(info, 0, s_bad);  // LAPACK convention: index of the first bad pivot
        return;
    }
    // rows below the block: one row per thread, x_c = (a_c - sum_{l<c} x_l L[c][l]) / L[c][c]
    const int row = j0 + jb + (blockIdx.x - 1) * 256 + tid;
    if (row >= n) return;
    T x[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) x[c] = c < jb ? A[(size_t) row + (size_t) (j0 + c) * lda] : T(0);
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
        if (c < jb) {
            T v = x[c];
#pragma unroll
            for (int l = 0; l < CH_NB; ++l)
                if (l < c) v = fma(-x[l], L[c][l], v);
            x[c] = v / L[c][c];
        }
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c)
        if (c < jb) A[(size_t) row + (size_t) (j0 + c) * lda] = x[c];
}

// Strict upper parts of the 32 x 32 diagonal blocks, saved before / restored after the blocked factorisation (its GEMM
// updates whole block rows; lapack::potrf must leave the other triangle as it was).  save: D[i][c] = A[i][blk(i) + c]
template<typename T>
__global__ void k_diag_upper(int restore, int n, T *__restrict__ A, int lda, T *__restrict__ D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b0 = (i / CH_NB) * CH_NB;
    for (int c = i - b0 + 1; c < CH_NB && b0 + c < n; ++c) {
        T *a = A + (size_t) i + (size_t) (b0 + c) * lda;
        if (restore) *a = D[(size_t) i + (size_t) c * n];
        else D[(size_t) i + (size_t) c * n] = *a;
    }
}

// One block step of X L^T = V (right side, lower, transposed) batched over problems: X (rows x n, ldx) in place, L the
// lower Cholesky factor (n x n, ldl).  Block column [j0, j0 + jb) of X has been updated with the block columns to its
// left (GEMM); this solves it against the diagonal block L_jj.  rows <= blockDim.x * gridDim.x.  grid = (row chunks, problems)
template<typename T>
struct TrsmProb {
    T *X;           // rows x n, ld ldx
    const T *L;     // n x n lower, ld ldl
    int rows, n, ldx, ldl;
};
template<typename T>
__global__ void __launch_bounds__(128) k_trsm_rlt_block(const TrsmProb<T> *__restrict__ probs, int j0, int jb) {
    const TrsmProb<T> p = probs[blockIdx.y];
    if (p.rows <= 0 || j0 >= p.n) return;
    __shared__ T L[CH_NB][CH_NB + 1];
    const int tid = threadIdx.x;
    const int jbb = min(jb, p.n - j0);
    for (int idx = tid; idx < CH_NB * CH_NB; idx += 128) {
        const int i = idx % CH_NB, j = idx / CH_NB;
        L[i][j] = (i < jbb && j < jbb && i >= j) ? p.L[(size_t) (j0 + i) + (size_t) (j0 + j) * p.ldl] : T(0);
    }
    __syncthreads();
    const int row = blockIdx.x * 128 + tid;
    if (row >= p.rows) return;
    T x[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) x[c] = c < jbb ? p.X[(size_t) row + (size_t) (j0 + c) * p.ldx] : T(0);
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
        if (c < jbb) {
            T v = x[c];
#pragma unroll
            for (int l = 0; l < CH_NB; ++l)
                if (l < c) v = fma(-x[l], L[c][l], v);
            x[c] = v / L[c][c];
        }
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c)
        if (c < jbb) p.X[(size_t) row + (size_t) (j0 + c) * p.ldx] = x[c];
}

// Descriptors of the blocked right-side solve V := V L^-T for a batch of compressed tiles whose rank lives on the
// device: per (block step, tile) one GEMM  X[:, jb] -= X[:, :j0] * L[jb, :j0]^T  and one TrsmProb.  Thread per (step, tile).
template<typename T>
__global__ void k_setup_trsm_tiles(const hcb_tile *__restrict__ tiles, const T *const *__restrict__ Ls, const int *__restrict__ ldls,
                                   int n_tiles, int nsteps, GemmProb<T> *__restrict__ gp, TrsmProb<T> *__restrict__ tp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_tiles * nsteps) return;
    const int t = idx % n_tiles, step = idx / n_tiles;
    const hcb_tile X = tiles[t];
    const int rk = *X.d_rank, n = X.n, j0 = step * CH_NB, jb = min(CH_NB, n - j0);
    T *V = reinterpret_cast<T *>(X.d_data) + (size_t) X.m * X.max_rank;  // V (rk x n, ld rk)
    GemmProb<T> g;
    g.A = V; g.B = Ls[t] + j0; g.C = V + (size_t) j0 * rk;
    g.m = (j0 > 0 && jb > 0) ? rk : 0; g.n = jb; g.k = j0;
    g.lda = rk; g.ldb = ldls[t]; g.ldc = rk; g.ta = 0; g.tb = 1; g.alpha = T(-1); g.beta = T(1);
    g.A2 = nullptr; g.k1 = j0; g.lda2 = 1;
    gp[idx] = g;
    if (step == 0) {
        tp[t] = TrsmProb<T>{V, Ls[t], rk, n, rk, ldls[t]};
        if (X.d_state) atomicAnd(X.d_state, ~HCB_STATE_ORTHO_V);  // V L^-T no longer has orthogonal rows (U is untouched)
    }
}

// Descriptors of the symmetric update of the diagonal tiles by the compressed tiles of a block column (HCore<T>::Syrk
// with a compressed A, HCore.cpp:484-575):  W = AV AV^T (k x k),  TT = AU W (m x k),  C := beta C + alpha TT AU^T.
// Thread per tile; W / TT live in the scratch slab of the tile.
template<typename T>
__global__ void k_setup_syrk_tiles(const hcb_tile *__restrict__ tiles, T *const *__restrict__ Cs, const int *__restrict__ ldcs,
                                   int n_tiles, T *__restrict__ ws, size_t slab, T alpha, T beta, GemmProb<T> *__restrict__ g1,
                                   GemmProb<T> *__restrict__ g2, GemmProb<T> *__restrict__ g3) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const hcb_tile X = tiles[t];
    const int k = *X.d_rank, m = X.m, n = X.n;
    const T *U = reinterpret_cast<const T *>(X.d_data), *V = U + (size_t) m * X.max_rank;
    T *W = ws + (size_t) t * slab, *TT = W + (((size_t) X.max_rank * X.max_rank + 31) / 32) * 32;
    auto mk = [](const T *A, int lda, int ta, const T *B, int ldb, int tb, T *C, int ldc, int mm, int nn, int kk, T al, T be) {
        GemmProb<T> g;
        g.A = A; g.B = B; g.C = C; g.m = mm; g.n = nn; g.k = kk; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.ta = ta; g.tb = tb;
        g.alpha = al; g.beta = be; g.A2 = nullptr; g.k1 = kk; g.lda2 = 1;
        return g;
    };
    g1[t] = mk(V, k, 0, V, k, 1, W, k, k, k, n, T(1), T(0));            // W  = V V^T
    g2[t] = mk(U, m, 0, W, k, 0, TT, m, m, k, k, T(1), T(0));           // TT = U W
    g3[t] = mk(TT, m, 0, U, m, 1, Cs[t], ldcs[t], m, m, k, alpha, beta);  // C  = beta C + alpha TT U^T
}

// Generic triangular solve for the compat entry (all side / uplo / trans / diag combinations, any shape): one thread
// per independent right-hand side (a column of B for side = Left, a row of B for side = Right), plain substitution with
// the triangle read through L2.  O(n^2) per thread: correct for every case, used by the reference-style tile-at-a-time
// flow; the TLR Cholesky driver uses the blocked, batched k_trsm_rlt_block + GEMM path instead.
template<typename T>
__global__ void k_trsm_generic(int right, int upper, int trans, int unit, int m, int n, T alpha, const T *__restrict__ A, int lda,
                               T *__restrict__ B, int ldb) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrhs = right ? m : n, len = right ? n : m;
    if (v >= nrhs) return;
    // unknown i of this right-hand side lives at B[off(i)]
    auto at = [&](int i) -> T & { return right ? B[(size_t) v + (size_t) i * ldb] : B[(size_t) i + (size_t) v * ldb]; };
    // Left:  op(A) x = alpha b  -> coefficient of x_j in equation i is op(A)(i, j)
    // Right: x op(A) = alpha b  -> equation i: sum_j x_j op(A)(j, i)  -> coefficient of x_j in equation i is op(A)(j, i)
    auto coef = [&](int i, int j) -> T {
        int r = right ? j : i, c = right ? i : j;      // element op(A)(r, c)
        if (trans) { const int tmp = r; r = c; c = tmp; }
        return A[(size_t) r + (size_t) c * lda];
    };
    // the system matrix M(i, j) = coef(i, j) is lower triangular when (upper XOR trans XOR right) is false
    const bool lower_sys = !(((upper != 0) != (trans != 0)) != (right != 0));
    for (int i = 0; i < len; ++i) at(i) *= alpha;
    if (lower_sys) {
        for (int i = 0; i < len; ++i) {
            T s = at(i);
            for (int j = 0; j < i; ++j) s = fma(-coef(i, j), at(j), s);
            at(i) = unit ? s : s / coef(i, i);
        }
    } else {
        for (int i = len - 1; i >= 0; --i) {
            T s = at(i);
            for (int j = i + 1; j < len; ++j) s = fma(-coef(i, j), at(j), s);
            at(i) = unit ? s : s / coef(i, i);
        }
    }
}

// C (uplo triangle only) := alpha * W + beta * C, W = op(A) op(A)^T already formed (n x n, ld n) -- blas::syrk semantics
template<typename T>
__global__ void k_syrk_combine(int upper, int n, T alpha, const T *__restrict__ W, T beta, T *__restrict__ Cm, int ldc) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= n || j >= n) return;
    if (upper ? (i > j) : (i < j)) return;
    T *c = Cm + (size_t) i + (size_t) j * ldc;
    *c = alpha * W[(size_t) i + (size_t) j * n] + (beta == T(0) ? T(0) : beta * *c);
}

// FillMatrixTriangle (omp/kernels.cpp:245-262): the STRICT `upper`/lower triangle of a square matrix := value
template<typename T>
__global__ void k_fill_triangle(int upper, int n, T *__restrict__ A, int lda, T value) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= n || j >= n) return;
    if (upper ? (i < j) : (i > j)) A[(size_t) i + (size_t) j * lda] = value;
}

// Symmetrize (omp/kernels.cpp:283-303): copy the `upper`/lower triangle onto the other one
template<typename T>
__global__ void k_symmetrize(int from_upper, int n, T *__restrict__ A, int lda) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= n || j >= n || i >= j) return;  // (i, j) with i < j is in the strict upper triangle
    if (from_upper) A[(size_t) j + (size_t) i * lda] = A[(size_t) i + (size_t) j * lda];
    else A[(size_t) i + (size_t) j * lda] = A[(size_t) j + (size_t) i * lda];
}

}  // namespace hcb
