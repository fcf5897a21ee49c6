// kernels_blas.cuh -- batched GEMM / copy kernels driven by DEVICE-resident problem descriptors.
//
// Replaces, for every dense contraction on the TLR-GEMM path, the per-tile cuBLAS calls of the reference
// (HCoreKernels::Gemm, src/kernels/cuda/kernels.cpp:18-24 -> blas::gemm(queue)) and its element-wise helpers
// (LaCpy / MultiplyByAlpha / ProcessVpointer / CalculateUVptr, src/kernels/cuda/CudaKernels.cu:37-230,371-429):
// one launch covers every tile of a batch, shapes are read from device memory (ranks are data-dependent).
#pragma once
#include "common.cuh"

namespace hcb {

// ---------------------------------------------------------------------------------------------------------------
// Batched GEMM, SIMT FMA path (FP64: DFMA; FP32: FFMA -- TF32 would cost ~1e-3 relative error).
// CTA tile BM x BN, 256 threads, each thread a TM x TN register micro-tile; operands staged in padded shared
// memory so both N and T operand layouts are read coalesced from global memory.
// grid = (tiles_bound, n_problems); a CTA whose tile lies outside its problem's m x n exits immediately.
// ---------------------------------------------------------------------------------------------------------------
template<typename T, int BM, int BN, int BK>
__global__ void __launch_bounds__(256) k_gemm_batched(const GemmProb<T> *__restrict__ probs) {
    constexpr int TM = BM / 16, TN = BN / 16;
    const GemmProb<T> p = probs[blockIdx.y];
    if (p.m <= 0 || p.n <= 0) return;
    const int tiles_m = (p.m + BM - 1) / BM, tiles_n = (p.n + BN - 1) / BN;
    // grid-stride over output tiles so that any grid.x >= 1 is correct (the host sizes it from rank bounds)
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const int row0 = (tile % tiles_m) * BM, col0 = (tile / tiles_m) * BN;
        T acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

        for (int k0 = 0; k0 < p.k; k0 += BK) {
            // ---- stage A tile: As[kk][r] = op(A)[row0 + r, k0 + kk]
            if (p.ta == 0) {
                for (int idx = tid; idx < BM * BK; idx += 256) {
                    const int r = idx % BM, kk = idx / BM;
                    const int gr = row0 + r, gk = k0 + kk;
                    As[kk][r] = (gr < p.m && gk < p.k) ? p.A[(size_t) gr + (size_t) gk * p.lda] : T(0);
                }
            } else {
                for (int idx = tid; idx < BM * BK; idx += 256) {
                    const int kk = idx % BK, r = idx / BK;
                    const int gr = row0 + r, gk = k0 + kk;
                    As[kk][r] = (gr < p.m && gk < p.k) ? p.A[(size_t) gk + (size_t) gr * p.lda] : T(0);
                }
            }
            // ---- stage B tile: Bs[kk][c] = op(B)[k0 + kk, col0 + c]
            if (p.tb == 0) {
                for (int idx = tid; idx < BN * BK; idx += 256) {
                    const int kk = idx % BK, c = idx / BK;
                    const int gc = col0 + c, gk = k0 + kk;
                    Bs[kk][c] = (gc < p.n && gk < p.k) ? p.B[(size_t) gk + (size_t) gc * p.ldb] : T(0);
                }
            } else {
                for (int idx = tid; idx < BN * BK; idx += 256) {
                    const int c = idx % BN, kk = idx / BN;
                    const int gc = col0 + c, gk = k0 + kk;
                    Bs[kk][c] = (gc < p.n && gk < p.k) ? p.B[(size_t) gc + (size_t) gk * p.ldb] : T(0);
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                T a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
        // ---- epilogue (BLAS semantics: C is not read when beta == 0, scratch need not be initialised)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gc = col0 + ty + 16 * j;
            if (gc >= p.n) continue;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const int gr = row0 + tx + 16 * i;
                if (gr >= p.m) continue;
                T *c = p.C + (size_t) gr + (size_t) gc * p.ldc;
                T v = p.alpha * acc[i][j];
                if (p.beta != T(0)) v = fma(p.beta, *c, v);
                *c = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Batched strided copy with scale and optional transpose (32x32 shared-memory tiles: both sides coalesced).
// grid = (tiles_bound, n_problems), block = (32, 8)
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
__global__ void __launch_bounds__(256) k_copy_batched(const CopyProb<T> *__restrict__ probs) {
    const CopyProb<T> p = probs[blockIdx.y];
    if (p.rows <= 0 || p.cols <= 0) return;
    __shared__ T tile[32][33];
    const int tr = (p.rows + 31) / 32, tc = (p.cols + 31) / 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
        const int r0 = (t % tr) * 32, c0 = (t / tr) * 32;
        if (!p.trans) {
            for (int j = ty; j < 32; j += 8) {
                const int r = r0 + tx, c = c0 + j;
                if (r < p.rows && c < p.cols)
                    p.dst[(size_t) r + (size_t) c * p.ldd] = p.scale * p.src[(size_t) r + (size_t) c * p.lds];
            }
        } else {
            // dst(r, c) = src(c, r): read src with its fast index (c) along threadIdx.x
            for (int j = ty; j < 32; j += 8) {
                const int c = c0 + tx, r = r0 + j;
                tile[j][tx] = (r < p.rows && c < p.cols) ? p.src[(size_t) c + (size_t) r * p.lds] : T(0);
            }
            __syncthreads();
            for (int j = ty; j < 32; j += 8) {
                const int r = r0 + tx, c = c0 + j;
                if (r < p.rows && c < p.cols) p.dst[(size_t) r + (size_t) c * p.ldd] = p.scale * tile[tx][j];
            }
            __syncthreads();
        }
    }
}

// lacpy / laset with triangle selection (compat entry points; MatrixType 'G','U','L')
template<typename T>
__global__ void k_lacpy(char type, int m, int n, const T *__restrict__ A, int lda, T *__restrict__ B, int ldb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= m || j >= n) return;
    if ((type == 'U' && i > j) || (type == 'L' && i < j)) return;
    B[(size_t) i + (size_t) j * ldb] = A[(size_t) i + (size_t) j * lda];
}

template<typename T>
__global__ void k_laset(char type, int m, int n, T offdiag, T diag, T *__restrict__ A, int lda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= m || j >= n) return;
    if (i == j) A[(size_t) i + (size_t) j * lda] = diag;
    else if (type == 'G' || (type == 'L' && i > j) || (type == 'U' && i < j)) A[(size_t) i + (size_t) j * lda] = offdiag;
}

// MultiplyByAlpha (omp/kernels.cpp:21-28): arr[m*rank + i] *= alpha, i < rows*cols
template<typename T>
__global__ void k_scale_flat(T *__restrict__ arr, size_t offset, size_t count, T alpha) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) arr[offset + i] *= alpha;
}

// CalculateVTnew (omp/kernels.cpp:116-130): VT[i + j*ld] *= sigma[i], i < rk, j < cols
template<typename T>
__global__ void k_scale_rows(T *__restrict__ VT, int ld, int rk, int cols, const T *__restrict__ sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < rk && j < cols) VT[(size_t) i + (size_t) j * ld] *= sigma[i];
}

// FillIdentityMatrix (omp/kernels.cpp:144-151): only the diagonal is written
template<typename T>
__global__ void k_fill_diag(T *__restrict__ A, int n, int ld, T v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(size_t) i + (size_t) i * ld] = v;
}

// CalculateNewRank (omp/kernels.cpp:82-104), on the device: first i >= 1 with sigma_i < thr, else size_s.
template<typename T>
__device__ __forceinline__ int new_rank_rule(const T *sigma, int size_s, T accuracy, int truncated) {
    const T thr = truncated ? accuracy * sigma[0] : accuracy;
    int rk = size_s;
    for (int i = 1; i < size_s; ++i)
        if (sigma[i] < thr) { rk = i; break; }
    return rk;
}

template<typename T>
__global__ void k_new_rank(const T *__restrict__ sigma, int size_s, T accuracy, int truncated, int *rank_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *rank_out = new_rank_rule(sigma, size_s, accuracy, truncated);
}

// upper/lower triangle of A (with optional unit diagonal) expanded into a dense n x n matrix (for trmm)
template<typename T>
__global__ void k_expand_tri(char uplo, char diag, int n, const T *__restrict__ A, int lda, T *__restrict__ D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n || j >= n) return;
    T v = T(0);
    if (i == j) v = (diag == 'U') ? T(1) : A[(size_t) i + (size_t) j * lda];
    else if ((uplo == 'U' && i < j) || (uplo == 'L' && i > j)) v = A[(size_t) i + (size_t) j * lda];
    D[(size_t) i + (size_t) j * n] = v;
}

// Element-wise precision conversion of whole buffers (FP32 tiles <-> their FP64 shadows, see t_tlr_gemm_promoted).
// grid = (chunks, n_problems)
template<typename S, typename D>
struct ConvProb { const S *src; D *dst; size_t n; };
template<typename S, typename D>
__global__ void __launch_bounds__(256) k_convert_batched(const ConvProb<S, D> *__restrict__ probs) {
    const ConvProb<S, D> p = probs[blockIdx.y];
    for (size_t i = (size_t) blockIdx.x * 256 + threadIdx.x; i < p.n; i += (size_t) gridDim.x * 256) p.dst[i] = (D) p.src[i];
}

}  // namespace hcb
