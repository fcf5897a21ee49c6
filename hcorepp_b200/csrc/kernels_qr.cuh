// kernels_qr.cuh -- batched Householder QR and reflector application (general path: any m, n, ld).
//
// Replaces cusolverDnXgeqrf / cusolverDn{S,D}orgqr / cusolverDnDormqr as called per tile by the reference
// (src/kernels/cuda/CudaKernels.cu:534-561, 962-1017, 732-768).  Output layout and sign convention are LAPACK's
// (dgeqr2/dlarfg: beta = -sign(alpha)*||x||, tau = (beta-alpha)/beta, v(0) = 1 implicit), so R and the reflectors
// match the reference's known answers (tests/kernels/TestKernels.cpp:470-511).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace hcb {

// One CTA per panel. Per column j: (1) CTA-wide norm of A[j+1:, j]; (2) thread 0 forms beta/tau; (3) v scaled in
// place; (4) every warp owns a set of trailing columns and applies H_j = I - tau v v^T to each of them with one
// fused pass (dot product, warp-shuffle reduction, rank-1 update); column reads are fully coalesced.
// 4-way unrolled strided dot product / axpy over rows [i0, m): independent accumulators keep 8 loads in flight per
// lane (the straight loop serialised on L2 latency: 1.06 ms per 512-panel launch, ncu launch list r01).
template<typename T>
__device__ __forceinline__ T lane_dot(const T *__restrict__ x, const T *__restrict__ y, int i0, int m, int lane) {
    T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);
    int i = i0 + lane;
    for (; i + 96 < m; i += 128) {
        d0 = fma(x[i], y[i], d0);
        d1 = fma(x[i + 32], y[i + 32], d1);
        d2 = fma(x[i + 64], y[i + 64], d2);
        d3 = fma(x[i + 96], y[i + 96], d3);
    }
    for (; i < m; i += 32) d0 = fma(x[i], y[i], d0);
    return (d0 + d1) + (d2 + d3);
}

template<typename T>
__device__ __forceinline__ void lane_axpy(T t, const T *__restrict__ x, T *__restrict__ y, int i0, int m, int lane) {
    int i = i0 + lane;
    for (; i + 96 < m; i += 128) {
        const T a0 = x[i], a1 = x[i + 32], a2 = x[i + 64], a3 = x[i + 96];
        const T b0 = y[i], b1 = y[i + 32], b2 = y[i + 64], b3 = y[i + 96];
        y[i] = fma(-t, a0, b0);
        y[i + 32] = fma(-t, a1, b1);
        y[i + 64] = fma(-t, a2, b2);
        y[i + 96] = fma(-t, a3, b3);
    }
    for (; i < m; i += 32) y[i] = fma(-t, x[i], y[i]);
}

template<typename T>
__global__ void __launch_bounds__(512) k_geqrf_batched(const QrProb<T> *__restrict__ probs) {
    const QrProb<T> p = probs[blockIdx.x];
    const int m = p.m, n = p.n, lda = p.lda;
    const int kmax = m < n ? m : n;
    if (kmax <= 0) return;
    __shared__ T red[33];
    __shared__ T s_tau, s_scale;
    T *A = p.A;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;

    for (int j = 0; j < kmax; ++j) {
        T *col = A + (size_t) j * lda;
        // (1) ||A[j+1:m, j]||^2
        T s0 = T(0), s1 = T(0);
        {
            int i = j + 1 + tid;
            for (; i + nthr < m; i += 2 * nthr) {
                const T x = col[i], y = col[i + nthr];
                s0 = fma(x, x, s0);
                s1 = fma(y, y, s1);
            }
            for (; i < m; i += nthr) { const T x = col[i]; s0 = fma(x, x, s0); }
        }
        const T ss = block_sum(s0 + s1, red);
        // (2) Householder scalars
        if (tid == 0) {
            const T alpha = col[j];
            if (ss == T(0)) {
                s_tau = T(0);  // H = I (dlarfg: xnorm == 0)
                s_scale = T(0);
            } else {
                const T nrm = t_sqrt(fma(alpha, alpha, ss));
                const T beta = alpha >= T(0) ? -nrm : nrm;
                s_tau = (beta - alpha) / beta;
                s_scale = T(1) / (alpha - beta);
                col[j] = beta;
            }
            p.tau[j] = s_tau;
        }
        __syncthreads();
        const T tau = s_tau, scale = s_scale;
        if (tau != T(0)) {
            // (3) v = x / (alpha - beta)
            for (int i = j + 1 + tid; i < m; i += nthr) col[i] *= scale;
            __syncthreads();
            // (4) trailing update, one warp per column
            for (int c = j + 1 + w; c < n; c += nw) {
                T *cc = A + (size_t) c * lda;
                T dot = lane_dot<T>(col, cc, j + 1, m, lane);
                dot = warp_sum(dot);
                dot += cc[j];  // v(0) = 1
                const T t = tau * dot;
                __syncwarp();  // every lane has read cc[j] before lane 0 overwrites it
                if (lane == 0) cc[j] -= t;
                lane_axpy<T>(t, col, cc, j + 1, m, lane);
            }
        }
        __syncthreads();
    }
}

// Apply k Householder reflectors (LAPACK layout) to C.
//   side_right == 0 : C := H.. C   -- every column of C is independent -> one warp per column
//   side_right == 1 : C := C H..   -- every row    of C is independent -> one warp per row
// forward == 1 applies H_0 first (Q^T C, or C Q); forward == 0 applies H_{k-1} first (Q C, or C Q^T).
// grid = (vector_blocks, n_problems); a warp loops over its vectors, no CTA-level synchronisation is needed.
template<typename T>
__global__ void __launch_bounds__(256) k_apply_reflectors(const ReflProb<T> *__restrict__ probs) {
    const ReflProb<T> p = probs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    int nvec = p.side_right ? p.mc : p.nc;
    if (p.nc_dev) nvec = *p.nc_dev;
    const int len = p.mv;  // reflector length == length of each transformed vector
    const size_t estride = p.side_right ? (size_t) p.ldc : 1;  // element stride inside one vector
    const size_t vstride = p.side_right ? 1 : (size_t) p.ldc;  // stride between vectors
    for (int vec = warp_global; vec < nvec; vec += warps_total) {
        T *x = p.C + (size_t) vec * vstride;
        for (int jj = 0; jj < p.k; ++jj) {
            const int j = p.forward ? jj : p.k - 1 - jj;
            const T tau = p.tau[j];
            if (tau == T(0)) continue;
            const T *v = p.V + (size_t) j * p.ldv;
            T dot = T(0);
            if (estride == 1) dot = lane_dot<T>(v, x, j + 1, len, lane);
            else for (int i = j + 1 + lane; i < len; i += 32) dot = fma(v[i], x[(size_t) i * estride], dot);
            dot = warp_sum(dot);
            dot += x[(size_t) j * estride];
            const T t = tau * dot;
            __syncwarp();
            if (lane == 0) x[(size_t) j * estride] -= t;
            if (estride == 1) lane_axpy<T>(t, v, x, j + 1, len, lane);
            else for (int i = j + 1 + lane; i < len; i += 32) x[(size_t) i * estride] = fma(-t, v[i], x[(size_t) i * estride]);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked (compact-WY) machinery: Q_b = H_j0 .. H_j0+jb-1 = I - V_b T_b V_b^T (LAPACK dlarft, forward/columnwise).
// The panel factorisation of each NBQ-column block stays level-2 (k_geqrf_batched on the sub-panel); everything else
// -- trailing update in the QR, and the whole rebuild Q*[X;0] -- becomes batched GEMMs (k_gemm_batched) with
// device-resident descriptors, which is where the flops are once the stacked rank r = kc + ka reaches the hundreds.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NBQ = 32;  // reflector block size

template<typename T>
struct LarftProb {
    const T *A;    // first element of the block inside the factored panel: &A[j0 + j0*lda]
    const T *tau;  // &tau[j0]
    T *Vc;         // &Vc[j0 + j0*ldvc]: receives the clean unit-lower-trapezoidal copy of the block's reflectors
    T *Tm;         // NBQ x NBQ (ld NBQ): receives T_b (upper triangular, zeros below)
    int lda, ldvc, rows, jb;  // rows = m - j0 ; jb = reflectors in this block (0: nothing to do)
};

// One CTA (256 threads) per block: (1) Vc = unit-lower-trapezoid(V_b); (2) G = Vc^T Vc (strict upper part);
// (3) T by the dlarft recurrence: T(i,i) = tau_i, T(0:i,i) = -tau_i * T(0:i,0:i) * G(0:i,i).
template<typename T>
__global__ void __launch_bounds__(256) k_larft_extract(const LarftProb<T> *__restrict__ probs) {
    const LarftProb<T> p = probs[blockIdx.x];
    const int jb = p.jb, rows = p.rows;
    if (jb <= 0 || rows <= 0) return;
    __shared__ T G[NBQ][NBQ + 1];
    __shared__ T Ts[NBQ][NBQ + 1];
    __shared__ T tcol[NBQ];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    for (int c = 0; c < jb; ++c) {
        const T *src = p.A + (size_t) c * p.lda;
        T *dst = p.Vc + (size_t) c * p.ldvc;
        for (int i = tid; i < rows; i += blockDim.x) dst[i] = (i < c) ? T(0) : (i == c ? T(1) : src[i]);
    }
    for (int idx = tid; idx < NBQ * NBQ; idx += blockDim.x) Ts[idx / NBQ][idx % NBQ] = T(0);
    __syncthreads();
    for (int pr = w; pr < jb * jb; pr += nw) {
        const int i = pr / jb, j = pr % jb;
        if (i >= j) continue;
        const T *vi = p.Vc + (size_t) i * p.ldvc, *vj = p.Vc + (size_t) j * p.ldvc;
        T d = lane_dot<T>(vi, vj, j, rows, lane);  // vj is zero above row j
        d = warp_sum(d);
        if (lane == 0) G[i][j] = d;
    }
    __syncthreads();
    if (w == 0) {
        for (int i = 0; i < jb; ++i) {
            const T ti = p.tau[i];
            if (lane < i) tcol[lane] = -ti * G[lane][i];
            __syncwarp();
            T acc = T(0);
            if (lane < i)
                for (int c = lane; c < i; ++c) acc = fma(Ts[lane][c], tcol[c], acc);
            __syncwarp();
            if (lane < i) Ts[lane][i] = acc;
            if (lane == i) Ts[i][i] = ti;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NBQ * NBQ; idx += blockDim.x) p.Tm[idx] = Ts[idx % NBQ][idx / NBQ];  // column-major
}

// ---------------------------------------------------------------------------------------------------------------
// On-chip panel factorisation: one thread-block CLUSTER per NBQ-column block.
//
// k_geqrf_batched + k_larft_extract work on the panel through L2 (8 + 4 MB of traffic per 1024 x 32 block; at
// r = 357 they were 45 % of the blocked-QR time and ran at the L2 bandwidth limit, ncu launch list r01).  Here the
// block lives in shared memory for the whole factorisation: each CTA of the cluster owns 512 rows x 32 columns
// (128 KB fp64), one row per thread.  Per column step every thread forms its row's 32 products x_j[row] * P[row][c],
// a transposed warp butterfly + one shared-memory pass reduce them per CTA, and the CTAs exchange their 32 partial
// sums (+ the pivot row) through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank) -- one vector per step instead of
// streaming the trailing columns through L2.  The same sums give ||x||^2 (c == j), the trailing-update dots (c > j)
// and V_prev^T v_j for the dlarft recurrence (c < j), so T_b and the clean V_b copy come out of the same kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PQ_ROWS = 512;  // rows (= threads) per CTA

// sum over the warp of pr[k] for every k in [0, 32); lane l returns the total for index l (31 shuffle steps)
template<typename T>
__device__ __forceinline__ T warp_reduce_32(T (&pr)[32], int lane) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const bool up = lane & 16;
        const T send = up ? pr[k] : pr[k + 16];
        const T keep = up ? pr[k + 16] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool up = lane & 8;
        const T send = up ? pr[k] : pr[k + 8];
        const T keep = up ? pr[k + 8] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool up = lane & 4;
        const T send = up ? pr[k] : pr[k + 4];
        const T keep = up ? pr[k + 4] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool up = lane & 2;
        const T send = up ? pr[k] : pr[k + 2];
        const T keep = up ? pr[k + 2] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const bool up = lane & 1;
        const T send = up ? pr[0] : pr[1];
        const T keep = up ? pr[1] : pr[0];
        pr[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    return pr[0];
}

// grid.x = cluster_size * n_panels, cluster dims (cluster_size, 1, 1), block = PQ_ROWS threads,
// dynamic shared memory = NBQ * PQ_ROWS * sizeof(T).
template<typename T>
__global__ void __launch_bounds__(PQ_ROWS) k_panel_qr_cluster(const QrProb<T> *__restrict__ qps,
                                                              const LarftProb<T> *__restrict__ lps) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const int panel = blockIdx.x / CS;
    const QrProb<T> q = qps[panel];
    const LarftProb<T> lp = lps[panel];
    const int m = q.m, jb = q.n < m ? q.n : m, lda = q.lda;
    if (m <= 0 || jb <= 0) return;  // uniform over the cluster (same descriptor)
    extern __shared__ __align__(16) unsigned char smem_raw_pq[];
    T *P = reinterpret_cast<T *>(smem_raw_pq);  // P[c * PQ_ROWS + t]: column c, local row t
    __shared__ T zpart[PQ_ROWS / 32][NBQ];
    __shared__ T zloc[2 * NBQ + 1];  // [0,32): partial dots; [32,64): pivot row (CTA 0 only); [64]: unused
    __shared__ T zsum[2 * NBQ];
    __shared__ T Ts[NBQ][NBQ + 1];
    __shared__ T tcol[NBQ];
    __shared__ T s_tau[NBQ];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int gr = crank * PQ_ROWS + t;  // row inside the block
    const bool have = gr < m;
    for (int c = 0; c < NBQ; ++c) P[c * PQ_ROWS + t] = (have && c < jb) ? q.A[(size_t) gr + (size_t) c * lda] : T(0);
    for (int idx = t; idx < NBQ * (NBQ + 1); idx += PQ_ROWS) (&Ts[0][0])[idx] = T(0);
    __syncthreads();

    for (int j = 0; j < jb; ++j) {
        // ---- partial sums z[c] = sum_{rows > j} x_j[row] * P[row][c]
        const T xv = (have && gr > j) ? P[j * PQ_ROWS + t] : T(0);
        T pr[NBQ];
#pragma unroll
        for (int c = 0; c < NBQ; ++c) pr[c] = xv * P[c * PQ_ROWS + t];
        const T mine = warp_reduce_32<T>(pr, lane);
        zpart[w][lane] = mine;
        __syncthreads();
        if (t < NBQ) {
            T sacc = T(0);
#pragma unroll
            for (int ww = 0; ww < PQ_ROWS / 32; ++ww) sacc += zpart[ww][t];
            zloc[t] = sacc;
            zloc[NBQ + t] = (crank == 0) ? P[t * PQ_ROWS + j] : T(0);  // pivot row j lives in CTA 0 (j < 32 <= PQ_ROWS)
        }
        if (CS > 1) {
            cluster.sync();
            if (t < 2 * NBQ) {
                T sacc = T(0);
                for (int rk = 0; rk < CS; ++rk) sacc += cluster.map_shared_rank(zloc, rk)[t];
                zsum[t] = sacc;
            }
            cluster.sync();  // everybody has read zloc before the next step overwrites it
        } else {
            __syncthreads();
            if (t < 2 * NBQ) zsum[t] = zloc[t];
            __syncthreads();
        }
        // ---- Householder scalars (every thread, from shared broadcasts)
        const T ss = zsum[j], alpha = zsum[NBQ + j];
        T tau = T(0), scale = T(0), beta = alpha;
        if (ss != T(0)) {
            const T nrm = t_sqrt(fma(alpha, alpha, ss));
            beta = alpha >= T(0) ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            scale = T(1) / (alpha - beta);
        }
        // ---- update my row: reflector entry + trailing columns
        if (tau != T(0)) {
            if (have && gr > j) {
                const T v = xv * scale;
                P[j * PQ_ROWS + t] = v;
                const T tv = tau * v;
                for (int c = j + 1; c < jb; ++c) {
                    const T wc = fma(scale, zsum[c], zsum[NBQ + c]);  // v^T A[:, c]
                    P[c * PQ_ROWS + t] = fma(-tv, wc, P[c * PQ_ROWS + t]);
                }
            } else if (have && gr == j) {
                P[j * PQ_ROWS + t] = beta;
                for (int c = j + 1; c < jb; ++c) {
                    const T wc = fma(scale, zsum[c], zsum[NBQ + c]);
                    P[c * PQ_ROWS + t] = fma(-tau, wc, P[c * PQ_ROWS + t]);
                }
            }
        }
        // ---- dlarft recurrence for column j of T (warp 0 of every CTA, only CTA 0 stores it)
        if (w == 0) {
            if (lane == 0) s_tau[j] = tau;
            if (lane < j) tcol[lane] = -tau * fma(scale, zsum[lane], zsum[NBQ + lane]);  // -tau * (V_prev^T v_j)
            __syncwarp();
            T acc = T(0);
            if (lane < j)
                for (int c = lane; c < j; ++c) acc = fma(Ts[lane][c], tcol[c], acc);
            __syncwarp();
            if (lane < j) Ts[lane][j] = acc;
            if (lane == j) Ts[j][j] = tau;
        }
        __syncthreads();
    }
    // ---- write back: factored block, tau, clean V_b, T_b
    if (have) {
        for (int c = 0; c < jb; ++c) {
            const T val = P[c * PQ_ROWS + t];
            q.A[(size_t) gr + (size_t) c * lda] = val;
            lp.Vc[(size_t) gr + (size_t) c * lp.ldvc] = (gr < c) ? T(0) : (gr == c ? T(1) : val);
        }
    }
    if (crank == 0) {
        if (t < jb) q.tau[t] = s_tau[t];
        for (int idx = t; idx < NBQ * NBQ; idx += PQ_ROWS) lp.Tm[idx] = Ts[idx % NBQ][idx / NBQ];
    }
}

}  // namespace hcb
