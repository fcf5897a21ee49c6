// kernels_qr.cuh -- batched Householder QR and reflector application (general path: any m, n, ld).
//
// Replaces cusolverDnXgeqrf / cusolverDn{S,D}orgqr / cusolverDnDormqr as called per tile by the reference
// (src/kernels/cuda/CudaKernels.cu:534-561, 962-1017, 732-768).  Output layout and sign convention are LAPACK's
// (dgeqr2/dlarfg: beta = -sign(alpha)*||x||, tau = (beta-alpha)/beta, v(0) = 1 implicit), so R and the reflectors
// match the reference's known answers (tests/kernels/TestKernels.cpp:470-511).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
#include <type_traits>

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
// k_panel_qr_regs: on-chip panel factorisation, one thread-block CLUSTER per NBQ-column block, the block in REGISTERS.
// (k_geqrf_batched + k_larft_extract work on the panel through L2: 8 + 4 MB of traffic per 1024 x 32 block, 1731 us per
// 512-panel launch at r = 357; a first on-chip version kept the block in shared memory -- 879 us; this one: ~650 us.)
//  256 threads per CTA, two rows per
// thread (local rows t and t + 256), 64 doubles of row data per thread.  The pivot column is always register slot 0:
// the trailing update writes column c into slot c-1, so the loop over the 32 columns stays a rolled loop with
// compile-time register indices.  Per step: 64 FMAs form the thread's products, two 16-value transposed butterflies
// reduce them over the warp, each warp stores its totals straight into every CTA of the cluster (DSMEM stores),
// ONE cluster barrier, 32 threads sum the 8*CS warp partials, ONE block barrier, update.  Finished reflector columns
// go to shared memory; V^T V for the dlarft recurrence is one DMMA pass over them at the end.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PR_THREADS = 256, PR_ROWS = 512, PR_MAXCS = 8;

// reciprocal square root to ~1 ulp without the library's special-case path (arguments are screened by the caller)
__device__ __forceinline__ double pq_rsqrt(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    y = fma(0.5 * y, fma(-q * y, y, 1.0), y);  // 2^-22 -> 2^-43
    y = fma(0.5 * y, fma(-q * y, y, 1.0), y);  //       -> rounding level
    return y;
}
__device__ __forceinline__ float pq_rsqrt(float q) { return rsqrtf(q); }

template<typename T>
__device__ __forceinline__ T warp_reduce_16(T (&pr)[16], int lane) {  // lane l returns the total of index (l >> 1) & 15
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool up = lane & 16;
        const T send = up ? pr[k] : pr[k + 8];
        const T keep = up ? pr[k + 8] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool up = lane & 8;
        const T send = up ? pr[k] : pr[k + 4];
        const T keep = up ? pr[k + 4] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool up = lane & 4;
        const T send = up ? pr[k] : pr[k + 2];
        const T keep = up ? pr[k + 2] : pr[k];
        pr[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = lane & 2;
        const T send = up ? pr[0] : pr[1];
        const T keep = up ? pr[1] : pr[0];
        pr[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    pr[0] += __shfl_xor_sync(0xffffffffu, pr[0], 1);
    return pr[0];
}

template<typename T>
constexpr size_t pr_smem_bytes() {
    return sizeof(T) * ((size_t) NBQ * PR_ROWS + 2 * PR_MAXCS * (PR_THREADS / 32) * NBQ + 2 * NBQ + 2 * NBQ +
                        3 * NBQ * (NBQ + 1) + NBQ * NBQ + 2 * NBQ);
}

__device__ __forceinline__ int pr_addr(int row, int col) { return col * PR_ROWS + ((row + 4 * col) & (PR_ROWS - 1)); }

template<typename T>
__global__ void __launch_bounds__(PR_THREADS, 1) k_panel_qr_regs(const QrProb<T> *__restrict__ qps,
                                                                 const LarftProb<T> *__restrict__ lps) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const int panel = blockIdx.x / CS;
    const QrProb<T> q = qps[panel];
    const LarftProb<T> lp = lps[panel];
    const int m = q.m, jb = q.n < m ? q.n : m, lda = q.lda;
    if (m <= 0 || jb <= 0) return;  // uniform over the cluster
    constexpr int NW = PR_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw_pr[];
    T *P = reinterpret_cast<T *>(smem_raw_pr);            // finished reflector columns, pr_addr layout
    T *zpart = P + NBQ * PR_ROWS;                         // [2][PR_MAXCS][NW][NBQ] warp partials from every CTA
    T *zrow = zpart + 2 * PR_MAXCS * NW * NBQ;            // [2][NBQ] pivot row (slot 0 = alpha)
    T *zsum = zrow + 2 * NBQ;                             // [NBQ] (+ NBQ spare)
    T *Ts = zsum + 2 * NBQ;                               // [NBQ][NBQ+1]
    T *Rs = Ts + NBQ * (NBQ + 1);                         // [NBQ][NBQ+1] R entries of the block (CTA 0)
    T *G = Rs + NBQ * (NBQ + 1);                          // [NBQ][NBQ+1] V^T V
    T *Gp = G + NBQ * (NBQ + 1);                          // [NBQ][NBQ] per-CTA partial of V^T V
    T *s_tau = Gp + NBQ * NBQ;                            // [NBQ]
    T *tcol = s_tau + NBQ;                                // [NBQ]
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int gr0 = crank * PR_ROWS + t, gr1 = gr0 + PR_THREADS;
    const bool have0 = gr0 < m, have1 = gr1 < m;
    T r0v[NBQ], r1v[NBQ];
#pragma unroll
    for (int c = 0; c < NBQ; ++c) {
        r0v[c] = (have0 && c < jb) ? q.A[(size_t) gr0 + (size_t) c * lda] : T(0);
        r1v[c] = (have1 && c < jb) ? q.A[(size_t) gr1 + (size_t) c * lda] : T(0);
    }
    for (int idx = t; idx < 3 * NBQ * (NBQ + 1); idx += PR_THREADS) Ts[idx] = T(0);  // Ts, Rs, G
    cluster.sync();

    for (int j = 0; j < jb; ++j) {
        const int par = j & 1;
        const T x0 = (have0 && gr0 > j) ? r0v[0] : T(0);
        const T x1 = (have1 && gr1 > j) ? r1v[0] : T(0);
        T tot;
        {
            T pr[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pr[c] = fma(x0, r0v[c], x1 * r1v[c]);
            const T lo = warp_reduce_16<T>(pr, lane);
#pragma unroll
            for (int c = 0; c < 16; ++c) pr[c] = fma(x0, r0v[c + 16], x1 * r1v[c + 16]);
            const T hi = warp_reduce_16<T>(pr, lane);
            tot = (lane & 1) ? hi : lo;  // even lanes: slot lane/2, odd lanes: slot 16 + lane/2
        }
        const int slot = (lane >> 1) + 16 * (lane & 1);
        T *zdst = zpart + ((par * PR_MAXCS + crank) * NW + w) * NBQ + slot;
        for (int rk = 0; rk < CS; ++rk) *cluster.map_shared_rank(zdst, rk) = tot;
        if (gr0 == j) {  // pivot row: slot 0 = alpha, slots 1.. = A[j][j+1..]
            for (int rk = 0; rk < CS; ++rk) {
                T *zr = cluster.map_shared_rank(zrow + par * NBQ, rk);
#pragma unroll
                for (int c = 0; c < NBQ; ++c) zr[c] = r0v[c];
            }
        }
        cluster.sync();
        if (t < NBQ) {  // four independent partial sums: this stretch is on the critical path of every column
            T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
            for (int src = 0; src < CS; ++src) {
                const T *zp = zpart + ((par * PR_MAXCS + src) * NW) * NBQ + t;
#pragma unroll
                for (int ww = 0; ww < NW; ww += 4) {
                    s0 += zp[(ww + 0) * NBQ]; s1 += zp[(ww + 1) * NBQ]; s2 += zp[(ww + 2) * NBQ]; s3 += zp[(ww + 3) * NBQ];
                }
            }
            zsum[t] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        const T ss = zsum[0], alpha = zrow[par * NBQ];
        T tau = T(0), scale = T(0), beta = alpha;
        if (ss != T(0)) {
            const T h = fma(alpha, alpha, ss);
            if (h > T(1e-30) && h < T(1e30)) {
                // one reciprocal square root gives both the norm and 1/beta; 1/(alpha - beta) = sign / (|alpha| + norm) from a
                // second one (the library sqrt + two divisions were ~500 cycles of dependent instructions per column)
                const T rs = pq_rsqrt(h), nrm = h * rs;
                beta = alpha >= T(0) ? -nrm : nrm;
                tau = (beta - alpha) * (alpha >= T(0) ? -rs : rs);
                const T rx = pq_rsqrt(t_abs(alpha) + nrm);
                scale = alpha >= T(0) ? rx * rx : -(rx * rx);
            } else {
                const T nrm = t_sqrt(h);
                beta = alpha >= T(0) ? -nrm : nrm;
                tau = (beta - alpha) / beta;
                scale = T(1) / (alpha - beta);
            }
        }
        // reflector entries of my rows (unit diagonal, zeros above) -> shared memory; trailing update + slot shift
        const T v0 = (gr0 > j) ? x0 * scale : (gr0 == j ? T(1) : T(0));
        const T v1 = x1 * scale;
        P[pr_addr(t, j)] = v0;
        P[pr_addr(t + PR_THREADS, j)] = v1;
        const T tv0 = (have0 && gr0 >= j) ? tau * v0 : T(0), tv1 = tau * v1;
        if (gr0 == j) {  // my row becomes row j of R
            Rs[j * (NBQ + 1) + j] = beta;
#pragma unroll
            for (int c = 1; c < NBQ; ++c) {
                const T wc = fma(scale, zsum[c], zrow[par * NBQ + c]);
                if (j + c < jb) Rs[j * (NBQ + 1) + j + c] = fma(-tau, wc, r0v[c]);
            }
        }
#pragma unroll
        for (int c = 1; c < NBQ; ++c) {
            const T wc = fma(scale, zsum[c], zrow[par * NBQ + c]);  // v^T A[:, j + c]
            r0v[c - 1] = fma(-tv0, wc, r0v[c]);
            r1v[c - 1] = fma(-tv1, wc, r1v[c]);
        }
        r0v[NBQ - 1] = T(0);
        r1v[NBQ - 1] = T(0);
        if (t == 0) s_tau[j] = tau;
        // zsum is rewritten only after the next cluster barrier; zrow / zpart are double buffered
    }
    for (int j = jb; j < NBQ; ++j) {  // unused reflector columns are zero
        P[pr_addr(t, j)] = T(0);
        P[pr_addr(t + PR_THREADS, j)] = T(0);
    }
    __syncthreads();
    // ---- G = V^T V: per-CTA partial over the local rows, cluster sum in CTA 0
    if constexpr (std::is_same<T, double>::value) {
        // DMMA: warp w owns the 8x8 tiles (ti, tj0) and (ti, tj0 + 1), K = 512 local rows
        const int g = lane >> 2, tq = lane & 3, ti = w >> 1, tj0 = 2 * (w & 1);
        double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        const int colA = 8 * ti + g, colB0 = 8 * tj0 + g, colB1 = colB0 + 8;
        const double *pa = P + colA * PR_ROWS, *pb0 = P + colB0 * PR_ROWS, *pb1 = P + colB1 * PR_ROWS;
        const int ra = 4 * colA + tq, rb0 = 4 * colB0 + tq, rb1 = 4 * colB1 + tq;
#pragma unroll 4
        for (int ks = 0; ks < PR_ROWS / 4; ks += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = 4 * (ks + h);
                const double a = pa[(rr + ra) & (PR_ROWS - 1)];
                const double b0 = pb0[(rr + rb0) & (PR_ROWS - 1)];
                const double b1 = pb1[(rr + rb1) & (PR_ROWS - 1)];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[0][h][0]), "+d"(acc[0][h][1]) : "d"(a), "d"(b0));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[1][h][0]), "+d"(acc[1][h][1]) : "d"(a), "d"(b1));
            }
        }
#pragma unroll
        for (int jx = 0; jx < 2; ++jx)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                Gp[(8 * ti + g) * NBQ + 8 * (tj0 + jx) + 2 * tq + h] = acc[jx][0][h] + acc[jx][1][h];
    } else {
        for (int qd = 0; qd < 4; ++qd) {
            const int idx = t + qd * PR_THREADS, i = idx / NBQ, jj = idx % NBQ;
            T acc = T(0);
            if (i < jj && jj < jb) {
                const T *pi = P + i * PR_ROWS, *pj = P + jj * PR_ROWS;
                const int ri = 4 * i, rj = 4 * jj;
                for (int r = 0; r < PR_ROWS; ++r) acc = fma(pi[(r + ri) & (PR_ROWS - 1)], pj[(r + rj) & (PR_ROWS - 1)], acc);
            }
            Gp[idx] = acc;
        }
    }
    cluster.sync();
    if (crank == 0) {
        for (int qd = 0; qd < 4; ++qd) {
            const int idx = t + qd * PR_THREADS, i = idx / NBQ, jj = idx % NBQ;
            T sacc = T(0);
            for (int rk = 0; rk < CS; ++rk) sacc += cluster.map_shared_rank(Gp, rk)[idx];
            G[i * (NBQ + 1) + jj] = sacc;
        }
    }
    cluster.sync();  // remote reads of Gp are done; nobody exits early
    // ---- write back the reflectors (A below the diagonal + clean copy)
#pragma unroll 1
    for (int c = 0; c < jb; ++c) {
        const T a0 = P[pr_addr(t, c)], a1 = P[pr_addr(t + PR_THREADS, c)];
        if (have0) {
            lp.Vc[(size_t) gr0 + (size_t) c * lp.ldvc] = a0;
            if (gr0 > c) q.A[(size_t) gr0 + (size_t) c * lda] = a0;
        }
        if (have1) {
            lp.Vc[(size_t) gr1 + (size_t) c * lp.ldvc] = a1;
            q.A[(size_t) gr1 + (size_t) c * lda] = a1;
        }
    }
    if (crank != 0) return;
    __syncthreads();
    // R entries (rows 0..jb-1, columns >= row) and tau
    for (int idx = t; idx < NBQ * NBQ; idx += PR_THREADS) {
        const int i = idx % NBQ, c = idx / NBQ;
        if (i <= c && c < jb && i < m) q.A[(size_t) i + (size_t) c * lda] = Rs[i * (NBQ + 1) + c];
    }
    if (t < jb) q.tau[t] = s_tau[t];
    if (w == 0) {
        for (int i = 0; i < jb; ++i) {
            const T ti = s_tau[i];
            if (lane < i) tcol[lane] = -ti * G[lane * (NBQ + 1) + i];
            __syncwarp();
            T acc = T(0);
            if (lane < i)
                for (int c = lane; c < i; ++c) acc = fma(Ts[lane * (NBQ + 1) + c], tcol[c], acc);
            __syncwarp();
            if (lane < i) Ts[lane * (NBQ + 1) + i] = acc;
            if (lane == i) Ts[i * (NBQ + 1) + i] = ti;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = t; idx < NBQ * NBQ; idx += PR_THREADS) lp.Tm[idx] = Ts[(idx % NBQ) * (NBQ + 1) + idx / NBQ];
}

}  // namespace hcb
