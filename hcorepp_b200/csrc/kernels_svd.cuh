// kernels_svd.cuh -- batched one-sided (Hestenes) Jacobi SVD with the matrix held in shared memory when it fits.
//
// Replaces cusolverDnXgesvd as called per tile by the reference (src/kernels/cuda/CudaKernels.cu:699-730) for the
// (kc+ka) x (kc+ka) recompression core (src/operators/concrete/Compressed.cpp:477-480) and for the initial
// compression of a dense tile (Compressed.cpp:99-102).  One-sided Jacobi computes small singular values to high
// RELATIVE accuracy, which is what the absolute-threshold rank rule (omp/kernels.cpp:97-102) needs.
#pragma once
#include "common.cuh"

namespace hcb {

// Round-robin (tournament) pairing: nb2 players (even), round in [0, nb2-1), slot in [0, nb2/2).
__device__ __forceinline__ void rr_pair(int nb2, int round, int slot, int &x, int &y) {
    const int mod = nb2 - 1;
    if (slot == 0) {
        x = mod;
        y = round % mod;
    } else {
        x = (round + slot) % mod;
        y = (round - slot + mod) % mod;
    }
    if (x > y) { const int t = x; x = y; y = t; }
}

// One CTA per problem. M (a x b, a >= b) = Uout diag(sigma) V^T; only the LEFT factor is produced here:
// the rotations are not accumulated.  The caller gets the scaled right factor V diag(sigma) = M^T Uout with one
// batched GEMM afterwards (exactly the quantity the recompression needs, Compressed.cpp:598-622), which halves the
// Jacobi work and its shared-memory footprint.  p.M is left untouched.
// dynamic shared memory: smem_elems elements of T. Layout when the problem fits: [M a*b | sig b]; otherwise only
// [sig b] lives in shared memory and the rotations work on a global copy of M (p.J, a*b elements).
template<typename T>
__global__ void __launch_bounds__(1024) k_jacobi_svd(const SvdProb<T> *__restrict__ probs, int smem_elems,
                                                     int max_sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    __shared__ int s_rot;
    const SvdProb<T> p = probs[blockIdx.x];
    const int a = p.a, b = p.b;
    if (a <= 0 || b <= 0) return;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
    const bool fits = (size_t) a * b + (size_t) b <= (size_t) smem_elems;
    T *M, *sig;
    int ldm;
    if (fits) {
        M = sm;
        sig = sm + (size_t) a * b;
        ldm = a;
    } else {
        M = p.J;
        sig = sm;
        ldm = a;
    }
    for (int idx = tid; idx < a * b; idx += nthr) M[idx] = p.M[(size_t) (idx % a) + (size_t) (idx / a) * p.ldm];
    __syncthreads();

    const T tol = Eps<T>::v() * t_sqrt((T) a);
    const int nb2 = (b + 1) & ~1;
    bool converged = (b < 2);
    int sweeps_used = 0;
    for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
        ++sweeps_used;
        __syncthreads();
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < nb2 - 1; ++round) {
            for (int slot = w; slot < nb2 / 2; slot += nw) {
                int x, y;
                rr_pair(nb2, round, slot, x, y);
                if (y >= b) continue;  // dummy player (odd b)
                T *mx = M + (size_t) x * ldm, *my = M + (size_t) y * ldm;
                T alpha = T(0), beta = T(0), gamma = T(0);
                for (int i = lane; i < a; i += 32) {
                    const T u = mx[i], v = my[i];
                    alpha = fma(u, u, alpha);
                    beta = fma(v, v, beta);
                    gamma = fma(u, v, gamma);
                }
                alpha = warp_sum(alpha);
                beta = warp_sum(beta);
                gamma = warp_sum(gamma);
                const T lim = tol * t_sqrt(alpha) * t_sqrt(beta);
                if (t_abs(gamma) > lim && gamma != T(0)) {
                    if (lane == 0) s_rot = 1;
                    const T zeta = (beta - alpha) / (T(2) * gamma);
                    const T t = (zeta >= T(0) ? T(1) : T(-1)) / (t_abs(zeta) + t_sqrt(fma(zeta, zeta, T(1))));
                    const T c = T(1) / t_sqrt(fma(t, t, T(1)));
                    const T s = c * t;
                    for (int i = lane; i < a; i += 32) {
                        const T u = mx[i], v = my[i];
                        mx[i] = c * u - s * v;
                        my[i] = s * u + c * v;
                    }
                }
            }
            __syncthreads();
        }
        converged = (s_rot == 0);
    }
    if (p.info && tid == 0) {
        if (!converged) atomicOr(p.info, 1);
        atomicOr(p.info, sweeps_used << 8);  // diagnostics: number of Jacobi sweeps in bits 8..15
    }

    // singular values = column norms
    for (int c = w; c < b; c += nw) {
        const T *mc = M + (size_t) c * ldm;
        T ss = T(0);
        for (int i = lane; i < a; i += 32) ss = fma(mc[i], mc[i], ss);
        ss = warp_sum(ss);
        if (lane == 0) sig[c] = t_sqrt(ss);
    }
    __syncthreads();
    // rank-sort (descending, stable) and scatter the normalised / permuted left factor
    for (int c = w; c < b; c += nw) {
        const T sc = sig[c];
        int pos = 0;
        for (int o = lane; o < b; o += 32) {
            const T so = sig[o];
            pos += (so > sc || (so == sc && o < c)) ? 1 : 0;
        }
        pos = warp_sum(pos);
        if (lane == 0) p.sigma[pos] = sc;
        const T *mc = M + (size_t) c * ldm;
        T *uo = p.Uout + (size_t) pos * p.ldu;
        for (int i = lane; i < a; i += 32) uo[i] = (sc > T(0)) ? mc[i] / sc : T(0);
    }
}

// Vout[:, c] /= sigma[c] (zero where sigma == 0): turns V diag(sigma) = M^T Uout into the orthonormal right factor.
template<typename T>
__global__ void k_unscale_cols(T *__restrict__ V, int ld, int rows, int cols, const T *__restrict__ sigma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < rows && j < cols) {
        const T s = sigma[j];
        V[(size_t) i + (size_t) j * ld] = s > T(0) ? V[(size_t) i + (size_t) j * ld] / s : T(0);
    }
}

}  // namespace hcb
