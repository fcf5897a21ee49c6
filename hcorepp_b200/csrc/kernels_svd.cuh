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

// One Jacobi rotation of the column pair (mx, my) of length a, done by one warp. Returns true if it rotated.
// noise2: columns whose squared norm is <= noise2 are left alone (currently 0: only exactly-zero columns -- a larger
// floor leaves the vectors of negligible singular values non-orthogonal, which the compat SVD entry point must not do).
// (de Rijk's norm-ordering swaps were tried and made the round-robin ordering converge SLOWER on
// the recompression cores -- 29 vs 18 sweeps in the numpy emulation -- so they are not used; what halves the sweep
// count is the LQ preconditioning done before this kernel, see k_extract_l.)
template<typename T>
__device__ __forceinline__ bool jacobi_rotate(T *__restrict__ mx, T *__restrict__ my, int a, int lane, T tol, T noise2) {
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
    if (!(t_abs(gamma) > lim) || gamma == T(0) || !(alpha > noise2) || !(beta > noise2)) return false;
    const T zeta = (beta - alpha) / (T(2) * gamma);
    const T t = (zeta >= T(0) ? T(1) : T(-1)) / (t_abs(zeta) + t_sqrt(fma(zeta, zeta, T(1))));
    const T c = T(1) / t_sqrt(fma(t, t, T(1)));
    const T s = c * t;
    for (int i = lane; i < a; i += 32) {
        const T u = mx[i], v = my[i];
        mx[i] = fma(-s, v, c * u);
        my[i] = fma(s, u, c * v);
    }
    return true;
}

// Register-resident variant of the rotation for columns stored with a zero-padded pitch of 64*NI rows (16-byte
// aligned): lane l owns rows {64 i + 2 l, 64 i + 2 l + 1}, i < NI.  Both columns are read ONCE with 128-bit shared
// loads, kept in registers across the reductions, rotated and written back with 128-bit stores -- the ncu profile of
// the loop-based version showed the FP64 pipe only 34 % busy with 63 % of the issued instructions being LDS/STS/loop
// overhead (profiles/r01_jacobi_full_summary.txt); this form issues ~7 FP64 instructions per 2 memory instructions.
template<typename T> struct alignas(2 * sizeof(T)) Vec2 { T x, y; };

template<typename T> __device__ __forceinline__ T t_rsqrt(T x);
template<> __device__ __forceinline__ double t_rsqrt(double x) { return rsqrt(x); }
template<> __device__ __forceinline__ float t_rsqrt(float x) { return rsqrtf(x); }

// nx2 / ny2: CACHED squared column norms (shared memory), so that only the inner product gamma has to be formed
// (LAPACK xGESVJ does the same): alpha' = alpha - t*gamma, beta' = beta + t*gamma, recomputed from the registers when
// the update cancels badly.  The rotation scalars use one reciprocal square root each instead of divisions and square
// roots:  t = 2*gamma / (d + sign(d)*sqrt(d^2 + 4*gamma^2)),  d = beta - alpha;  c = rsqrt(1 + t^2);  s = c*t.
template<typename T, int NI>
__device__ __forceinline__ int jacobi_rotate_reg(T *__restrict__ mx, T *__restrict__ my, T *nx2, T *ny2, int lane, T tol2,
                                                 T big2 = T(3.0e38)) {  // returns 1 if rotated, | 2 if cos^2 > big2
    Vec2<T> u[NI], v[NI];
    T gamma = T(0);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        u[i] = *reinterpret_cast<const Vec2<T> *>(mx + 64 * i + 2 * lane);
        v[i] = *reinterpret_cast<const Vec2<T> *>(my + 64 * i + 2 * lane);
        gamma = fma(u[i].x, v[i].x, gamma);
        gamma = fma(u[i].y, v[i].y, gamma);
    }
    const T alpha = *nx2, beta = *ny2;
    gamma = warp_sum(gamma);
    // |gamma| > tol * sqrt(alpha * beta)  <=>  gamma^2 > tol^2 * alpha * beta
    if (!(gamma * gamma > tol2 * alpha * beta)) return 0;
    const int ret = (gamma * gamma > big2 * alpha * beta) ? 3 : 1;
    const T d = beta - alpha, g2 = gamma + gamma;
    const T h = fma(d, d, g2 * g2);          // > 0 because gamma != 0 here
    const T den = t_abs(d) + h * t_rsqrt(h);  // |d| + sqrt(d^2 + 4 gamma^2) > 0
    const T rd = t_rsqrt(den);
    const T t = (d >= T(0) ? g2 : -g2) * (rd * rd);
    const T c = t_rsqrt(fma(t, t, T(1)));
    const T s = c * t;
    T na = T(0), nb = T(0);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        Vec2<T> nu, nv;
        nu.x = fma(-s, v[i].x, c * u[i].x); nu.y = fma(-s, v[i].y, c * u[i].y);
        nv.x = fma(s, u[i].x, c * v[i].x);  nv.y = fma(s, u[i].y, c * v[i].y);
        *reinterpret_cast<Vec2<T> *>(mx + 64 * i + 2 * lane) = nu;
        *reinterpret_cast<Vec2<T> *>(my + 64 * i + 2 * lane) = nv;
        u[i] = nu;
        v[i] = nv;
    }
    const T tg = t * gamma;
    T a2 = alpha - tg, b2 = beta + tg;
    if (a2 < T(0.01) * alpha || b2 < T(0.01) * beta) {  // cancellation: recompute both norms from the registers
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            na = fma(u[i].x, u[i].x, na); na = fma(u[i].y, u[i].y, na);
            nb = fma(v[i].x, v[i].x, nb); nb = fma(v[i].y, v[i].y, nb);
        }
        a2 = warp_sum(na);
        b2 = warp_sum(nb);
    }
    __syncwarp();  // every lane has read the cached norms before lane 0 replaces them (racecheck: write-after-read)
    if (lane == 0) { *nx2 = a2; *ny2 = b2; }
    return ret;
}

// Epilogue shared by the Jacobi kernels: singular values = column norms of the rotated copy M (ld ldm), rank-sort
// (descending, stable) and scatter of the normalised left factor.  sig: b elements of shared memory.
template<typename T, bool L2 = false>  // L2: read M with ld.global.cg (written by other SMs in earlier sweeps)
__device__ void jacobi_finish(const T *M, int ldm, T *sig, const SvdProb<T> &p) {
    auto ld = [](const T *q) { return L2 ? __ldcg(q) : *q; };
    const int a = p.a, b = p.b;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
    (void) tid;
    // singular values = column norms
    for (int c = w; c < b; c += nw) {
        const T *mc = M + (size_t) c * ldm;
        T ss = T(0);
        for (int i = lane; i < a; i += 32) { const T v = ld(mc + i); ss = fma(v, v, ss); }
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
        for (int i = lane; i < a; i += 32) uo[i] = (sc > T(0)) ? ld(mc + i) / sc : T(0);
    }
}

// The sweeps.  NI > 0: shared-memory columns have pitch 64*NI (zero padded), rotations are register resident.
// NI == 0: generic form, pitch = a, loop-based rotations (any size; also the global-memory fallback).
//
// M (a x b, a >= b) = Uout diag(sigma) V^T; only the LEFT factor is produced here: the rotations are not accumulated.
// The caller gets the scaled right factor V diag(sigma) = M^T Uout with one batched GEMM afterwards (exactly the
// quantity the recompression needs, Compressed.cpp:598-622), which halves the Jacobi work and its shared-memory
// footprint.  p.M is left untouched.
//
// Three regimes, chosen per problem from its true size (dynamic shared memory = smem_elems elements of T):
//   (A) pitch*b + b fits        : the whole matrix lives in shared memory, cyclic (round-robin) one-sided Jacobi;
//   (B) two column blocks fit   : BLOCK one-sided Jacobi -- the rotated copy of M lives in global memory (p.J, L2
//                                 resident), pairs of w-column blocks are staged in shared memory, all w*w cross pairs
//                                 (and, once per sweep, the pairs inside each block) are rotated there, blocks go back;
//   (C) otherwise               : rotations directly on the global copy (slow, correctness-only fallback).
template<typename T, int NI>
__device__ void jacobi_sweeps(T *sm, const SvdProb<T> &p, int smem_elems, int max_sweeps) {
    __shared__ int s_rot;
    const int a = p.a, b = p.b;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
    const int P = NI > 0 ? 64 * NI : a;  // shared-memory column pitch
    // shared-memory budget: columns + sigma (b) + cached squared norms (register form only: b, or 2*bw when blocked)
    const size_t nrm_a = NI > 0 ? (size_t) b : 0;
    const bool fits = (size_t) P * b + (size_t) b + nrm_a <= (size_t) smem_elems;
    int bw = 0;  // block width of regime (B)
    if (!fits) {
        bw = 32;
        while (bw >= 2 && (size_t) 2 * P * bw + (size_t) b + (NI > 0 ? (size_t) 2 * bw : 0) > (size_t) smem_elems) bw >>= 1;
        if (bw < 2) bw = 0;
    }
    const bool in_smem = fits, blocked = !fits && bw > 0;
    T *M = in_smem ? sm : p.J;              // the rotated copy
    const int ldm = in_smem ? P : a;
    T *sig = in_smem ? sm + (size_t) P * b : (blocked ? sm + (size_t) 2 * P * bw : sm);
    T *nrm = sig + b;  // cached squared norms (NI > 0 only): b entries in regime A, 2*bw in regime B
    if (in_smem) {
        for (int idx = tid; idx < P * b; idx += nthr) {
            const int i = idx % P, c = idx / P;
            M[idx] = i < a ? p.M[(size_t) i + (size_t) c * p.ldm] : T(0);
        }
    } else {
        for (int idx = tid; idx < a * b; idx += nthr) M[idx] = p.M[(size_t) (idx % a) + (size_t) (idx / a) * p.ldm];
    }
    __syncthreads();

    const T tol = Eps<T>::v() * t_sqrt((T) a);  // threshold on |cos(angle)| of a column pair
    const T tol2 = tol * tol;
    auto rotate = [&](T *cx, T *cy, T *nx, T *ny) -> bool {
        if constexpr (NI > 0) return jacobi_rotate_reg<T, NI>(cx, cy, nx, ny, lane, tol2);
        else return jacobi_rotate<T>(cx, cy, a, lane, tol, T(0));
    };
    // squared norms of `cnt` staged columns (pitch P) into out[]
    auto norms_of = [&](const T *cols, int cnt, T *out) {
        for (int c = w; c < cnt; c += nw) {
            const T *mc = cols + (size_t) c * P;
            T ss = T(0);
            for (int i = lane; i < P; i += 32) ss = fma(mc[i], mc[i], ss);
            ss = warp_sum(ss);
            if (lane == 0) out[c] = ss;
        }
    };
    bool converged = (b < 2);
    int sweeps_used = 0;
    if (!blocked) {
        // ---- regimes (A) and (C): cyclic one-sided Jacobi over all column pairs, one warp per pair.
        // (C) with NI > 0 cannot use the padded register form on the unpadded global copy -> generic rotation.
        const int nb2 = (b + 1) & ~1;
        for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
            ++sweeps_used;
            __syncthreads();
            if (tid == 0) s_rot = 0;
            if (NI > 0 && in_smem) norms_of(M, b, nrm);  // refresh the cached norms once per sweep
            __syncthreads();
            for (int round = 0; round < nb2 - 1; ++round) {
                for (int slot = w; slot < nb2 / 2; slot += nw) {
                    int x, y;
                    rr_pair(nb2, round, slot, x, y);
                    if (y >= b) continue;  // dummy player (odd b)
                    T *cx = M + (size_t) x * ldm, *cy = M + (size_t) y * ldm;
                    const bool r = in_smem ? rotate(cx, cy, nrm + x, nrm + y) : jacobi_rotate<T>(cx, cy, a, lane, tol, T(0));
                    if (r && lane == 0) s_rot = 1;
                }
                __syncthreads();
            }
            converged = (s_rot == 0);
        }
    } else {
        // ---- regime (B): block one-sided Jacobi, row-cyclic over column blocks: block I stays in shared memory while
        // every later block J streams through (stage J, rotate the bw x bw cross pairs, write J back).
        T *BA = sm, *BB = sm + (size_t) P * bw;
        const int nblk = (b + bw - 1) / bw;
        auto stage = [&](T *dst, int c0, int wc) {
            for (int idx = tid; idx < P * wc; idx += nthr) {
                const int i = idx % P, c = idx / P;
                dst[idx] = i < a ? M[(size_t) (c0 + c) * a + i] : T(0);
            }
        };
        auto unstage = [&](const T *src, int c0, int wc) {
            for (int idx = tid; idx < P * wc; idx += nthr) {
                const int i = idx % P, c = idx / P;
                if (i < a) M[(size_t) (c0 + c) * a + i] = src[idx];
            }
        };
        for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
            ++sweeps_used;
            __syncthreads();
            if (tid == 0) s_rot = 0;
            __syncthreads();
            for (int bi = 0; bi < nblk; ++bi) {
                const int ci0 = bi * bw, wi = min(bw, b - ci0);
                stage(BA, ci0, wi);
                __syncthreads();
                if (NI > 0) {
                    norms_of(BA, wi, nrm);
                    __syncthreads();
                }
                // pairs inside block I (round-robin over its wi columns)
                const int nu2 = (wi + 1) & ~1;
                for (int round = 0; round < nu2 - 1; ++round) {
                    for (int slot = w; slot < nu2 / 2; slot += nw) {
                        int x, y;
                        rr_pair(nu2, round, slot, x, y);
                        if (y >= wi) continue;
                        if (rotate(BA + (size_t) x * P, BA + (size_t) y * P, nrm + x, nrm + y) && lane == 0) s_rot = 1;
                    }
                    __syncthreads();
                }
                for (int bj = bi + 1; bj < nblk; ++bj) {
                    const int cj0 = bj * bw, wj = min(bw, b - cj0);
                    stage(BB, cj0, wj);
                    __syncthreads();
                    if (NI > 0) {
                        norms_of(BB, wj, nrm + bw);
                        __syncthreads();
                    }
                    // cross pairs: round t pairs column i of BA with column (i + t) mod bw of BB
                    for (int t = 0; t < bw; ++t) {
                        for (int i = w; i < wi; i += nw) {
                            const int j = (i + t) % bw;
                            if (j >= wj) continue;
                            if (rotate(BA + (size_t) i * P, BB + (size_t) j * P, nrm + i, nrm + bw + j) && lane == 0) s_rot = 1;
                        }
                        __syncthreads();
                    }
                    unstage(BB, cj0, wj);
                    __syncthreads();
                }
                unstage(BA, ci0, wi);
                __syncthreads();
            }
            converged = (s_rot == 0);
        }
    }
    if (p.info && tid == 0) {
        if (!converged) atomicOr(p.info, 1);
        info_max_sweeps(p.info, sweeps_used);  // diagnostics: number of Jacobi sweeps in bits 8..15
    }
    __syncthreads();

    jacobi_finish<T>(M, ldm, sig, p);
}

// One CTA per problem.  THREADS = 1024 (64 registers/thread: register-resident rotations up to 128 rows) or 512
// (128 registers/thread: up to 512 rows); taller problems use the generic loop-based rotations.
template<typename T, int THREADS>
__global__ void __launch_bounds__(THREADS) k_jacobi_svd(const SvdProb<T> *__restrict__ probs, int smem_elems,
                                                        int max_sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const SvdProb<T> p = probs[blockIdx.x];
    if (p.a <= 0 || p.b <= 0) return;
    const int ni = (p.a + 63) / 64;
    if constexpr (THREADS == 1024) {
        switch (ni) {
            case 1: jacobi_sweeps<T, 1>(sm, p, smem_elems, max_sweeps); break;
            case 2: jacobi_sweeps<T, 2>(sm, p, smem_elems, max_sweeps); break;
            default: jacobi_sweeps<T, 0>(sm, p, smem_elems, max_sweeps); break;
        }
    } else {
        switch (ni) {
            case 1: jacobi_sweeps<T, 1>(sm, p, smem_elems, max_sweeps); break;
            case 2: jacobi_sweeps<T, 2>(sm, p, smem_elems, max_sweeps); break;
            case 3: jacobi_sweeps<T, 3>(sm, p, smem_elems, max_sweeps); break;
            case 4: jacobi_sweeps<T, 4>(sm, p, smem_elems, max_sweeps); break;
            case 5: jacobi_sweeps<T, 5>(sm, p, smem_elems, max_sweeps); break;
            case 6: jacobi_sweeps<T, 6>(sm, p, smem_elems, max_sweeps); break;
            case 7: jacobi_sweeps<T, 7>(sm, p, smem_elems, max_sweeps); break;
            case 8: jacobi_sweeps<T, 8>(sm, p, smem_elems, max_sweeps); break;
            default: jacobi_sweeps<T, 0>(sm, p, smem_elems, max_sweeps); break;
        }
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
