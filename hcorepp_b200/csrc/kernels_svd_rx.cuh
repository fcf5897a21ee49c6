// kernels_svd_rx.cuh -- block one-sided Jacobi with the pivot block in REGISTERS (a <= 384 rows, any b).
//
// The shared-memory block Jacobi of kernels_svd.cuh reads and writes both columns of every pair from shared memory:
// 4 * 8 * a bytes per rotation, i.e. ~96 cycles of the SM's 128 B/clk shared-memory port per rotation at a = 357 --
// that, not the FP64 pipe (~50 cycles), is what its ~110 cycles per rotation were made of (ncu r01_jacobi_full).
// Here the 32 columns of the pivot block I live in the registers of the 16 warps (2 columns each) for the whole pass
// over the later blocks J; a warp loads TWO columns of J, rotates them against its two x columns (4 rotations, as two
// steps of two independent pairs, so two dependency chains are in flight per warp) and stores them back:
// 4 * 8 * a bytes of shared-memory traffic per 4 rotations.  The rotations are "fast" (scaled) rotations in the
// SEQUENTIAL (in-place) form
//     x' = x - (t dy/dx) y,   y' = y + (t c^2 dx/dy) x',   dx' = c dx,  dy' = dy / c,
// two FMAs per element instead of four and no copy of the old x (the simultaneous form y' = y + (t dx/dy) x_old made
// the compiler rotate the x registers through a spare slot: one MOV per DFMA, 9 % of the kernel's instructions --
// ncu source view, round 2); the per-column scales are folded back when a block leaves shared memory / registers.  Block J+1 is prefetched with cp.async into the other shared-memory region while block J is rotated.
// The pairs INSIDE a block are done once per sweep in shared memory with jacobi_rotate_reg (kernels_svd.cuh).
#pragma once
#include "common.cuh"
#include "kernels_svd.cuh"

namespace hcb {

// Geometry: RX_THREADS / 32 warps x 2 register-resident columns = RX_BW columns per block, one CTA per SM.
// (256 threads / 16-column blocks with two CTAs per SM was measured: 190 ms vs 162 ms per step -- twice the block
// staging per rotation outweighs the latency hiding.)
constexpr int RX_THREADS = 512;
constexpr int RX_BW = RX_THREADS / 16;  // block width (columns)
constexpr int RX_MAX_NI = 6;        // two columns per warp up to 64 * 6 = 384 rows
constexpr int RX_MAX_NI_TALL = 12;  // one column per warp up to 768 rows

// per-column metadata in shared memory: squared norm (true), scale d, 1/d
// t = 2 gamma / (d + sign(d) sqrt(d^2 + 4 gamma^2)) needs sqrt and a reciprocal; both from the MUFU approximations
// plus ONE Newton step (~1e-12 relative): an inexact angle only leaves a residual cosine of that size for the next
// sweep, orthogonality depends on c = rsqrt(1 + t^2) alone, which stays at full precision (and off the critical path:
// the column updates need only t).
__device__ __forceinline__ double rx_tangent(double d, double g2) {
    const double h = fma(d, d, g2 * g2);
    double rs, r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(h));
    double sh = h * rs;
    sh = fma(0.5 * rs, fma(-sh, sh, h), sh);
    const double den = fabs(d) + sh;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    return (d >= 0.0 ? g2 : -g2) * r;
}
__device__ __forceinline__ float rx_tangent(float d, float g2) {
    const float h = fmaf(d, d, g2 * g2);
    const float den = fabsf(d) + h * rsqrtf(h);
    return (d >= 0.0f ? g2 : -g2) / den;
}

template<typename T>
struct RxMeta {
    T *n2, *d, *id;
};

// full-precision reciprocal square root without the library's special-case path (argument is 1 + t^2 >= 1)
__device__ __forceinline__ double rx_rsqrt(double q) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    y = fma(0.5 * y, fma(-q * y, y, 1.0), y);
    y = fma(0.5 * y, fma(-q * y, y, 1.0), y);
    return y;
}
__device__ __forceinline__ float rx_rsqrt(float q) { return rsqrtf(q); }

template<typename T, int NI>
__device__ __forceinline__ T rx_sumsq(const Vec2<T> (&x)[NI]) {
    T s = T(0);
#pragma unroll
    for (int i = 0; i < NI; ++i) { s = fma(x[i].x, x[i].x, s); s = fma(x[i].y, x[i].y, s); }
    return warp_sum(s);
}

// Two independent pairs (xa, ya) and (xb, yb) at once; ix*/iy* index the metadata arrays.  The scalar part of the two
// rotations is PACKED: lanes 0-15 carry pair a, lanes 16-31 pair b through one branch-free instruction stream (the
// scalars were half of the instructions of a rotation when every lane computed both), then the four factors are
// broadcast.  Returns bit 0: something rotated; bit 1: some pair was still above big2 (see the sweep loop).
template<typename T, int NI>
__device__ __forceinline__ int rx_duo(Vec2<T> (&xa)[NI], int ixa, Vec2<T> (&ya)[NI], int iya, Vec2<T> (&xb)[NI], int ixb,
                                      Vec2<T> (&yb)[NI], int iyb, const RxMeta<T> &mt, int lane, T tol2, T big2) {
    T ga = T(0), gb = T(0), ga2 = T(0), gb2 = T(0);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        ga = fma(xa[i].x, ya[i].x, ga); ga2 = fma(xa[i].y, ya[i].y, ga2);
        gb = fma(xb[i].x, yb[i].x, gb); gb2 = fma(xb[i].y, yb[i].y, gb2);
    }
    ga += ga2;
    gb += gb2;
    const bool hi = lane & 16;
    // both warp sums with 5 shuffles: each half of the warp ends up with the total of ITS pair
    T g = (hi ? gb : ga) + __shfl_xor_sync(0xffffffffu, hi ? ga : gb, 16);
    g += __shfl_xor_sync(0xffffffffu, g, 8);
    g += __shfl_xor_sync(0xffffffffu, g, 4);
    g += __shfl_xor_sync(0xffffffffu, g, 2);
    g += __shfl_xor_sync(0xffffffffu, g, 1);
    const int ix = hi ? ixb : ixa, iy = hi ? iyb : iya;
    const T alpha = mt.n2[ix], beta = mt.n2[iy];
    const T dx = mt.d[ix], dy = mt.d[iy], idx = mt.id[ix], idy = mt.id[iy];
    const T gam = g * dx * dy, gg = gam * gam, ab = alpha * beta;
    const bool rot = gg > tol2 * ab;
    const bool big = gg > big2 * ab;
    const T d = rot ? beta - alpha : T(1), g2 = rot ? gam + gam : T(0);  // benign inputs where nothing rotates
    const T t = rx_tangent(d, g2);
    const T q = fma(t, t, T(1));
    const T c = rx_rsqrt(q);
    const T fx = t * dy * idx, fy = t * (c * c) * dx * idy;  // zero when !rot (t == 0)
    const T tg = t * gam, a2 = alpha - tg, b2 = beta + tg, rc = q * c;  // rc = 1 / c
    const bool redo = rot && ((a2 < T(0.01) * alpha) || (b2 < T(0.01) * beta));
    __syncwarp();
    if (rot && (lane & 15) == 0) {
        mt.n2[ix] = a2; mt.n2[iy] = b2;
        mt.d[ix] = c * dx; mt.d[iy] = rc * dy;
        mt.id[ix] = rc * idx; mt.id[iy] = c * idy;
    }
    __syncwarp();  // lanes 0 / 16 wrote the metadata that lane 0 may re-read in the (rare) redo path below
    const unsigned brot = __ballot_sync(0xffffffffu, rot), bredo = __ballot_sync(0xffffffffu, redo);
    const unsigned bbig = __ballot_sync(0xffffffffu, big);
    const T fxa = __shfl_sync(0xffffffffu, fx, 0), fya = __shfl_sync(0xffffffffu, fy, 0);
    const T fxb = __shfl_sync(0xffffffffu, fx, 16), fyb = __shfl_sync(0xffffffffu, fy, 16);
    if (brot & 1u) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            xa[i].x = fma(-fxa, ya[i].x, xa[i].x); xa[i].y = fma(-fxa, ya[i].y, xa[i].y);
            ya[i].x = fma(fya, xa[i].x, ya[i].x);  ya[i].y = fma(fya, xa[i].y, ya[i].y);
        }
    }
    if (brot & 0x10000u) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            xb[i].x = fma(-fxb, yb[i].x, xb[i].x); xb[i].y = fma(-fxb, yb[i].y, xb[i].y);
            yb[i].x = fma(fyb, xb[i].x, yb[i].x);  yb[i].y = fma(fyb, xb[i].y, yb[i].y);
        }
    }
    if (bredo & 1u) {  // rare: the norm update cancelled -- recompute the true squared norms d^2 * |stored|^2
        const T sx = rx_sumsq<T, NI>(xa), sy = rx_sumsq<T, NI>(ya);
        if (lane == 0) { mt.n2[ixa] = mt.d[ixa] * mt.d[ixa] * sx; mt.n2[iya] = mt.d[iya] * mt.d[iya] * sy; }
    }
    if (bredo & 0x10000u) {
        const T sx = rx_sumsq<T, NI>(xb), sy = rx_sumsq<T, NI>(yb);
        if (lane == 0) { mt.n2[ixb] = mt.d[ixb] * mt.d[ixb] * sx; mt.n2[iyb] = mt.d[iyb] * mt.d[iyb] * sy; }
    }
    __syncwarp();
    return (brot ? 1 : 0) | (bbig ? 2 : 0);
}

// One pair (x, y), every lane computes the scalars (used by the one-column-per-warp geometry of tall problems).
template<typename T, int NI>
__device__ __forceinline__ int rx_solo(Vec2<T> (&x)[NI], int ix, Vec2<T> (&y)[NI], int iy, const RxMeta<T> &mt, int lane, T tol2,
                                       T big2) {
    T g = T(0), g2a = T(0);
#pragma unroll
    for (int i = 0; i < NI; ++i) { g = fma(x[i].x, y[i].x, g); g2a = fma(x[i].y, y[i].y, g2a); }
    g = warp_sum(g + g2a);
    const T alpha = mt.n2[ix], beta = mt.n2[iy];
    const T dx = mt.d[ix], dy = mt.d[iy], idx = mt.id[ix], idy = mt.id[iy];
    const T gam = g * dx * dy, gg = gam * gam, ab = alpha * beta;
    if (!(gg > tol2 * ab)) return 0;
    const int ret = (gg > big2 * ab) ? 3 : 1;
    const T t = rx_tangent(beta - alpha, gam + gam);
    const T q = fma(t, t, T(1));
    const T c = rx_rsqrt(q);
    const T fx = t * dy * idx, fy = t * (c * c) * dx * idy;
    const T tg = t * gam, rc = q * c;
    T a2 = alpha - tg, b2 = beta + tg;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        x[i].x = fma(-fx, y[i].x, x[i].x); x[i].y = fma(-fx, y[i].y, x[i].y);
        y[i].x = fma(fy, x[i].x, y[i].x);  y[i].y = fma(fy, x[i].y, y[i].y);
    }
    const T ndx = c * dx, ndy = rc * dy;
    if (a2 < T(0.01) * alpha || b2 < T(0.01) * beta) {  // rare: recompute the true squared norms
        a2 = ndx * ndx * rx_sumsq<T, NI>(x);
        b2 = ndy * ndy * rx_sumsq<T, NI>(y);
    }
    __syncwarp();
    if (lane == 0) {
        mt.n2[ix] = a2; mt.n2[iy] = b2;
        mt.d[ix] = ndx; mt.d[iy] = ndy;
        mt.id[ix] = rc * idx; mt.id[iy] = c * idy;
    }
    __syncwarp();
    return ret;
}

// ONE sweep over all column pairs of problem p (the rotated copy lives in p.J between sweeps; `first` makes it from
// p.M).  Returns bit 0: something rotated, bit 1: some pair was above the predictive-stop level.  When `last_allowed`
// or the sweep converged, the epilogue (sigma, sort, scatter of the left factor, info) runs too and bit 2 is set.
// A sweep is cut into RX_PARTS work items (consecutive ranges of pivot blocks with about equal numbers of block pairs).
// The item's description lives in shared memory (the kernel is at its 128-register cap): part index, the rotated / big
// flags carried over from the earlier parts of this sweep.  A part that is not the last one returns bit 3 set and the
// flags so far in bits 0-1.
constexpr int RX_PARTS = 4;
struct RxItem { int part, carry, bi_lo, bi_hi; };
template<typename T, int NI, int XPW>  // XPW: register-resident columns per warp (2; 1 for tall problems, NI > 6)
__device__ int jacobi_sweep_rx(T *sm, const SvdProb<T> &p, int sweep, volatile RxItem *it, int max_sweeps, T stop2) {
    constexpr int P = 64 * NI, BW = (RX_THREADS / 32) * XPW;
    __shared__ int s_rot, s_big;
    const int a = p.a, b = p.b;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = RX_THREADS / 32;
    T *R0 = sm, *R1 = sm + (size_t) P * BW;      // two column-block regions
    T *sig = R1 + (size_t) P * BW;               // b singular values
    T *meta = sig + b;                           // 3 x 64 metadata: x columns [0,32), y columns [32,64)
    RxMeta<T> mt{meta, meta + 2 * BW, meta + 4 * BW};
    T *M = p.J;  // rotated copy (global / L2), EVEN ld: every column starts 16-byte aligned for the cp.async staging
    const int lda = (a + 1) & ~1;  // (p.J holds (a + 1) * b elements)
    if (sweep == 0 && it->part == 0) {
        for (int idx = tid; idx < a * b; idx += RX_THREADS) {
            const int i = idx % a, c = idx / a;
            M[(size_t) c * lda + i] = p.M[(size_t) i + (size_t) c * p.ldm];
        }
        if (lda > a)  // the pad row stays zero for the whole solve (nothing writes rows >= a)
            for (int c = tid; c < b; c += RX_THREADS) M[(size_t) c * lda + a] = T(0);
    }
    __syncthreads();
    const T tol = Eps<T>::v() * t_sqrt((T) a);
    const T tol2 = tol * tol;
    // Predictive stop: the sweep after one in which every pair was already below cos = sqrt(tol) would only verify
    // (quadratic convergence leaves residual cosines ~ tol), so it is skipped.
    // stop2 > 0 (fused path, opt-in): accuracy-aware stop -- end after the sweep whose max cos^2 is below the requested
    // compression accuracy (residual non-orthogonality ~0.05 * accuracy; see NOTES_NEXT_ROUND.md).  0: machine precision.
    const T big2 = stop2 > T(16) * tol ? stop2 : T(16) * tol;  // cos < 4 sqrt(tol) = 2.5e-7: the residual after that sweep is ~6e-14
    const int nblk = (b + BW - 1) / BW;

    // asynchronous staging of columns [c0, c0 + wc) into dst (pitch P, zero padded rows and columns)
    auto stage_async = [&](T *dst, int c0, int wc) {
        for (int q = tid; q < BW * P / 2; q += RX_THREADS) {
            const int c = q / (P / 2), row = 2 * (q % (P / 2));
            T *d = dst + (size_t) c * P + row;
            const T *src = M + (size_t) (c0 + c) * lda + row;
            const bool v0 = c < wc && row < a, v1 = c < wc && row + 1 < a;
            if (sizeof(T) == 8) {
                // even ld + zero pad row: every row pair of a live column is one aligned 16-byte copy (with ld = a, odd a
                // sent every other column through synchronous scalar loads: 3 % of the kernel's stall samples)
                if (v0) cp_async_16(d, src);
                else { d[0] = T(0); d[1] = T(0); }
            } else {
                d[0] = v0 ? __ldcg(src) : T(0);
                d[1] = v1 ? __ldcg(src + 1) : T(0);
            }
        }
        cp_async_commit();
    };
    // squared norms of the 32 staged columns -> metadata slots [m0, m0 + 32), scales reset to 1
    auto init_meta = [&](const T *cols, int m0) {
        for (int c = w; c < BW; c += NW) {
            const T *mc = cols + (size_t) c * P;
            T ss = T(0);
            for (int i = lane; i < P; i += 32) ss = fma(mc[i], mc[i], ss);
            ss = warp_sum(ss);
            if (lane == 0) { mt.n2[m0 + c] = ss; mt.d[m0 + c] = T(1); mt.id[m0 + c] = T(1); }
        }
    };
    // columns [c0, c0 + wc) of a staged block back to the global copy, scales folded in
    auto unstage = [&](const T *src, int c0, int wc, int m0) {
        for (int idx = tid; idx < P * wc; idx += RX_THREADS) {
            const int i = idx % P, c = idx / P;
            if (i < a) M[(size_t) (c0 + c) * lda + i] = src[idx] * mt.d[m0 + c];
        }
    };

    bool converged = (b < 2);
    if (!converged) {
        __syncthreads();
        if (tid == 0) {
            s_rot = it->carry & 1; s_big = (it->carry >> 1) & 1;
            // pivot blocks of this part: block bi costs (nblk - bi) block passes; the total is cut into RX_PARTS shares
            const int total = nblk * (nblk + 1) / 2, part = it->part;
            int acc = 0, q = 0, lo = nblk, hi = 0;
            for (int bi = 0; bi < nblk; ++bi) {  // block bi belongs to the first part whose share is not yet full
                while (q + 1 < RX_PARTS && acc * RX_PARTS >= (q + 1) * total) ++q;
                if (q == part) { lo = bi < lo ? bi : lo; hi = bi + 1; }
                acc += nblk - bi;
            }
            it->bi_lo = lo; it->bi_hi = hi;  // (lo = nblk, hi = 0: empty part)
        }
        __syncthreads();
        for (int bi = it->bi_lo; bi < it->bi_hi; ++bi) {
            const int ci0 = bi * BW, wi = min(BW, b - ci0);
            // block I -> R0 (and block I+1 -> R1 right behind it)
            stage_async(R0, ci0, wi);
            if (bi + 1 < nblk) stage_async(R1, ci0 + BW, min(BW, b - ci0 - BW));
            cp_async_wait_all();
            __syncthreads();
            init_meta(R0, 0);
            __syncthreads();
            // ---- pairs inside block I: round-robin in shared memory (plain rotations on the cached norms)
            {
                const int nu2 = (wi + 1) & ~1;
                for (int round = 0; round < nu2 - 1; ++round) {
                    for (int slot = w; slot < nu2 / 2; slot += NW) {
                        int x, y;
                        rr_pair(nu2, round, slot, x, y);
                        if (y >= wi) continue;
                        const int rr = jacobi_rotate_reg<T, NI>(R0 + (size_t) x * P, R0 + (size_t) y * P, mt.n2 + x, mt.n2 + y,
                                                                lane, tol2, big2);
                        if ((rr & 1) && lane == 0) s_rot = 1;
                        if ((rr & 2) && lane == 0) s_big = 1;
                    }
                    __syncthreads();
                }
            }
            if (bi + 1 == nblk) {  // last block: nothing to pair it with
                unstage(R0, ci0, wi, 0);
                __syncthreads();
                continue;
            }
            // ---- block I -> registers: warp w owns columns XPW w .. XPW w + XPW - 1
            Vec2<T> x0[NI], x1[XPW == 2 ? NI : 1];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                x0[i] = *reinterpret_cast<const Vec2<T> *>(R0 + (size_t) (XPW * w + 0) * P + 64 * i + 2 * lane);
                if constexpr (XPW == 2) x1[i] = *reinterpret_cast<const Vec2<T> *>(R0 + (size_t) (2 * w + 1) * P + 64 * i + 2 * lane);
            }
            __syncthreads();  // R0 is free from here on
            int any = 0;
            for (int bj = bi + 1; bj < nblk; ++bj) {
                const int k = bj - bi - 1;             // pass number: block J sits in R1 for even k, R0 for odd k
                T *BB = (k & 1) ? R0 : R1, *other = (k & 1) ? R1 : R0;
                const int cj0 = bj * BW, wj = min(BW, b - cj0);
                if (k > 0) {  // (pass 0's block was staged together with block I)
                    cp_async_wait_all();
                    __syncthreads();
                }
                if (bj + 1 < nblk) stage_async(other, cj0 + BW, min(BW, b - cj0 - BW));  // flies during this pass
                init_meta(BB, BW);
                __syncthreads();
                if constexpr (XPW == 2) {
                    for (int s = 0; s < BW / 2; ++s) {
                        const int q = (w + s) & (BW / 2 - 1), ja = 2 * q, jb = 2 * q + 1;
                        Vec2<T> ya[NI], yb[NI];
                        T *pa = BB + (size_t) ja * P + 2 * lane, *pb = BB + (size_t) jb * P + 2 * lane;
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            ya[i] = *reinterpret_cast<const Vec2<T> *>(pa + 64 * i);
                            yb[i] = *reinterpret_cast<const Vec2<T> *>(pb + 64 * i);
                        }
                        const int ix = 2 * w, iya = BW + ja, iyb = BW + jb;
                        any |= rx_duo<T, NI>(x0, ix + 0, ya, iya, x1, ix + 1, yb, iyb, mt, lane, tol2, big2);
                        any |= rx_duo<T, NI>(x1, ix + 1, ya, iya, x0, ix + 0, yb, iyb, mt, lane, tol2, big2);
#pragma unroll
                        for (int i = 0; i < NI; ++i) {
                            *reinterpret_cast<Vec2<T> *>(pa + 64 * i) = ya[i];
                            *reinterpret_cast<Vec2<T> *>(pb + 64 * i) = yb[i];
                        }
                        // (a per-pair hand-off between neighbouring warps -- mbarriers, then ticket counters -- was tried in
                        // place of this block barrier: correct with monotone tickets, but the polling warps cost more issue
                        // slots than the barrier stalls they removed: 176 ms vs 162 ms per step)
                        __syncthreads();
                    }
                } else {
                    for (int s = 0; s < BW; ++s) {  // one x column per warp, one y column per round
                        const int j = (w + s) & (BW - 1);
                        Vec2<T> y[NI];
                        T *py = BB + (size_t) j * P + 2 * lane;
#pragma unroll
                        for (int i = 0; i < NI; ++i) y[i] = *reinterpret_cast<const Vec2<T> *>(py + 64 * i);
                        any |= rx_solo<T, NI>(x0, w, y, BW + j, mt, lane, tol2, big2);
#pragma unroll
                        for (int i = 0; i < NI; ++i) *reinterpret_cast<Vec2<T> *>(py + 64 * i) = y[i];
                        __syncthreads();
                    }
                }
                unstage(BB, cj0, wj, BW);
                // fold the x scales so that they cannot drift far from 1
                {
                    const T d0 = mt.d[XPW * w], d1 = mt.d[XPW * w + XPW - 1];
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        x0[i].x *= d0; x0[i].y *= d0;
                        if constexpr (XPW == 2) { x1[i].x *= d1; x1[i].y *= d1; }
                    }
                    __syncwarp();
                    if (lane < XPW) { mt.d[XPW * w + lane] = T(1); mt.id[XPW * w + lane] = T(1); }
                }
                __syncthreads();
            }
            if ((any & 1) && lane == 0) s_rot = 1;
            if ((any & 2) && lane == 0) s_big = 1;
            // ---- block I: registers -> global copy (scales are 1 after the last fold)
            {
                auto put = [&](const Vec2<T> (&x)[NI], int c) {
                    if (c >= wi) return;
                    T *dst = M + (size_t) (ci0 + c) * lda;
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const int r = 64 * i + 2 * lane;
                        if (r < a) dst[r] = x[i].x;
                        if (r + 1 < a) dst[r + 1] = x[i].y;
                    }
                };
                put(x0, XPW * w);
                if constexpr (XPW == 2) put(x1, 2 * w + 1);
            }
            __syncthreads();
        }
        if (it->part + 1 < RX_PARTS) return 8 | (s_rot ? 1 : 0) | (s_big ? 2 : 0);  // the sweep goes on in the next part
        converged = (s_rot == 0) || (s_big == 0);
    } else if (it->part + 1 < RX_PARTS) {
        return 8;
    }
    const int flags = converged ? 0 : 3;
    if (!converged && sweep + 1 < max_sweeps) return flags;
    if (p.info && tid == 0) {
        if (!converged) atomicOr(p.info, 1);
        info_max_sweeps(p.info, b < 2 ? 0 : sweep + 1);
    }
    __syncthreads();
    jacobi_finish<T, true>(M, lda, sig, p);
    return flags | 4;
}

template<typename T>
constexpr size_t rx_smem_bytes(int ni, int b_bound) {  // two block regions (32 columns up to 384 rows, 16 columns above)
    const size_t wide = (size_t) 2 * 64 * (ni < RX_MAX_NI ? ni : RX_MAX_NI) * RX_BW, tall = ni > RX_MAX_NI ? (size_t) 2 * 64 * ni * (RX_BW / 2) : 0;
    return sizeof(T) * ((wide > tall ? wide : tall) + (size_t) b_bound + 6 * RX_BW);
}

// Persistent kernel: grid = min(n_probs, resident CTAs) CTAs of 512 threads; a <= 384.  Work items are (sweep, part,
// problem) triples -- a quarter of a sweep of one problem -- handed out (sweep, part)-major through an atomic counter, so
// that the 256 problems of a batch do not run as 1.73 "waves" of whole problems on 148 SMs (the second one 73 % full): a
// CTA that finishes an item takes the next one, whichever problem it belongs to.  (Whole sweeps as items, round 1: the
// last sweep of the ~180 of 256 problems that need it ran as 148 + 32 -- ncu showed 12.9 % of the kernel's warp samples
// at the scheduler barrier; quarter sweeps make that tail four times shorter.)  sched[0] = next item, sched[1] =
// finished problems, sched[2 + t] = state of problem t: number of completed parts, or -1 when finished,
// sched[2 + n_probs + t] = rotated / big flags of the parts of t's current sweep.  Items are claimed in increasing
// order and only resident CTAs claim, so waiting for a problem's previous part (claimed earlier, hence running or done)
// cannot deadlock.
// Dynamic shared memory: rx_smem_bytes(ceil(a_bound / 64), b_bound).  sched must be zeroed before the launch.
template<typename T>
__global__ void __launch_bounds__(RX_THREADS, 1) k_jacobi_svd_rx(const SvdProb<T> *__restrict__ probs, int n_probs,
                                                                 int max_sweeps, int *__restrict__ sched, T stop2) {
    extern __shared__ __align__(16) unsigned char smem_raw_rx[];
    T *sm = reinterpret_cast<T *>(smem_raw_rx);
    __shared__ int s_item, s_state;
    __shared__ RxItem s_it;
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int item = -1;
            if (*reinterpret_cast<volatile int *>(sched + 1) < n_probs) item = atomicAdd(sched, 1);
            int state = 0;
            if (item >= 0 && item < n_probs * max_sweeps * RX_PARTS) {
                const int pr = item % n_probs, sp = item / n_probs;  // sp = sweep * RX_PARTS + part
                volatile int *st = reinterpret_cast<volatile int *>(sched + 2 + pr);
                unsigned spins = 0;
                while ((state = *st) >= 0 && state < sp) {  // previous part still running elsewhere
                    __nanosleep(200);
                    if (++spins > (1u << 26)) __trap();  // watchdog (> 10 s): abort instead of hanging the GPU
                }
                __threadfence();
                s_it.part = sp % RX_PARTS;
                s_it.carry = (sp % RX_PARTS) ? *reinterpret_cast<volatile int *>(sched + 2 + n_probs + pr) : 0;
            } else {
                item = -1;
            }
            s_item = item;
            s_state = state;
        }
        __syncthreads();
        const int item = s_item;
        if (item < 0) break;
        if (s_state < 0) continue;  // problem already finished
        const int pr = item % n_probs, sw = item / n_probs / RX_PARTS;
        const SvdProb<T> p = probs[pr];
        int flags = 4;
        if (p.a > 0 && p.b > 0) {
            switch ((p.a + 63) / 64) {
                case 1: flags = jacobi_sweep_rx<T, 1, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 2: flags = jacobi_sweep_rx<T, 2, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 3: flags = jacobi_sweep_rx<T, 3, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 4: flags = jacobi_sweep_rx<T, 4, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 5: flags = jacobi_sweep_rx<T, 5, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 6: flags = jacobi_sweep_rx<T, 6, 2>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 7: flags = jacobi_sweep_rx<T, 7, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 8: flags = jacobi_sweep_rx<T, 8, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 9: flags = jacobi_sweep_rx<T, 9, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 10: flags = jacobi_sweep_rx<T, 10, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                case 11: flags = jacobi_sweep_rx<T, 11, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
                default: flags = jacobi_sweep_rx<T, 12, 1>(sm, p, sw, &s_it, max_sweeps, stop2); break;
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();  // the rotated copy / outputs are visible before the state moves on
            if (flags & 4) {
                *reinterpret_cast<volatile int *>(sched + 2 + pr) = -1;
                atomicAdd(sched + 1, 1);
            } else {
                if (flags & 8) {
                    *reinterpret_cast<volatile int *>(sched + 2 + n_probs + pr) = flags & 3;
                    __threadfence();
                }
                *reinterpret_cast<volatile int *>(sched + 2 + pr) = item / n_probs + 1;
            }
        }
    }
}

}  // namespace hcb
