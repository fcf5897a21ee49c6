// kernels_strip.cuh -- strip-resident block-reflector application (FP64, DMMA) for the two GEMM-shaped halves of the
// recompression: the trailing updates of the blocked Householder QR and the rebuild C := Q [X; 0].
//
// Right-looking blocked QR streams the whole trailing matrix through HBM once per 32-column block (three batched
// GEMMs with K = 32: ~4 flop/byte, ncu launch list r01: k_gemm_dmma<2,2> at ~1.2 TB/s and 26 % of the step).  Here a
// 32-column strip (all rows) is loaded ONCE into the shared memory of a thread-block cluster -- 256 rows per CTA --
// and every block reflector is applied to it while it stays on chip:
//     W  = V_p^T S      (32 x 32, K = rows)   DMMA, per-CTA partial, summed across the cluster through DSMEM
//     W2 = op(T_p) W    (32 x 32 x 32)        SIMT, in shared memory
//     S -= V_p W2       (rows x 32, K = 32)   DMMA, accumulators loaded from / stored to the resident strip
// V_p (256 x 32 per CTA) is prefetched with cp.async into the other half of a double buffer while the current block
// is applied; it comes from L2 (the panel's reflectors were just written, or are shared by the strips of one panel).
// Used left-looking by the QR (strip = block b, reflectors 0..b-1, Q^T) and last-to-first by the rebuild (strip = 32
// columns of C, all blocks, Q): HBM traffic is one read + one write of the strip instead of 3 passes per block.
#pragma once
#include "common.cuh"
#include "kernels_dmma.cuh"
#include "kernels_qr.cuh"
#include <cooperative_groups.h>

namespace hcb {

constexpr int SK_WP = 36;    // pitch of the W2 matrix (conflict-free DMMA B-fragment loads)
constexpr int SK_MAXCS = 8;  // portable cluster size limit

struct StripJob {
    double *S;         // strip: column 0, row 0 (ld lds); ncols <= NBQ columns, m rows
    const double *Vc;  // clean reflector panel (ld ldv): block p = columns [32p, 32p+32), rows >= 32p
    const double *Tb;  // T factors, NBQ*NBQ per block (column-major, ld NBQ)
    int lds, ldv, m, ncols;
    int kmax;                      // reflectors in the panel
    int p_first, p_count, p_step;  // blocks applied in this order: p_first + i * p_step
    int trans_t;                   // 1: Q^T (W2 = T^T W), 0: Q (W2 = T W)
};

// distributed shared memory through explicit shared::cluster addresses (a generic pointer from map_shared_rank makes
// the compiler emit LD.E through the global load path: lg_throttle was the top stall of such a version)
__device__ __forceinline__ unsigned dsmem_addr(const void *smem_ptr, unsigned rank) {
    const unsigned a = (unsigned) __cvta_generic_to_shared(smem_ptr);
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ double ld_dsmem(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// ROWS rows per CTA (ROWS / 32 warps), VBUFS reflector buffers.
//   <256, 2>: one CTA per SM, next block prefetched while the current one is applied; 1024 rows = cluster of 4
//   <128, 1>: two CTAs per SM; 1024 rows = cluster of 8 (4x the DSMEM traffic per block: measured slower)
template<int ROWS, int VBUFS>
struct StripCfg {
    static constexpr int THREADS = ROWS;  // one warp per 32 rows
    static constexpr size_t SMEM = sizeof(double) * ((1 + VBUFS) * (size_t) NBQ * ROWS + 2 * NBQ * NBQ + NBQ * SK_WP + NBQ * NBQ);
};

// grid.x = cluster_size * n_jobs, cluster (cluster_size,1,1), block ROWS threads, dynamic smem StripCfg::SMEM
template<int ROWS, int VBUFS>
__global__ void __launch_bounds__(ROWS, ROWS == 128 ? 2 : 1) k_strip_reflect(const StripJob *__restrict__ jobs) {
    namespace cg = cooperative_groups;
    constexpr int THREADS = ROWS, NWARP = ROWS / 32, CPW = NBQ / NWARP;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const StripJob jb_ = jobs[blockIdx.x / CS];
    if (jb_.ncols <= 0 || jb_.p_count <= 0 || jb_.m <= 0) return;  // uniform over the cluster
    extern __shared__ __align__(16) unsigned char smem_raw_sk[];
    double *Sb = reinterpret_cast<double *>(smem_raw_sk);
    double *Vb0 = Sb + NBQ * ROWS;               // reflector buffer(s)
    double *Wp = Vb0 + VBUFS * NBQ * ROWS;       // two partial-W buffers, TRANSPOSED ([S column][reflector]), read remotely
    double *Wf = Wp + 2 * NBQ * NBQ;             // summed W, then -W2 ([reflector][S column], pitch SK_WP)
    double *Ts = Wf + NBQ * SK_WP;               // op(T_p), stored so that lane i reads op(T)[i][k] at Ts[k*32+i]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = jb_.m, ncols = jb_.ncols;
    // shared-memory position of (row, col) of a ROWS x 32 block: column-major with the rows of column c rotated by 4c:
    // both DMMA fragment patterns touch 16 distinct 8-byte bank pairs per half-warp without padding
    auto sk = [](int row, int col) { return col * ROWS + ((row + 4 * col) & (ROWS - 1)); };
    // rows are dealt to the CTAs of the cluster in groups of 32 (group gg -> CTA gg % CS): the reflector blocks are zero
    // above their diagonal, so a contiguous split would leave the first CTAs idle for the later blocks
    auto grow = [&](int lr) { return 32 * ((lr >> 5) * CS + crank) + (lr & 31); };

    // thread-invariant part of the block copies: thread handles the row pair (prow, prow + 1) of columns pc0, pc0 + 2, ..
    const int prow = 2 * (tid % (ROWS / 2)), pc0 = tid / (ROWS / 2), pgr = grow(prow);
    auto copy_cols = [&](double *dstb, const double *srcb, int ld, int ncol_valid, bool r0, bool r1) {
        const double *src = srcb + (size_t) pc0 * ld + pgr;
        for (int col = pc0; col < NBQ; col += THREADS / (ROWS / 2), src += (size_t) (THREADS / (ROWS / 2)) * ld) {
            double *dst = dstb + col * ROWS + ((prow + 4 * col) & (ROWS - 1));
            const bool v0 = r0 && col < ncol_valid, v1 = r1 && col < ncol_valid;
            if (v0 && v1 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) cp_async_16(dst, src);
            else {
                if (v0) cp_async_8(dst, src); else dst[0] = 0.0;
                if (v1) cp_async_8(dst + 1, src + 1); else dst[1] = 0.0;
            }
        }
    };
    // ---- strip -> shared memory (zero padded), asynchronously: everything in flight at once (a scalar
    // load/store loop was 30 % of the first version's stall samples, profiles/r01_ncu_strip_reflect.txt)
    copy_cols(Sb, jb_.S, jb_.lds, ncols, pgr < m, pgr + 1 < m);

    auto block_cols = [&](int p) { const int left = jb_.kmax - p * NBQ; return left < NBQ ? left : NBQ; };
    auto cta_active = [&](int p) { return grow(ROWS - 1) >= p * NBQ && 32 * crank < m; };
    // cp.async load of V_p, local rows (rows above the block, beyond m and columns >= jb are 0)
    auto load_v = [&](int p, double *Vb) {
        if (!cta_active(p)) return;
        const int j0 = p * NBQ;
        copy_cols(Vb, jb_.Vc + (size_t) j0 * jb_.ldv, jb_.ldv, block_cols(p), pgr >= j0 && pgr < m, pgr + 1 >= j0 && pgr + 1 < m);
    };
    constexpr int TQ = NBQ * NBQ / THREADS;  // T elements per thread
    double treg[TQ];
    auto fetch_t = [&](int p) {
        const double *Tg = jb_.Tb + (size_t) p * NBQ * NBQ;
#pragma unroll
        for (int q = 0; q < TQ; ++q) treg[q] = Tg[tid + q * THREADS];
    };
    auto store_t = [&]() {
#pragma unroll
        for (int q = 0; q < TQ; ++q) {
            const int idx = tid + q * THREADS, i = idx % NBQ, k = idx / NBQ;  // treg = T[i][k]
            if (jb_.trans_t) Ts[i * NBQ + k] = treg[q];  // op(T)[k][i] = T[i][k]
            else Ts[k * NBQ + i] = treg[q];
        }
    };

    cp_async_commit();
    int p = jb_.p_first;
    load_v(p, Vb0);
    cp_async_commit();
    fetch_t(p);
    store_t();
    // the warp's 32 x 32 piece of the strip as DMMA accumulator fragments
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    double sacc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) sacc[i][j][h] = Sb[sk(32 * w + 8 * i + g, 8 * j + 2 * t + h)];

    for (int it = 0; it < jb_.p_count; ++it, p += jb_.p_step) {
        const double *Vb = Vb0 + (VBUFS == 2 ? (it & 1) : 0) * NBQ * ROWS;
        double *Wpb = Wp + (it & 1) * NBQ * NBQ;
        const int j0 = p * NBQ;
        const bool active = cta_active(p);
        const bool more = it + 1 < jb_.p_count;
        if (more) fetch_t(p + jb_.p_step);  // next T: in registers until this block's phase 2 is over
        cp_async_wait_all();
        __syncthreads();  // S (previous phase 3), V_p and op(T_p) are in place
        if (VBUFS == 2 && more) {  // next block's reflectors fly in while this one is applied
            load_v(p + jb_.p_step, Vb0 + ((it + 1) & 1) * NBQ * ROWS);
            cp_async_commit();
        }

        // ---- phase 1: partial W = V_loc^T S_loc.  Warp (kh, tg) = (w / TG, w % TG) accumulates a 2x2 block (or, with
        // four warps, a 2x4 block) of the 16 output tiles over its share of the local rows: 2 A + 2 B fragment loads feed
        // 4 DMMAs (one A + one B per tile would put more wavefronts on the shared-memory port than the DMMAs take).
        {
            constexpr int KH = NWARP >= 8 ? NWARP / 4 : 1, TG = NWARP / KH;       // K splits, tile groups (4)
            constexpr int NTJ = 16 / TG / 2;                                       // tile columns per group (2 or 4)
            const int kh = w / TG, tg = w % TG;
            const int ti0 = 2 * (tg / (4 / NTJ)), tj0 = NTJ * (tg % (4 / NTJ));
            double acc[2][NTJ][2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NTJ; ++j) acc[i][j][0][0] = acc[i][j][0][1] = acc[i][j][1][0] = acc[i][j][1][1] = 0.0;
            if (active) {
                // V is zero above row j0: skip the local 32-row groups that lie entirely above it
                const int gfirst = j0 / 32 - crank;
                const int ks_lo = gfirst > 0 ? 8 * ((gfirst + CS - 1) / CS) : 0;
                constexpr int KSPAN = ROWS / 4 / KH;  // k-steps per K split
                const int ks_beg = ks_lo > kh * KSPAN ? ks_lo : kh * KSPAN, ks_end = (kh + 1) * KSPAN;
#pragma unroll 2
                for (int ks = ks_beg; ks < ks_end; ks += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = 4 * (ks + h) + t;
                        double a[2], b[NTJ];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int ca = 8 * (ti0 + i) + g;
                            a[i] = Vb[ca * ROWS + ((rr + 4 * ca) & (ROWS - 1))];
                        }
#pragma unroll
                        for (int j = 0; j < NTJ; ++j) {
                            const int cb = 8 * (tj0 + j) + g;
                            b[j] = Sb[cb * ROWS + ((rr + 4 * cb) & (ROWS - 1))];
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < NTJ; ++j) dmma_m8n8k4(acc[i][j][h][0], acc[i][j][h][1], a[i], b[j]);
                    }
                }
            }
            // K splits are added through a scratch area (aliases Wf, free until the cluster sum), pair-wise barriers
            double *scr = Wf + tg * (2 * NTJ * 64);
            for (int step = KH - 1; step >= 1; --step) {
                if (kh == step) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < NTJ; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h) scr[((i * NTJ + j) * 2 + h) * 32 + lane] = acc[i][j][0][h] + acc[i][j][1][h];
                }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + tg), "r"(32 * KH) : "memory");
                if (kh == step - 1) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < NTJ; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h) acc[i][j][0][h] += scr[((i * NTJ + j) * 2 + h) * 32 + lane];
                }
                if (step > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + tg), "r"(32 * KH) : "memory");
            }
            if (kh == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < NTJ; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            Wpb[(8 * (tj0 + j) + 2 * t + h) * NBQ + 8 * (ti0 + i) + g] = acc[i][j][0][h] + acc[i][j][1][h];
            }
        }
        cluster.sync();
        // ---- cluster sum + phase 2 without a block barrier in between: warp w sums and transforms ITS CPW columns
        {
            double wsum[CPW];
#pragma unroll
            for (int c = 0; c < CPW; ++c) wsum[c] = 0.0;
            if (CS == 4) {  // the common case (1024 rows): all 4 x CPW remote loads in flight before the first add
                double v[4][CPW];
#pragma unroll
                for (int rk = 0; rk < 4; ++rk) {
                    const unsigned ra = dsmem_addr(Wpb + (CPW * w) * NBQ + lane, (unsigned) rk);
#pragma unroll
                    for (int c = 0; c < CPW; ++c) v[rk][c] = ld_dsmem(ra + (unsigned) (c * NBQ * sizeof(double)));
                }
#pragma unroll
                for (int c = 0; c < CPW; ++c) wsum[c] = (v[0][c] + v[1][c]) + (v[2][c] + v[3][c]);
            } else if (CS == 8) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    double v[4][CPW];
#pragma unroll
                    for (int rk = 0; rk < 4; ++rk) {
                        const unsigned ra = dsmem_addr(Wpb + (CPW * w) * NBQ + lane, (unsigned) (4 * half + rk));
#pragma unroll
                        for (int c = 0; c < CPW; ++c) v[rk][c] = ld_dsmem(ra + (unsigned) (c * NBQ * sizeof(double)));
                    }
#pragma unroll
                    for (int c = 0; c < CPW; ++c) wsum[c] += (v[0][c] + v[1][c]) + (v[2][c] + v[3][c]);
                }
            } else {
                for (int rk = 0; rk < CS; ++rk) {
                    const unsigned ra = dsmem_addr(Wpb + (CPW * w) * NBQ + lane, (unsigned) rk);
#pragma unroll
                    for (int c = 0; c < CPW; ++c) wsum[c] += ld_dsmem(ra + (unsigned) (c * NBQ * sizeof(double)));
                }
            }
#pragma unroll
            for (int c = 0; c < CPW; ++c) Wf[lane * SK_WP + CPW * w + c] = wsum[c];
            __syncwarp();
            double o[CPW];
#pragma unroll
            for (int c = 0; c < CPW; ++c) o[c] = 0.0;
#pragma unroll 4
            for (int k = 0; k < NBQ; ++k) {
                const double tv = Ts[k * NBQ + lane];
#pragma unroll
                for (int c = 0; c < CPW; ++c) o[c] = fma(tv, Wf[k * SK_WP + CPW * w + c], o[c]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CPW; ++c) Wf[lane * SK_WP + CPW * w + c] = -o[c];
        }
        __syncthreads();
        if (more) store_t();  // op(T) of the next block (Ts is not read in phase 3)
        // ---- phase 3: S_loc += V_loc (-W2); warp w owns rows 32w..32w+31, whose fragments stay in registers (sacc) over
        // all blocks: they are only STORED to the shared copy (the B operand of the next phase 1), never re-loaded
        if (active && grow(32 * w) + 32 > j0 && grow(32 * w) < m) {
#pragma unroll
            for (int ks = 0; ks < NBQ / 4; ++ks) {
                const int kc = 4 * ks + t;
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Vb[sk(32 * w + 8 * i + g, kc)];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Wf[kc * SK_WP + 8 * j + g];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_m8n8k4(sacc[i][j][0], sacc[i][j][1], a[i], b[j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Sb[sk(32 * w + 8 * i + g, 8 * j + 2 * t + h)] = sacc[i][j][h];
        }
        if (VBUFS == 1 && more) {
            __syncthreads();  // everybody is done with V_p
            load_v(p + jb_.p_step, Vb0);
            cp_async_commit();
        }
    }
    __syncthreads();
    // ---- strip back to global memory
    for (int idx = tid; idx < NBQ * ROWS; idx += THREADS) {
        const int col = idx / ROWS, row = idx % ROWS, gr = grow(row);
        if (gr < m && col < ncols) jb_.S[(size_t) gr + (size_t) col * jb_.lds] = Sb[sk(row, col)];
    }
    cluster.sync();  // nobody leaves while a neighbour may still read its partial sums
}

}  // namespace hcb
