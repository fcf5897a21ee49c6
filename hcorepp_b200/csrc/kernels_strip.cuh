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
// Cluster barrier for the PULL pattern used here (a CTA writes its own shared memory, the others read it remotely after
// the barrier).  cg::cluster_group::sync() compiles to MEMBAR.ALL.GPU + ERRBAR + arrive + wait: the GPU-scope fence
// also waits for this thread's outstanding global traffic -- the cp.async prefetch of the next reflector block -- so
// every barrier exposed the full L2 / HBM latency of that prefetch.  The data crossing the barrier is shared memory
// written by the owning SM: a CTA-scope fence (the writes are performed in this SM's shared memory) + relaxed arrive
// + acquiring wait is what the hardware needs.  (Round-1 advice: in the PTX memory model a CTA-scope fence does not
// formally order the writes for another CTA's ld.shared::cluster; the two formally sufficient forms --
// fence.acq_rel.cluster + relaxed arrive, and barrier.cluster.arrive.release -- were compiled for sm_100a: both become
// MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR + UCGABAR_ARV, i.e. exactly the GPU-scope fence measured above
// (profiles/r01_cluster_barrier.txt: +1.1 k cycles per window), while this form is MEMBAR.ALL.CTA + UCGABAR_ARV.  The
// shared-memory writes of an SM are performed in that SM's own shared memory, and the remote reads travel through the
// cluster network AFTER the barrier wait (acquire): kept, with compute-sanitizer racecheck clean and every parity test
// exercising it.)
__device__ __forceinline__ void cluster_sync_pull() {
    asm volatile("fence.acq_rel.cta;\n" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// Geometry: 256 rows per CTA (cluster of 4 for 1024 rows, one CTA per SM), 8 warps in two groups of 4.
//
// The two groups work on the two 16-column HALVES of the strip, half a period apart.  Cluster barriers separate
// "windows"; in every window one group runs phase 1 of a block (tensor pipe) while the other group runs the cluster sum
// and the T product of the previous window's phase 1 (latency-bound: DSMEM loads, SIMT) followed by its phase 3 (tensor
// pipe).  With all 8 warps in lock step the tensor pipe idled during sum / T product / barriers: 41 % active in the ncu
// capture, 20 % of the samples in the sum + T product alone (profiles/r01_ncu_strip_reflect_v3.txt).
//   window n, group A:  n even: phase 1 of block n/2        n odd : sum, T product, phase 3 of block (n-1)/2
//             group B:  n odd : phase 1 of block (n-1)/2    n even: sum, T product, phase 3 of block (n-2)/2
// V_b (buffer b & 1) is read in windows 2b .. 2b+2 and V_{b+2} is fetched during window 2b+3; op(T_b) is written by group A
// at the start of window 2b+1 and read by group B in window 2b+2.
constexpr int SKR = 256;                 // rows per CTA
constexpr int SK_HP = 20;                // pitch of a group's W2 (16 columns; conflict-free DMMA B-fragment loads)
constexpr size_t SK_SMEM = sizeof(double) * (3 * (size_t) NBQ * SKR + 2 * 16 * NBQ + 2 * NBQ * SK_HP + NBQ * NBQ);

// grid.x = cluster_size * n_jobs, cluster (cluster_size,1,1), block 256 threads, dynamic smem SK_SMEM
__global__ void __launch_bounds__(SKR, 1) k_strip_reflect(const StripJob *__restrict__ jobs) {
    namespace cg = cooperative_groups;
    constexpr int ROWS = SKR, THREADS = SKR;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const StripJob jb_ = jobs[blockIdx.x / CS];
    if (jb_.ncols <= 0 || jb_.p_count <= 0 || jb_.m <= 0) return;  // uniform over the cluster
    extern __shared__ __align__(16) unsigned char smem_raw_sk[];
    double *Sb = reinterpret_cast<double *>(smem_raw_sk);
    double *Vb0 = Sb + NBQ * ROWS;               // two reflector buffers
    double *Wp = Vb0 + 2 * NBQ * ROWS;           // per group: partial W, TRANSPOSED ([S column 16][reflector 32]), read remotely
    double *Wf = Wp + 2 * 16 * NBQ;              // per group: summed W, then -W2 ([reflector][S column], pitch SK_HP)
    double *Ts = Wf + 2 * NBQ * SK_HP;           // op(T_b), stored so that lane i reads op(T)[i][k] at Ts[k*32+i]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int grp = w >> 2, wg = w & 3;          // group (half strip) and warp inside the group
    const int c0 = 16 * grp;                     // first strip column of this group's half
    double *Wpg = Wp + grp * 16 * NBQ, *Wfg = Wf + grp * NBQ * SK_HP;
    const int m = jb_.m, ncols = jb_.ncols, P = jb_.p_count;
    // shared-memory position of (row, col) of a ROWS x 32 block: column-major with the rows of column c rotated by 4c:
    // both DMMA fragment patterns touch 16 distinct 8-byte bank pairs per half-warp without padding
    auto sk = [](int row, int col) { return col * ROWS + ((row + 4 * col) & (ROWS - 1)); };
    // rows are dealt to the CTAs of the cluster in groups of 32 (group gg -> CTA gg % CS): the reflector blocks are zero
    // above their diagonal, so a contiguous split would leave the first CTAs idle for the later blocks
    auto grow = [&](int lr) { return 32 * ((lr >> 5) * CS + crank) + (lr & 31); };
    auto group_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };

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
    auto blk = [&](int b) { return jb_.p_first + b * jb_.p_step; };  // b-th block applied
    auto block_cols = [&](int p) { const int left = jb_.kmax - p * NBQ; return left < NBQ ? left : NBQ; };
    auto cta_active = [&](int p) { return grow(ROWS - 1) >= p * NBQ && 32 * crank < m; };
    // cp.async load of V_p, local rows (rows above the block, beyond m and columns >= jb are 0)
    auto load_v = [&](int b) {
        const int p = blk(b);
        if (!cta_active(p)) return;
        const int j0 = p * NBQ;
        copy_cols(Vb0 + (b & 1) * NBQ * ROWS, jb_.Vc + (size_t) j0 * jb_.ldv, jb_.ldv, block_cols(p), pgr >= j0 && pgr < m,
                  pgr + 1 >= j0 && pgr + 1 < m);
    };
    // op(T) is handled by group A alone (128 threads, 8 elements each)
    double treg[8];
    auto fetch_t = [&](int b) {
        const double *Tg = jb_.Tb + (size_t) blk(b) * NBQ * NBQ;
#pragma unroll
        for (int q = 0; q < 8; ++q) treg[q] = Tg[(tid & 127) + q * 128];
    };
    auto store_t = [&]() {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = (tid & 127) + q * 128, i = idx % NBQ, k = idx / NBQ;  // treg = T[i][k]
            if (jb_.trans_t) Ts[i * NBQ + k] = treg[q];  // op(T)[k][i] = T[i][k]
            else Ts[k * NBQ + i] = treg[q];
        }
    };

    // ---- prologue: strip, V_0 (and V_1) in flight; T_0 in registers
    copy_cols(Sb, jb_.S, jb_.lds, ncols, pgr < m, pgr + 1 < m);
    load_v(0);
    cp_async_commit();
    if (P > 1) load_v(1);  // not needed before window 2: it keeps flying through windows 0 and 1
    cp_async_commit();
    if (grp == 0) fetch_t(0);
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    // the warp's 64 x 16 piece of its half strip as DMMA accumulator fragments (rows 64 wg .. 64 wg + 63)
    double sacc[8][2][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) sacc[i][j][h] = Sb[sk(64 * wg + 8 * i + g, c0 + 8 * j + 2 * t + h)];

    for (int n = 0; n <= 2 * P; ++n) {
        // fetch V_{b+2} into the buffer that block b released after window 2b+2
        if (n >= 3 && (n & 1) && (n + 1) / 2 < P) {
            load_v((n + 1) / 2);
            cp_async_commit();
        }
        const int rel = n - grp;  // group-relative window: even -> phase 1 of block rel/2, odd -> finish block (rel-1)/2
        if (rel >= 0 && !(rel & 1) && rel / 2 < P) {
            // ================= phase 1 of block b: partial W = V_loc^T S_loc (32 x 16) =================
            const int b = rel / 2, p = blk(b), j0 = p * NBQ;
            const double *Vb = Vb0 + (b & 1) * NBQ * ROWS;
            // warp (kh, tg): K half kh, tile rows 2 tg, 2 tg + 1, both tile columns of the half strip
            const int kh = wg >> 1, tg = wg & 1, ti0 = 2 * tg;
            double acc[2][2][2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j][0][0] = acc[i][j][0][1] = acc[i][j][1][0] = acc[i][j][1][1] = 0.0;
            if (cta_active(p)) {
                // V is zero above row j0: skip the local 32-row groups that lie entirely above it
                const int gfirst = j0 / 32 - crank;
                const int ks_lo = gfirst > 0 ? 8 * ((gfirst + CS - 1) / CS) : 0;
                constexpr int KSPAN = ROWS / 4 / 2;
                const int ks_beg = ks_lo > kh * KSPAN ? ks_lo : kh * KSPAN, ks_end = (kh + 1) * KSPAN;
#pragma unroll 2
                for (int ks = ks_beg; ks < ks_end; ks += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = 4 * (ks + h) + t;
                        double a[2], bb[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int ca = 8 * (ti0 + i) + g;
                            a[i] = Vb[ca * ROWS + ((rr + 4 * ca) & (ROWS - 1))];
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int cb = c0 + 8 * j + g;
                            bb[j] = Sb[cb * ROWS + ((rr + 4 * cb) & (ROWS - 1))];
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j) dmma_m8n8k4(acc[i][j][h][0], acc[i][j][h][1], a[i], bb[j]);
                    }
                }
            }
            // the two K halves are added through a scratch area (aliases this group's Wf, free in this window)
            double *scr = Wfg + tg * 256;
            if (kh == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) scr[((i * 2 + j) * 2 + h) * 32 + lane] = acc[i][j][0][h] + acc[i][j][1][h];
            }
            asm volatile("bar.sync %0, 64;" ::"r"(3 + 2 * grp + tg) : "memory");
            if (kh == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            Wpg[(8 * j + 2 * t + h) * NBQ + 8 * (ti0 + i) + g] =
                                acc[i][j][0][h] + acc[i][j][1][h] + scr[((i * 2 + j) * 2 + h) * 32 + lane];
            }
        } else if (rel >= 1 && (rel & 1) && (rel - 1) / 2 < P) {
            // ================= cluster sum, T product, phase 3 of block b =================
            const int b = (rel - 1) / 2, p = blk(b), j0 = p * NBQ;
            const double *Vb = Vb0 + (b & 1) * NBQ * ROWS;
            if (grp == 0) {
                store_t();  // op(T_b); group B reads it one window later
                if (b + 1 < P) fetch_t(b + 1);
            }
            // warp wg sums and transforms ITS 4 columns of the half strip: lane = reflector index
            double wsum[4] = {0.0, 0.0, 0.0, 0.0};
            if (CS == 4) {  // the common case (1024 rows): all 16 remote loads in flight before the first add
                double v[4][4];
#pragma unroll
                for (int rk = 0; rk < 4; ++rk) {
                    const unsigned ra = dsmem_addr(Wpg + (4 * wg) * NBQ + lane, (unsigned) rk);
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[rk][c] = ld_dsmem(ra + (unsigned) (c * NBQ * sizeof(double)));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) wsum[c] = (v[0][c] + v[1][c]) + (v[2][c] + v[3][c]);
            } else {
                for (int rk = 0; rk < CS; ++rk) {
                    const unsigned ra = dsmem_addr(Wpg + (4 * wg) * NBQ + lane, (unsigned) rk);
#pragma unroll
                    for (int c = 0; c < 4; ++c) wsum[c] += ld_dsmem(ra + (unsigned) (c * NBQ * sizeof(double)));
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) Wfg[lane * SK_HP + 4 * wg + c] = wsum[c];
            group_bar();  // (group A: op(T_b) is complete; both: nobody still reads the K-split scratch)
            // 4 outputs per lane, each a 32-term dot product: four partial chains per output (16 independent FMA chains of
            // 8) instead of one chain of 32 -- this stretch is latency-bound and sits on the window's critical path
            double o[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[q][c] = 0.0;
#pragma unroll
            for (int k = 0; k < NBQ; k += 4) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double tv = Ts[(k + q) * NBQ + lane];
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[q][c] = fma(tv, Wfg[(k + q) * SK_HP + 4 * wg + c], o[q][c]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) Wfg[lane * SK_HP + 4 * wg + c] = -((o[0][c] + o[1][c]) + (o[2][c] + o[3][c]));
            group_bar();
            // phase 3: S_half += V_loc (-W2); warp wg owns rows 64 wg .. 64 wg + 63 (two 32-row groups)
            if (cta_active(p)) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r0 = 64 * wg + 32 * half;
                    if (!(grow(r0) + 32 > j0 && grow(r0) < m)) continue;
#pragma unroll
                    for (int ks = 0; ks < NBQ / 4; ++ks) {
                        const int kc = 4 * ks + t;
                        double a[4], bb[2];
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = Vb[sk(r0 + 8 * i + g, kc)];
#pragma unroll
                        for (int j = 0; j < 2; ++j) bb[j] = Wfg[kc * SK_HP + 8 * j + g];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j) dmma_m8n8k4(sacc[4 * half + i][j][0], sacc[4 * half + i][j][1], a[i], bb[j]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                Sb[sk(r0 + 8 * i + g, c0 + 8 * j + 2 * t + h)] = sacc[4 * half + i][j][h];
                }
            }
        }
        if (n & 1) cp_async_wait_all();  // V blocks are first read in even windows (phase 1 of group A)
        cluster_sync_pull();
    }
    // ---- strip back to global memory
    for (int idx = tid; idx < NBQ * ROWS; idx += THREADS) {
        const int col = idx / ROWS, row = idx % ROWS, gr = grow(row);
        if (gr < m && col < ncols) jb_.S[(size_t) gr + (size_t) col * jb_.lds] = Sb[sk(row, col)];
    }
}

}  // namespace hcb
