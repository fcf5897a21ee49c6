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

constexpr int SK_ROWS = 128;     // rows per CTA
constexpr int SK_THREADS = 128;  // 4 warps; two CTAs (of different clusters) share an SM and fill each other's barrier gaps
constexpr int SK_WP = 36;        // pitch of the W / W2 matrix (conflict-free DMMA B-fragment loads)
constexpr int SK_MAXCS = 8;      // portable cluster size limit: up to 1024 rows per strip

struct StripJob {
    double *S;         // strip: column 0, row 0 (ld lds); ncols <= NBQ columns, m rows
    const double *Vc;  // clean reflector panel (ld ldv): block p = columns [32p, 32p+32), rows >= 32p
    const double *Tb;  // T factors, NBQ*NBQ per block (column-major, ld NBQ)
    int lds, ldv, m, ncols;
    int kmax;                      // reflectors in the panel
    int p_first, p_count, p_step;  // blocks applied in this order: p_first + i * p_step
    int trans_t;                   // 1: Q^T (W2 = T^T W), 0: Q (W2 = T W)
};

// shared-memory position of element (row, col) of a 128 x 32 block: column-major with the rows of column c rotated by
// 4c, so that both DMMA fragment patterns (4 rows x 4..8 columns and 8 rows x 4 columns) touch 16 distinct 8-byte bank
// pairs per half-warp without padding.
__device__ __forceinline__ int sk_addr(int row, int col) { return col * SK_ROWS + ((row + 4 * col) & (SK_ROWS - 1)); }

constexpr size_t SK_SMEM_BYTES =
    sizeof(double) * (2 * (size_t) NBQ * SK_ROWS + 2 * NBQ * NBQ + NBQ * SK_WP + NBQ * NBQ);  // 97 KB: two CTAs per SM

// grid.x = cluster_size * n_jobs, cluster (cluster_size,1,1), block SK_THREADS, dynamic smem SK_SMEM_BYTES
__global__ void __launch_bounds__(SK_THREADS, 2) k_strip_reflect(const StripJob *__restrict__ jobs) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const StripJob jb_ = jobs[blockIdx.x / CS];
    if (jb_.ncols <= 0 || jb_.p_count <= 0 || jb_.m <= 0) return;  // uniform over the cluster
    extern __shared__ __align__(16) unsigned char smem_raw_sk[];
    double *Sb = reinterpret_cast<double *>(smem_raw_sk);
    double *Vb = Sb + NBQ * SK_ROWS;             // reflector block (single buffer: the co-resident CTA hides its load)
    double *Wp = Vb + NBQ * SK_ROWS;             // two partial-W buffers (read by the other CTAs of the cluster)
    double *Wf = Wp + 2 * NBQ * NBQ;             // summed W, then -W2 (pitch SK_WP)
    double *Ts = Wf + NBQ * SK_WP;               // op(T_p), stored so that lane i reads op(T)[i][k] at Ts[k*32+i]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = jb_.m, ncols = jb_.ncols;
    // rows are dealt to the CTAs of the cluster in groups of 32 (group gg -> CTA gg % CS): the reflector blocks are zero
    // above their diagonal, so a contiguous split would leave the first CTAs idle for the later blocks
    auto grow = [&](int lr) { return 32 * ((lr >> 5) * CS + crank) + (lr & 31); };

    // ---- strip -> shared memory (zero padded), asynchronously: everything in flight at once (a scalar
    // load/store loop was 30 % of the first version's stall samples, profiles/r01_ncu_strip_reflect.txt)
    for (int q = tid; q < NBQ * SK_ROWS / 2; q += SK_THREADS) {
        const int col = q / (SK_ROWS / 2), row = 2 * (q % (SK_ROWS / 2)), gr = grow(row);
        const double *src = jb_.S + (size_t) col * jb_.lds + gr;
        double *dst = Sb + sk_addr(row, col);
        const bool v0 = col < ncols && gr < m, v1 = col < ncols && gr + 1 < m;
        if (v0 && v1 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) cp_async_16(dst, src);
        else {
            if (v0) cp_async_8(dst, src); else dst[0] = 0.0;
            if (v1) cp_async_8(dst + 1, src + 1); else dst[1] = 0.0;
        }
    }

    auto block_cols = [&](int p) { const int left = jb_.kmax - p * NBQ; return left < NBQ ? left : NBQ; };
    auto cta_active = [&](int p) { return grow(SK_ROWS - 1) >= p * NBQ && 32 * crank < m; };
    // cp.async load of V_p, local rows, into Vb (rows above the block, beyond m and columns >= jb are 0)
    auto load_v = [&](int p) {
        if (!cta_active(p)) return;
        const int j0 = p * NBQ, jb = block_cols(p);
        for (int q = tid; q < NBQ * SK_ROWS / 2; q += SK_THREADS) {
            const int col = q / (SK_ROWS / 2), row = 2 * (q % (SK_ROWS / 2)), gr = grow(row);
            const double *src = jb_.Vc + (size_t) (j0 + col) * jb_.ldv + gr;
            double *dst = Vb + sk_addr(row, col);
            const bool v0 = col < jb && gr >= j0 && gr < m, v1 = col < jb && gr + 1 >= j0 && gr + 1 < m;
            if (v0 && v1 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) cp_async_16(dst, src);
            else {
                if (v0) cp_async_8(dst, src); else dst[0] = 0.0;
                if (v1) cp_async_8(dst + 1, src + 1); else dst[1] = 0.0;
            }
        }
    };
    constexpr int TQ = NBQ * NBQ / SK_THREADS;  // T elements per thread
    double treg[TQ];
    auto fetch_t = [&](int p) {
        const double *Tg = jb_.Tb + (size_t) p * NBQ * NBQ;
#pragma unroll
        for (int q = 0; q < TQ; ++q) treg[q] = Tg[tid + q * SK_THREADS];
    };
    auto store_t = [&]() {
#pragma unroll
        for (int q = 0; q < TQ; ++q) {
            const int idx = tid + q * SK_THREADS, i = idx % NBQ, k = idx / NBQ;  // treg = T[i][k]
            if (jb_.trans_t) Ts[i * NBQ + k] = treg[q];  // op(T)[k][i] = T[i][k]
            else Ts[k * NBQ + i] = treg[q];
        }
    };

    int p = jb_.p_first;
    load_v(p);
    cp_async_commit();
    fetch_t(p);
    store_t();

    for (int it = 0; it < jb_.p_count; ++it, p += jb_.p_step) {
        double *Wpb = Wp + (it & 1) * NBQ * NBQ;
        const int j0 = p * NBQ;
        const bool active = cta_active(p);
        if (it + 1 < jb_.p_count) fetch_t(p + jb_.p_step);  // next T: in registers until this block's phase 2 is over
        cp_async_wait_all();
        __syncthreads();  // S, V_p and op(T_p) are in place (previous phase 3 finished)

        // ---- phase 1: partial W = V_loc^T S_loc; warp w owns the output tiles (w, 0..3)
        {
            double acc[4][2][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
            if (active) {
                const int colA = 8 * w + g;
                const double *pa = Vb + colA * SK_ROWS;
                const int ra = 4 * colA + t;
                // V is zero above row j0: skip the local 32-row groups that lie entirely above it
                const int gfirst = j0 / 32 - crank;
                const int ks0 = gfirst > 0 ? 8 * ((gfirst + CS - 1) / CS) : 0;
#pragma unroll 2
                for (int ks = ks0; ks < SK_ROWS / 4; ks += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = 4 * (ks + h);
                        const double a = pa[(rr + ra) & (SK_ROWS - 1)];
                        double b[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) b[j] = Sb[(8 * j + g) * SK_ROWS + ((rr + 4 * (8 * j + g) + t) & (SK_ROWS - 1))];
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[j][h][0], acc[j][h][1], a, b[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) Wpb[(8 * w + g) * NBQ + 8 * j + 2 * t + h] = acc[j][0][h] + acc[j][1][h];
        }
        cluster.sync();
        // ---- cluster sum + phase 2 without a block barrier in between: warp w sums and transforms ITS 8 columns
        {
            double wsum[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) wsum[c] = 0.0;
            for (int rk = 0; rk < CS; ++rk) {
                const double *rp = cluster.map_shared_rank(Wpb, rk) + lane * NBQ + 8 * w;
#pragma unroll
                for (int c = 0; c < 8; ++c) wsum[c] += rp[c];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) Wf[lane * SK_WP + 8 * w + c] = wsum[c];
            __syncwarp();
            double o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = 0.0;
#pragma unroll 4
            for (int k = 0; k < NBQ; ++k) {
                const double tv = Ts[k * NBQ + lane];
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = fma(tv, Wf[k * SK_WP + 8 * w + c], o[c]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) Wf[lane * SK_WP + 8 * w + c] = -o[c];
        }
        __syncthreads();
        if (it + 1 < jb_.p_count) store_t();  // op(T) of the next block (Ts is not read in phase 3)
        // ---- phase 3: S_loc += V_loc (-W2); warp w owns rows 32w..32w+31
        if (active && grow(32 * w) + 32 > j0 && grow(32 * w) < m) {
            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) acc[i][j][h] = Sb[sk_addr(32 * w + 8 * i + g, 8 * j + 2 * t + h)];
#pragma unroll
            for (int ks = 0; ks < NBQ / 4; ++ks) {
                const int kc = 4 * ks + t;
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Vb[sk_addr(32 * w + 8 * i + g, kc)];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Wf[kc * SK_WP + 8 * j + g];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Sb[sk_addr(32 * w + 8 * i + g, 8 * j + 2 * t + h)] = acc[i][j][h];
        }
        if (it + 1 < jb_.p_count) {
            __syncthreads();  // everybody is done with V_p
            load_v(p + jb_.p_step);
            cp_async_commit();
        }
    }
    __syncthreads();
    // ---- strip back to global memory
    for (int idx = tid; idx < NBQ * SK_ROWS; idx += SK_THREADS) {
        const int col = idx / SK_ROWS, row = idx % SK_ROWS, gr = grow(row);
        if (gr < m && col < ncols) jb_.S[(size_t) gr + (size_t) col * jb_.lds] = Sb[sk_addr(row, col)];
    }
    cluster.sync();  // nobody leaves while a neighbour may still read its partial sums
}

}  // namespace hcb
