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

constexpr int SK_ROWS = 256;     // rows per CTA
constexpr int SK_THREADS = 256;  // 8 warps
constexpr int SK_WP = 36;        // pitch of the W / W2 matrix (conflict-free DMMA B-fragment loads)

struct StripJob {
    double *S;         // strip: column 0, row 0 (ld lds); ncols <= NBQ columns, m rows
    const double *Vc;  // clean reflector panel (ld ldv): block p = columns [32p, 32p+32), rows >= 32p
    const double *Tb;  // T factors, NBQ*NBQ per block (column-major, ld NBQ)
    int lds, ldv, m, ncols;
    int kmax;                      // reflectors in the panel
    int p_first, p_count, p_step;  // blocks applied in this order: p_first + i * p_step
    int trans_t;                   // 1: Q^T (W2 = T^T W), 0: Q (W2 = T W)
};

// shared-memory position of element (row, col) of a 256 x 32 block: column-major with the rows of column c rotated by
// 4c, so that both DMMA fragment patterns (4 rows x 4..8 columns and 8 rows x 4 columns) touch 16 distinct 8-byte bank
// pairs per half-warp without padding (3 x 64 KB blocks + the small matrices must fit 227 KB).
__device__ __forceinline__ int sk_addr(int row, int col) { return col * SK_ROWS + ((row + 4 * col) & (SK_ROWS - 1)); }

constexpr size_t SK_SMEM_BYTES =
    sizeof(double) * (3 * (size_t) NBQ * SK_ROWS + 2 * NBQ * NBQ + NBQ * SK_WP + NBQ * NBQ);

// grid.x = cluster_size * n_jobs, cluster (cluster_size,1,1), block SK_THREADS, dynamic smem SK_SMEM_BYTES
__global__ void __launch_bounds__(SK_THREADS, 1) k_strip_reflect(const StripJob *__restrict__ jobs) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const StripJob jb_ = jobs[blockIdx.x / CS];
    if (jb_.ncols <= 0 || jb_.p_count <= 0 || jb_.m <= 0) return;  // uniform over the cluster
    extern __shared__ __align__(16) unsigned char smem_raw_sk[];
    double *Sb = reinterpret_cast<double *>(smem_raw_sk);
    double *Vb0 = Sb + NBQ * SK_ROWS;            // two reflector buffers
    double *Wp = Vb0 + 2 * NBQ * SK_ROWS;        // two partial-W buffers (read by the other CTAs of the cluster)
    double *Wf = Wp + 2 * NBQ * NBQ;             // summed W, then -W2 (pitch SK_WP)
    double *Ts = Wf + NBQ * SK_WP;               // op(T_p), stored so that lane i reads op(T)[i][k] at Ts[k*32+i]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = jb_.m, ncols = jb_.ncols;
    // rows are dealt to the CTAs of the cluster in groups of 32 (group gg -> CTA gg % CS): the reflector blocks are zero
    // above their diagonal, so a contiguous split would leave the first CTAs idle for the later blocks
    auto grow = [&](int lr) { return 32 * ((lr >> 5) * CS + crank) + (lr & 31); };

    // ---- strip -> shared memory (zero padded), asynchronously: all 64 KB in flight at once (the scalar
    // load/store loop was 30 % of the kernel's stall samples, profiles/r01_strip_reflect_ncu.txt)
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
    // cp.async prefetch of V_p rows [r0, r0+256) into buffer `buf` (rows above the block, beyond m and columns >= jb are 0)
    auto prefetch_v = [&](int p, int buf) {
        if (!cta_active(p)) return;
        double *Vb = Vb0 + buf * NBQ * SK_ROWS;
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
    double treg[4];
    auto fetch_t = [&](int p) {
        const double *Tg = jb_.Tb + (size_t) p * NBQ * NBQ;
#pragma unroll
        for (int q = 0; q < 4; ++q) treg[q] = Tg[tid + q * SK_THREADS];
    };

    int p = jb_.p_first;
    prefetch_v(p, 0);
    cp_async_commit();
    fetch_t(p);

    for (int it = 0; it < jb_.p_count; ++it, p += jb_.p_step) {
        const int buf = it & 1;
        const double *Vb = Vb0 + buf * NBQ * SK_ROWS;
        double *Wpb = Wp + buf * NBQ * NBQ;
        const int j0 = p * NBQ;
        const bool active = cta_active(p);
        cp_async_wait_all();
        // op(T_p) -> Ts
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * SK_THREADS, i = idx % NBQ, k = idx / NBQ;  // treg = T[i][k]
            if (jb_.trans_t) Ts[i * NBQ + k] = treg[q];  // op(T)[k][i] = T[i][k]
            else Ts[k * NBQ + i] = treg[q];
        }
        __syncthreads();
        if (it + 1 < jb_.p_count) {  // next block's reflectors fly in while this one is applied
            prefetch_v(p + jb_.p_step, buf ^ 1);
            fetch_t(p + jb_.p_step);
        }
        cp_async_commit();

        // ---- phase 1: partial W = V_loc^T S_loc; warp w owns output tiles (ti, tj0) and (ti, tj0+1)
        {
            const int ti = w >> 1, tj0 = 2 * (w & 1);
            double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
            if (active) {
                const int colA = 8 * ti + g, colB0 = 8 * tj0 + g, colB1 = colB0 + 8;
                const double *pa = Vb + colA * SK_ROWS, *pb0 = Sb + colB0 * SK_ROWS, *pb1 = Sb + colB1 * SK_ROWS;
                const int ra = 4 * colA + t, rb0 = 4 * colB0 + t, rb1 = 4 * colB1 + t;
                // V is zero above row j0: skip the local 32-row groups that lie entirely above it
                const int gfirst = j0 / 32 - crank;
                const int ks0 = gfirst > 0 ? 8 * ((gfirst + CS - 1) / CS) : 0;
#pragma unroll 4
                for (int ks = ks0; ks < SK_ROWS / 4; ks += 2) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = 4 * (ks + h);
                        const double a = pa[(rr + ra) & (SK_ROWS - 1)];
                        const double b0 = pb0[(rr + rb0) & (SK_ROWS - 1)];
                        const double b1 = pb1[(rr + rb1) & (SK_ROWS - 1)];
                        dmma_m8n8k4(acc[0][h][0], acc[0][h][1], a, b0);
                        dmma_m8n8k4(acc[1][h][0], acc[1][h][1], a, b1);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    Wpb[(8 * ti + g) * NBQ + 8 * (tj0 + j) + 2 * t + h] = acc[j][0][h] + acc[j][1][h];
        }
        cluster.sync();
        // ---- cluster sum of the partials
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * SK_THREADS, i = idx / NBQ, j = idx % NBQ;
            double s = 0.0;
            for (int rk = 0; rk < CS; ++rk) s += cluster.map_shared_rank(Wpb, rk)[idx];
            Wf[i * SK_WP + j] = s;
        }
        __syncthreads();
        // ---- phase 2: W2 = op(T) W, stored negated in place; warp w owns columns 4w..4w+3, lane = row
        {
            double o[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
            for (int k = 0; k < NBQ; ++k) {
                const double tv = Ts[k * NBQ + lane];
#pragma unroll
                for (int c = 0; c < 4; ++c) o[c] = fma(tv, Wf[k * SK_WP + 4 * w + c], o[c]);
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) Wf[lane * SK_WP + 4 * w + c] = -o[c];
        }
        __syncthreads();
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
        __syncthreads();
    }
    cp_async_wait_all();
    // ---- strip back to global memory
    for (int idx = tid; idx < NBQ * SK_ROWS; idx += SK_THREADS) {
        const int col = idx / SK_ROWS, row = idx % SK_ROWS, gr = grow(row);
        if (gr < m && col < ncols) jb_.S[(size_t) gr + (size_t) col * jb_.lds] = Sb[sk_addr(row, col)];
    }
    cluster.sync();  // nobody leaves while a neighbour may still read its partial sums
}

}  // namespace hcb
