// kernels_dmma.cuh -- batched FP64 GEMM on the tensor pipe (DMMA, mma.sync.m8n8k4.f64) for the GEMM-shaped bulk of
// the path: the rank-space contractions, the block Gram-Schmidt of the incremental recompression, the core product,
// V*Sigma = M^T U and the rebuild [CU | Q2] * Us.
//
// FP64 has no tcgen05 kind on sm_100a (kinds: f16/tf32/f8f6f4/i8/mx*), so TMEM-accumulator MMA does not apply to this
// path; the FP64 tensor instruction is the warp-level mma.sync DMMA (SASS: DMMA), issued once per 16 cycles per SM
// sub-partition.  Operand feed (round 2): an operand whose tile rows are CONTIGUOUS in global memory (A not transposed,
// B transposed: 256..1024-byte segments per k) is brought into its shared-memory stage by the TMA engine --
// cp.async.bulk (SASS: UBLKCP), one bulk copy per k column, completion counted in bytes on an mbarrier (expect_tx /
// complete_tx) -- so the big left operands of the path (CU in the Gram-Schmidt update and in the rebuild, 1024-row
// columns) cost no issue slots and no registers.  Operands that are contiguous along k (128-byte segments at BK = 16:
// a measured loss as bulk copies, profiles/r02_gemm_variants.txt) or that the bulk engine cannot take (16-byte
// alignment: odd leading dimensions such as V factors stored with ld = rank, odd tails) go global -> register ->
// shared with the prefetch of the next stage overlapping the DMMAs of the current one.
// (Tensor-map TMA, cp.async.bulk.tensor, would need one CUtensorMap per operand of every batched problem, built on the
// device for data-dependent shapes; the 1-D bulk form needs no descriptor and lets every column land at the padded
// shared-memory pitch that makes the DMMA fragment loads bank-conflict free.)
// Same device-resident GemmProb descriptors as k_gemm_batched.
#pragma once
#include "common.cuh"

namespace hcb {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- TMA bulk copy (global -> shared, completion on an mbarrier) ------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned) __cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned) __cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// CTA tile (32*WM) x (32*WN), WM*WN == 4 warps, each warp a 32x32 tile = 4x4 DMMA tiles (32 accumulator doubles).
// BK = 16: four k-steps of 4 per shared-memory stage (two stages); operands staged k-major with a row pitch of tile+4
// doubles so that the 8-byte fragment loads of a half-warp hit 16 distinct bank pairs.
// Optional second A segment (p.A2): op(A) = [op(A) | op(A2)] along k, split at p.k1 -- the rebuild
// CU' = [CU | Q2] * Us reads CU from the tile and Q2 from scratch without a second accumulate pass over C.
// grid = (tiles_bound, n_problems), grid-stride over output tiles.
template<int WM, int WN>
__global__ void __launch_bounds__(128, 3) k_gemm_dmma(const GemmProb<double> *__restrict__ probs) {
    constexpr int BM = 32 * WM, BN = 32 * WN, BK = 16, LDA = BM + 4, LDB = BN + 4;
    constexpr int A_PER_THR = BM * BK / 128, B_PER_THR = BN * BK / 128;
    static_assert(WM * WN == 4, "four warps per CTA");
    const GemmProb<double> p = probs[blockIdx.y];
    if (p.m <= 0 || p.n <= 0) return;
    __shared__ __align__(16) double As[2][BK][LDA];
    __shared__ __align__(16) double Bs[2][BK][LDB];
    __shared__ __align__(8) unsigned long long full[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, c = lane & 3;
    const int tiles_m = (p.m + BM - 1) / BM, tiles_n = (p.n + BN - 1) / BN;
    const int nk = (p.k + BK - 1) / BK;
    if (tid == 0) {
        mbar_init(full + 0, 1);
        mbar_init(full + 1, 1);
        fence_mbar_init();
    }
    // bulk-copy eligibility (uniform over the CTA): rows contiguous, 16-byte aligned base, even leading dimension
    const bool a_al = p.ta == 0 && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && ((p.lda & 1) == 0) &&
                      (p.A2 == nullptr || (((reinterpret_cast<uintptr_t>(p.A2) & 15) == 0) && ((p.lda2 & 1) == 0)));
    const bool b_al = p.tb != 0 && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && ((p.ldb & 1) == 0);
    const int k1 = p.A2 ? p.k1 : p.k;
    unsigned git = 0;  // stage uses so far: buffer = git & 1, mbarrier parity = (git >> 1) & 1
    __syncthreads();

    for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const int row0 = (tile % tiles_m) * BM, col0 = (tile / tiles_m) * BN;
        const int mrem = min(BM, p.m - row0), nrem = min(BN, p.n - col0);
        const bool a_bulk = a_al && ((mrem & 1) == 0), b_bulk = b_al && ((nrem & 1) == 0);  // (odd tail tiles: registers)
        const bool any_bulk = a_bulk || b_bulk;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        double ra[A_PER_THR], rb[B_PER_THR];

        // TMA part of stage `kt` into buffer `buf`: one bulk copy per k column of every bulk operand; the columns of a K
        // tail that do not exist are zeroed by the threads.  One mbarrier phase per stage use (arrive even with 0 bytes).
        auto issue = [&](int buf, int kt) {
            const int k0 = kt * BK, krem = min(BK, p.k - k0);
            if (!any_bulk) return;  // register-fed tile: no mbarrier traffic at all
            if (warp == 0) {
                const unsigned bytes = (a_bulk ? (unsigned) (krem * mrem) * 8u : 0u) + (b_bulk ? (unsigned) (krem * nrem) * 8u : 0u);
                if (lane == 0) mbar_expect_tx(full + buf, bytes);
                __syncwarp();
                if (a_bulk && lane < krem) {
                    const int gk = k0 + lane;
                    const double *src = gk < k1 ? p.A + (size_t) row0 + (size_t) gk * p.lda
                                                : p.A2 + (size_t) row0 + (size_t) (gk - k1) * p.lda2;
                    bulk_g2s(&As[buf][lane][0], src, (unsigned) mrem * 8u, full + buf);
                }
                if (b_bulk && lane >= 16 && lane - 16 < krem) {
                    const int kk = lane - 16;
                    bulk_g2s(&Bs[buf][kk][0], p.B + (size_t) col0 + (size_t) (k0 + kk) * p.ldb, (unsigned) nrem * 8u, full + buf);
                }
            }
            if (krem < BK) {
                if (a_bulk)
                    for (int idx = tid; idx < (BK - krem) * BM; idx += 128) As[buf][krem + idx / BM][idx % BM] = 0.0;
                if (b_bulk)
                    for (int idx = tid; idx < (BK - krem) * BN; idx += 128) Bs[buf][krem + idx / BN][idx % BN] = 0.0;
                fence_proxy_async();
            }
        };
        auto fetch = [&](int k0) {
            if (!a_bulk) {
#pragma unroll
                for (int q = 0; q < A_PER_THR; ++q) {
                    const int idx = tid + q * 128;
                    int r, kk;
                    if (p.ta == 0) { r = idx % BM; kk = idx / BM; } else { kk = idx % BK; r = idx / BK; }
                    const int gr = row0 + r, gk = k0 + kk;
                    double v = 0.0;
                    if (gr < p.m && gk < p.k) {
                        if (p.ta != 0) v = gk < k1 ? p.A[(size_t) gk + (size_t) gr * p.lda] : p.A2[(size_t) (gk - k1) + (size_t) gr * p.lda2];
                        else v = gk < k1 ? p.A[(size_t) gr + (size_t) gk * p.lda] : p.A2[(size_t) gr + (size_t) (gk - k1) * p.lda2];
                    }
                    ra[q] = v;
                }
            }
            if (!b_bulk) {
#pragma unroll
                for (int q = 0; q < B_PER_THR; ++q) {
                    const int idx = tid + q * 128;
                    int cc, kk;
                    if (p.tb == 0) { kk = idx % BK; cc = idx / BK; } else { cc = idx % BN; kk = idx / BN; }
                    const int gc = col0 + cc, gk = k0 + kk;
                    rb[q] = (gc < p.n && gk < p.k)
                                ? (p.tb == 0 ? p.B[(size_t) gk + (size_t) gc * p.ldb] : p.B[(size_t) gc + (size_t) gk * p.ldb])
                                : 0.0;
                }
            }
        };
        auto stash = [&](int buf) {
            if (!a_bulk) {
#pragma unroll
                for (int q = 0; q < A_PER_THR; ++q) {
                    const int idx = tid + q * 128;
                    int r, kk;
                    if (p.ta == 0) { r = idx % BM; kk = idx / BM; } else { kk = idx % BK; r = idx / BK; }
                    As[buf][kk][r] = ra[q];
                }
            }
            if (!b_bulk) {
#pragma unroll
                for (int q = 0; q < B_PER_THR; ++q) {
                    const int idx = tid + q * 128;
                    int cc, kk;
                    if (p.tb == 0) { kk = idx % BK; cc = idx / BK; } else { cc = idx % BN; kk = idx / BN; }
                    Bs[buf][kk][cc] = rb[q];
                }
            }
        };

        __syncthreads();  // previous tile's readers are done with the buffers
        if (nk > 0) {
            const int b0 = any_bulk ? (int) (git & 1) : 0;
            issue(b0, 0);
            fetch(0);
            stash(b0);
        }
        __syncthreads();
        for (int it = 0; it < nk; ++it, git += any_bulk ? 1u : 0u) {
            const int buf = any_bulk ? (int) (git & 1) : (it & 1);
            if (it + 1 < nk) {
                issue(buf ^ 1, it + 1);     // the other buffer was last read in iteration it-1, fenced by the barrier below
                fetch((it + 1) * BK);       // global loads in flight while the tensor pipe works
            }
            if (any_bulk) mbar_wait(full + buf, (git >> 1) & 1);
#pragma unroll
            for (int ks = 0; ks < BK; ks += 4) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[buf][ks + c][wm * 32 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[buf][ks + c][wn * 32 + j * 8 + g];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            if (it + 1 < nk) stash(buf ^ 1);
            __syncthreads();
        }
        // epilogue: lane holds C[g][2c], C[g][2c+1] of every 8x8 tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gc = col0 + wn * 32 + j * 8 + 2 * c + h;
                if (gc >= p.n) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int gr = row0 + wm * 32 + i * 8 + g;
                    if (gr >= p.m) continue;
                    double *cp = p.C + (size_t) gr + (size_t) gc * p.ldc;
                    double v = p.alpha * acc[i][j][h];
                    if (p.beta != 0.0) v = fma(p.beta, *cp, v);
                    *cp = v;
                }
            }
        }
    }
}

}  // namespace hcb
