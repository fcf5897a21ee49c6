// kernels_dmma.cuh -- batched FP64 GEMM on the tensor pipe (DMMA, mma.sync.m8n8k4.f64) for the GEMM-shaped bulk of
// the path: the rank-space contractions, the block Gram-Schmidt of the incremental recompression, the core product,
// V*Sigma = M^T U and the rebuild [CU | Q2] * Us.
//
// FP64 has no tcgen05 kind on sm_100a (kinds: f16/tf32/f8f6f4/i8/mx*), so TMEM-accumulator MMA does not apply to this
// path; the FP64 tensor instruction is the warp-level mma.sync DMMA (SASS: DMMA), issued once per 16 cycles per SM
// sub-partition.  Operand feed (round 2): a three-stage shared-memory ring filled asynchronously -- the left operand by
// the TMA engine when a k column of its tile is one contiguous, 16-byte aligned run of the global matrix
// (cp.async.bulk, SASS UBLKCP, one bulk copy per k column, completion counted in bytes on an mbarrier: the big left
// operands CU / [CU | Q2] / AU, 512-byte columns), everything else by zero-filling cp.async (LDGSTS; 16-byte chunks when
// alignment allows, 8-byte otherwise: V factors are stored with ld = rank) -- so no thread holds operand data in
// registers and two stages of copies are in flight while the tensor pipe works on the third.  (Operands contiguous
// along k would be 128-byte bulk copies at BK = 16: a measured loss, profiles/r02_gemm_variants.txt.)
// (Tensor-map TMA, cp.async.bulk.tensor, would need one CUtensorMap per operand of every batched problem, built on the
// device for data-dependent shapes; the 1-D bulk form needs no descriptor and lets every column land at the padded
// shared-memory pitch that makes the DMMA fragment loads bank-conflict free.)
// Same device-resident GemmProb descriptors as k_gemm_batched.
#pragma once
#include "common.cuh"
#include <type_traits>

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

// zero-filling asynchronous copies (LDGSTS): src_bytes < size zero-fills the rest of the destination
__device__ __forceinline__ void cp_async_zfill_16(void *smem, const void *gmem, unsigned src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_zfill_8(void *smem, const void *gmem, unsigned src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((unsigned) __cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes));
}
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int DG_BK = 16, DG_STAGES = 3;
template<int WM, int WN, int JT = 4>
constexpr size_t dg_smem_bytes() {
    constexpr int BM = 32 * WM, BN = 8 * JT * WN;
    constexpr size_t ea = (size_t) (DG_BK * (BM + 4) > BM * (DG_BK + 4) ? DG_BK * (BM + 4) : BM * (DG_BK + 4));
    constexpr size_t eb = (size_t) (DG_BK * (BN + 4) > BN * (DG_BK + 4) ? DG_BK * (BN + 4) : BN * (DG_BK + 4));
    return sizeof(double) * DG_STAGES * (ea + eb) + 8 * DG_STAGES + 16;
}

// CTA tile (32*WM) x (8*JT*WN), WM*WN == 4 warps, each warp a 32 x (8*JT) tile = 4 x JT DMMA tiles (JT = 4: 32x32, 32
// accumulator doubles).  <4, 1, 6> is the 128 x 48 tile of the SKINNY products of the incremental recompression
// (n = kp = 44 new columns: block Gram-Schmidt, the contraction's U_AB' / V_AB'): with 64-column tiles the second warp
// column owns 12 of its 32 columns, so two of the four SM sub-partitions' tensor pipes idle half of the time; here the
// four warps split M and each covers all 48 columns (92 % of the issued DMMAs are useful instead of 69 %).
// BK = 16 per stage, THREE stages in a ring: while the tensor pipe works on stage s, the copies of stages s+1 and s+2
// are in flight and no thread holds operand data in registers.  Every stage keeps the orientation of the GLOBAL operand,
//     A: ta == 0 -> As[k][m] (pitch BM + 4),  ta == 1 -> As[m][k] (pitch BK + 4)
//     B: tb == 0 -> Bs[n][k] (pitch BK + 4),  tb == 1 -> Bs[k][n] (pitch BN + 4)
// so that global runs stay contiguous in shared memory; both DMMA fragment patterns then touch 16 distinct 8-byte banks
// per half-warp.  Feed per operand and tile:
//   * TMA bulk copies (cp.async.bulk, mbarrier expect_tx) when a k column of the operand tile is one contiguous,
//     16-byte aligned run of the global matrix (A not transposed: the big left operands CU / [CU | Q2] / AU);
//   * otherwise zero-filling cp.async (LDGSTS), 16 bytes when the operand is 16-byte aligned with an even leading
//     dimension, 8 bytes else (V factors stored with ld = rank, odd ranks) -- edges and the K tail need no special path.
// Optional second A segment (p.A2): op(A) = [op(A) | op(A2)] along k, split at p.k1.
// grid = (tiles_bound, n_problems), grid-stride over output tiles.
template<int WM, int WN, int JT = 4>
__global__ void __launch_bounds__(128, 3) k_gemm_dmma(const GemmProb<double> *__restrict__ probs) {
    constexpr int BM = 32 * WM, BN = 8 * JT * WN, WTN = 8 * JT, BK = DG_BK, S = DG_STAGES;
    constexpr int PA0 = BM + 4, PA1 = BK + 4, PB0 = BK + 4, PB1 = BN + 4;
    constexpr int EA = BK * PA0 > BM * PA1 ? BK * PA0 : BM * PA1, EB = BK * PB1 > BN * PB0 ? BK * PB1 : BN * PB0;
    static_assert(WM * WN == 4, "four warps per CTA");
    const GemmProb<double> p = probs[blockIdx.y];
    if (p.m <= 0 || p.n <= 0) return;
    extern __shared__ __align__(16) unsigned char dg_smem[];
    double *As = reinterpret_cast<double *>(dg_smem);
    double *Bs = As + (size_t) S * EA;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(Bs + (size_t) S * EB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, c = lane & 3;
    const int tiles_m = (p.m + BM - 1) / BM, tiles_n = (p.n + BN - 1) / BN;
    const int nk = (p.k + BK - 1) / BK;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    // operand properties (uniform over the CTA)
    const bool a2 = p.A2 != nullptr;
    const bool a_16 = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && ((p.lda & 1) == 0) &&
                      (!a2 || (((reinterpret_cast<uintptr_t>(p.A2) & 15) == 0) && ((p.lda2 & 1) == 0) && ((p.k1 & 1) == 0 || p.ta == 0)));
    const bool b_16 = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && ((p.ldb & 1) == 0);
    const int k1 = a2 ? p.k1 : p.k;
    unsigned phase_bits = 0;  // bit s = parity of the next TMA-fed use of stage s (only bulk-fed uses arm the mbarrier)
    __syncthreads();

    for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const int row0 = (tile % tiles_m) * BM, col0 = (tile / tiles_m) * BN;
        const int mrem = min(BM, p.m - row0);
        const bool a_bulk = p.ta == 0 && a_16 && ((mrem & 1) == 0);   // whole k columns of the A tile by TMA
        const int nrem = min(BN, p.n - col0);
        const int imax = min(4, max(0, (mrem - wm * 32 + 7) >> 3)), jmax = min(JT, max(0, (nrem - wn * WTN + 7) >> 3));
        double acc[4][JT][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < JT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // issue the copies of k-tile kt into stage st (all threads; one cp.async group per call)
        auto issue = [&](int st, int kt) {
            const int k0 = kt * BK, krem = min(BK, p.k - k0);
            double *as = As + (size_t) st * EA, *bs = Bs + (size_t) st * EB;
            // ---- A
            if (a_bulk) {
                if (warp == 0) {
                    if (lane == 0) mbar_expect_tx(full + st, (unsigned) (krem * mrem) * 8u);
                    __syncwarp();
                    if (lane < krem) {
                        const int gk = k0 + lane;
                        const double *src = gk < k1 ? p.A + (size_t) row0 + (size_t) gk * p.lda
                                                    : p.A2 + (size_t) row0 + (size_t) (gk - k1) * p.lda2;
                        bulk_g2s(as + lane * PA0, src, (unsigned) mrem * 8u, full + st);
                    }
                }
                if (krem < BK) {  // K tail: the missing k columns are zeroed by the threads
                    for (int idx = tid; idx < (BK - krem) * BM; idx += 128) as[(krem + idx / BM) * PA0 + idx % BM] = 0.0;
                    fence_proxy_async();
                }
            } else if (p.ta == 0) {   // As[k][m]: runs along m
                if (a_16) {
                    for (int q = tid; q < BK * BM / 2; q += 128) {
                        const int r = 2 * (q % (BM / 2)), kk = q / (BM / 2), gr = row0 + r, gk = k0 + kk;
                        const bool kv = gk < p.k;
                        const double *src = !kv ? p.A : (gk < k1 ? p.A + (size_t) gr + (size_t) gk * p.lda : p.A2 + (size_t) gr + (size_t) (gk - k1) * p.lda2);
                        const int left = kv ? p.m - gr : 0;
                        cp_async_zfill_16(as + kk * PA0 + r, left > 0 ? src : p.A, left >= 2 ? 16u : (left == 1 ? 8u : 0u));
                    }
                } else {
                    for (int q = tid; q < BK * BM; q += 128) {
                        const int r = q % BM, kk = q / BM, gr = row0 + r, gk = k0 + kk;
                        const bool v = gr < p.m && gk < p.k;
                        const double *src = !v ? p.A : (gk < k1 ? p.A + (size_t) gr + (size_t) gk * p.lda : p.A2 + (size_t) gr + (size_t) (gk - k1) * p.lda2);
                        cp_async_zfill_8(as + kk * PA0 + r, src, v ? 8u : 0u);
                    }
                }
            } else {                  // As[m][k]: runs along k
                if (a_16) {
                    for (int q = tid; q < BM * BK / 2; q += 128) {
                        const int kk = 2 * (q % (BK / 2)), r = q / (BK / 2), gr = row0 + r, gk = k0 + kk;
                        const bool rv = gr < p.m;
                        // (k1 even or no second segment: a 16-byte chunk never straddles the split)
                        const double *src = !rv ? p.A : (gk < k1 ? p.A + (size_t) gk + (size_t) gr * p.lda : p.A2 + (size_t) (gk - k1) + (size_t) gr * p.lda2);
                        const int left = rv ? p.k - gk : 0;
                        cp_async_zfill_16(as + r * PA1 + kk, left > 0 ? src : p.A, left >= 2 ? 16u : (left == 1 ? 8u : 0u));
                    }
                } else {
                    for (int q = tid; q < BM * BK; q += 128) {
                        const int kk = q % BK, r = q / BK, gr = row0 + r, gk = k0 + kk;
                        const bool v = gr < p.m && gk < p.k;
                        const double *src = !v ? p.A : (gk < k1 ? p.A + (size_t) gk + (size_t) gr * p.lda : p.A2 + (size_t) (gk - k1) + (size_t) gr * p.lda2);
                        cp_async_zfill_8(as + r * PA1 + kk, src, v ? 8u : 0u);
                    }
                }
            }
            // ---- B
            if (p.tb == 0) {          // Bs[n][k]: runs along k
                if (b_16) {
                    for (int q = tid; q < BN * BK / 2; q += 128) {
                        const int kk = 2 * (q % (BK / 2)), cc = q / (BK / 2), gc = col0 + cc, gk = k0 + kk;
                        const int left = gc < p.n ? p.k - gk : 0;
                        cp_async_zfill_16(bs + cc * PB0 + kk, left > 0 ? p.B + (size_t) gk + (size_t) gc * p.ldb : p.B,
                                          left >= 2 ? 16u : (left == 1 ? 8u : 0u));
                    }
                } else {
                    for (int q = tid; q < BN * BK; q += 128) {
                        const int kk = q % BK, cc = q / BK, gc = col0 + cc, gk = k0 + kk;
                        const bool v = gc < p.n && gk < p.k;
                        cp_async_zfill_8(bs + cc * PB0 + kk, v ? p.B + (size_t) gk + (size_t) gc * p.ldb : p.B, v ? 8u : 0u);
                    }
                }
            } else {                  // Bs[k][n]: runs along n
                if (b_16) {
                    for (int q = tid; q < BK * BN / 2; q += 128) {
                        const int cc = 2 * (q % (BN / 2)), kk = q / (BN / 2), gc = col0 + cc, gk = k0 + kk;
                        const int left = gk < p.k ? p.n - gc : 0;
                        cp_async_zfill_16(bs + kk * PB1 + cc, left > 0 ? p.B + (size_t) gc + (size_t) gk * p.ldb : p.B,
                                          left >= 2 ? 16u : (left == 1 ? 8u : 0u));
                    }
                } else {
                    for (int q = tid; q < BK * BN; q += 128) {
                        const int cc = q % BN, kk = q / BN, gc = col0 + cc, gk = k0 + kk;
                        const bool v = gc < p.n && gk < p.k;
                        cp_async_zfill_8(bs + kk * PB1 + cc, v ? p.B + (size_t) gc + (size_t) gk * p.ldb : p.B, v ? 8u : 0u);
                    }
                }
            }
            cp_async_commit();
        };

        __syncthreads();  // the previous tile's readers are done with every stage
        // prologue: stages 0 .. S-2 (an empty group is committed when there is nothing to load: group counting stays uniform)
        for (int s = 0; s < S - 1; ++s) {
            if (s < nk) issue(s, s);
            else cp_async_commit();
        }
        for (int it = 0; it < nk; ++it) {
            const int st = it % S;
            cp_async_wait<S - 2>();          // this thread's copies of stage `it` have landed ...
            if (a_bulk) {                    // ... and so have the TMA bytes of its A part
                mbar_wait(full + st, (phase_bits >> st) & 1u);
                phase_bits ^= 1u << st;
            }
            __syncthreads();                 // ... everybody's; and everybody is done reading stage it-1
            if (it + S - 1 < nk) issue((it + S - 1) % S, it + S - 1);   // refill the stage read in iteration it-1
            else cp_async_commit();
            const double *as = As + (size_t) st * EA, *bs = Bs + (size_t) st * EB;
            // Edge tiles: a warp only issues the DMMAs of the 8 x 8 sub-tiles that intersect the matrix (warp-uniform
            // bounds imax / jmax).  The skinny products of the incremental recompression have n = kp = 44 of BN = 64
            // columns: the second warp column then issues 2 of its 4 tile columns -- a quarter of the CTA's DMMAs saved.
            auto mma_stage = [&](auto jm_tag) {
                constexpr int JM = decltype(jm_tag)::value;
#pragma unroll
                for (int ks = 0; ks < BK; ks += 4) {
                    double a[4], b[JM];
                    if (p.ta == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = as[(ks + c) * PA0 + wm * 32 + i * 8 + g];
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = as[(wm * 32 + i * 8 + g) * PA1 + ks + c];
                    }
                    if (p.tb == 0) {
#pragma unroll
                        for (int j = 0; j < JM; ++j) b[j] = bs[(wn * WTN + j * 8 + g) * PB0 + ks + c];
                    } else {
#pragma unroll
                        for (int j = 0; j < JM; ++j) b[j] = bs[(ks + c) * PB1 + wn * WTN + j * 8 + g];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (i < imax) {
#pragma unroll
                            for (int j = 0; j < JM; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                        }
                    }
                }
            };
            switch (jmax) {
                case 6: if constexpr (JT >= 6) mma_stage(std::integral_constant<int, 6>{}); break;
                case 5: if constexpr (JT >= 5) mma_stage(std::integral_constant<int, 5>{}); break;
                case 4: mma_stage(std::integral_constant<int, 4>{}); break;
                case 3: mma_stage(std::integral_constant<int, 3>{}); break;
                case 2: mma_stage(std::integral_constant<int, 2>{}); break;
                case 1: mma_stage(std::integral_constant<int, 1>{}); break;
                default: break;
            }
        }
        cp_async_wait<0>();
        // epilogue: lane holds C[g][2c], C[g][2c+1] of every 8x8 tile
#pragma unroll
        for (int j = 0; j < JT; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int gc = col0 + wn * WTN + j * 8 + 2 * c + h;
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
