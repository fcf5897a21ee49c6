// kernels_dmma.cuh -- batched FP64 GEMM on the tensor pipe (DMMA, mma.sync.m8n8k4.f64) for the GEMM-shaped bulk of
// the recompression: blocked-QR trailing updates, the compact-WY rebuild, V*Sigma = M^T U, the rank-space products.
//
// FP64 has no tcgen05 kind on sm_100a (kinds: f16/tf32/f8f6f4/i8/mx*), so TMEM-accumulator MMA does not apply to this
// path; the FP64 tensor instruction is the warp-level mma.sync DMMA (SASS: DMMA), issued once per 16 cycles per SM
// sub-partition, which leaves the issue slots free for operand traffic (the SIMT DFMA path needs one issue slot every
// 2 cycles just for the math).  Same device-resident GemmProb descriptors as k_gemm_batched.
#pragma once
#include "common.cuh"

namespace hcb {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// CTA tile (32*WM) x (32*WN), WM*WN == 4 warps, each warp a 32x32 tile = 4x4 DMMA tiles (32 accumulator doubles).
// BK = 16: four k-steps of 4 per shared-memory stage; operands staged k-major with a row pitch of tile+4 doubles so
// that the 8-byte fragment loads of a half-warp hit 16 distinct bank pairs.  Global -> register prefetch of the next
// stage overlaps the DMMAs of the current one.  grid = (tiles_bound, n_problems), grid-stride over output tiles.
template<int WM, int WN>
__global__ void __launch_bounds__(128, 3) k_gemm_dmma(const GemmProb<double> *__restrict__ probs) {
    constexpr int BM = 32 * WM, BN = 32 * WN, BK = 16, LDA = BM + 4, LDB = BN + 4;
    constexpr int A_PER_THR = BM * BK / 128, B_PER_THR = BN * BK / 128;
    static_assert(WM * WN == 4, "four warps per CTA");
    const GemmProb<double> p = probs[blockIdx.y];
    if (p.m <= 0 || p.n <= 0) return;
    __shared__ double As[2][BK][LDA];
    __shared__ double Bs[2][BK][LDB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, c = lane & 3;
    const int tiles_m = (p.m + BM - 1) / BM, tiles_n = (p.n + BN - 1) / BN;
    const int nk = (p.k + BK - 1) / BK;

    for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const int row0 = (tile % tiles_m) * BM, col0 = (tile / tiles_m) * BN;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        double ra[A_PER_THR], rb[B_PER_THR];

        auto fetch = [&](int k0) {
#pragma unroll
            for (int q = 0; q < A_PER_THR; ++q) {
                const int idx = tid + q * 128;
                int r, kk;
                if (p.ta == 0) { r = idx % BM; kk = idx / BM; } else { kk = idx % BK; r = idx / BK; }
                const int gr = row0 + r, gk = k0 + kk;
                ra[q] = (gr < p.m && gk < p.k)
                            ? (p.ta == 0 ? p.A[(size_t) gr + (size_t) gk * p.lda] : p.A[(size_t) gk + (size_t) gr * p.lda])
                            : 0.0;
            }
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
        };
        auto stash = [&](int buf) {
#pragma unroll
            for (int q = 0; q < A_PER_THR; ++q) {
                const int idx = tid + q * 128;
                int r, kk;
                if (p.ta == 0) { r = idx % BM; kk = idx / BM; } else { kk = idx % BK; r = idx / BK; }
                As[buf][kk][r] = ra[q];
            }
#pragma unroll
            for (int q = 0; q < B_PER_THR; ++q) {
                const int idx = tid + q * 128;
                int cc, kk;
                if (p.tb == 0) { kk = idx % BK; cc = idx / BK; } else { cc = idx % BN; kk = idx / BN; }
                Bs[buf][kk][cc] = rb[q];
            }
        };

        __syncthreads();  // previous tile's readers are done with the buffers
        if (nk > 0) {
            fetch(0);
            stash(0);
        }
        __syncthreads();
        for (int it = 0; it < nk; ++it) {
            const int buf = it & 1;
            if (it + 1 < nk) fetch((it + 1) * BK);  // global loads in flight while the tensor pipe works
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
            if (it + 1 < nk) {
                stash(buf ^ 1);  // the other buffer was last read in iteration it-1, fenced by the barrier below
            }
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
