// kernels_tlr.cuh -- the fused, batched TLR-GEMM flow: device-side problem setup + the recompression glue kernels.
//
// Mirrors, for a whole batch of (A,B,C) tile triples at once and with all ranks resident on the device,
//   HCore<T>::Gemm              src/api/HCore.cpp:22-344       (operand-mix decision table, temp products)
//   CompressedTile<T>::Gemm     src/operators/concrete/Compressed.cpp:208-694 (stack, QR x2, core SVD, truncate, rebuild)
// The reference runs ~20 launches + 5 cuSOLVER calls + 1 host sync PER TILE; here one launch per phase covers the batch.
#pragma once
#include "common.cuh"
#include "kernels_strip.cuh"
#include <type_traits>
#include "kernels_blas.cuh"
#include "kernels_dmma.cuh"
#include "kernels_qr.cuh"
#include "kernels_svd_rx.cuh"

namespace hcb {

// CholeskyQR2 of the new-column panels (k_cholqr_pass): up to CQ_KP = 48 columns (6 DMMA tiles), padded smem pitch
constexpr int CQ_NT = 6, CQ_KP = 8 * CQ_NT, CQ_P = CQ_KP + 1;

enum Mix { DDD = 0, DDC = 1, DCD = 2, DCC = 3, CDD = 4, CDC = 5, CCD = 6, CCC = 7 };

template<typename T>
struct RecompProb {
    T *UW, *VW;        // QR'd stacks: UW m x r (ld m), VW n x r (ld n)
    T *tauU, *tauV;
    T *M;              // core (p x q) or its transpose, stored a x b (ld a), a >= b
    T *Us, *Vs, *sigma;  // Jacobi outputs: Us a x b (ld a), Vs b x b (ld b)
    T *CU, *CV;        // output tile factors (CU ld m; CV ld = new rank)
    T *VN;             // n x rank scratch (ld n): V factor before the final transpose
    int *rank_ptr;     // C tile's device-resident rank
    int *rk_new;       // scratch: rank chosen by the truncation rule
    int *info;
    int m, n, r, p, q, a, b, transposed, max_rank, active;
    // blocked-QR scratch (per side: 0 = U stack, 1 = V stack)
    T *VC[2];   // clean reflector panels (same shape / ld as UW, VW)
    T *TB[2];   // T blocks: NBQ*NBQ elements per block, block b at TB + b*NBQ*NBQ
    T *WB[2];   // NBQ x (r or rank) GEMM temporaries (ld NBQ), two per side: W at WB, W2 at WB + NBQ*wcols
    int wcols;  // columns each W temporary can hold
    // LQ preconditioning of the core before the Jacobi sweeps: MT = M^T (b x a, ld b) is QR-factored, L = R^T
    T *MT, *tauM, *Lb;  // Lb: a x b (ld a) lower-trapezoidal factor handed to the Jacobi kernel
    // column ordering of the stacks: they are assembled unsorted in SU0 / SV0 (the VC buffers, free until the QR) and
    // gathered into UW / VW sorted by decreasing V-stack column norm; pos[c] = destination column of stack column c
    T *SU0, *SV0;
    int *pos;
    // Incremental U side (round 2): when the tile's state says "CU has orthonormal columns" (it does after every
    // recompression, Compressed.cpp:558-560) only the kp NEW columns P are orthogonalised against CU (block classical
    // Gram-Schmidt, twice) and QR-factored:  [CU | P] = [CU | Q2] * [[I, G], [0, R2]].  The 2 m r^2 Householder QR of
    // the U stack becomes 8 m kc kp GEMM flops + a kp-column panel, and the rebuild CU' = [CU | Q2] * Us is one GEMM.
    int inc;            // 1: U side incremental
    int kc, kp;         // old rank / new columns (r = kc + kp)
    T *Pn;              // P (m x kp, ld m): orthogonalised in place against CU, then QR-factored in place
    T *Gu, *Gu2;        // first / second pass coefficients CU^T P (kc x kp, ld kc)
    T *Q2;              // explicit Q2 (m x kp, ld m)
    T *TU;              // rebuild target [CU | Q2] * Us[:, :rk] (m x rk, ld m), copied into CU by k_finalize
    // Incremental V side (round 2): when the state says "the rows of CV are mutually orthogonal" (CV = diag(sigma) W^T
    // after every recompression, Compressed.cpp:598-622) the new right columns Y are orthogonalised against W the same way:
    //   [beta CV^T | Y] = [W | Q2v] * RV,  RV = [[beta S, Gv], [0, R2v]]  (r x r).
    // The column-sorted, graded triangular factor the Jacobi sweeps need (5 sweeps instead of ~17) comes from a
    // Householder QR of the SMALL r x r matrix RV * Pi instead of the n x r stack (R only: no Q is ever applied), and
    // V' = [W | Q2v] * (RV RU^T Us) is GEMMs.
    int vinc;           // 1: V side incremental
    int ldvw;           // leading dimension of the QR-factored V panel k_extract_r reads R from (n, or r when vinc)
    T *sig0;            // row norms of CV = the old singular values (kc)
    T *Yn;              // Y (n x kp, ld n): orthogonalised in place against W, then QR-factored in place
    T *Hv, *Zv;         // Hv = sum over the two passes of CV * Y (kc x kp, ld kc);  Zv = diag(sigma)^-2 * (CV * Y) of the pass
    T *Q2v, *Q2vT;      // explicit Q2v (n x kp, ld n) and its transpose (kp x n, ld kp)
    T *RVp;             // copy of RV * Pi (r x r, ld r) kept for V S = RV Pi (RU Pi)^T Us (the QR overwrites its own copy)
    T *T1;              // (RU Pi)^T Us (r x b, ld r)
    T beta_c;           // beta of the call (scales the old singular values in RV)
    // Both sides incremental: K = RU RV^T = diag(beta S, 0) + Xu Xv^T is a rank-kp update of a diagonal matrix, so the
    // three r^3 GEMMs around the Jacobi kernel collapse to rank-kp ones:
    //   K' = (RU Pi) R'^T :  row i < kc of K' is column pos[i] of R' (gather), plus  Xu * (R'[:, pos[kc..]])^T
    //   V S' = RV RU^T Us  =  diag(beta S, 0) Us  +  Xv (Xu^T Us)
    int both;           // 1: inc && vinc
    T *Xu, *Xv;         // [Gu ; R2u], [Gv ; R2v]  (r x kp, ld r)
    T *Rn;              // R'[:, pos[kc + l]], l < kp  (r x kp, ld r)
    T *T1b;             // Xu^T Us  (kp x b, ld kp)
    int lp, lq;         // leading dimensions of the extracted triangles MT (p x r) / Lb (q x r): p, q rounded up to even
                        // so that the core GEMM's row-contiguous operands qualify for TMA bulk copies
    int ldus;           // leading dimension of Us: a rounded up to even, so that Us is a 16-byte-copy operand of the GEMMs
    // CholeskyQR2 of the kp new columns (see k_cholqr_factor): scratch (4 kp x kp matrices: G, S, R1, G2), the second Q buffer,
    // the per-side "CholQR failed, Householder takes over" flags and "explicit Q2 is already in place"
    T *CQw[2], *CQt[2];
    int *cq_fail;       // [2] per tile
    int q2_done[2];
    unsigned long long *stats;  // context counters (hcb_ctx_stats): [0] CholQR2 sides done, [1] fell back in pass 0, [2] in pass 1,
                                // [3] Gram-Schmidt second passes skipped, [4] run, [5] negligible new columns deflated,
                                // [6] r x r triangular factors by Cholesky, [7] fell back to the Householder R-only QR
    int *state;         // C tile's device state word (may be null)
    int fixed_rank;     // per-tile fixed rank (0: batch value)
};

// Preconditioning of the product term (CCC): an orthogonal J (ka x ka) that makes the rows of J^T (T1 * BR) mutually
// orthogonal.  U_AB' = AL * J and V_AB' = J^T * T2 give the SAME product AL * T2, but the V stack then has orthogonal
// columns inside each of its two parts, and -- with the columns sorted by norm -- the core K = RU * RV^T needs ~6
// Jacobi sweeps instead of ~20 (11 with LQ preconditioning); numpy emulation of the exact kernel logic, r = 126..204.
template<typename T>
struct PrecondProb {
    const T *T1;   // ka x kb (ld ka)
    const T *BR;   // right factor of op(B): kb x n as (ptr, ld, trans)
    T *J;          // ka x ka (ld ka)
    T *T1J;        // kb x ka (ld kb) = T1^T * J
    int ka, kb, n, ldbr, tbr, active;
};

// One panel to be QR-factored by the blocked machinery (a stack panel of the recompression, or the transposed core).
template<typename T>
struct PanelDesc {
    T *A, *tau;   // m x n panel (ld m), tau[min(m,n)]
    T *VC;        // clean reflector panel (same shape / ld)
    T *TB;        // NBQ*NBQ per block
    T *WB;        // 2 * NBQ * wcols
    int m, n, wcols, active;
};

template<typename T>
struct SetupArgs {
    const hcb_tile *A, *B, *C;  // device copies of the descriptor arrays
    int n_tiles, mix, opA, opB;
    T alpha, beta;
    // per-tile scratch (element offsets inside one tile's slab) and slab stride
    T *ws;
    size_t slab, o_w1, o_w2, o_uw, o_vw, o_tauu, o_tauv, o_m, o_j, o_us, o_vs, o_sig, o_vn;
    size_t o_vcu, o_vcv, o_tbu, o_tbv, o_wbu, o_wbv;  // blocked-QR scratch
    size_t o_mt, o_taum, o_lb, o_vcm, o_tbm, o_wbm;     // core LQ preconditioning scratch
    int wcols;
    int use_lq;              // 1: LQ-precondition the core before the Jacobi sweeps
    PanelDesc<T> *pd_stack;  // 2 per tile (U stack, V stack)
    PanelDesc<T> *pd_core;   // 1 per tile (transposed core)
    PrecondProb<T> *pc;      // 1 per tile (CCC only)
    size_t o_pj, o_pos;      // J (kA_b^2 elements) and pos (r_b ints, stored in T-sized slots)
    QrProb<T> *qr_core;      // 1 per tile: unblocked QR of the transposed core (small-rank path)
    LqProb<T> *lq;           // 1 per tile
    int kA_b, kB_b, kC_b, r_b;  // rank bounds the scratch was sized for
    int *rk_new;                // n_tiles ints
    int *info;                  // n_tiles ints (may be null)
    GemmProb<T> *g1, *g2, *g3, *gv;  // gv: V diag(sigma) = M^T Us after the Jacobi kernel
    GemmProb<T> *gc;                 // core = RU RV^T (or its transpose) from the extracted triangles, see k_extract_r
    CopyProb<T> *cp;            // 4 per tile
    QrProb<T> *qr;              // 2 per tile
    ReflProb<T> *rf;            // 2 per tile
    SvdProb<T> *svd;
    RecompProb<T> *rc;
    // incremental U side
    int inc_enabled;            // host: blocked fp64 strip path in use
    size_t o_gu, o_gu2, o_q2, o_tu;   // scratch offsets: Gu, Gu2 (kC_b x kp_b), [VCp | Q2] (2 m kp_b), TU (m x rank bound)
    GemmProb<T> *gi;            // 4 per tile, arrays of n: gi + q*n_tiles, q = 0..3 (G = CU^T P, P -= CU G, twice)
    int *cq_fail;               // [2 * n_tiles] flags (zeroed here)
    unsigned long long *stats;  // context counters
    int cq_enabled;
    PanelDesc<T> *pd_inc;       // 1 per tile: QR of the orthogonalised P
    StripJob *inc_sj;           // nst_inc per tile: explicit Q2 = H_0 .. H_{kp-1} [I; 0]
    int nst_inc, kp_b;
    int inc_refresh;            // consecutive incremental updates allowed before one full re-factorisation
    // incremental V side
    int vinc_enabled;
    size_t o_sig0, o_hv, o_zv, o_q2v, o_q2vt, o_rvp, o_t1;
    GemmProb<T> *giv;           // 4 arrays of n (Hv = CV Y, Y -= CV^T Zv, twice)
    GemmProb<T> *gt1, *gx;      // T1 = (RU Pi)^T Us ;  X = (RV Pi) T1  (arrays of n)
    PanelDesc<T> *pd_vcore;     // 1 per tile: the r x r panel RV * Pi (R-only QR)
    size_t o_xu, o_xv, o_rn, o_t1b;
    int *err_flag;              // context-level sticky error word (bit 2: a rank exceeded its bound)
};

template<typename T>
__device__ __forceinline__ GemmProb<T> mk_gemm(const T *A, int lda, int ta, const T *B, int ldb, int tb, T *C, int ldc,
                                               int m, int n, int k, T alpha, T beta) {
    GemmProb<T> g;
    g.A = A; g.B = B; g.C = C; g.m = m; g.n = n; g.k = k; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.ta = ta; g.tb = tb; g.alpha = alpha; g.beta = beta;
    g.A2 = nullptr; g.k1 = k; g.lda2 = 1;
    return g;
}

template<typename T>
__device__ __forceinline__ CopyProb<T> mk_copy(const T *src, int lds, T *dst, int ldd, int rows, int cols, int trans,
                                               T scale) {
    CopyProb<T> c;
    c.src = src; c.dst = dst; c.rows = rows; c.cols = cols; c.lds = lds; c.ldd = ldd; c.trans = trans; c.scale = scale;
    return c;
}

// One thread per tile triple: reads the (device-resident) ranks and writes every problem descriptor of the step.
// Operand factors (HCore.cpp:57-110): a compressed operand X = XU * XV contributes, under op,
//   left factor  L (rows x kx) = op ? XV^T : XU        right factor R (kx x cols) = op ? XU^T : XV.
template<typename T>
__global__ void k_setup_tlr(SetupArgs<T> s) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.n_tiles) return;
    const hcb_tile A = s.A[t], B = s.B[t], C = s.C[t];
    const int m = C.m, n = C.n;
    const int kdim = s.opA ? A.m : A.n;  // inner dimension of op(A) op(B)
    const bool ac = A.type == HCB_TILE_COMPRESSED, bc = B.type == HCB_TILE_COMPRESSED, cc = C.type == HCB_TILE_COMPRESSED;
    const int ka = ac ? *A.d_rank : 0, kb = bc ? *B.d_rank : 0, kc = cc ? *C.d_rank : 0;
    T *slab = s.ws + (size_t) t * s.slab;
    GemmProb<T> g1 = mk_gemm<T>(nullptr, 1, 0, nullptr, 1, 0, nullptr, 1, 0, 0, 0, T(0), T(0)), g2 = g1, g3 = g1, gv = g1, gc = g1;
    CopyProb<T> c0 = mk_copy<T>(nullptr, 1, nullptr, 1, 0, 0, 0, T(0)), c1 = c0, c2 = c0, c3 = c0;
    QrProb<T> q0{nullptr, nullptr, 0, 0, 1}, q1 = q0;
    ReflProb<T> r0{nullptr, nullptr, nullptr, 0, 0, 1, 0, 0, 1, 0, 0, nullptr}, r1 = r0;
    SvdProb<T> sv{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 1, 1};
    RecompProb<T> rc;
    memset(&rc, 0, sizeof(rc));
    PanelDesc<T> pdu, pdv, pdm;
    memset(&pdu, 0, sizeof(pdu)); memset(&pdv, 0, sizeof(pdv)); memset(&pdm, 0, sizeof(pdm));
    QrProb<T> qm{nullptr, nullptr, 0, 0, 1};
    LqProb<T> lq{nullptr, nullptr, 0, 0};
    int bad = 0;
    if ((ac && (ka > s.kA_b || ka < 0)) || (bc && (kb > s.kB_b || kb < 0)) || (cc && (kc > s.kC_b || kc < 0))) bad = 1;

    // operand views
    const T *Ad = (const T *) A.d_data, *Bd = (const T *) B.d_data;
    T *Cd = (T *) C.d_data;
    // dense: (ptr, ld, trans) ; compressed: left factor AL (m x ka), right factor AR (ka x kdim)
    const T *AU = Ad, *AV = Ad + (size_t) A.m * A.max_rank;
    const T *BU = Bd, *BV = Bd + (size_t) B.m * B.max_rank;
    const T *ALp = s.opA ? AV : AU; const int ALld = s.opA ? ka : A.m, ALt = s.opA;
    const T *ARp = s.opA ? AU : AV; const int ARld = s.opA ? A.m : ka, ARt = s.opA;
    const T *BLp = s.opB ? BV : BU; const int BLld = s.opB ? kb : B.m, BLt = s.opB;
    const T *BRp = s.opB ? BU : BV; const int BRld = s.opB ? B.m : kb, BRt = s.opB;
    T *CU = Cd, *CV = Cd + (size_t) C.m * C.max_rank;
    T *W1 = slab + s.o_w1, *W2 = slab + s.o_w2, *UW = slab + s.o_uw, *VW = slab + s.o_vw;
    T *SU0 = slab + s.o_vcu, *SV0 = slab + s.o_vcv;  // unsorted stacks live in the (still unused) VC buffers
    T *PJ = slab + s.o_pj;
    PrecondProb<T> pc;
    memset(&pc, 0, sizeof(pc));
    const T one = T(1), zero = T(0);
    int kp = 0;  // rank of the product term entering the recompression
    GemmProb<T> gi0 = g1, gi1 = g1, gi2 = g1, gi3 = g1;
    GemmProb<T> gv0 = g1, gv1 = g1, gv2 = g1, gv3 = g1, gt1 = g1, gx = g1;
    PanelDesc<T> pdi, pdiv, pdvc;
    memset(&pdi, 0, sizeof(pdi)); memset(&pdiv, 0, sizeof(pdiv)); memset(&pdvc, 0, sizeof(pdvc));
    int inc = 0, vinc = 0;

    if (!bad) {
        switch (s.mix) {
            case DDD:  // HCore.cpp:300-313 -> DenseTile::Gemm (Dense.cpp:46-108)
                g1 = mk_gemm<T>(Ad, A.ld, s.opA, Bd, B.ld, s.opB, Cd, C.ld, m, n, kdim, s.alpha, s.beta);
                break;
            case CDD:  // T = AR*op(B) (ka x n) ; C = alpha*AL*T + beta*C  (HCore.cpp:246-258)
                g1 = mk_gemm<T>(ARp, ARld, ARt, Bd, B.ld, s.opB, W1, ka, ka, n, kdim, one, zero);
                g2 = mk_gemm<T>(ALp, ALld, ALt, W1, ka, 0, Cd, C.ld, m, n, ka, s.alpha, s.beta);
                break;
            case DCD:  // T = op(A)*BL (m x kb) ; C = alpha*T*BR + beta*C  (HCore.cpp:234-245)
                g1 = mk_gemm<T>(Ad, A.ld, s.opA, BLp, BLld, BLt, W1, m, m, kb, kdim, one, zero);
                g2 = mk_gemm<T>(W1, m, 0, BRp, BRld, BRt, Cd, C.ld, m, n, kb, s.alpha, s.beta);
                break;
            case CCD:  // T1 = AR*BL ; T2 = T1*BR ; C = alpha*AL*T2 + beta*C  (HCore.cpp:221-232)
                g1 = mk_gemm<T>(ARp, ARld, ARt, BLp, BLld, BLt, W1, ka, ka, kb, kdim, one, zero);
                g2 = mk_gemm<T>(W1, ka, 0, BRp, BRld, BRt, W2, ka, ka, n, kb, one, zero);
                g3 = mk_gemm<T>(ALp, ALld, ALt, W2, ka, 0, Cd, C.ld, m, n, ka, s.alpha, s.beta);
                break;
            case DDC:  // T = alpha*op(A)*op(B) ; T += beta*CU*CV ; C := (U = T, V = I)  (HCore.cpp:272-299)
                g1 = mk_gemm<T>(Ad, A.ld, s.opA, Bd, B.ld, s.opB, W1, m, m, n, kdim, s.alpha, zero);
                g2 = mk_gemm<T>(CU, m, 0, CV, kc, 0, W1, m, m, n, kc, s.beta, one);
                break;
            case CDC:  // U_AB = AL (m x ka) ; V_AB^T = op(B)^T * AR^T (n x ka)
                kp = ka;
                c1 = mk_copy<T>(ALp, ALld, SU0 + (size_t) m * kc, m, m, ka, ALt, s.alpha);
                g1 = mk_gemm<T>(Bd, B.ld, !s.opB, ARp, ARld, !ARt, SV0 + (size_t) n * kc, n, n, ka, kdim, one, zero);
                break;
            case DCC:  // U_AB = alpha*op(A)*BL (m x kb) ; V_AB^T = BR^T (n x kb)
                kp = kb;
                g1 = mk_gemm<T>(Ad, A.ld, s.opA, BLp, BLld, BLt, SU0 + (size_t) m * kc, m, m, kb, kdim, s.alpha, zero);
                c1 = mk_copy<T>(BRp, BRld, SV0 + (size_t) n * kc, n, n, kb, !BRt, one);
                break;
            case CCC:  // T1 = AR*BL (ka x kb) ; V_AB^T = BR^T * T1^T (n x ka) ; U_AB = AL  (HCore.cpp:221-232,300-313)
                kp = ka;
                g1 = mk_gemm<T>(ARp, ARld, ARt, BLp, BLld, BLt, W1, ka, ka, kb, kdim, one, zero);
                // k_precond_product: J, T1J = T1^T J (W2).  V_AB'^T = BR^T * T1J, U_AB' = alpha * AL * J
                pc = PrecondProb<T>{W1, BRp, PJ, W2, ka, kb, n, BRld, BRt, 1};
                g2 = mk_gemm<T>(BRp, BRld, !BRt, W2, kb, 0, SV0 + (size_t) n * kc, n, n, ka, kb, one, zero);
                g3 = mk_gemm<T>(ALp, ALld, ALt, PJ, ka, 0, SU0 + (size_t) m * kc, m, m, ka, ka, s.alpha, zero);
                break;
        }
        if (cc && s.mix != DDC) {
            // stacks (Compressed.cpp:332-349, 378-379): UW = [CU | alpha*U_AB], VW = [beta*CV^T | V_AB^T]
            const int r = kc + kp;
            if (r > s.r_b) {
                bad = 1;
            } else {
                c0 = mk_copy<T>(CU, m, SU0, m, m, kc, 0, one);
                c2 = mk_copy<T>(CV, kc, SV0, n, n, kc, 1, s.beta);
                rc.SU0 = SU0; rc.SV0 = SV0;
                rc.pos = reinterpret_cast<int *>(slab + s.o_pos);
                const int p = m < r ? m : r, q = n < r ? n : r;
                q0 = QrProb<T>{UW, slab + s.o_tauu, m, r, m};
                q1 = QrProb<T>{VW, slab + s.o_tauv, n, r, n};
                rc.UW = UW; rc.VW = VW; rc.tauU = slab + s.o_tauu; rc.tauV = slab + s.o_tauv;
                rc.M = slab + s.o_m; rc.Us = slab + s.o_us; rc.Vs = slab + s.o_vs; rc.sigma = slab + s.o_sig;
                rc.CU = CU; rc.CV = CV; rc.VN = slab + s.o_vn; rc.rank_ptr = C.d_rank; rc.rk_new = s.rk_new + t;
                rc.info = s.info ? s.info + t : nullptr;
                rc.m = m; rc.n = n; rc.r = r; rc.p = p; rc.q = q; rc.max_rank = C.max_rank; rc.active = 1;
                rc.VC[0] = slab + s.o_vcu; rc.VC[1] = slab + s.o_vcv;
                rc.TB[0] = slab + s.o_tbu; rc.TB[1] = slab + s.o_tbv;
                rc.WB[0] = slab + s.o_wbu; rc.WB[1] = slab + s.o_wbv;
                rc.wcols = s.wcols;
                rc.MT = slab + s.o_mt; rc.tauM = slab + s.o_taum; rc.Lb = slab + s.o_lb;
                rc.transposed = p < q;
                rc.a = rc.transposed ? q : p;
                rc.b = rc.transposed ? p : q;
                // Jacobi runs on M itself when the stacks are sorted (use_lq == 0), else on its LQ factor L
                rc.ldus = (rc.a + 1) & ~1;
                sv = SvdProb<T>{s.use_lq ? rc.Lb : rc.M, slab + s.o_j, rc.Us, rc.Vs, rc.sigma, rc.info, rc.a, rc.b, rc.a, rc.ldus, rc.b};
                pdu = PanelDesc<T>{UW, rc.tauU, rc.VC[0], rc.TB[0], rc.WB[0], m, r, s.wcols, 1};
                pdv = PanelDesc<T>{VW, rc.tauV, rc.VC[1], rc.TB[1], rc.WB[1], n, r, s.wcols, 1};
                pdm = PanelDesc<T>{rc.MT, rc.tauM, slab + s.o_vcm, slab + s.o_tbm, slab + s.o_wbm, rc.b, rc.a, s.wcols, 1};
                qm = QrProb<T>{rc.MT, rc.tauM, rc.b, rc.a, rc.b};
                lq = LqProb<T>{rc.MT, rc.Lb, rc.a, rc.b};
                gv = mk_gemm<T>(rc.M, rc.a, 1, rc.Us, rc.ldus, 0, rc.Vs, rc.b, rc.b, rc.b, rc.a, one, zero);
                // core from the extracted triangles RU (p x r, ld p, in MT) and RV (q x r, ld q, in Lb)
                rc.lp = s.use_lq ? p : ((p + 1) & ~1);
                rc.lq = s.use_lq ? q : ((q + 1) & ~1);
                if (rc.transposed) gc = mk_gemm<T>(rc.Lb, rc.lq, 0, rc.MT, rc.lp, 1, rc.M, rc.a, q, p, r, one, zero);
                else gc = mk_gemm<T>(rc.MT, rc.lp, 0, rc.Lb, rc.lq, 1, rc.M, rc.a, p, q, r, one, zero);
                // rebuild (Compressed.cpp:551-560, 611-628): CU = Q_U [Unew;0], VN = Q_V [Vfac;0], rank read on device
                r0 = ReflProb<T>{UW, rc.tauU, CU, m, p, m, m, 0, m, 0, 0, rc.rk_new};
                r1 = ReflProb<T>{VW, rc.tauV, rc.VN, n, q, n, n, 0, n, 0, 0, rc.rk_new};
                rc.state = C.d_state;
                rc.fixed_rank = C.fixed_rank;
                rc.kc = kc; rc.kp = kp;
                // incremental U side: CU is known to be orthonormal, the new columns fit (r <= m, kp within the scratch)
                // (bits 8.. of the state word count the consecutive incremental updates: every s.inc_refresh-th update
                // takes the full path, so that the loss of orthogonality of CU -- it adds up, ~0.05 * accuracy per
                // update with the accuracy-aware Jacobi stop -- is reset)
                const int st_word = C.d_state ? *C.d_state : 0;
                inc = s.inc_enabled && (st_word & HCB_STATE_ORTHO_U) && (st_word >> 8) < s.inc_refresh && kc >= 1 &&
                      kp >= 1 && r <= m && r <= n && kp <= s.kp_b;
                if (inc) {
                    rc.inc = 1;
                    rc.Pn = SU0 + (size_t) m * kc;
                    rc.Gu = slab + s.o_gu; rc.Gu2 = slab + s.o_gu2;
                    T *VCp = slab + s.o_q2;
                    rc.Q2 = VCp + (size_t) m * s.kp_b;
                    rc.TU = slab + s.o_tu;
                    c0.rows = 0;          // CU is read in place
                    pdu.active = 0;       // no Householder QR of the U stack
                    q0.m = 0;
                    gi0 = mk_gemm<T>(CU, m, 1, rc.Pn, m, 0, rc.Gu, kc, kc, kp, m, one, zero);     // G  = CU^T P
                    gi1 = mk_gemm<T>(CU, m, 0, rc.Gu, kc, 0, rc.Pn, m, m, kp, kc, -one, one);     // P -= CU G
                    gi2 = mk_gemm<T>(CU, m, 1, rc.Pn, m, 0, rc.Gu2, kc, kc, kp, m, one, zero);    // G2 = CU^T P
                    gi3 = mk_gemm<T>(CU, m, 0, rc.Gu2, kc, 0, rc.Pn, m, m, kp, kc, -one, one);    // P -= CU G2
                    pdi = PanelDesc<T>{rc.Pn, rc.tauU, VCp, rc.TB[0], rc.WB[0], m, kp, s.wcols, 1};
                }
                // incremental V side: the rows of CV are known to be orthogonal
                rc.ldvw = n;
                rc.beta_c = s.beta;
                vinc = s.vinc_enabled && (st_word & HCB_STATE_ORTHO_V) && (st_word >> 8) < s.inc_refresh && kc >= 1 && kp >= 1 &&
                       r <= m && r <= n && kp <= s.kp_b && s.beta != T(0);
                if (vinc) {
                    rc.vinc = 1;
                    rc.ldvw = r;
                    rc.sig0 = slab + s.o_sig0;
                    rc.Yn = SV0 + (size_t) n * kc;
                    rc.Hv = slab + s.o_hv; rc.Zv = slab + s.o_zv;
                    T *VCy = slab + s.o_q2v;
                    rc.Q2v = VCy + (size_t) n * s.kp_b;
                    rc.Q2vT = slab + s.o_q2vt;
                    rc.RVp = slab + s.o_rvp;
                    rc.T1 = slab + s.o_t1;
                    c2.rows = 0;          // CV is read in place
                    pdv.active = 0;       // no Householder QR of the n x r stack
                    q1.m = 0;
                    gv0 = mk_gemm<T>(CV, kc, 0, rc.Yn, n, 0, rc.Hv, kc, kc, kp, n, one, zero);    // Hv  = CV Y
                    gv1 = mk_gemm<T>(CV, kc, 1, rc.Zv, kc, 0, rc.Yn, n, n, kp, kc, -one, one);    // Y  -= CV^T Zv
                    gv2 = mk_gemm<T>(CV, kc, 0, rc.Yn, n, 0, rc.Zv, kc, kc, kp, n, one, zero);    // Zv  = CV Y   (second pass)
                    gv3 = gv1;
                    pdiv = PanelDesc<T>{rc.Yn, rc.tauV, VCy, rc.TB[1], rc.WB[1], n, kp, s.wcols, 1};
                    // R-only QR of the r x r matrix RV * Pi, assembled in the (otherwise unused) VW buffer with ld r; its clean
                    // reflectors / T blocks reuse the V stack's scratch (free once the Y panel has been expanded into Q2v)
                    pdvc = PanelDesc<T>{VW, slab + s.o_taum, slab + s.o_vcm, slab + s.o_tbm, slab + s.o_wbm, r, r, s.wcols, 1};
                    // V S' = (RV Pi) (RU Pi)^T Us  instead of  K'^T Us  (K' = K O carries the QR's orthogonal factor on the right)
                    gv.m = 0;
                    gt1 = mk_gemm<T>(rc.MT, rc.lp, 1, rc.Us, rc.ldus, 0, rc.T1, r, r, rc.b, p, one, zero);
                    gx = mk_gemm<T>(rc.RVp, r, 0, rc.T1, r, 0, rc.Vs, rc.b, r, rc.b, r, one, zero);
                    if (inc) {  // both sides incremental: rank-kp forms of the core and of V S'
                        rc.both = 1;
                        rc.Xu = slab + s.o_xu; rc.Xv = slab + s.o_xv; rc.Rn = slab + s.o_rn; rc.T1b = slab + s.o_t1b;
                        gc = mk_gemm<T>(rc.Xu, r, 0, rc.Rn, r, 1, rc.M, rc.a, r, r, kp, one, one);          // K' += Xu Rn^T
                        gt1 = mk_gemm<T>(rc.Xu, r, 1, rc.Us, rc.ldus, 0, rc.T1b, kp, kp, rc.b, r, one, zero);  // T1b = Xu^T Us
                        gx = mk_gemm<T>(rc.Xv, r, 0, rc.T1b, kp, 0, rc.Vs, rc.b, r, rc.b, kp, one, zero);   // X   = Xv T1b
                    }
                }
            }
        }
    }
    if (bad) {
        g1.m = g2.m = g3.m = gv.m = gc.m = 0;
        c0.rows = c1.rows = c2.rows = c3.rows = 0;
        q0.m = q1.m = 0;
        r0.k = r1.k = 0; r0.nc = r1.nc = 0; r0.nc_dev = r1.nc_dev = nullptr;
        sv.a = sv.b = 0;
        rc.active = 0;
        pdu.active = pdv.active = pdm.active = 0;
        pc.active = 0;
        qm.m = qm.n = 0;
        lq.a = lq.b = 0;
        gi0.m = gi1.m = gi2.m = gi3.m = 0;
        gv0.m = gv1.m = gv2.m = gv3.m = gt1.m = gx.m = 0;
        pdi.active = pdiv.active = pdvc.active = 0;
        inc = vinc = 0;
        rc.inc = rc.vinc = rc.both = 0;
        // rank exceeded the bound the scratch was sized for: tile left untouched.  Sticky in d_info (flags are OR-ed, the
        // caller of the entry point zeroes them) AND in the context's error word, so that a caller without an info buffer
        // still hears about it at the next hcb_ctx_sync.
        if (s.info) atomicOr(s.info + t, 4);
        if (s.err_flag) atomicOr(s.err_flag, 4);
    }
    {   // CholeskyQR2 of the two new-column panels: scratch, flags
        rc.cq_fail = s.cq_fail ? s.cq_fail + 2 * (size_t) t : nullptr;
        rc.stats = s.stats;
        rc.q2_done[0] = rc.q2_done[1] = 0;
        for (int side = 0; side < 2; ++side) {
            const bool on = s.cq_enabled && rc.active && (side ? vinc : inc) && rc.kp >= 1 && rc.kp <= CQ_KP &&
                            rc.kp * rc.kp <= 2 * NBQ * s.wcols;
            rc.CQw[side] = on ? rc.WB[side] : nullptr;
            rc.CQt[side] = on ? (side ? pdiv.VC : pdi.VC) : nullptr;
            if (rc.cq_fail) rc.cq_fail[side] = on ? 0 : 1;
        }
    }
    for (int q = 0; q < 4; ++q) s.gi[(size_t) q * s.n_tiles + t] = q == 0 ? gi0 : (q == 1 ? gi1 : (q == 2 ? gi2 : gi3));
    for (int q = 0; q < 4; ++q) s.giv[(size_t) q * s.n_tiles + t] = q == 0 ? gv0 : (q == 1 ? gv1 : (q == 2 ? gv2 : gv3));
    s.gt1[t] = gt1; s.gx[t] = gx;
    s.pd_inc[t] = pdi;
    s.pd_inc[s.n_tiles + t] = pdiv;
    s.pd_vcore[t] = pdvc;
    if constexpr (std::is_same<T, double>::value) {
        // explicit Q2 (U side: jobs [0, n*nst), V side: jobs [n*nst, 2*n*nst)): strip st of [I; 0] takes the panel's blocks
        // last to first
        for (int side = 0; side < 2; ++side) {
            const bool on = side ? vinc : inc;
            const int rows = side ? n : m;
            const PanelDesc<T> &pd = side ? pdiv : pdi;
            T *Q = side ? rc.Q2v : rc.Q2;
            for (int st = 0; st < s.nst_inc; ++st) {
                StripJob j{nullptr, nullptr, nullptr, 1, 1, 0, 0, 0, 0, 0, -1, 0};
                if (on && st * NBQ < kp) {
                    const int nbp = (kp + NBQ - 1) / NBQ, nc = (kp - st * NBQ) < NBQ ? (kp - st * NBQ) : NBQ;
                    j = StripJob{Q + (size_t) st * NBQ * rows, pd.VC, pd.TB, rows, rows, rows, nc, kp, nbp - 1, nbp, -1, 0};
                }
                s.inc_sj[((size_t) side * s.n_tiles + t) * s.nst_inc + st] = j;
            }
        }
    }
    s.g1[t] = g1; s.g2[t] = g2; s.g3[t] = g3; s.gv[t] = gv; s.gc[t] = gc;
    s.cp[4 * t + 0] = c0; s.cp[4 * t + 1] = c1; s.cp[4 * t + 2] = c2; s.cp[4 * t + 3] = c3;
    s.qr[2 * t + 0] = q0; s.qr[2 * t + 1] = q1;
    s.rf[2 * t + 0] = r0; s.rf[2 * t + 1] = r1;
    s.svd[t] = sv;
    s.rc[t] = rc;
    s.pd_stack[2 * t + 0] = pdu; s.pd_stack[2 * t + 1] = pdv;
    s.pd_core[t] = pdm;
    s.qr_core[t] = qm;
    s.pc[t] = pc;
    s.lq[t] = lq;
}

// One CTA (256 threads) per tile.  N = (T1 * D)^T with D = row norms of BR (kb x ka, columns = rows of T1*D) is
// orthogonalised by one-sided Jacobi while the rotations are accumulated in J (orthogonal to rounding by
// construction); then T1J = T1^T * J.  Everything lives in shared memory; if it does not fit, J = I (no
// preconditioning, same product).
template<typename T>
__global__ void __launch_bounds__(1024) k_precond_product(const PrecondProb<T> *__restrict__ probs, int smem_elems,
                                                           int max_sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw_pc[];
    T *sm = reinterpret_cast<T *>(smem_raw_pc);
    const PrecondProb<T> p = probs[blockIdx.x];
    if (!p.active) return;
    const int ka = p.ka, kb = p.kb;
    if (ka <= 0 || kb <= 0) return;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
    __shared__ int s_rot;
    const bool fits = (size_t) ka * kb + (size_t) ka * ka + (size_t) kb <= (size_t) smem_elems;
    if (!fits) {
        for (int idx = tid; idx < ka * ka; idx += nthr) p.J[idx] = (idx % ka == idx / ka) ? T(1) : T(0);
        for (int idx = tid; idx < kb * ka; idx += nthr) p.T1J[idx] = p.T1[(size_t) (idx / kb) + (size_t) (idx % kb) * ka];
        return;
    }
    T *N = sm, *J = sm + (size_t) ka * kb, *d = J + (size_t) ka * ka;
    // row norms of BR (kb x n).  Stored N (ld = kb): consecutive threads walk down a column (rows j contiguous), a
    // group of kbp threads per column, nthr / kbp columns in flight; stored T: lanes run along the row.
    for (int j = tid; j < kb; j += nthr) d[j] = T(0);
    __syncthreads();
    if (!p.tbr) {
        int kbp = 32;
        while (kbp < kb && kbp < nthr) kbp <<= 1;      // threads per column (power of two >= min(kb, nthr))
        const int groups = nthr / kbp, gidx = tid / kbp, jl = tid % kbp;
        for (int j = jl; j < kb && gidx < groups; j += kbp) {  // (a partial last group sits out)
            T ss = T(0);
            for (int c = gidx; c < p.n; c += groups) {
                const T x = p.BR[(size_t) j + (size_t) c * p.ldbr];
                ss = fma(x, x, ss);
            }
            atomicAdd(&d[j], ss);
        }
    } else {
        for (int j = w; j < kb; j += nw) {
            T ss = T(0);
            for (int c = lane; c < p.n; c += 32) {
                const T x = p.BR[(size_t) c + (size_t) j * p.ldbr];
                ss = fma(x, x, ss);
            }
            ss = warp_sum(ss);
            if (lane == 0) d[j] = ss;
        }
    }
    __syncthreads();
    for (int j = tid; j < kb; j += nthr) d[j] = t_sqrt(d[j]);
    for (int idx = tid; idx < ka * ka; idx += nthr) J[idx] = (idx % ka == idx / ka) ? T(1) : T(0);
    __syncthreads();
    for (int idx = tid; idx < ka * kb; idx += nthr) {
        const int j = idx % kb, i = idx / kb;  // N(j, i) = T1(i, j) * d(j)
        N[idx] = p.T1[(size_t) i + (size_t) j * ka] * d[j];
    }
    __syncthreads();
    // J only has to be a good preconditioner (ANY orthogonal J gives the same product), so the sweeps stop at a loose
    // relative orthogonality of sqrt(eps), ignore columns at rounding level, and are capped (max_sweeps).
    // Instruction diet (the ncu capture showed 364 instructions per rotation, issue-bound): round-robin indices without
    // integer division, squared-threshold test without square roots, cached squared norms (one warp sum per rotation
    // instead of three, refreshed every sweep, recomputed when the update cancels).
    const T tol = t_sqrt(Eps<T>::v()), tol2 = tol * tol;
    T t1max = T(0);
    for (int idx = lane; idx < ka * kb; idx += 32) t1max = t_abs(N[idx]) > t1max ? t_abs(N[idx]) : t1max;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const T other = __shfl_xor_sync(0xffffffffu, t1max, o); t1max = other > t1max ? other : t1max; }
    const T noise2 = (Eps<T>::v() * t1max) * (Eps<T>::v() * t1max) * (T) kb;
    const int nb2 = (ka + 1) & ~1, mod = nb2 - 1;
    __shared__ T nrm[256];                     // cached squared column norms of N
    bool converged = ka < 2 || ka > 256;       // (wider product terms keep J = I: no preconditioning, same product)
    for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
        __syncthreads();
        if (tid == 0) s_rot = 0;
        for (int c = w; c < ka; c += nw) {  // refresh the cached squared norms
            const T *nc = N + (size_t) c * kb;
            T ss = T(0);
            for (int i = lane; i < kb; i += 32) ss = fma(nc[i], nc[i], ss);
            ss = warp_sum(ss);
            if (lane == 0) nrm[c] = ss;
        }
        __syncthreads();
        for (int round = 0; round < nb2 - 1; ++round) {
            for (int slot = w; slot < nb2 / 2; slot += nw) {
                int x, y;
                if (slot == 0) { x = mod; y = round; }
                else {
                    x = round + slot; if (x >= mod) x -= mod;
                    y = round - slot; if (y < 0) y += mod;
                }
                if (x > y) { const int tt = x; x = y; y = tt; }
                if (y >= ka) continue;
                T *nx = N + (size_t) x * kb, *ny = N + (size_t) y * kb;
                T gamma = T(0);
                for (int i = lane; i < kb; i += 32) gamma = fma(nx[i], ny[i], gamma);
                gamma = warp_sum(gamma);
                const T alpha = nrm[x], beta = nrm[y];
                if (!(gamma * gamma > tol2 * alpha * beta)) continue;
                if (!(alpha > noise2) || !(beta > noise2)) continue;
                if (lane == 0) s_rot = 1;
                const T t = rx_tangent(beta - alpha, gamma + gamma);
                const T c = t_rsqrt(fma(t, t, T(1))), sn = c * t;
                T na = T(0), nb = T(0);
                for (int i = lane; i < kb; i += 32) {
                    const T u = nx[i], v = ny[i];
                    const T u2 = fma(-sn, v, c * u), v2 = fma(sn, u, c * v);
                    nx[i] = u2;
                    ny[i] = v2;
                    na = fma(u2, u2, na);
                    nb = fma(v2, v2, nb);
                }
                T *jx = J + (size_t) x * ka, *jy = J + (size_t) y * ka;
                for (int i = lane; i < ka; i += 32) {
                    const T u = jx[i], v = jy[i];
                    jx[i] = fma(-sn, v, c * u);
                    jy[i] = fma(sn, u, c * v);
                }
                const T tg = t * gamma;
                T a2 = alpha - tg, b2 = beta + tg;
                if (a2 < T(0.01) * alpha || b2 < T(0.01) * beta) { a2 = warp_sum(na); b2 = warp_sum(nb); }
                __syncwarp();  // all lanes have read nrm[x], nrm[y]
                if (lane == 0) { nrm[x] = a2; nrm[y] = b2; }
            }
            __syncthreads();
        }
        converged = (s_rot == 0);
    }
    __syncthreads();
    for (int idx = tid; idx < ka * ka; idx += nthr) p.J[idx] = J[idx];
    for (int idx = tid; idx < kb * ka; idx += nthr) {  // T1J(j, i') = sum_i T1(i, j) * J(i, i')
        const int j = idx % kb, ip = idx / kb;
        const T *t1 = p.T1 + (size_t) j * ka, *jc = J + (size_t) ip * ka;
        T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
        int i = 0;
        for (; i + 3 < ka; i += 4) {
            a0 = fma(t1[i], jc[i], a0);
            a1 = fma(t1[i + 1], jc[i + 1], a1);
            a2 = fma(t1[i + 2], jc[i + 2], a2);
            a3 = fma(t1[i + 3], jc[i + 3], a3);
        }
        for (; i < ka; ++i) a0 = fma(t1[i], jc[i], a0);
        p.T1J[idx] = (a0 + a1) + (a2 + a3);
    }
}

// Column order of the stacks: pos[c] = rank of V-stack column c by decreasing norm (stable).  One CTA per tile.
template<typename T>
__global__ void __launch_bounds__(256) k_stack_order(const RecompProb<T> *__restrict__ probs, int smem_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw_so[];
    T *nrm = reinterpret_cast<T *>(smem_raw_so);
    const RecompProb<T> p = probs[blockIdx.x];
    if (!p.active) return;
    const int r = p.r, n = p.n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (r > smem_elems) {  // cannot rank in shared memory: keep the natural order
        for (int c = threadIdx.x; c < r; c += blockDim.x) p.pos[c] = c;
        return;
    }
    for (int c = w; c < r; c += nw) {
        if (p.vinc && c < p.kc) {  // incremental V side: the old columns are beta * sigma_c * w_c, never assembled
            if (lane == 0) { const T v = p.beta_c * p.sig0[c]; nrm[c] = v * v; }
            continue;
        }
        const T *col = p.SV0 + (size_t) c * n;
        T ss = T(0);
        for (int i = lane; i < n; i += 32) ss = fma(col[i], col[i], ss);
        ss = warp_sum(ss);
        if (lane == 0) nrm[c] = ss;
    }
    __syncthreads();
    for (int c = w; c < r; c += nw) {
        const T sc = nrm[c];
        int pos = 0;
        for (int o = lane; o < r; o += 32) {
            const T so = nrm[o];
            pos += (so > sc || (so == sc && o < c)) ? 1 : 0;
        }
        pos = warp_sum(pos);
        if (lane == 0) p.pos[c] = pos;
    }
}

// UW[:, pos[c]] = SU0[:, c], VW[:, pos[c]] = SV0[:, c]  -- the same permutation on both stacks leaves SU * SV^T
// unchanged.  grid = (column chunks, 2 * n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_permute_stacks(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y >> 1];
    if (!p.active) return;
    const int side = blockIdx.y & 1;
    if (p.inc && side == 0) return;  // incremental U side: CU is used in place, P stays where the contraction wrote it
    if (p.vinc && side == 1) return; // incremental V side: likewise for CV and Y
    const int rows = side ? p.n : p.m;
    const T *src = side ? p.SV0 : p.SU0;
    T *dst = side ? p.VW : p.UW;
    for (int c = blockIdx.x; c < p.r; c += gridDim.x) {
        const T *s = src + (size_t) c * rows;
        T *d = dst + (size_t) p.pos[c] * rows;
        for (int i = threadIdx.x; i < rows; i += blockDim.x) d[i] = s[i];
    }
}

// Descriptors of the blocked QR of both stacks, for every NBQ-column block at once.  One thread per (block, panel);
// arrays are indexed [blk * npan + pan], pan = 2*tile + side.
template<typename T>
struct QrBlockArrays {
    QrProb<T> *qr;
    LarftProb<T> *lf;
    GemmProb<T> *gw, *gw2, *gup;
    int nblk, npan;
    StripJob *sj;  // fp64 strip-resident path (left-looking updates): one job per (block, panel); nullptr otherwise
};

template<typename T>
__global__ void k_setup_qr_blocks(const PanelDesc<T> *__restrict__ pds, QrBlockArrays<T> o) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= o.nblk * o.npan) return;
    const int blk = idx / o.npan, pan = idx % o.npan;
    const PanelDesc<T> pd = pds[pan];
    QrProb<T> q{nullptr, nullptr, 0, 0, 1};
    LarftProb<T> lf{nullptr, nullptr, nullptr, nullptr, 1, 1, 0, 0};
    GemmProb<T> gz = mk_gemm<T>(nullptr, 1, 0, nullptr, 1, 0, nullptr, 1, 0, 0, 0, T(0), T(0)), gw = gz, gw2 = gz, gup = gz;
    if (pd.active) {
        const int m = pd.m, r = pd.n, kmax = m < r ? m : r, j0 = blk * NBQ;
        if (j0 < kmax) {
            const int jb = (kmax - j0) < NBQ ? (kmax - j0) : NBQ;
            T *blkA = pd.A + (size_t) j0 + (size_t) j0 * m;
            T *Vc = pd.VC + (size_t) j0 + (size_t) j0 * m;
            T *Tm = pd.TB + (size_t) blk * NBQ * NBQ;
            T *W = pd.WB, *W2 = W + (size_t) NBQ * pd.wcols;
            q = QrProb<T>{blkA, pd.tau + j0, m - j0, jb, m};
            lf = LarftProb<T>{blkA, pd.tau + j0, Vc, Tm, m, m, m - j0, jb};
            const int nt = r - (j0 + jb);  // trailing columns
            if (nt > 0) {
                T *A2 = pd.A + (size_t) j0 + (size_t) (j0 + jb) * m;
                gw = mk_gemm<T>(Vc, m, 1, A2, m, 0, W, NBQ, jb, nt, m - j0, T(1), T(0));       // W  = Vc^T A2
                gw2 = mk_gemm<T>(Tm, NBQ, 1, W, NBQ, 0, W2, NBQ, jb, nt, jb, T(1), T(0));       // W2 = T^T W
                gup = mk_gemm<T>(Vc, m, 0, W2, NBQ, 0, A2, m, m - j0, nt, jb, T(-1), T(1));     // A2 -= Vc W2
            }
        }
    }
    o.qr[idx] = q; o.lf[idx] = lf; o.gw[idx] = gw; o.gw2[idx] = gw2; o.gup[idx] = gup;
    if constexpr (std::is_same<T, double>::value) {
        if (o.sj) {  // block `blk` of the panel, brought up to date with reflector blocks 0..blk-1 (Q^T)
            StripJob j{nullptr, nullptr, nullptr, 1, 1, 0, 0, 0, 0, 0, 1, 1};
            if (pd.active) {
                const int m = pd.m, r = pd.n, kmax = m < r ? m : r, j0 = blk * NBQ;
                if (j0 < r && blk > 0) {
                    const int nc = (r - j0) < NBQ ? (r - j0) : NBQ;
                    const int np = blk < (kmax + NBQ - 1) / NBQ ? blk : (kmax + NBQ - 1) / NBQ;
                    j = StripJob{pd.A + (size_t) j0 * m, pd.VC, pd.TB, m, m, m, nc, kmax, 0, np, 1, 1};
                }
            }
            o.sj[idx] = j;
        }
    }
}

// Descriptors of the blocked rebuild C := Q [X;0] (C = CU for side 0, VN for side 1), blocks applied last-to-first:
// W = Vc_b^T C[j0:,:] ; W2 = T_b W ; C[j0:,:] -= Vc_b W2.  The column count is the rank chosen on the device.
template<typename T>
struct ApplyBlockArrays {
    GemmProb<T> *gw, *gw2, *gup;
    int nblk, npan;
};

// Strip jobs of the rebuild: strip s (32 columns) of CU / VN, all reflector blocks last-to-first (Q).
template<typename T>
__global__ void k_setup_apply_strips(const RecompProb<T> *__restrict__ rcs, StripJob *__restrict__ out, int nstrips, int npan,
                                     GemmProb<T> *__restrict__ gru) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < npan / 2) {
        // incremental U side: CU' = [CU | Q2] * Us[:, :rk] as two GEMMs into TU (the rank was just chosen on the device)
        const RecompProb<T> rc = rcs[idx];
        GemmProb<T> ga = mk_gemm<T>(nullptr, 1, 0, nullptr, 1, 0, nullptr, 1, 0, 0, 0, T(0), T(0)), gb = ga;
        if (rc.active && rc.inc) {
            int rk = *rc.rk_new;
            if (rk > rc.wcols) rk = 0;
            if (std::is_same<T, double>::value) {  // one pass: op(A) = [CU | Q2] (two-segment A of k_gemm_dmma)
                ga = mk_gemm<T>(rc.CU, rc.m, 0, rc.Us, rc.ldus, 0, rc.TU, rc.m, rc.m, rk, rc.kc + rc.kp, T(1), T(0));
                ga.A2 = rc.Q2; ga.k1 = rc.kc; ga.lda2 = rc.m;
            } else {
                ga = mk_gemm<T>(rc.CU, rc.m, 0, rc.Us, rc.ldus, 0, rc.TU, rc.m, rc.m, rk, rc.kc, T(1), T(0));
                gb = mk_gemm<T>(rc.Q2, rc.m, 0, rc.Us + rc.kc, rc.ldus, 0, rc.TU, rc.m, rc.m, rk, rc.kp, T(1), T(1));
            }
        }
        gru[idx] = ga;
        gru[npan / 2 + idx] = gb;
        // incremental V side: VN = [CV^T | Q2v] * Vs[:, :rk] = [CV ; Q2v^T]^T * Vs (two-segment transposed A) -> VN (n x rk)
        GemmProb<T> gn = mk_gemm<T>(nullptr, 1, 0, nullptr, 1, 0, nullptr, 1, 0, 0, 0, T(0), T(0));
        if (rc.active && rc.vinc) {
            int rk = *rc.rk_new;
            if (rk > rc.wcols) rk = 0;
            gn = mk_gemm<T>(rc.CV, rc.kc, 1, rc.Vs, rc.b, 0, rc.VN, rc.n, rc.n, rk, rc.kc + rc.kp, T(1), T(0));
            gn.A2 = rc.Q2vT; gn.k1 = rc.kc; gn.lda2 = rc.kp;
        }
        gru[2 * (npan / 2) + idx] = gn;
    }
    if (idx >= nstrips * npan) return;
    // the strips of one panel are neighbours in the grid: they run at the same time and share the panel's reflector
    // blocks through L2 (strip-major order re-read them from HBM: 10 GB per launch in the ncu capture)
    const int st = idx % nstrips, pan = idx / nstrips, side = pan & 1;
    const RecompProb<T> rc = rcs[pan >> 1];
    StripJob j{nullptr, nullptr, nullptr, 1, 1, 0, 0, 0, 0, 0, -1, 0};
    if constexpr (std::is_same<T, double>::value) {
        if (rc.active && !(rc.inc && side == 0) && !(rc.vinc && side == 1)) {
            const int m = side ? rc.n : rc.m, r = rc.r, kmax = m < r ? m : r;
            int rk = *rc.rk_new;
            if (rk > rc.wcols) rk = 0;
            const int c0 = st * NBQ;
            if (c0 < rk && kmax > 0) {
                const int nb = (kmax + NBQ - 1) / NBQ;
                j = StripJob{(side ? rc.VN : rc.CU) + (size_t) c0 * m, rc.VC[side], rc.TB[side], m, m, m,
                             (rk - c0) < NBQ ? (rk - c0) : NBQ, kmax, nb - 1, nb, -1, 0};
            }
        }
    }
    out[idx] = j;
}

template<typename T>
__global__ void k_setup_apply_blocks(const RecompProb<T> *__restrict__ rcs, ApplyBlockArrays<T> o) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= o.nblk * o.npan) return;
    const int blk = idx / o.npan, pan = idx % o.npan, side = pan & 1;
    const RecompProb<T> rc = rcs[pan >> 1];
    GemmProb<T> gz = mk_gemm<T>(nullptr, 1, 0, nullptr, 1, 0, nullptr, 1, 0, 0, 0, T(0), T(0)), gw = gz, gw2 = gz, gup = gz;
    if (rc.active) {
        const int m = side ? rc.n : rc.m, r = rc.r, kmax = m < r ? m : r, j0 = blk * NBQ;
        int rk = *rc.rk_new;
        if (rk > rc.wcols) rk = 0;  // cannot happen (wcols >= max_rank bound); guard against scratch overrun
        if (j0 < kmax && rk > 0) {
            const int jb = (kmax - j0) < NBQ ? (kmax - j0) : NBQ;
            T *Cm = (side ? rc.VN : rc.CU) + j0;
            T *Vc = rc.VC[side] + (size_t) j0 + (size_t) j0 * m;
            T *Tm = rc.TB[side] + (size_t) blk * NBQ * NBQ;
            T *W = rc.WB[side], *W2 = W + (size_t) NBQ * rc.wcols;
            gw = mk_gemm<T>(Vc, m, 1, Cm, m, 0, W, NBQ, jb, rk, m - j0, T(1), T(0));
            gw2 = mk_gemm<T>(Tm, NBQ, 0, W, NBQ, 0, W2, NBQ, jb, rk, jb, T(1), T(0));
            gup = mk_gemm<T>(Vc, m, 0, W2, NBQ, 0, Cm, m, m - j0, rk, jb, T(-1), T(1));
        }
    }
    o.gw[idx] = gw; o.gw2[idx] = gw2; o.gup[idx] = gup;
}

// core = RU * RV^T with RU = triu(UW[:p, :r]), RV = triu(VW[:q, :r]) (Compressed.cpp:363-370, 442-461), written as
// M (a x b, ld a): core itself when p >= q, its transpose otherwise.  grid = (chunks, n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_core_build(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active) return;
    const int total = p.p * p.q;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx % p.p, j = idx / p.p;
        T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
        int l = (i > j ? i : j);
        const T *pu = p.UW + (size_t) i, *pv = p.VW + (size_t) j;
        for (; l + 3 < p.r; l += 4) {  // independent accumulators: 8 loads in flight
            acc0 = fma(pu[(size_t) l * p.m], pv[(size_t) l * p.n], acc0);
            acc1 = fma(pu[(size_t) (l + 1) * p.m], pv[(size_t) (l + 1) * p.n], acc1);
            acc2 = fma(pu[(size_t) (l + 2) * p.m], pv[(size_t) (l + 2) * p.n], acc2);
            acc3 = fma(pu[(size_t) (l + 3) * p.m], pv[(size_t) (l + 3) * p.n], acc3);
        }
        for (; l < p.r; ++l) acc0 = fma(pu[(size_t) l * p.m], pv[(size_t) l * p.n], acc0);
        const T acc = (acc0 + acc1) + (acc2 + acc3);
        // M (a x b, ld a) and its transpose MT (b x a, ld b), which is the one that gets QR-factored
        if (p.transposed) {
            p.M[(size_t) j + (size_t) i * p.a] = acc;
            p.MT[(size_t) i + (size_t) j * p.b] = acc;
        } else {
            p.M[(size_t) i + (size_t) j * p.a] = acc;
            p.MT[(size_t) j + (size_t) i * p.b] = acc;
        }
    }
}

// RU = triu(UW[:p, :r]) -> MT (p x r, ld p), RV = triu(VW[:q, :r]) -> Lb (q x r, ld q): the triangles of the two QR'd
// stacks with the reflectors below the diagonal masked out, so that the core is ONE batched DMMA GEMM (k_core_build's
// strided dot products ran at 7.5 ms per 256 tiles at r = 357, ncu launch list r01).  grid = (chunks, n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_extract_r(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active || p.both) return;  // (both sides incremental: the core is assembled by k_both_core, no triangles needed)
    const int nu = p.p * p.r, total = nu + p.q * p.r;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (idx < nu) {
            const int i = idx % p.p, l = idx / p.p;
            if (p.inc) {
                // RU = [[I, G], [0, R2]] (r x r) in the column order of the sorted V stack: column l goes to pos[l]
                T v;
                if (l < p.kc) v = (i == l) ? T(1) : T(0);
                else {
                    const int c = l - p.kc;
                    if (i < p.kc) v = p.Gu[(size_t) i + (size_t) c * p.kc] + p.Gu2[(size_t) i + (size_t) c * p.kc];
                    else v = (i - p.kc <= c) ? p.Pn[(size_t) (i - p.kc) + (size_t) c * p.m] : T(0);
                }
                p.MT[(size_t) i + (size_t) p.pos[l] * p.lp] = v;
            } else {
                p.MT[(size_t) i + (size_t) l * p.lp] = i <= l ? p.UW[(size_t) i + (size_t) l * p.m] : T(0);
            }
        } else {
            const int e = idx - nu, j = e % p.q, l = e / p.q;
            p.Lb[(size_t) j + (size_t) l * p.lq] = j <= l ? p.VW[(size_t) j + (size_t) l * p.ldvw] : T(0);
        }
    }
}

// LQ preconditioning (Drmac-Veselic style, without pivoting): M^T = Q R  =>  M = L Q^T with L = R^T (a x b, lower
// trapezoidal).  L has the same singular values and LEFT singular vectors as M, and one-sided Jacobi on a triangular
// factor needs far fewer sweeps (numpy emulation: 20 -> 11 on a 126 x 126 recompression core, 42 -> 11 on a 200 x 200
// matrix with singular values spread over 14 decades).
// grid = (chunks, n_problems)
template<typename T>
__global__ void __launch_bounds__(256) k_extract_l(const LqProb<T> *__restrict__ probs) {
    const LqProb<T> p = probs[blockIdx.y];
    const int total = p.a * p.b;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx % p.a, j = idx / p.a;  // L(i, j) = R(j, i) for j <= i
        p.Lb[idx] = (j <= i) ? p.MT[(size_t) j + (size_t) i * p.b] : T(0);
    }
}

// Rank truncation on the device (Compressed.cpp:509-527 + omp/kernels.cpp:82-104) and scatter of the truncated
// core factors: CU[:, :rk] = [Ufac ; 0] (m x rk), VN[:, :rk] = [Vfac * diag(sigma) ; 0] (n x rk).  One CTA per tile.
template<typename T>
__global__ void __launch_bounds__(256) k_truncate(const RecompProb<T> *__restrict__ probs, T accuracy, int truncated,
                                                  int fixed_rank) {
    const RecompProb<T> p = probs[blockIdx.x];
    if (!p.active) return;
    __shared__ int s_rk;
    if (threadIdx.x == 0) {
        int rk;
        const int fr = p.fixed_rank > 0 ? p.fixed_rank : fixed_rank;  // per-tile replay rank overrides the batch value
        if (fr > 0) rk = fr < p.r ? fr : p.r;
        else rk = new_rank_rule(p.sigma, p.b, accuracy, truncated);
        if (rk > p.b) rk = p.b;  // only b = min(m, n, r) singular triplets exist
        if (rk < 1) rk = 1;
        if (rk > p.max_rank) {
            rk = p.max_rank;
            if (p.info) atomicOr(p.info, 2);
        }
        s_rk = rk;
        *p.rk_new = rk;
    }
    __syncthreads();
    const int rk = s_rk;
    // Us = normalised left factor of M, Vs = (right factor of M) * diag(sigma)  [= M^T Us].
    // not transposed (M = core)  : Ufac = Us (p x b),            Vfac*S = Vs (q x b)
    // transposed (M = core^T)    : core = (Vs/S) S Us^T -> Ufac = Vs / sigma (p x b, b == p), Vfac*S = Us * sigma (q x b)
    if (!p.transposed) {
        // (incremental U side: Us is consumed by the rebuild GEMMs [CU | Q2] * Us, CU must stay intact until then)
        for (int idx = threadIdx.x; idx < (p.inc ? 0 : p.m * rk); idx += blockDim.x) {
            const int i = idx % p.m, c = idx / p.m;
            p.CU[(size_t) i + (size_t) c * p.m] = (i < p.p) ? p.Us[(size_t) i + (size_t) c * p.ldus] : T(0);
        }
        // (incremental V side: Vs = V S' in [W | Q2v] coordinates is consumed by the rebuild GEMM; its first kc rows are
        // divided by the old singular values here because the GEMM multiplies by CV^T = W diag(sigma), not by W)
        if (p.vinc) {
            for (int idx = threadIdx.x; idx < p.kc * rk; idx += blockDim.x) {
                const int i = idx % p.kc, c = idx / p.kc;
                const size_t o = (size_t) i + (size_t) c * p.b;
                // (rank-kp form: Vs holds Xv (Xu^T Us) only; the diagonal part beta S Us, divided by S, is beta Us)
                p.Vs[o] = p.both ? p.Vs[o] / p.sig0[i] + p.beta_c * p.Us[(size_t) i + (size_t) c * p.ldus] : p.Vs[o] / p.sig0[i];
            }
        }
        for (int idx = threadIdx.x; idx < (p.vinc ? 0 : p.n * rk); idx += blockDim.x) {
            const int i = idx % p.n, c = idx / p.n;
            p.VN[(size_t) i + (size_t) c * p.n] = (i < p.q) ? p.Vs[(size_t) i + (size_t) c * p.b] : T(0);
        }
    } else {
        for (int idx = threadIdx.x; idx < p.m * rk; idx += blockDim.x) {
            const int i = idx % p.m, c = idx / p.m;
            const T sg = p.sigma[c];
            p.CU[(size_t) i + (size_t) c * p.m] = (i < p.p && sg > T(0)) ? p.Vs[(size_t) i + (size_t) c * p.b] / sg : T(0);
        }
        for (int idx = threadIdx.x; idx < p.n * rk; idx += blockDim.x) {
            const int i = idx % p.n, c = idx / p.n;
            p.VN[(size_t) i + (size_t) c * p.n] = (i < p.q) ? p.sigma[c] * p.Us[(size_t) i + (size_t) c * p.ldus] : T(0);
        }
    }
}

// CV (rk x n, ld rk) = VN^T (CalculateUVptr, omp/kernels.cpp:106-114) and commit the new rank (Compressed.cpp:656-662).
// grid = (tile_chunks, n_tiles), block = (32, 8)
template<typename T>
__global__ void __launch_bounds__(256) k_finalize(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active) return;
    __shared__ T tile[32][33];
    const int rk = *p.rk_new;
    const int tr = (rk + 31) / 32, tc = (p.n + 31) / 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
        const int r0 = (t % tr) * 32, c0 = (t / tr) * 32;  // r: rank index, c: column of the tile (row of VN)
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + tx, r = r0 + j;
            tile[j][tx] = (r < rk && c < p.n) ? p.VN[(size_t) c + (size_t) r * p.n] : T(0);
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int r = r0 + tx, c = c0 + j;
            if (r < rk && c < p.n) p.CV[(size_t) r + (size_t) c * rk] = tile[tx][j];
        }
        __syncthreads();
    }
    if (p.inc) {  // incremental U side: the rebuilt factor sits in TU
        const int nthr = 256, tid = ty * 32 + tx;
        for (size_t idx = (size_t) blockIdx.x * nthr + tid; idx < (size_t) p.m * rk; idx += (size_t) gridDim.x * nthr)
            p.CU[idx] = p.TU[idx];
    }
    if (blockIdx.x == 0 && tx == 0 && ty == 0) {
        *p.rank_ptr = rk;
        // U = Q_U * (normalised left vectors of the core): orthonormal columns unless a kept singular value vanished
        if (p.state) {
            const bool ortho = !p.transposed && p.sigma[rk - 1] > T(1e-30) * p.sigma[0] && p.sigma[0] > T(0);
            const int count = (p.inc || p.vinc) ? ((*p.state >> 8) + 1) : 0;  // consecutive incremental updates so far
            *p.state = ortho ? (HCB_STATE_ORTHO_U | HCB_STATE_ORTHO_V | (count << 8)) : 0;
        }
    }
}

// [I; 0] in the explicit-Q2 buffers of the incremental tiles (U side and V side).  grid = (chunks, 2 * n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_inc_eye(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y >> 1];
    const int side = blockIdx.y & 1;
    if (!p.active || !(side ? p.vinc : p.inc) || p.q2_done[side]) return;   // (CholeskyQR2 already left the explicit Q2 there)
    const int rows = side ? p.n : p.m;
    T *Q = side ? p.Q2v : p.Q2;
    const size_t total = (size_t) rows * p.kp;
    for (size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x)
        Q[idx] = (idx % rows == idx / rows) ? T(1) : T(0);
}

// Incremental V side: sig0[i] = || CV(i, :) ||, the old singular values (CV = diag(sigma) W^T).  One CTA per tile;
// thread (i, g) of a 32 x 8 block walks row i + 32a over the columns g, g + 8, ... (rows are contiguous in memory).
template<typename T>
__global__ void __launch_bounds__(256) k_vinc_sig0(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.x];
    if (!p.active || !p.vinc) return;
    __shared__ T part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i0 = 0; i0 < p.kc; i0 += 32) {
        const int i = i0 + tx;
        T ss = T(0);
        if (i < p.kc)
            for (int j = ty; j < p.n; j += 8) { const T v = p.CV[(size_t) i + (size_t) j * p.kc]; ss = fma(v, v, ss); }
        part[ty][tx] = ss;
        __syncthreads();
        if (ty == 0 && i < p.kc) {
            T tot = T(0);
#pragma unroll
            for (int g = 0; g < 8; ++g) tot += part[g][tx];
            p.sig0[i] = t_sqrt(tot);
        }
        __syncthreads();
    }
}

// CholeskyQR2 of the kp new columns X (rows x kp, after the block Gram-Schmidt against the old basis): X = Q2 R2 with
//   pass 0:  G = X^T X,  Gs = D G D with D = diag(G)^-1/2 (column scaling: the new columns span 8 decades),
//            Gs = Rc^T Rc (Cholesky),  Q1 = X (D Rc^-1),  R1 = Rc D^-1
//   pass 1:  G2 = Q1^T Q1,  G2 = Rc2^T Rc2,  Q2 = Q1 Rc2^-1,  R2 = Rc2 R1
// ONE kernel per pass, one CTA (8 warps) per tile side: every warp forms the Gram matrix of its 1/8 of the rows with DMMA
// (m8n8k4: both fragments are the same X[r0 + c][8 t + g] loads; the 21 upper 8 x 8 tile pairs stay in registers), the
// partial sums meet in shared memory, the kp x kp Cholesky and the triangular inverse run there, and the same warps
// multiply their rows by S = D Rc^-1 with DMMA again.  No per-column cluster barrier, no explicit-Q strip pass: 2 x 0.4 ms
// per k-step of 256 tiles (measured, ncu launch list r02) instead of ~2.9 ms for the register-panel Householder QR + the
// strips that form Q2.
// CholeskyQR squares the condition number, so the fast path is taken only where it is provably safe: every pivot of the
// SCALED Gram matrix must be >= 1e-6 (the column keeps at least 1e-3 of its norm after orthogonalisation against the
// earlier ones: condition <~ 5e4, first-pass loss of orthogonality <~ 1e-7, removed by the second pass) and every pivot
// of the second pass must lie in [1/4, 4]; otherwise the side's flag is raised and the Householder panel path -- whose
// descriptors stay active, X untouched -- factors it as before.  Independent directions (the BASELINE generator's tiles)
// take the fast path; nearly dependent new columns (saturated tiles) fall back.  On success pass 1 writes R2 into the top
// triangle of X (what the core assembly reads), leaves Q2 in its buffer and switches the side's Householder panel and
// explicit-Q strip jobs off.  kp <= 48.  grid = 2 * n_tiles (problem = side * n_tiles + tile), 256 threads.
template<typename T>
__global__ void __launch_bounds__(256, 2) k_cholqr_pass(RecompProb<T> *__restrict__ probs, PanelDesc<T> *__restrict__ pd_inc,
                                                        StripJob *__restrict__ inc_sj, int nst_inc, int n_tiles, int pass) {
    static_assert(std::is_same<T, double>::value, "FP64 tensor path");
    const int prob = blockIdx.x, side = prob / n_tiles, t = prob % n_tiles;
    const RecompProb<T> p = probs[t];
    if (!p.active || !p.cq_fail || p.cq_fail[side] || !p.CQw[side]) return;
    const int kp = p.kp, rows = side ? p.n : p.m, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    T *Xg = side ? p.Yn : p.Pn;                       // the panel (its top triangle receives R2 at the end)
    const T *In = pass == 0 ? Xg : p.CQt[side];       // pass 0 reads X, pass 1 reads Q1
    T *Out = pass == 0 ? p.CQt[side] : (side ? p.Q2v : p.Q2);
    T *R1 = p.CQw[side];                              // kp x kp (ld kp), kept between the passes
    __shared__ T R[CQ_KP][CQ_P];                      // Gram matrix -> Cholesky factor (upper)
    __shared__ T S[CQ_KP][CQ_P];                      // D Rc^-1 (upper)
    __shared__ T sd[CQ_KP];
    __shared__ int s_bad;
    for (int idx = tid; idx < CQ_KP * CQ_P; idx += 256) (&R[0][0])[idx] = T(0);
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // ---- Gram matrix: warp w takes the row chunks w, w + 8, ... of 32 rows (8 DMMA k-steps each)
    {
        T acc[CQ_NT * (CQ_NT + 1) / 2][2];
#pragma unroll
        for (int q = 0; q < CQ_NT * (CQ_NT + 1) / 2; ++q) acc[q][0] = acc[q][1] = T(0);
        for (int r0 = w * 32; r0 < rows; r0 += 256) {
#pragma unroll 2
            for (int ks = 0; ks < 32; ks += 4) {
                const int r = r0 + ks + c;
                T v[CQ_NT];
#pragma unroll
                for (int tt = 0; tt < CQ_NT; ++tt) {
                    const int col = 8 * tt + g;
                    v[tt] = (r < rows && col < kp) ? In[(size_t) r + (size_t) col * rows] : T(0);
                }
                int q = 0;
#pragma unroll
                for (int ti = 0; ti < CQ_NT; ++ti)
#pragma unroll
                    for (int tj = ti; tj < CQ_NT; ++tj, ++q) dmma_m8n8k4(acc[q][0], acc[q][1], v[ti], v[tj]);
            }
        }
        int q = 0;
#pragma unroll
        for (int ti = 0; ti < CQ_NT; ++ti)
#pragma unroll
            for (int tj = ti; tj < CQ_NT; ++tj, ++q) {
                atomicAdd(&R[8 * ti + g][8 * tj + 2 * c], acc[q][0]);
                atomicAdd(&R[8 * ti + g][8 * tj + 2 * c + 1], acc[q][1]);
            }
    }
    __syncthreads();
    // ---- scaling, Cholesky (upper, in place), checks
    // Negligible columns are DEFLATED instead of failing the test: a new column whose norm is below 1e-13 of the largest new
    // column (the product term's singular values reach 1e-16 of its largest; such columns are rounding noise of the
    // contraction, often mutually parallel, and 5 decades below any accuracy the truncation can ask for) is replaced by a
    // zero column -- Q[:, j] = 0, R[j, :] = R[:, j] = 0: the core then has an exactly zero row there, which no retained
    // singular vector touches.  Pass 1 recognises them by their exactly zero Gram diagonal.
    __shared__ T s_gmax;
    if (tid < 32) {
        T mx = T(0);
        for (int j = tid; j < kp; j += 32) { const T v = R[j][j]; mx = (v > mx) ? v : mx; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const T other = __shfl_xor_sync(0xffffffffu, mx, o); mx = other > mx ? other : mx; }
        if (tid == 0) s_gmax = mx;
    }
    __syncthreads();
    for (int j = tid; j < kp; j += 256) {
        const T gjj = R[j][j];
        if (!(gjj >= T(0)) || !(gjj < T(1e300))) s_bad = 1;                  // NaN / Inf
        const bool defl = pass == 0 ? !(gjj > T(1e-26) * s_gmax) : (gjj == T(0));
        sd[j] = defl ? T(0) : (pass == 0 ? T(1) / t_sqrt(gjj) : T(1));
        if (defl && pass == 0 && p.stats) atomicAdd(p.stats + 5, 1ull);
    }
    __syncthreads();
    for (int idx = tid; idx < kp * kp; idx += 256) {
        const int i = idx % kp, j = idx / kp;
        if (i <= j) {
            const bool di = sd[i] == T(0), dj = sd[j] == T(0);
            R[i][j] = (di || dj) ? ((i == j) ? T(1) : T(0)) : R[i][j] * sd[i] * sd[j];   // deflated: identity row / column
        }
    }
    __syncthreads();
    const T lo = pass == 0 ? T(1e-6) : T(0.25), hi = T(4);
    if (!s_bad) {
        // right-looking Cholesky with ONE barrier per step: the trailing update uses the unscaled pivot row (read-only during
        // the step), R[i][l] -= R[j][i] R[j][l] / piv; the rows are divided by sqrt(pivot) once at the end
        for (int j = 0; j < kp; ++j) {
            const T piv = R[j][j];
            if (!(piv >= lo) || !(piv <= hi)) { if (tid == 0) s_bad = 1; break; }   // (uniform: everybody reads the same value)
            const T inv = T(1) / piv;
            const int nt = kp - j - 1;
            for (int idx = tid; idx < nt * nt; idx += 256) {
                const int i = j + 1 + idx % nt, l = j + 1 + idx / nt;
                if (i <= l) R[i][l] = fma(-R[j][i] * inv, R[j][l], R[i][l]);
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (!s_bad) {
        for (int idx = tid; idx < kp * kp; idx += 256) {
            const int i = idx % kp, j = idx / kp;
            if (i < j) R[i][j] *= T(1) / t_sqrt(R[i][i]);       // (diagonal entries are still the pivots here)
        }
        __syncthreads();
        for (int j = tid; j < kp; j += 256) R[j][j] = t_sqrt(R[j][j]);
    }
    __syncthreads();
    if (s_bad) {   // not safely well conditioned: Householder takes over (X is intact)
        if (tid == 0) {
            p.cq_fail[side] = 1;
            if (p.stats) atomicAdd(p.stats + 1 + pass, 1ull);
        }
        return;
    }
    // ---- S = D Rc^-1 (upper): thread j solves Rc y = e_j by back substitution, straight into shared memory
    for (int idx = tid; idx < CQ_KP * CQ_P; idx += 256) (&S[0][0])[idx] = T(0);
    __syncthreads();
    for (int j = tid; j < kp; j += 256) {
        for (int i = j; i >= 0; --i) {
            T a = (i == j) ? T(1) : T(0);
            for (int l = i + 1; l <= j; ++l) a = fma(-R[i][l], S[l][j], a);
            S[i][j] = a / R[i][i];
        }
        for (int i = 0; i <= j; ++i) S[i][j] *= sd[i];
    }
    __syncthreads();
    // ---- Out = In * S for this warp's rows: 8-row tiles, N = 48 (6 tiles), K = kp padded to a multiple of 4
    for (int r0 = w * 8; r0 < rows; r0 += 64) {
        T acc[CQ_NT][2];
#pragma unroll
        for (int tt = 0; tt < CQ_NT; ++tt) acc[tt][0] = acc[tt][1] = T(0);
        const int r = r0 + g;
        T a[CQ_KP / 4];                                   // all A fragments of this row tile first: 12 independent loads in flight
#pragma unroll
        for (int q = 0; q < CQ_KP / 4; ++q) {
            const int kk = 4 * q + c;
            a[q] = (r < rows && kk < kp) ? In[(size_t) r + (size_t) kk * rows] : T(0);
        }
#pragma unroll
        for (int q = 0; q < CQ_KP / 4; ++q) {
            const int kk = 4 * q + c;
#pragma unroll
            for (int tt = 0; tt < CQ_NT; ++tt) dmma_m8n8k4(acc[tt][0], acc[tt][1], a[q], S[kk][8 * tt + g]);
        }
        if (r < rows) {
#pragma unroll
            for (int tt = 0; tt < CQ_NT; ++tt) {
                const int col = 8 * tt + 2 * c;
                if (col < kp) Out[(size_t) r + (size_t) col * rows] = acc[tt][0];
                if (col + 1 < kp) Out[(size_t) r + (size_t) (col + 1) * rows] = acc[tt][1];
            }
        }
    }
    // ---- triangular factors
    if (pass == 0) {
        for (int idx = tid; idx < kp * kp; idx += 256) {
            const int i = idx % kp, j = idx / kp;
            R1[idx] = (i <= j && sd[i] != T(0) && sd[j] != T(0)) ? R[i][j] / sd[j] : T(0);   // R1 = Rc D^-1 (zero row / column where deflated)
        }
    } else {
        // (the top kp rows of X are read by nobody in this pass: pass 1 reads Q1)
        for (int idx = tid; idx < kp * kp; idx += 256) {
            const int i = idx % kp, j = idx / kp;
            if (i > j) continue;
            T a = T(0);
            for (int l = i; l <= j; ++l) a = fma(R[i][l], R1[(size_t) l + (size_t) j * kp], a);
            Xg[(size_t) i + (size_t) j * rows] = a;                             // R2 = Rc2 R1
        }
        if (tid == 0) {
            if (p.stats) atomicAdd(p.stats + 0, 1ull);
            probs[t].q2_done[side] = 1;
            pd_inc[(size_t) side * n_tiles + t].active = 0;
            for (int st = 0; st < nst_inc; ++st)
                inc_sj[((size_t) side * n_tiles + t) * nst_inc + st] = StripJob{nullptr, nullptr, nullptr, 1, 1, 0, 0, 0, 0, 0, -1, 0};
        }
    }
}

// Selective re-orthogonalisation (Daniel-Gragg-Kaufman-Stewart / "twice is enough"): after the FIRST block Gram-Schmidt
// pass  P' = P - CU G  (CU orthonormal, so |p_j|^2 = |p'_j|^2 + |g_j|^2 without cancellation), a second pass is only
// needed for columns that lost more than half of their squared norm; when EVERY new column of a tile side kept
//   |p'_j|^2 >= |g_j|^2     (rho^2 >= 1/2: the orthogonality of Q2 against CU is already eps / rho <= 1.5 eps)
// the two second-pass GEMMs of that side are switched off on the device (descriptor m = 0) and their coefficient
// buffer is zeroed.  New directions that lie mostly inside span(CU) -- the saturated tiles of smooth kernels -- fail the
// test and take both passes; independent subspaces (the BASELINE generator's tiles: |g|^2 / |p|^2 ~ kc / nb <= 0.31) pass.
// V side: the coefficients in the orthonormal basis W = CV^T S^-1 are S^-1 Hv, |.|^2 = sum_i Hv_ij^2 / sigma_i^2.
// grid = 2 * n_tiles (side = blockIdx.x & 1), 256 threads; gi / giv are the [4][n_tiles] descriptor arrays.
template<typename T>
__global__ void __launch_bounds__(256) k_inc_gate(const RecompProb<T> *__restrict__ probs, GemmProb<T> *__restrict__ gi,
                                                  GemmProb<T> *__restrict__ giv, int n_tiles, int force_once) {
    const int t = blockIdx.x >> 1, side = blockIdx.x & 1;
    const RecompProb<T> p = probs[t];
    if (!p.active || !(side ? p.vinc : p.inc)) return;
    const int rows = side ? p.n : p.m, kc = p.kc, kp = p.kp;
    const T *X = side ? p.Yn : p.Pn;        // the once-orthogonalised new columns (rows x kp, ld rows)
    const T *G = side ? p.Hv : p.Gu;        // kc x kp, ld kc
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    for (int j = w; j < kp; j += nw) {
        const T *x = X + (size_t) j * rows, *g = G + (size_t) j * kc;
        T a0 = T(0), a1 = T(0), b = T(0);
        int i = lane;
        for (; i + 32 < rows; i += 64) { a0 = fma(x[i], x[i], a0); a1 = fma(x[i + 32], x[i + 32], a1); }
        for (; i < rows; i += 32) a0 = fma(x[i], x[i], a0);
        for (int l = lane; l < kc; l += 32) {
            const T c = side ? g[l] / p.sig0[l] : g[l];
            b = fma(c, c, b);
        }
        const T rest = warp_sum(a0 + a1), coef = warp_sum(b);
        if (lane == 0 && !(rest >= coef)) s_fail = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && p.stats) atomicAdd(p.stats + ((s_fail && !force_once) ? 4 : 3), 1ull);
    if (s_fail && !force_once) return;      // second pass stays on (force_once: test switch, see HCB_GS_FORCE_ONCE)
    T *Z = side ? p.Zv : p.Gu2;             // second-pass coefficients: none
    for (int idx = threadIdx.x; idx < kc * kp; idx += blockDim.x) Z[idx] = T(0);
    if (threadIdx.x == 0) {
        GemmProb<T> *arr = side ? giv : gi;
        arr[(size_t) 2 * n_tiles + t].m = 0;
        arr[(size_t) 3 * n_tiles + t].m = 0;
    }
}

// Incremental V side: Zv = diag(sigma)^-2 * H of the CURRENT pass.  pass 0: H = Hv (first GEMM);  pass 1: the second
// GEMM left H2 = CV * Y1 in Zv -- it is added to Hv (Gv = S^-1 (H + H2)) and scaled in place.  grid = (chunks, n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_vinc_scale(const RecompProb<T> *__restrict__ probs, int pass) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active || !p.vinc) return;
    const int total = p.kc * p.kp;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const T sg = p.sig0[idx % p.kc], s2 = sg * sg;
        T h;
        if (pass == 0) h = p.Hv[idx];
        else { h = p.Zv[idx]; p.Hv[idx] += h; }
        p.Zv[idx] = h / s2;
    }
}

// Incremental V side: RV * Pi (r x r) into the VW buffer (ld r, QR-factored next) and into RVp (kept):
//   RV = [[beta S, Gv], [0, R2v]],  Gv = S^-1 Hv,  R2v = the triangle of the QR-factored Y panel;  column c goes to pos[c].
// grid = (chunks, n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_vinc_build_rv(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active || !p.vinc) return;
    const int r = p.r, total = r * r;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx % r, c = idx / r;
        T v;
        if (c < p.kc) v = (i == c) ? p.beta_c * p.sig0[c] : T(0);
        else {
            const int l = c - p.kc;
            if (i < p.kc) v = p.Hv[(size_t) i + (size_t) l * p.kc] / p.sig0[i];
            else v = (i - p.kc <= l) ? p.Yn[(size_t) (i - p.kc) + (size_t) l * p.n] : T(0);
        }
        const size_t o = (size_t) i + (size_t) p.pos[c] * r;
        p.VW[o] = v;
        p.RVp[o] = v;
        if (p.both && c >= p.kc) {   // Xv = [Gv ; R2v], Xu = [Gu ; R2u]  (r x kp, natural order)
            const int l = c - p.kc;
            p.Xv[(size_t) i + (size_t) l * r] = v;
            T u;
            if (i < p.kc) u = p.Gu[(size_t) i + (size_t) l * p.kc] + p.Gu2[(size_t) i + (size_t) l * p.kc];
            else u = (i - p.kc <= l) ? p.Pn[(size_t) (i - p.kc) + (size_t) l * p.m] : T(0);
            p.Xu[(size_t) i + (size_t) l * r] = u;
        }
    }
}

// Incremental V side: the graded triangular factor R' of M = RV * Pi (r x r) WITHOUT a Householder QR.  R'^T R' = M^T M,
// and M = (well-conditioned) * (column scaling): its columns are kc "spikes" beta sigma_j e_j and kp dense columns whose
// mutual angles are those of the (preconditioned, hence nearly orthogonal) new right columns, so the Gram matrix of the
// NORMALISED columns is far from singular (independent subspaces: condition ~ 3) although M itself spans 8 decades --
// exactly the situation in which a Cholesky factorisation is accurate column by column.  One CTA per tile:
//   (a) the scaled Gram matrix Gs = D M^T M D, D = diag(|M e_c|)^-1, is ASSEMBLED from the structure (spike-spike: identity,
//       spike-dense: one product, dense-dense: kp (kp + 1) / 2 dot products) into the scratch S (r x r, upper, ld r);
//   (b) blocked left-looking Cholesky of Gs in S, 32 columns at a time: block-row update by DMMA (A operand staged in
//       shared memory, B streamed from L2), 32 x 32 diagonal factorisation with the pivot test (>= 1e-6: a column keeps
//       at least 1e-3 of its norm against the earlier ones), triangular solve of the block row;
//   (c) R' = Rc D^-1 into the panel buffer VW and the tile's Householder panel descriptor is switched off.
// A failed pivot test (nearly dependent columns) leaves VW untouched and the Householder R-only QR factors the tile as
// before.  4x fewer flops than the Householder QR (r^3 / 3), GEMM-shaped, one launch instead of 2 per 32 columns.
// Dynamic shared memory: vcore_chol_smem(r_bound).  grid = n_tiles, 256 threads.
constexpr int VC_NB = 32;
inline size_t vcore_chol_smem(int r_bound) {
    const size_t rp = (size_t) ((r_bound + 31) / 32) * 32;
    return sizeof(double) * (rp /*dscale*/ + (size_t) CQ_KP * CQ_P /*N*/ + (size_t) VC_NB * (VC_NB + 1) /*diag block*/ +
                             (size_t) VC_NB * (rp + 4) /*A operand*/) + sizeof(int) * rp /*ipos*/ + 64;
}
template<typename T>
__global__ void __launch_bounds__(256) k_vcore_chol(RecompProb<T> *__restrict__ probs, PanelDesc<T> *__restrict__ pd_vcore, int r_bound) {
    static_assert(std::is_same<T, double>::value, "FP64 tensor path");
    const int t = blockIdx.x;
    const RecompProb<T> p = probs[t];
    if (!p.active || !p.vinc || !pd_vcore[t].active) return;
    const int r = p.r, kc = p.kc, kp = p.kp;
    if (kp > CQ_KP || r > r_bound || r < 2) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, c4 = lane & 3;
    extern __shared__ __align__(16) unsigned char vc_smem[];
    const int rp = ((r_bound + 31) / 32) * 32, AP = rp + 4;
    T *dscale = reinterpret_cast<T *>(vc_smem);
    T *Nm = dscale + rp;                          // [CQ_KP][CQ_P] dense-dense Gram (unscaled)
    T *Dk = Nm + CQ_KP * CQ_P;                    // [32][33] diagonal block
    T *As = Dk + VC_NB * (VC_NB + 1);             // [32][AP]: As[i][k] = S[k][n0 + i]
    int *ipos = reinterpret_cast<int *>(As + (size_t) VC_NB * AP);
    __shared__ int s_bad;
    const T *M = p.RVp;                           // RV * Pi (kept copy), ld r
    T *S = p.T1;                                  // scratch r x r (free until the V S' products)
    if (tid == 0) s_bad = 0;
    for (int o = tid; o < r; o += 256) ipos[p.pos[o]] = o;
    __syncthreads();
    // ---- (a) column norms / scales: spikes from their single entry, dense columns by a warp each
    for (int cs = tid; cs < r; cs += 256) {
        const int o = ipos[cs];
        if (o < kc) { const T v = M[(size_t) o + (size_t) cs * r]; dscale[cs] = v != T(0) ? T(1) / t_abs(v) : T(0); }
    }
    // dense-dense Gram matrix N = Dn^T Dn (Dn = the kp dense columns of M, r rows) by DMMA, as in k_cholqr_pass: warp w takes
    // the row chunks w, w + 8, ... of 32 rows; the 21 upper tile pairs stay in registers and meet in shared memory
    for (int idx = tid; idx < CQ_KP * CQ_P; idx += 256) Nm[idx] = T(0);
    __syncthreads();
    {
        T acc[CQ_NT * (CQ_NT + 1) / 2][2];
#pragma unroll
        for (int q = 0; q < CQ_NT * (CQ_NT + 1) / 2; ++q) acc[q][0] = acc[q][1] = T(0);
        const T *colp[CQ_NT];
        bool colv[CQ_NT];
#pragma unroll
        for (int tt = 0; tt < CQ_NT; ++tt) {
            const int l = 8 * tt + g;
            colv[tt] = l < kp;
            colp[tt] = M + (size_t) (colv[tt] ? p.pos[kc + l] : 0) * r;
        }
        for (int r0 = w * 32; r0 < r; r0 += 256) {
#pragma unroll 2
            for (int ks = 0; ks < 32; ks += 4) {
                const int row = r0 + ks + c4;
                T v[CQ_NT];
#pragma unroll
                for (int tt = 0; tt < CQ_NT; ++tt) v[tt] = (row < r && colv[tt]) ? colp[tt][row] : T(0);
                int q = 0;
#pragma unroll
                for (int ti = 0; ti < CQ_NT; ++ti)
#pragma unroll
                    for (int tj = ti; tj < CQ_NT; ++tj, ++q) dmma_m8n8k4(acc[q][0], acc[q][1], v[ti], v[tj]);
            }
        }
        int q = 0;
#pragma unroll
        for (int ti = 0; ti < CQ_NT; ++ti)
#pragma unroll
            for (int tj = ti; tj < CQ_NT; ++tj, ++q) {
                atomicAdd(&Nm[(8 * ti + g) * CQ_P + 8 * tj + 2 * c4], acc[q][0]);
                atomicAdd(&Nm[(8 * ti + g) * CQ_P + 8 * tj + 2 * c4 + 1], acc[q][1]);
            }
    }
    __syncthreads();
    for (int idx = tid; idx < kp * kp; idx += 256) {   // mirror the upper triangle (the diagonal tiles hold both halves already)
        const int i = idx % kp, j = idx / kp;
        if (i > j && (i >> 3) != (j >> 3)) Nm[i * CQ_P + j] = Nm[j * CQ_P + i];
    }
    __syncthreads();
    for (int l = tid; l < kp; l += 256) { const T v = Nm[l * CQ_P + l]; dscale[p.pos[kc + l]] = v > T(0) ? T(1) / t_sqrt(v) : T(0); }
    __syncthreads();
    // Negligible columns are deflated (as in k_cholqr_pass): a column below 1e-13 of the largest one -- in practice the new
    // columns that CholeskyQR2 deflated: their part outside span(W) is zero, so they are exact combinations of the spikes
    // and would give an exactly zero pivot -- gets an identity row / column here and a zero column in R'.
    __shared__ T s_dmin;
    if (tid < 32) {
        T mn = T(1e300);
        for (int cs = tid; cs < r; cs += 32) { const T d = dscale[cs]; if (d > T(0) && d < mn) mn = d; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const T other = __shfl_xor_sync(0xffffffffu, mn, o); mn = other < mn ? other : mn; }
        if (tid == 0) s_dmin = mn;                 // 1 / (largest column norm)
    }
    __syncthreads();
    for (int cs = tid; cs < r; cs += 256)
        if (dscale[cs] > T(1e13) * s_dmin) dscale[cs] = T(0);
    __syncthreads();
    // scaled Gram matrix, upper triangle in sorted order (a zero column -- deflated -- gets an identity row / column)
    for (int idx = tid; idx < r * r; idx += 256) {
        const int i = idx % r, j = idx / r;
        if (i > j) continue;
        const int oi = ipos[i], oj = ipos[j];
        T v;
        if (i == j) v = T(1);
        else if (dscale[i] == T(0) || dscale[j] == T(0)) v = T(0);
        else if (oi < kc && oj < kc) v = T(0);
        else if (oi >= kc && oj >= kc) v = Nm[(oi - kc) * CQ_P + (oj - kc)] * dscale[i] * dscale[j];
        else {  // spike (row o_s) against a dense column: M[o_s][spike col] * M[o_s][dense col]
            const int cs = oi < kc ? i : j, cd = oi < kc ? j : i, os = oi < kc ? oi : oj;
            v = M[(size_t) os + (size_t) cs * r] * M[(size_t) os + (size_t) cd * r] * dscale[i] * dscale[j];
        }
        S[idx] = v;
    }
    __syncthreads();
    // ---- (b) blocked left-looking Cholesky (upper factor) of S
    for (int n0 = 0; n0 < r && !s_bad; n0 += VC_NB) {
        const int jw = min(VC_NB, r - n0), K = n0;
        if (K > 0) {
            for (int idx = tid; idx < VC_NB * K; idx += 256) {
                const int k = idx % K, i = idx / K;
                As[(size_t) i * AP + k] = (i < jw) ? S[(size_t) k + (size_t) (n0 + i) * r] : T(0);
            }
            __syncthreads();
            const int ntiles = (r - n0 + 7) / 8;
            for (int nt = w; nt < ntiles; nt += 8) {
                T acc[4][2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) acc[mi][0] = acc[mi][1] = T(0);
                const int colb = n0 + nt * 8 + g;
                const T *bcol = S + (size_t) (colb < r ? colb : 0) * r;
#pragma unroll 4
                for (int k0 = 0; k0 < K; k0 += 4) {
                    const T b = colb < r ? bcol[k0 + c4] : T(0);
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) dmma_m8n8k4(acc[mi][0], acc[mi][1], As[(size_t) (mi * 8 + g) * AP + k0 + c4], b);
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = n0 + mi * 8 + g, col = n0 + nt * 8 + 2 * c4 + h;
                        if (row < n0 + jw && col < r && row <= col) S[(size_t) row + (size_t) col * r] -= acc[mi][h];
                    }
            }
            __syncthreads();
        }
        // diagonal block -> shared memory, Cholesky with the pivot test
        for (int idx = tid; idx < VC_NB * VC_NB; idx += 256) {
            const int i = idx % VC_NB, j = idx / VC_NB;
            Dk[i * (VC_NB + 1) + j] = (i <= j && j < jw) ? S[(size_t) (n0 + i) + (size_t) (n0 + j) * r] : T(0);
        }
        __syncthreads();
        if (w == 0) {   // 32 x 32: one warp, warp-level barriers only (lane l owns column l of the trailing block)
            for (int j = 0; j < jw; ++j) {
                const T piv = Dk[j * (VC_NB + 1) + j];
                if (!(piv >= T(1e-6)) || !(piv <= T(4))) { if (lane == 0) s_bad = 1; break; }
                const T inv = T(1) / piv;
                if (lane > j && lane < jw) {
                    const T f = Dk[j * (VC_NB + 1) + lane] * inv;
                    for (int i = j + 1; i <= lane; ++i) Dk[i * (VC_NB + 1) + lane] = fma(-Dk[j * (VC_NB + 1) + i], f, Dk[i * (VC_NB + 1) + lane]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (s_bad) break;
        for (int idx = tid; idx < jw * jw; idx += 256) {   // rows divided by sqrt(pivot)
            const int i = idx % jw, j = idx / jw;
            if (i < j) Dk[i * (VC_NB + 1) + j] *= T(1) / t_sqrt(Dk[i * (VC_NB + 1) + i]);
        }
        __syncthreads();
        for (int j = tid; j < jw; j += 256) Dk[j * (VC_NB + 1) + j] = t_sqrt(Dk[j * (VC_NB + 1) + j]);
        __syncthreads();
        for (int idx = tid; idx < jw * jw; idx += 256) {
            const int i = idx % jw, j = idx / jw;
            if (i <= j) S[(size_t) (n0 + i) + (size_t) (n0 + j) * r] = Dk[i * (VC_NB + 1) + j];
        }
        // block row: x = Rd^-T s for every column to the right (forward substitution, a thread per column)
        for (int col = n0 + jw + tid; col < r; col += 256) {
            T x[VC_NB];
            T *sc = S + (size_t) n0 + (size_t) col * r;
            for (int i = 0; i < jw; ++i) {
                T a = sc[i];
                for (int l = 0; l < i; ++l) a = fma(-Dk[l * (VC_NB + 1) + i], x[l], a);
                x[i] = a / Dk[i * (VC_NB + 1) + i];
            }
            for (int i = 0; i < jw; ++i) sc[i] = x[i];
        }
        __syncthreads();
    }
    __syncthreads();
    if (s_bad) {   // nearly dependent columns: the Householder R-only QR takes this tile (VW still holds M)
        if (tid == 0 && p.stats) atomicAdd(p.stats + 7, 1ull);
        return;
    }
    // ---- (c) R' = Rc D^-1 into the panel buffer (upper triangle; the lower part is masked by its readers)
    for (int idx = tid; idx < r * r; idx += 256) {
        const int i = idx % r, j = idx / r;
        if (i > j) continue;
        const T d = dscale[j];
        p.VW[idx] = d != T(0) ? S[idx] / d : T(0);
    }
    if (tid == 0) {
        pd_vcore[t].active = 0;
        if (p.stats) atomicAdd(p.stats + 6, 1ull);
    }
}

// Both sides incremental: the part of the core K' = (RU Pi) R'^T that is a pure gather -- row i < kc of K' is column
// pos[i] of the triangular factor R' (in VW, ld r, reflectors below the diagonal masked), rows >= kc are zero -- and
// Rn = the kp columns pos[kc + l] of R' for the rank-kp GEMM K' += Xu Rn^T that follows.  32 x 32 tiles through shared
// memory (coalesced on both sides).  grid = (tile chunks, n_tiles), block (32, 8)
template<typename T>
__global__ void __launch_bounds__(256) k_both_core(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active || !p.both) return;
    __shared__ T tile[32][33];
    const int r = p.r, nt = (r + 31) / 32, tx = threadIdx.x, ty = threadIdx.y;
    for (int t = blockIdx.x; t < nt * nt; t += gridDim.x) {
        const int i0 = (t % nt) * 32, j0 = (t / nt) * 32;   // block of K' rows [i0, i0+32) x columns [j0, j0+32)
        for (int q = ty; q < 32; q += 8) {                   // read: column pos[i0 + q] of R', rows j0 + tx
            const int i = i0 + q, j = j0 + tx;
            T v = T(0);
            if (i < p.kc && j < r) { const int c = p.pos[i]; v = j <= c ? p.VW[(size_t) j + (size_t) c * r] : T(0); }
            tile[q][tx] = v;
        }
        __syncthreads();
        for (int q = ty; q < 32; q += 8) {                   // write: K'(i0 + tx, j0 + q)
            const int i = i0 + tx, j = j0 + q;
            if (i < r && j < r) p.M[(size_t) i + (size_t) j * p.a] = tile[tx][q];
        }
        __syncthreads();
    }
    const int total = r * p.kp;
    for (int idx = blockIdx.x * 256 + ty * 32 + tx; idx < total; idx += gridDim.x * 256) {
        const int j = idx % r, l = idx / r, c = p.pos[p.kc + l];
        p.Rn[idx] = j <= c ? p.VW[(size_t) j + (size_t) c * r] : T(0);
    }
}

// Incremental V side: Q2vT (kp x n, ld kp) = Q2v^T, so that [CV ; Q2vT] is one transposed two-segment GEMM operand.
// grid = (tile chunks, n_tiles), block (32, 8)
template<typename T>
__global__ void __launch_bounds__(256) k_vinc_transpose_q2(const RecompProb<T> *__restrict__ probs) {
    const RecompProb<T> p = probs[blockIdx.y];
    if (!p.active || !p.vinc) return;
    __shared__ T tile[32][33];
    const int tr = (p.n + 31) / 32, tc = (p.kp + 31) / 32, tx = threadIdx.x, ty = threadIdx.y;
    for (int t = blockIdx.x; t < tr * tc; t += gridDim.x) {
        const int r0 = (t % tr) * 32, c0 = (t / tr) * 32;  // r: row of Q2v (0..n), c: column (0..kp)
        for (int j = ty; j < 32; j += 8) {
            const int r = r0 + tx, c = c0 + j;
            tile[j][tx] = (r < p.n && c < p.kp) ? p.Q2v[(size_t) r + (size_t) c * p.n] : T(0);
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int c = c0 + tx, r = r0 + j;
            if (r < p.n && c < p.kp) p.Q2vT[(size_t) c + (size_t) r * p.kp] = tile[tx][j];
        }
        __syncthreads();
    }
}

// DDC epilogue (HCore.cpp:291-298 -> CompressedTile::ReadjustTile, Compressed.cpp:696-734): C becomes full rank,
// rank = min(m, n), with the dense result T held in one factor and the identity in the other.  For m >= n this is
// exactly the reference's (U = T, V = I).  For m < n the reference's index arithmetic is wrong (its own comment:
// "not handled correctly", HCore.cpp:296); here the mathematically equivalent (U = I, V = T) is written instead.
// grid = (chunks, n_tiles)
template<typename T>
__global__ void __launch_bounds__(256) k_ddc_finalize(const hcb_tile *__restrict__ Ctiles, const T *__restrict__ ws,
                                                      size_t slab, size_t o_w1, const int *__restrict__ info) {
    const hcb_tile C = Ctiles[blockIdx.y];
    if (info && info[blockIdx.y] != 0) return;
    const int m = C.m, n = C.n, rank = m < n ? m : n;
    const T *W = ws + (size_t) blockIdx.y * slab + o_w1;  // T, m x n, ld m
    T *CU = (T *) C.d_data, *CV = CU + (size_t) m * C.max_rank;
    const int gstride = gridDim.x * blockDim.x, g0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) {
        for (int idx = g0; idx < m * rank; idx += gstride) CU[idx] = W[idx];
        for (int idx = g0; idx < rank * n; idx += gstride) CV[idx] = (idx % rank == idx / rank) ? T(1) : T(0);
    } else {
        for (int idx = g0; idx < m * rank; idx += gstride) CU[idx] = (idx % m == idx / m) ? T(1) : T(0);
        for (int idx = g0; idx < rank * n; idx += gstride) CV[idx] = W[idx];  // V = T (rank == m, ld m)
    }
    if (g0 == 0) {
        *C.d_rank = rank;
        if (C.d_state) *C.d_state = 0;  // (U = T, V = I) is not an orthonormal-U representation
    }
}

// Initial compression epilogue (Compressed.cpp:103-135): rank rule, U = Uf[:, :rk], V = diag(sigma) Vf[:, :rk]^T.
// Us/Vs are the Jacobi outputs for M = A (m >= n) or M = A^T (m < n): Us normalised (a x s), Vs = V diag(sigma) (s x s).
template<typename T>
struct CompressProb {
    const T *Us; const T *Vs; const T *sigma;
    T *U; T *V; int *rank_ptr; int *info;
    int m, n, s, a, transposed, max_rank;
    int ld_us, ld_vs;  // leading dimensions of Us / Vs
    int tail_check;    // > 0 (sketched SVD): accept only if sigma[s - tail_check] is far below the threshold, else
                       // *info = 8 and the tile is left untouched (the caller re-does it with the full SVD)
    int *uinfo;        // optional, the CALLER's info word of this tile: |= 2 when the rank was clipped at max_rank
};

// Uniform(-1, 1) test matrix for the range finder (counter-based: splitmix64 of the element index).
template<typename T>
__global__ void k_fill_uniform(T *__restrict__ out, size_t n, unsigned long long seed) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        out[i] = (T) ((double) (z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
    }
}

// Sketched compression glue, one CTA per problem.  mode 0: Mr = R^T (k x k lower triangle) from the QR-factored B^T
// panel (R in its upper triangle, reflectors below); mode 1: Vn = [Vs ; 0] (n x k, ld n) from Vs (k x k).
template<typename T>
struct SketchGlue {
    const T *src;
    T *dst;
    int rows, k, ld_src;
};
template<typename T>
__global__ void k_sketch_glue(const SketchGlue<T> *__restrict__ gs, int mode) {
    const SketchGlue<T> g = gs[blockIdx.x];
    if (mode == 0) {
        for (int idx = threadIdx.x; idx < g.k * g.k; idx += blockDim.x) {
            const int i = idx % g.k, j = idx / g.k;  // Mr(i, j) = R(j, i) for j <= i
            g.dst[idx] = (j <= i) ? g.src[(size_t) j + (size_t) i * g.ld_src] : T(0);
        }
    } else {
        for (size_t idx = threadIdx.x; idx < (size_t) g.rows * g.k; idx += blockDim.x) {
            const int i = (int) (idx % g.rows), c = (int) (idx / g.rows);
            g.dst[idx] = i < g.k ? g.src[(size_t) i + (size_t) c * g.ld_src] : T(0);
        }
    }
}

// Q_explicit start: [I_k ; 0] (m x k, ld m), one CTA per problem
template<typename T>
__global__ void k_eye_batched(T *const *__restrict__ mats, const int *__restrict__ ms, int k) {
    T *A = mats[blockIdx.x];
    const int m = ms[blockIdx.x];
    for (size_t idx = threadIdx.x; idx < (size_t) m * k; idx += blockDim.x) A[idx] = (idx % m == idx / m) ? T(1) : T(0);
}

template<typename T>
__global__ void __launch_bounds__(256) k_compress_finalize(const CompressProb<T> *__restrict__ probs, T accuracy,
                                                           int truncated, int fixed_rank) {
    const CompressProb<T> p = probs[blockIdx.x];
    __shared__ int s_rk;
    if (threadIdx.x == 0) {
        int rk;
        if (fixed_rank > 0) rk = fixed_rank < p.s ? fixed_rank : p.s;
        else rk = new_rank_rule(p.sigma, p.s, accuracy, truncated);
        if (rk < 1) rk = 1;
        const bool clipped = rk > p.max_rank;
        if (clipped) rk = p.max_rank;  // Compressed.cpp:117-119 (silent clamp in the reference; reported here)
        if (p.tail_check > 0) {  // sketched SVD: the captured spectrum must have decayed well below the threshold
            const T thr = T(0.01) * accuracy * (truncated ? p.sigma[0] : T(1));
            const bool ok = rk + p.tail_check <= p.s && p.sigma[p.s - p.tail_check] <= thr;
            if (p.info) *p.info = ok ? 0 : 8;
            if (!ok) rk = -1;
        }
        s_rk = rk;
        if (rk > 0) *p.rank_ptr = rk;
        if (rk > 0 && clipped && p.uinfo) atomicOr(p.uinfo, 2);
    }
    __syncthreads();
    const int rk = s_rk;
    if (rk < 0) return;
    if (!p.transposed) {  // A = Us S V^T : U = Us[:, :rk], V = (V S)^T = Vs^T
        for (int idx = threadIdx.x; idx < p.m * rk; idx += blockDim.x) {
            const int i = idx % p.m, c = idx / p.m;
            p.U[(size_t) i + (size_t) c * p.m] = p.Us[(size_t) i + (size_t) c * p.ld_us];
        }
        for (int idx = threadIdx.x; idx < rk * p.n; idx += blockDim.x) {
            const int c = idx % rk, j = idx / rk;
            p.V[(size_t) c + (size_t) j * rk] = p.Vs[(size_t) j + (size_t) c * p.ld_vs];
        }
    } else {  // A^T = Us S V^T -> A = V S Us^T : U = Vs / sigma (m x rk), V = sigma * Us^T (rk x n)
        for (int idx = threadIdx.x; idx < p.m * rk; idx += blockDim.x) {
            const int i = idx % p.m, c = idx / p.m;
            const T sg = p.sigma[c];
            p.U[(size_t) i + (size_t) c * p.m] = sg > T(0) ? p.Vs[(size_t) i + (size_t) c * p.ld_vs] / sg : T(0);
        }
        for (int idx = threadIdx.x; idx < rk * p.n; idx += blockDim.x) {
            const int c = idx % rk, j = idx / rk;
            p.V[(size_t) c + (size_t) j * rk] = p.sigma[c] * p.Us[(size_t) j + (size_t) c * p.ld_us];
        }
    }
}

}  // namespace hcb
