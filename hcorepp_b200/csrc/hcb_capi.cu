// hcb_capi.cu -- the C ABI of libhcore_b200.so (declared in include/hcore_b200.h) and the host-side orchestration
// of the fused batched TLR-GEMM flow.  Host code here only builds descriptors and enqueues kernels: there is no
// CPU arithmetic path and no fallback -- without a CUDA device every compute entry point returns HCB_ENODEVICE.
#include "common.cuh"
#include "kernels_blas.cuh"
#include "kernels_dmma.cuh"
#include "kernels_qr.cuh"
#include "kernels_svd.cuh"
#include "kernels_svd_rx.cuh"
#include "kernels_tlr.cuh"
#include "kernels_chol.cuh"

#include <algorithm>
#include <type_traits>
#include <unordered_map>

namespace hcb {
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int ensure_ws(hcb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return HCB_OK;
    // grow-only arena: growing synchronises (in-flight kernels may still use the old arena) -- callers reserve up
    // front (hcb_ctx_reserve_workspace) so that steady-state calls never get here.
    HCB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->ws) HCB_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    const size_t want = align_up(bytes + bytes / 8, 1 << 20);
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) return fail(HCB_ENOMEM, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e));
    ctx->ws_bytes = want;
    return HCB_OK;
}

int ring_upload(hcb_ctx *ctx, const void *host, size_t bytes, void **d_out) {
    ParamRing &r = ctx->ring;
    const size_t need = align_up(bytes, 256);
    if (need * ParamRing::NSEG > r.cap) {  // (re)allocate: a segment must hold the largest upload
        HCB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (r.h) HCB_CUDA(cudaFreeHost(r.h));
        if (r.d) HCB_CUDA(cudaFree(r.d));
        r.cap = align_up(std::max(need * 2, (size_t) 1 << 20), 4096) * ParamRing::NSEG;
        r.off = 0;
        r.seg = 0;
        HCB_CUDA(cudaMallocHost((void **) &r.h, r.cap));
        HCB_CUDA(cudaMalloc((void **) &r.d, r.cap));
        for (int i = 0; i < ParamRing::NSEG; ++i) {
            if (!r.ev[i]) HCB_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
            r.ev_pending[i] = false;
        }
    }
    const size_t seg_bytes = r.cap / ParamRing::NSEG;
    if (r.off + need > (size_t) (r.seg + 1) * seg_bytes) {
        // leave this segment: everything issued from it so far is covered by its event; enter the next one once ITS
        // previous lap has drained
        HCB_CUDA(cudaEventRecord(r.ev[r.seg], ctx->stream));
        r.ev_pending[r.seg] = true;
        r.seg = (r.seg + 1) % ParamRing::NSEG;
        r.off = (size_t) r.seg * seg_bytes;
        if (r.ev_pending[r.seg]) {
            HCB_CUDA(cudaEventSynchronize(r.ev[r.seg]));
            r.ev_pending[r.seg] = false;
        }
    }
    std::memcpy(r.h + r.off, host, bytes);
    HCB_CUDA(cudaMemcpyAsync(r.d + r.off, r.h + r.off, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d_out = r.d + r.off;
    r.off += need;
    return HCB_OK;
}

struct PhaseScope {  // records begin/end events around one phase when timing is enabled
    hcb_ctx *ctx;
    cudaEvent_t end = nullptr;
    PhaseScope(hcb_ctx *c, int phase) : ctx(c) {
        if (!c->timing) return;
        auto get = [&]() {
            if (c->ev_used == c->ev_pool.size()) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                c->ev_pool.push_back(e);
            }
            return c->ev_pool[c->ev_used++];
        };
        cudaEvent_t beg = get();
        end = get();
        cudaEventRecord(beg, c->stream);
        c->phase_recs.push_back({phase, beg, end});
    }
    ~PhaseScope() {
        if (end) cudaEventRecord(end, ctx->stream);
    }
};

static int check_ctx(hcb_ctx *ctx) {
    if (!ctx) return fail(HCB_EINVAL, "null context");
    HCB_CUDA(cudaSetDevice(ctx->device));
    return HCB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// launch helpers (device-resident problem arrays)
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
int launch_gemm(hcb_ctx *ctx, const GemmProb<T> *d_probs, int n_probs, int m_bound, int n_bound) {
    if (n_probs <= 0 || m_bound <= 0 || n_bound <= 0) return HCB_OK;
    for (int off = 0; off < n_probs; off += 65535) {
        const int cnt = std::min(65535, n_probs - off);
        if (m_bound <= 32 && n_bound <= 32) {
            dim3 grid(1, cnt);
            k_gemm_batched<T, 32, 32, 16><<<grid, 256, 0, ctx->stream>>>(d_probs + off);
            HCB_LAUNCH_CHECK("k_gemm_batched");
            continue;
        }
        if constexpr (std::is_same<T, double>::value) {
            // FP64: tensor pipe (DMMA), three-stage ring fed by TMA bulk copies / cp.async. Tile shape follows the skinny side.
            auto go = [&](auto kern, size_t smem, int tiles) -> int {
                HCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
                dim3 grid(std::max(1, tiles), cnt);
                kern<<<grid, 128, smem, ctx->stream>>>(d_probs + off);
                return HCB_OK;
            };
            if (m_bound <= 32) HCB_TRY(go(k_gemm_dmma<1, 4>, dg_smem_bytes<1, 4>(), cdiv(n_bound, 128)));
            else if (n_bound <= 32) HCB_TRY(go(k_gemm_dmma<4, 1>, dg_smem_bytes<4, 1>(), cdiv(m_bound, 128)));
            else if (n_bound <= 48 && m_bound > 64)  // skinny right side (n = kp = 44): 128 x 48 tiles, all four warps along M
                HCB_TRY(go(k_gemm_dmma<4, 1, 6>, dg_smem_bytes<4, 1, 6>(), cdiv(m_bound, 128)));
            else HCB_TRY(go(k_gemm_dmma<2, 2>, dg_smem_bytes<2, 2>(), cdiv(m_bound, 64) * cdiv(n_bound, 64)));
            HCB_LAUNCH_CHECK("k_gemm_dmma");
        } else {
            // FP32: FFMA (TF32 tensor MMA would cost ~1e-3 relative error, incompatible with accuracy <= 1e-4)
            if (m_bound <= 32) {
                dim3 grid(std::max(1, cdiv(n_bound, 64)), cnt);
                k_gemm_batched<T, 32, 64, 16><<<grid, 256, 0, ctx->stream>>>(d_probs + off);
            } else {
                dim3 grid(std::max(1, cdiv(m_bound, 64) * cdiv(n_bound, 64)), cnt);
                k_gemm_batched<T, 64, 64, 16><<<grid, 256, 0, ctx->stream>>>(d_probs + off);
            }
            HCB_LAUNCH_CHECK("k_gemm_batched");
        }
    }
    return HCB_OK;
}

template<typename T>
int launch_copy(hcb_ctx *ctx, const CopyProb<T> *d_probs, int n_probs, int rows_bound, int cols_bound) {
    if (n_probs <= 0 || rows_bound <= 0 || cols_bound <= 0) return HCB_OK;
    for (int off = 0; off < n_probs; off += 65535) {
        const int cnt = std::min(65535, n_probs - off);
        dim3 grid(std::max(1, std::min(1024, cdiv(rows_bound, 32) * cdiv(cols_bound, 32))), cnt);
        k_copy_batched<T><<<grid, dim3(32, 8), 0, ctx->stream>>>(d_probs + off);
        HCB_LAUNCH_CHECK("k_copy_batched");
    }
    return HCB_OK;
}

template<typename T>
int launch_qr(hcb_ctx *ctx, const QrProb<T> *d_probs, int n_probs) {
    if (n_probs <= 0) return HCB_OK;
    k_geqrf_batched<T><<<n_probs, 512, 0, ctx->stream>>>(d_probs);
    HCB_LAUNCH_CHECK("k_geqrf_batched");
    return HCB_OK;
}

template<typename T>
int launch_refl(hcb_ctx *ctx, const ReflProb<T> *d_probs, int n_probs, int nvec_bound) {
    if (n_probs <= 0 || nvec_bound <= 0) return HCB_OK;
    for (int off = 0; off < n_probs; off += 65535) {
        const int cnt = std::min(65535, n_probs - off);
        dim3 grid(std::max(1, cdiv(nvec_bound, 8)), cnt);
        k_apply_reflectors<T><<<grid, 256, 0, ctx->stream>>>(d_probs + off);
        HCB_LAUNCH_CHECK("k_apply_reflectors");
    }
    return HCB_OK;
}

template<typename T>
int launch_svd(hcb_ctx *ctx, const SvdProb<T> *d_probs, int n_probs, int a_bound, int b_bound, double stop2 = 0.0) {
    if (n_probs <= 0) return HCB_OK;
    // shared memory: enough for the whole bound-sized problem (zero-padded column pitch of 64 rows), capped at the
    // opt-in limit; the kernel decides per problem (from its true a, b) which regime applies.
    const size_t pitch = (size_t) cdiv(a_bound, 64) * 64;
    size_t want = (pitch * b_bound + (size_t) b_bound) * sizeof(T);
    const size_t cap = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
    if (a_bound > 128 && a_bound <= 64 * RX_MAX_NI_TALL && want + (size_t) b_bound * sizeof(T) > cap) {
        // does not fit shared memory as a whole: block Jacobi with the pivot block in registers
        const size_t rx = align_up(rx_smem_bytes<T>(cdiv(a_bound, 64), b_bound), 16);
        if (rx <= cap) {
            HCB_CUDA(cudaFuncSetAttribute(k_jacobi_svd_rx<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) rx));
            const size_t need = (size_t) 2 * n_probs + 2;  // item counter, finished counter, per-problem state, per-problem flags
            if (need > ctx->svd_sched_n) {
                if (ctx->svd_sched) { HCB_CUDA(cudaStreamSynchronize(ctx->stream)); HCB_CUDA(cudaFree(ctx->svd_sched)); }
                ctx->svd_sched = nullptr; ctx->svd_sched_n = 0;
                HCB_CUDA(cudaMalloc(&ctx->svd_sched, need * 2 * sizeof(int)));
                ctx->svd_sched_n = need * 2;
            }
            HCB_CUDA(cudaMemsetAsync(ctx->svd_sched, 0, need * sizeof(int), ctx->stream));
            // persistent: as many CTAs as are resident at once (the item scheduler relies on it), (sweep, problem) items
            // from an atomic counter
            int per_sm = 1;
            HCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_jacobi_svd_rx<T>, RX_THREADS, rx));
            const int grid = std::max(1, std::min(n_probs, ctx->sm_count * std::max(per_sm, 1)));
            k_jacobi_svd_rx<T><<<grid, RX_THREADS, rx, ctx->stream>>>(d_probs, n_probs, 40, ctx->svd_sched, (T) stop2);
            HCB_LAUNCH_CHECK("k_jacobi_svd_rx");
            return HCB_OK;
        }
    }
    if (want > cap) want = cap / 16 * 16;
    if ((size_t) std::max(b_bound, 1) * sizeof(T) > want) return fail(HCB_EUNSUPPORTED, "svd: problem too large");
    want = align_up(want, 16);
    if (a_bound <= 128) {  // one warp per column pair, 64 registers per thread
        HCB_CUDA(cudaFuncSetAttribute(k_jacobi_svd<T, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) want));
        const int threads = b_bound >= 48 ? 1024 : (b_bound >= 24 ? 512 : 256);
        k_jacobi_svd<T, 1024><<<n_probs, threads, want, ctx->stream>>>(d_probs, (int) (want / sizeof(T)), 40);
    } else {               // 128 registers per thread: both columns of a pair stay in registers up to 512 rows
        HCB_CUDA(cudaFuncSetAttribute(k_jacobi_svd<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) want));
        k_jacobi_svd<T, 512><<<n_probs, 512, want, ctx->stream>>>(d_probs, (int) (want / sizeof(T)), 40);
    }
    HCB_LAUNCH_CHECK("k_jacobi_svd");
    return HCB_OK;
}

// Blocked Householder QR of `cnt` panels described on the device (pds): per NBQ-column block -- panel factorisation,
// T_b + clean V_b, then the trailing update A2 -= V_b T_b^T (V_b^T A2) as three batched GEMMs.
template<typename T>
size_t qr_block_desc_bytes(int cnt, int cols_bound) {  // one descriptor set per (column block, panel)
    const size_t nb = (size_t) cnt * cdiv(cols_bound, NBQ);
    return align_up(sizeof(QrProb<T>) * nb, 256) + align_up(sizeof(LarftProb<T>) * nb, 256) +
           3 * align_up(sizeof(GemmProb<T>) * nb, 256) + align_up(sizeof(StripJob) * nb, 256);
}

// cluster launch of the strip-resident block-reflector kernel: `cnt` jobs, rows_bound rows
inline bool strip_path_ok(const hcb_ctx *ctx, int rows_bound) {
    return rows_bound <= SK_MAXCS * SKR && SK_SMEM <= ctx->smem_optin;
}
inline int launch_strips(hcb_ctx *ctx, const StripJob *jobs, int cnt, int rows_bound) {
    if (cnt <= 0) return HCB_OK;
    HCB_CUDA(cudaFuncSetAttribute(k_strip_reflect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SK_SMEM));
    int cs = 1;
    while (cs * SKR < rows_bound) cs <<= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) (cs * cnt));
    cfg.blockDim = dim3(SKR);
    cfg.dynamicSmemBytes = SK_SMEM;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned) cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    HCB_CUDA(cudaLaunchKernelEx(&cfg, k_strip_reflect, jobs));
    HCB_LAUNCH_CHECK("k_strip_reflect");
    return HCB_OK;
}

// tail_cols / tail_cnt (optional): only the FIRST tail_cnt panels can have more than tail_cols columns (the recompression
// passes [2n stack panels | 2n kp-column panels of the incremental path] in one call): column blocks beyond tail_cols are
// launched over those first panels only.
template<typename T>
int run_blocked_qr(hcb_ctx *ctx, const PanelDesc<T> *d_pds, int cnt, int rows_bound, int cols_bound, char *desc_store,
                   int tail_cols = 0, int tail_cnt = 0) {
    const int nfac = cdiv(std::min(rows_bound, cols_bound), NBQ);  // blocks that hold reflectors
    const int nblk = cdiv(cols_bound, NBQ);                         // column blocks (>= nfac for a wide panel)
    const size_t nb = (size_t) cnt * nblk;
    const bool strips = std::is_same<T, double>::value && strip_path_ok(ctx, rows_bound);
    QrBlockArrays<T> q;
    char *p = desc_store;
    q.qr = reinterpret_cast<QrProb<T> *>(p); p += align_up(sizeof(QrProb<T>) * nb, 256);
    q.lf = reinterpret_cast<LarftProb<T> *>(p); p += align_up(sizeof(LarftProb<T>) * nb, 256);
    q.gw = reinterpret_cast<GemmProb<T> *>(p); p += align_up(sizeof(GemmProb<T>) * nb, 256);
    q.gw2 = reinterpret_cast<GemmProb<T> *>(p); p += align_up(sizeof(GemmProb<T>) * nb, 256);
    q.gup = reinterpret_cast<GemmProb<T> *>(p); p += align_up(sizeof(GemmProb<T>) * nb, 256);
    q.sj = strips ? reinterpret_cast<StripJob *>(p) : nullptr;
    q.nblk = nblk;
    q.npan = cnt;
    k_setup_qr_blocks<T><<<cdiv(nblk * cnt, 128), 128, 0, ctx->stream>>>(d_pds, q);
    HCB_LAUNCH_CHECK("k_setup_qr_blocks");
    const size_t pq_smem = pr_smem_bytes<T>();
    const bool on_chip = rows_bound <= PR_MAXCS * PR_ROWS && pq_smem + 1024 <= ctx->smem_optin;
    if (on_chip) HCB_CUDA(cudaFuncSetAttribute(k_panel_qr_regs<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pq_smem));
    const int cnt_all = cnt;
    for (int b = 0; b < nblk; ++b) {
        const size_t o = (size_t) b * cnt_all;
        if (tail_cnt > 0 && b * NBQ >= tail_cols) cnt = std::min(cnt_all, tail_cnt);   // only the wide panels reach this block
        // left-looking: bring column block b up to date with blocks 0..b-1 while it sits in cluster shared memory
        if (strips && b > 0) HCB_TRY(launch_strips(ctx, q.sj + o, cnt, rows_bound));
        if (b >= nfac) continue;
        if (on_chip) {
            // cluster of 1/2/4/8 CTAs per panel block, 512 rows each, block resident in registers
            const int rows_b = std::max(1, rows_bound - b * NBQ);
            int cs = 1;
            while (cs * PR_ROWS < rows_b) cs <<= 1;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned) (cs * cnt));
            cfg.blockDim = dim3(PR_THREADS);
            cfg.dynamicSmemBytes = pq_smem;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned) cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const QrProb<T> *qarg = q.qr + o;
            const LarftProb<T> *larg = q.lf + o;
            HCB_CUDA(cudaLaunchKernelEx(&cfg, k_panel_qr_regs<T>, qarg, larg));
            HCB_LAUNCH_CHECK("k_panel_qr_regs");
        } else {
            HCB_TRY(launch_qr<T>(ctx, q.qr + o, cnt));
            k_larft_extract<T><<<cnt, 256, 0, ctx->stream>>>(q.lf + o);
            HCB_LAUNCH_CHECK("k_larft_extract");
        }
        const int nt_b = cols_bound - (b + 1) * NBQ;
        if (!strips && nt_b > 0) {
            HCB_TRY(launch_gemm<T>(ctx, q.gw + o, cnt, NBQ, nt_b));
            HCB_TRY(launch_gemm<T>(ctx, q.gw2 + o, cnt, NBQ, nt_b));
            HCB_TRY(launch_gemm<T>(ctx, q.gup + o, cnt, rows_bound - b * NBQ, nt_b));
        }
    }
    return HCB_OK;
}

// single host-side problem -> device (through the pinned ring) -> batched kernel with a batch of one
template<typename P>
int upload_one(hcb_ctx *ctx, const P &prob, const P **d_out) {
    void *d = nullptr;
    HCB_TRY(ring_upload(ctx, &prob, sizeof(P), &d));
    *d_out = reinterpret_cast<const P *>(d);
    return HCB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Batched SVD pipeline shared by the compat SVD entry point and the compressing constructor:
//   M = op(src) (a x b, a >= b)  ->  MT = M^T, QR(MT), L = R^T  ->  Jacobi on L (left vectors Us, sigma)
//   ->  Vs = M^T Us  (= V diag(sigma)).
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
struct SvdJob {
    const T *src;  // dense source
    int ld_src;
    int trans_src;  // 1: M = src^T
    int a, b;
    T *Us, *Vs, *sigma;  // outputs: Us a x b (ld a), Vs b x b (ld b), sigma b
    int *info = nullptr; // optional device info word: |= 1 Jacobi not converged, bits 8.. = sweeps
};

template<typename T>
struct SvdJobLayout {
    size_t eM, eTau, eTB, eWB, slab;
    int nblk;
    SvdJobLayout(int a, int b) {
        nblk = cdiv(b, NBQ);
        eM = align_up((size_t) (a + 1) * b, 32);  // (+1: the Jacobi work copy uses an even leading dimension)
        eTau = align_up((size_t) b, 32);
        eTB = align_up((size_t) NBQ * NBQ * nblk, 32);
        eWB = align_up((size_t) 2 * NBQ * a, 32);
        slab = 5 * eM + eTau + eTB + eWB;  // M | MT | Lb | Jwork | VC | tau | TB | WB
    }
};

template<typename P>
int stage_array(hcb_ctx *ctx, const std::vector<P> &host, P *d_dst) {
    void *st = nullptr;
    HCB_TRY(ring_upload(ctx, host.data(), sizeof(P) * host.size(), &st));
    HCB_CUDA(cudaMemcpyAsync(d_dst, st, sizeof(P) * host.size(), cudaMemcpyDeviceToDevice, ctx->stream));
    return HCB_OK;
}

// scratch: `ws` holds cnt slabs of SvdJobLayout(a_bound, b_bound).slab elements; desc: device bytes for descriptors
template<typename T>
size_t svd_jobs_desc_bytes(int cnt, int a_bound, int b_bound) {
    const SvdJobLayout<T> L(a_bound, b_bound);
    return 2 * align_up(sizeof(CopyProb<T>) * cnt, 256) + align_up(sizeof(PanelDesc<T>) * cnt, 256) +
           align_up(sizeof(QrProb<T>) * cnt, 256) + align_up(sizeof(LqProb<T>) * cnt, 256) +
           align_up(sizeof(SvdProb<T>) * cnt, 256) + align_up(sizeof(GemmProb<T>) * cnt, 256) +
           qr_block_desc_bytes<T>(cnt, a_bound) + 256;
}

template<typename T>
int run_svd_jobs(hcb_ctx *ctx, const std::vector<SvdJob<T>> &jobs, int a_bound, int b_bound, T *ws, char *desc) {
    const int cnt = (int) jobs.size();
    if (cnt == 0) return HCB_OK;
    const SvdJobLayout<T> L(a_bound, b_bound);
    std::vector<CopyProb<T>> c1(cnt), c2(cnt);
    std::vector<PanelDesc<T>> pd(cnt);
    std::vector<QrProb<T>> qp(cnt);
    std::vector<LqProb<T>> lq(cnt);
    std::vector<SvdProb<T>> sv(cnt);
    std::vector<GemmProb<T>> gv(cnt);
    for (int t = 0; t < cnt; ++t) {
        const SvdJob<T> &j = jobs[t];
        T *M = ws + (size_t) t * L.slab, *MT = M + L.eM, *Lb = MT + L.eM, *Jw = Lb + L.eM, *VC = Jw + L.eM, *tau = VC + L.eM,
          *TB = tau + L.eTau, *WB = TB + L.eTB;
        c1[t] = CopyProb<T>{j.src, M, j.a, j.b, j.ld_src, j.a, j.trans_src, T(1)};           // M  = op(src)   (a x b)
        c2[t] = CopyProb<T>{j.src, MT, j.b, j.a, j.ld_src, j.b, j.trans_src ? 0 : 1, T(1)};  // MT = op(src)^T (b x a)
        pd[t] = PanelDesc<T>{MT, tau, VC, TB, WB, j.b, j.a, a_bound, 1};
        qp[t] = QrProb<T>{MT, tau, j.b, j.a, j.b};
        lq[t] = LqProb<T>{MT, Lb, j.a, j.b};
        sv[t] = SvdProb<T>{Lb, Jw, j.Us, j.Vs, j.sigma, j.info, j.a, j.b, j.a, j.a, j.b};
        gv[t] = GemmProb<T>{M, j.Us, j.Vs, j.b, j.b, j.a, j.a, j.a, j.b, 1, 0, T(1), T(0)};  // Vs = M^T Us = V diag(sigma)
    }
    char *p = desc;
    auto carve = [&](size_t bytes) { char *q = p; p += align_up(bytes, 256); return q; };
    auto *d_c1 = reinterpret_cast<CopyProb<T> *>(carve(sizeof(CopyProb<T>) * cnt));
    auto *d_c2 = reinterpret_cast<CopyProb<T> *>(carve(sizeof(CopyProb<T>) * cnt));
    auto *d_pd = reinterpret_cast<PanelDesc<T> *>(carve(sizeof(PanelDesc<T>) * cnt));
    auto *d_qp = reinterpret_cast<QrProb<T> *>(carve(sizeof(QrProb<T>) * cnt));
    auto *d_lq = reinterpret_cast<LqProb<T> *>(carve(sizeof(LqProb<T>) * cnt));
    auto *d_sv = reinterpret_cast<SvdProb<T> *>(carve(sizeof(SvdProb<T>) * cnt));
    auto *d_gv = reinterpret_cast<GemmProb<T> *>(carve(sizeof(GemmProb<T>) * cnt));
    char *d_blocks = p;
    HCB_TRY(stage_array(ctx, c1, d_c1));
    HCB_TRY(stage_array(ctx, c2, d_c2));
    HCB_TRY(stage_array(ctx, pd, d_pd));
    HCB_TRY(stage_array(ctx, qp, d_qp));
    HCB_TRY(stage_array(ctx, lq, d_lq));
    HCB_TRY(stage_array(ctx, sv, d_sv));
    HCB_TRY(stage_array(ctx, gv, d_gv));
    HCB_TRY(launch_copy<T>(ctx, d_c1, cnt, a_bound, b_bound));
    HCB_TRY(launch_copy<T>(ctx, d_c2, cnt, b_bound, a_bound));
    if (b_bound > 2 * NBQ) HCB_TRY(run_blocked_qr<T>(ctx, d_pd, cnt, b_bound, a_bound, d_blocks));
    else HCB_TRY(launch_qr<T>(ctx, d_qp, cnt));
    {
        dim3 grid(std::max(1, std::min(64, cdiv((long long) a_bound * b_bound, 256))), cnt);
        k_extract_l<T><<<grid, 256, 0, ctx->stream>>>(d_lq);
        HCB_LAUNCH_CHECK("k_extract_l");
    }
    HCB_TRY(launch_svd<T>(ctx, d_sv, cnt, a_bound, b_bound));
    return launch_gemm<T>(ctx, d_gv, cnt, b_bound, b_bound);
}

// ---------------------------------------------------------------------------------------------------------------
// (2) fine-grained kernel table
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
int t_gemm(hcb_ctx *ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, T alpha, const T *A, int64_t lda, const T *B,
           int64_t ldb, T beta, T *C, int64_t ldc) {
    HCB_TRY(check_ctx(ctx));
    if (m < 0 || n < 0 || k < 0) return fail(HCB_EINVAL, "gemm: negative dimension");
    if (m == 0 || n == 0) return HCB_OK;
    GemmProb<T> g{A, B, C, (int) m, (int) n, (int) k, (int) lda, (int) ldb, (int) ldc, ta ? 1 : 0, tb ? 1 : 0, alpha, beta};
    const GemmProb<T> *d;
    HCB_TRY(upload_one(ctx, g, &d));
    return launch_gemm<T>(ctx, d, 1, (int) m, (int) n);
}

template<typename T>
int t_copy(hcb_ctx *ctx, const T *src, int lds, T *dst, int ldd, int rows, int cols, int trans, T scale) {
    if (rows <= 0 || cols <= 0) return HCB_OK;
    CopyProb<T> c{src, dst, rows, cols, lds, ldd, trans, scale};
    const CopyProb<T> *d;
    HCB_TRY(upload_one(ctx, c, &d));
    return launch_copy<T>(ctx, d, 1, rows, cols);
}

template<typename T>
int t_multiply_by_alpha(hcb_ctx *ctx, T *arr, int64_t rows, int64_t cols, int64_t m, int64_t rank, T alpha) {
    HCB_TRY(check_ctx(ctx));
    const size_t count = (size_t) rows * cols;
    if (count == 0) return HCB_OK;
    k_scale_flat<T><<<cdiv(count, 256), 256, 0, ctx->stream>>>(arr, (size_t) m * rank, count, alpha);
    HCB_LAUNCH_CHECK("k_scale_flat");
    return HCB_OK;
}

template<typename T>
int t_process_v(hcb_ctx *ctx, int64_t n, int64_t crank, int ungqr, int64_t vm, T beta, const T *CV, int64_t ldcv, T *V,
                int64_t arank, const T *B, int cholesky) {
    HCB_TRY(check_ctx(ctx));
    (void) ungqr;  // conj is the identity for real T (Definitions.hpp:10-13 instantiates float/double only)
    if (cholesky) {
        // omp/kernels.cpp:35-55: V (ld ldcv) = beta*CV, Vptr (ld arank) = B -- plain scaled copies
        HCB_TRY(t_copy<T>(ctx, CV, (int) ldcv, V, (int) ldcv, (int) crank, (int) n, 0, beta));
        return t_copy<T>(ctx, B, (int) arank, V + n * crank, (int) arank, (int) arank, (int) n, 0, T(1));
    }
    // omp/kernels.cpp:57-77: V[j + i*vm] = beta*CV[i + j*ldcv] ; Vptr[j + i*vm] = B[i + j*arank]
    HCB_TRY(t_copy<T>(ctx, CV, (int) ldcv, V, (int) vm, (int) n, (int) crank, 1, beta));
    return t_copy<T>(ctx, B, (int) arank, V + n * crank, (int) vm, (int) n, (int) arank, 1, T(1));
}

template<typename T>
int t_new_rank_device(hcb_ctx *ctx, int truncated, const T *sigma, int64_t size_s, T acc, int32_t *d_rank) {
    HCB_TRY(check_ctx(ctx));
    k_new_rank<T><<<1, 32, 0, ctx->stream>>>(sigma, (int) size_s, acc, truncated, d_rank);
    HCB_LAUNCH_CHECK("k_new_rank");
    return HCB_OK;
}

template<typename T>
int t_new_rank(hcb_ctx *ctx, int truncated, const T *sigma, int64_t size_s, T acc, int64_t *host_rank) {
    HCB_TRY(check_ctx(ctx));
    HCB_TRY(ensure_ws(ctx, 256));
    int32_t *d = reinterpret_cast<int32_t *>(ctx->ws);
    HCB_TRY(t_new_rank_device<T>(ctx, truncated, sigma, size_s, acc, d));
    int32_t h = 0;
    HCB_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    HCB_CUDA(cudaStreamSynchronize(ctx->stream));  // host-visible rank: the reference syncs here too
    *host_rank = h;
    return HCB_OK;
}

template<typename T>
int t_uvptr(hcb_ctx *ctx, int64_t rank, int64_t vm, T *UV, const T *Vnew) {
    HCB_TRY(check_ctx(ctx));  // UV (rank x vm, ld rank) = Vnew (vm x rank, ld vm)^T
    return t_copy<T>(ctx, Vnew, (int) vm, UV, (int) rank, (int) rank, (int) vm, 1, T(1));
}

template<typename T>
int t_vtnew(hcb_ctx *ctx, int64_t rk, int ungqr, int64_t min_vm_vn, const T *sigma, T *VT, int64_t size_s, int64_t vm) {
    HCB_TRY(check_ctx(ctx));
    const int cols = (int) (ungqr ? min_vm_vn : vm);
    if (rk <= 0 || cols <= 0) return HCB_OK;
    dim3 block(32, 8), grid(cdiv(rk, 32), cdiv(cols, 8));
    k_scale_rows<T><<<grid, block, 0, ctx->stream>>>(VT, (int) size_s, (int) rk, cols, sigma);
    HCB_LAUNCH_CHECK("k_scale_rows");
    return HCB_OK;
}

template<typename T>
int t_fill_identity(hcb_ctx *ctx, int64_t n, T *A) {
    HCB_TRY(check_ctx(ctx));
    if (n <= 0) return HCB_OK;
    k_fill_diag<T><<<cdiv(n, 256), 256, 0, ctx->stream>>>(A, (int) n, (int) n, T(1));
    HCB_LAUNCH_CHECK("k_fill_diag");
    return HCB_OK;
}

template<typename T>
int t_lacpy(hcb_ctx *ctx, int type, int64_t m, int64_t n, const T *A, int64_t lda, T *B, int64_t ldb) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    dim3 block(32, 8), grid(cdiv(m, 32), cdiv(n, 8));
    k_lacpy<T><<<grid, block, 0, ctx->stream>>>((char) type, (int) m, (int) n, A, (int) lda, B, (int) ldb);
    HCB_LAUNCH_CHECK("k_lacpy");
    return HCB_OK;
}

template<typename T>
int t_laset(hcb_ctx *ctx, int type, int64_t m, int64_t n, T off, T diag, T *A, int64_t lda) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    dim3 block(32, 8), grid(cdiv(m, 32), cdiv(n, 8));
    k_laset<T><<<grid, block, 0, ctx->stream>>>((char) type, (int) m, (int) n, off, diag, A, (int) lda);
    HCB_LAUNCH_CHECK("k_laset");
    return HCB_OK;
}

template<typename T>
int t_geqrf(hcb_ctx *ctx, int64_t m, int64_t n, T *A, int64_t lda, T *tau) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    QrProb<T> q{A, tau, (int) m, (int) n, (int) lda};
    const QrProb<T> *d;
    HCB_TRY(upload_one(ctx, q, &d));
    return launch_qr<T>(ctx, d, 1);
}

template<typename T>
int t_unmqr(hcb_ctx *ctx, int side, int trans, int64_t m, int64_t n, int64_t k, const T *A, int64_t lda, const T *tau,
            T *C, int64_t ldc) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0 || k <= 0) return HCB_OK;
    const bool right = (side == 1 || side == 'R');
    const bool tr = (trans != 0 && trans != 'N');
    // Q = H_0 ... H_{k-1}.  Left: Q C applies H_{k-1} first, Q^T C applies H_0 first.
    //                        Right: C Q applies H_0 first, C Q^T applies H_{k-1} first.
    const int forward = right ? (tr ? 0 : 1) : (tr ? 1 : 0);
    ReflProb<T> r{A, tau, C, (int) (right ? n : m), (int) k, (int) lda, (int) m, (int) n, (int) ldc, right ? 1 : 0,
                  forward, nullptr};
    const ReflProb<T> *d;
    HCB_TRY(upload_one(ctx, r, &d));
    return launch_refl<T>(ctx, d, 1, (int) (right ? m : n));
}

template<typename T>
int t_ungqr(hcb_ctx *ctx, int64_t m, int64_t n, int64_t k, T *A, int64_t lda, const T *tau) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    // explicit Q (m x n) = H_0..H_{k-1} [I;0]: keep the reflectors in scratch, overwrite A with [I;0], apply.
    const size_t bytes = align_up((size_t) m * k * sizeof(T), 256);
    HCB_TRY(ensure_ws(ctx, bytes));
    T *Vw = reinterpret_cast<T *>(ctx->ws);
    HCB_TRY(t_copy<T>(ctx, A, (int) lda, Vw, (int) m, (int) m, (int) k, 0, T(1)));
    HCB_TRY(t_laset<T>(ctx, 'G', m, n, T(0), T(1), A, lda));
    ReflProb<T> r{Vw, tau, A, (int) m, (int) k, (int) m, (int) m, (int) n, (int) lda, 0, 0, nullptr};
    const ReflProb<T> *d;
    HCB_TRY(upload_one(ctx, r, &d));
    return launch_refl<T>(ctx, d, 1, (int) n);
}

template<typename T>
int t_svd(hcb_ctx *ctx, int64_t m, int64_t n, T *A, int64_t lda, T *S, T *U, int64_t ldu, T *VT, int64_t ldvt) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    const int a = (int) std::max(m, n), b = (int) std::min(m, n);
    const bool transposed = m < n;
    const SvdJobLayout<T> L(a, b);
    const size_t eOut = align_up((size_t) a * b, 32) + align_up((size_t) b * b, 32);
    const size_t desc = svd_jobs_desc_bytes<T>(1, a, b);
    HCB_TRY(ensure_ws(ctx, desc + (L.slab + eOut) * sizeof(T) + 1024));
    char *base = reinterpret_cast<char *>(ctx->ws);
    T *ws = reinterpret_cast<T *>(base + align_up(desc, 256));
    T *Us = ws + L.slab, *Vs = Us + align_up((size_t) a * b, 32);
    std::vector<SvdJob<T>> jobs(1);
    jobs[0] = SvdJob<T>{A, (int) lda, transposed ? 1 : 0, a, b, Us, Vs, S};
    HCB_TRY(run_svd_jobs<T>(ctx, jobs, a, b, ws, base));
    {   // Vs = V diag(S): normalise its columns to get V
        dim3 block(32, 8), grid(cdiv(b, 32), cdiv(b, 8));
        k_unscale_cols<T><<<grid, block, 0, ctx->stream>>>(Vs, b, b, b, S);
        HCB_LAUNCH_CHECK("k_unscale_cols");
    }
    if (!transposed) {  // A = Us S Vs^T : U = Us (m x n), VT = Vs^T (n x n)
        HCB_TRY(t_copy<T>(ctx, Us, a, U, (int) ldu, (int) m, b, 0, T(1)));
        return t_copy<T>(ctx, Vs, b, VT, (int) ldvt, b, (int) n, 1, T(1));
    }
    // A^T = Us S Vs^T -> A = Vs S Us^T : U = Vs (m x m), VT = Us^T (m x n)
    HCB_TRY(t_copy<T>(ctx, Vs, b, U, (int) ldu, (int) m, b, 0, T(1)));
    return t_copy<T>(ctx, Us, a, VT, (int) ldvt, b, (int) n, 1, T(1));
}

template<typename T>
int t_trmm(hcb_ctx *ctx, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *A, int64_t lda,
           T *B, int64_t ldb) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    // B := alpha op(tri(A)) B (Left) or alpha B op(tri(A)) (Right): expand the triangle, GEMM into scratch, copy back.
    const bool right = (side == 'R' || side == 1);
    const int na = (int) (right ? n : m);
    const size_t eD = (size_t) na * na, eO = (size_t) m * n;
    HCB_TRY(ensure_ws(ctx, (eD + eO) * sizeof(T) + 1024));
    T *D = reinterpret_cast<T *>(ctx->ws), *O = D + eD;
    dim3 block(32, 8), grid(cdiv(na, 32), cdiv(na, 8));
    k_expand_tri<T><<<grid, block, 0, ctx->stream>>>((char) uplo, (char) diag, na, A, (int) lda, D);
    HCB_LAUNCH_CHECK("k_expand_tri");
    const int tr = (trans != 0 && trans != 'N') ? 1 : 0;
    GemmProb<T> g = right ? GemmProb<T>{B, D, O, (int) m, (int) n, na, (int) ldb, na, (int) m, 0, tr, alpha, T(0)}
                          : GemmProb<T>{D, B, O, (int) m, (int) n, na, na, (int) ldb, (int) m, tr, 0, alpha, T(0)};
    const GemmProb<T> *d;
    HCB_TRY(upload_one(ctx, g, &d));
    HCB_TRY(launch_gemm<T>(ctx, d, 1, (int) m, (int) n));
    return t_copy<T>(ctx, O, (int) m, B, (int) ldb, (int) m, (int) n, 0, T(1));
}

// ---------------------------------------------------------------------------------------------------------------
// (3) fused batched TLR GEMM
// ---------------------------------------------------------------------------------------------------------------
struct BatchShape {
    int m = 0, n = 0, k = 0;           // maxima over the batch
    int kA = 0, kB = 0, kC = 0, maxrankC = 0;  // rank bounds
    int mix = 0;
    int r_force = 0;  // > 0: stacked-rank bound given directly (workspace query) instead of kC + kA
};

static inline int bound_of(const hcb_tile &t) {
    if (t.type != HCB_TILE_COMPRESSED) return 0;
    return t.rank_bound > 0 ? std::min(t.rank_bound, t.max_rank) : t.max_rank;
}

template<typename T>
struct Layout {
    size_t slab = 0, o_w1 = 0, o_w2 = 0, o_uw = 0, o_vw = 0, o_tauu = 0, o_tauv = 0, o_m = 0, o_j = 0, o_us = 0,
           o_vs = 0, o_sig = 0, o_vn = 0, o_vcu = 0, o_vcv = 0, o_tbu = 0, o_tbv = 0, o_wbu = 0, o_wbv = 0, o_mt = 0,
           o_taum = 0, o_lb = 0, o_vcm = 0, o_tbm = 0, o_wbm = 0, o_pj = 0, o_pos = 0, o_gu = 0, o_gu2 = 0, o_q2 = 0,
           o_tu = 0, o_sig0 = 0, o_hv = 0, o_zv = 0, o_q2v = 0, o_q2vt = 0, o_rvp = 0, o_t1 = 0, o_xu = 0, o_xv = 0, o_rn = 0,
           o_t1b = 0;
    int r_b = 0, pq_b = 0, wcols = 0, nblk = 0, kp_b = 0;
};

template<typename T>
Layout<T> make_layout(const BatchShape &s) {
    Layout<T> L;
    size_t off = 0;
    auto take = [&](size_t elems) { size_t o = off; off += align_up(std::max<size_t>(elems, 1), 32); return o; };
    size_t w1 = 0, w2 = 0;
    switch (s.mix) {
        case CCC: w1 = (size_t) s.kA * s.kB; w2 = (size_t) s.kA * s.kB; break;
        case CCD: w1 = (size_t) s.kA * s.kB; w2 = (size_t) s.kA * s.n; break;
        case CDD: w1 = (size_t) s.kA * s.n; break;
        case DCD: w1 = (size_t) s.m * s.kB; break;
        case DDC: w1 = (size_t) s.m * s.n; break;
        default: break;
    }
    L.o_w1 = take(w1);
    L.o_w2 = take(w2);
    const bool recomp = (s.mix == CCC || s.mix == CDC || s.mix == DCC);
    if (recomp) {
        const int kp = (s.mix == DCC) ? s.kB : s.kA;
        L.r_b = s.r_force > 0 ? s.r_force : s.kC + kp;
        const int p_b = std::min(s.m, L.r_b), q_b = std::min(s.n, L.r_b);
        L.pq_b = std::max(p_b, q_b);
        const size_t sq = (size_t) L.pq_b * L.pq_b;
        L.o_uw = take((size_t) s.m * L.r_b);
        L.o_vw = take((size_t) s.n * L.r_b);
        L.o_tauu = take(L.r_b);
        L.o_tauv = take(L.r_b);
        L.o_m = take(sq);
        L.o_j = take(sq + L.pq_b);  // rotated copy of the core with an even leading dimension
        L.o_us = take(sq + L.pq_b);  // Us with an even leading dimension
        L.o_vs = take(sq);
        L.o_sig = take(L.pq_b);
        L.o_vn = take((size_t) s.n * std::min(L.pq_b, std::max(s.maxrankC, 1)));
        // blocked QR: clean reflector panels, T blocks, GEMM temporaries
        L.nblk = cdiv(L.pq_b, NBQ);
        L.wcols = std::max(L.r_b, std::max(s.maxrankC, 1));
        L.o_vcu = take((size_t) s.m * L.r_b);
        L.o_vcv = take((size_t) s.n * L.r_b);
        L.o_tbu = take((size_t) NBQ * NBQ * L.nblk);
        L.o_tbv = take((size_t) NBQ * NBQ * L.nblk);
        L.o_wbu = take((size_t) 2 * NBQ * L.wcols);
        L.o_wbv = take((size_t) 2 * NBQ * L.wcols);
        // LQ preconditioning of the core
        L.o_mt = take((size_t) (L.pq_b + 1) * L.r_b);  // M^T for the LQ path, or RU (p x r, even ld) for the GEMM core
        L.o_taum = take(L.pq_b);
        L.o_lb = take((size_t) (L.pq_b + 1) * L.r_b);  // L for the LQ path, or RV (q x r, even ld)
        L.o_vcm = take(sq);
        L.o_tbm = take((size_t) NBQ * NBQ * L.nblk);
        L.o_wbm = take((size_t) 2 * NBQ * L.wcols);
        L.o_pj = take((size_t) s.kA * s.kA);
        L.o_pos = take((size_t) L.r_b);  // r_b ints in T-sized slots
        // incremental U side: Gram-Schmidt coefficients, [clean reflectors | explicit Q2] of the new columns, rebuilt CU
        L.kp_b = kp;
        L.o_gu = take((size_t) s.kC * kp);
        L.o_gu2 = take((size_t) s.kC * kp);
        L.o_q2 = take((size_t) 2 * s.m * kp);
        L.o_tu = take((size_t) s.m * std::min(L.pq_b, std::max(s.maxrankC, 1)));
        // incremental V side
        L.o_sig0 = take(s.kC);
        L.o_hv = take((size_t) s.kC * kp);
        L.o_zv = take((size_t) s.kC * kp);
        L.o_q2v = take((size_t) 2 * s.n * kp);
        L.o_q2vt = take((size_t) s.n * kp);
        L.o_rvp = take((size_t) L.r_b * L.r_b);
        L.o_t1 = take((size_t) L.r_b * L.r_b);
        L.o_xu = take((size_t) L.r_b * kp);
        L.o_xv = take((size_t) L.r_b * kp);
        L.o_rn = take((size_t) L.r_b * kp);
        L.o_t1b = take((size_t) L.r_b * kp);
    }
    L.slab = off;
    return L;
}

template<typename T>
struct DescArrays {  // device arrays living at the front of the scratch arena
    size_t bytes = 0;
    size_t o_g1, o_g2, o_g3, o_gv, o_gc, o_cp, o_qr, o_rf, o_svd, o_rc, o_rk, o_tiles;
    size_t o_bqr = 0, o_blf = 0, o_bgw = 0, o_bgw2 = 0, o_bgup = 0, o_agw = 0, o_agw2 = 0, o_agup = 0;
    size_t o_pds = 0, o_pdc = 0, o_qrc = 0, o_lq = 0, o_pc = 0, o_asj = 0, o_gi = 0, o_isj = 0, o_gru = 0, o_giv = 0, o_pvc = 0, o_bqr2 = 0;
    size_t o_cqf = 0;
    int nst_inc = 0;
    explicit DescArrays(int n, int nblk = 0, int qr_cols = 0, int rk_bound = 0, int kp_b = 0) {
        size_t off = 0;
        auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return o; };
        o_g1 = take(sizeof(GemmProb<T>) * n);
        o_g2 = take(sizeof(GemmProb<T>) * n);
        o_g3 = take(sizeof(GemmProb<T>) * n);
        o_gv = take(sizeof(GemmProb<T>) * n);
        o_gc = take(sizeof(GemmProb<T>) * n);
        o_cp = take(sizeof(CopyProb<T>) * 4 * n);
        o_qr = take(sizeof(QrProb<T>) * 2 * n);
        o_rf = take(sizeof(ReflProb<T>) * 2 * n);
        o_svd = take(sizeof(SvdProb<T>) * n);
        o_rc = take(sizeof(RecompProb<T>) * n);
        o_rk = take(sizeof(int) * n);
        o_tiles = take(sizeof(hcb_tile) * 3 * n);
        o_pds = take(sizeof(PanelDesc<T>) * 4 * n);  // U stack, V stack per tile, then the incremental P and Y panels
        o_pdc = take(sizeof(PanelDesc<T>) * n);
        o_gi = take(sizeof(GemmProb<T>) * 4 * n);
        o_giv = take(sizeof(GemmProb<T>) * 6 * n);   // 4 Gram-Schmidt GEMMs + T1 + X of the incremental V side
        o_gru = take(sizeof(GemmProb<T>) * 3 * n);
        o_pvc = take(sizeof(PanelDesc<T>) * n);
        o_cqf = take(sizeof(int) * 2 * n);          // CholeskyQR2 of the new-column panels: per-side fallback flags
        nst_inc = std::max(1, cdiv(std::max(kp_b, 1), NBQ));
        o_isj = take(sizeof(StripJob) * (size_t) nst_inc * 2 * n);
        o_qrc = take(sizeof(QrProb<T>) * n);
        o_lq = take(sizeof(LqProb<T>) * n);
        o_pc = take(sizeof(PrecondProb<T>) * n);
        if (nblk > 0) {
            const size_t nb = (size_t) nblk * 2 * n;
            o_bqr = take(qr_block_desc_bytes<T>(4 * n, qr_cols));
            o_bqr2 = take(qr_block_desc_bytes<T>(n, qr_cols));   // second pass: the r x r panels of the incremental V side
            o_asj = take(sizeof(StripJob) * (size_t) cdiv(std::max(rk_bound, 1), NBQ) * 2 * n);
            o_agw = take(sizeof(GemmProb<T>) * nb);
            o_agw2 = take(sizeof(GemmProb<T>) * nb);
            o_agup = take(sizeof(GemmProb<T>) * nb);
        }
        bytes = off;
    }
};

static int classify(const hcb_tile *A, const hcb_tile *B, const hcb_tile *C, int64_t n, int opA, int opB, BatchShape &s) {
    if (n <= 0) return HCB_OK;
    const int ta = A[0].type, tb = B[0].type, tc = C[0].type;
    s.mix = (ta == HCB_TILE_COMPRESSED ? 4 : 0) | (tb == HCB_TILE_COMPRESSED ? 2 : 0) | (tc == HCB_TILE_COMPRESSED ? 1 : 0);
    for (int64_t t = 0; t < n; ++t) {
        if (A[t].type != ta || B[t].type != tb || C[t].type != tc)
            return fail(HCB_EINVAL, "tlr_gemm_batched: the batch must be homogeneous in its Dense/Compressed mix");
        const int am = opA ? A[t].n : A[t].m, ak = opA ? A[t].m : A[t].n;
        const int bk = opB ? B[t].n : B[t].m, bn = opB ? B[t].m : B[t].n;
        if (am != C[t].m || bn != C[t].n || ak != bk)
            return fail(HCB_EINVAL, "tlr_gemm_batched: op(A) op(B) does not conform with C");
        for (const hcb_tile *x : {&A[t], &B[t], &C[t]}) {
            if (!x->d_data) return fail(HCB_EINVAL, "tlr_gemm_batched: null tile buffer");
            if (x->type == HCB_TILE_COMPRESSED && (!x->d_rank || x->max_rank < 1))
                return fail(HCB_EINVAL, "tlr_gemm_batched: compressed tile needs d_rank and max_rank >= 1");
            if (x->type == HCB_TILE_DENSE && x->ld < x->m) return fail(HCB_EINVAL, "tlr_gemm_batched: dense ld < m");
        }
        if (s.mix == DDC && C[t].max_rank < std::min(C[t].m, C[t].n))
            return fail(HCB_EINVAL, "tlr_gemm_batched: Dense*Dense -> Compressed makes C full rank "
                                    "(HCore.cpp:291-298): max_rank must be >= min(m, n)");
        s.m = std::max(s.m, C[t].m);
        s.n = std::max(s.n, C[t].n);
        s.k = std::max(s.k, ak);
        s.kA = std::max(s.kA, bound_of(A[t]));
        s.kB = std::max(s.kB, bound_of(B[t]));
        s.kC = std::max(s.kC, bound_of(C[t]));
        s.maxrankC = std::max(s.maxrankC, C[t].max_rank);
    }
    return HCB_OK;
}

int tlr_gemm_promoted(hcb_ctx *ctx, int n, const hcb_tile *A, int opA, const hcb_tile *B, int opB, const hcb_tile *C,
                      float alpha, float beta, const hcb_compress_params *prm, int32_t *d_info, bool reset_info);

template<typename T>
int t_tlr_gemm_batched(hcb_ctx *ctx, int64_t n64, const hcb_tile *A, int opA, const hcb_tile *B, int opB,
                       const hcb_tile *C, T alpha, T beta, const hcb_compress_params *prm, int32_t *d_info,
                       bool reset_info = true) {
    HCB_TRY(check_ctx(ctx));
    if (n64 <= 0) return HCB_OK;
    if (!A || !B || !C || !prm) return fail(HCB_EINVAL, "tlr_gemm_batched: null argument");
    if (n64 > (1 << 24)) return fail(HCB_EINVAL, "tlr_gemm_batched: batch too large");
    const int n = (int) n64;
    BatchShape s;
    HCB_TRY(classify(A, B, C, n, opA, opB, s));
    if constexpr (std::is_same<T, float>::value) {
        // FP32 tiles whose recompression is big enough for the blocked path run on the FP64 machinery (DMMA GEMMs, strip
        // reflectors, register Jacobi, incremental recompression): the tiles are converted to FP64 shadows, the FP64 path
        // runs on them (ranks / state words are shared), C is converted back.  The native FP32 kernels (SIMT FFMA GEMM,
        // GEMM-blocked QR, shared-memory Jacobi) stay for small tiles and under HCB_FP32_NATIVE=1.
        const bool native = getenv("HCB_FP32_NATIVE") && atoi(getenv("HCB_FP32_NATIVE")) != 0;  // (read per call: tests toggle it)
        const bool recomp = (s.mix == CCC || s.mix == CDC || s.mix == DCC);
        const int kp = (s.mix == DCC) ? s.kB : s.kA;
        if (!native && recomp && s.kC + kp > 2 * NBQ && strip_path_ok(ctx, std::max(s.m, s.n)))
            return tlr_gemm_promoted(ctx, n, A, opA, B, opB, C, alpha, beta, prm, d_info, reset_info);
    }
    const Layout<T> L = make_layout<T>(s);
    const bool blocked = L.r_b > 2 * NBQ;  // compact-WY path once the stacked rank spans more than two blocks
    const int rk_bound = std::max(1, std::min(L.pq_b, s.maxrankC));
    const DescArrays<T> D(n, blocked ? L.nblk : 0, std::max(L.r_b, L.pq_b), rk_bound, L.kp_b);
    // incremental U side (tiles whose state says "U orthonormal"): needs the blocked fp64 strip machinery
    const bool inc_enabled = blocked && std::is_same<T, double>::value && strip_path_ok(ctx, std::max(s.m, s.n)) &&
                             !getenv("HCB_NO_INCREMENTAL");
    const size_t total = D.bytes + L.slab * sizeof(T) * (size_t) n + 256;
    HCB_TRY(ensure_ws(ctx, total));
    char *base = reinterpret_cast<char *>(ctx->ws);

    // descriptors: host arrays -> pinned ring -> device
    std::vector<hcb_tile> packed(3 * (size_t) n);
    std::copy(A, A + n, packed.begin());
    std::copy(B, B + n, packed.begin() + n);
    std::copy(C, C + n, packed.begin() + 2 * (size_t) n);
    hcb_tile *d_tiles = reinterpret_cast<hcb_tile *>(base + D.o_tiles);
    {
        void *staged = nullptr;
        HCB_TRY(ring_upload(ctx, packed.data(), packed.size() * sizeof(hcb_tile), &staged));
        HCB_CUDA(cudaMemcpyAsync(d_tiles, staged, packed.size() * sizeof(hcb_tile), cudaMemcpyDeviceToDevice, ctx->stream));
    }

    SetupArgs<T> sa;
    sa.A = d_tiles; sa.B = d_tiles + n; sa.C = d_tiles + 2 * (size_t) n;
    sa.n_tiles = n; sa.mix = s.mix; sa.opA = opA ? 1 : 0; sa.opB = opB ? 1 : 0;
    sa.alpha = alpha; sa.beta = beta;
    sa.ws = reinterpret_cast<T *>(base + align_up(D.bytes, 256));
    sa.slab = L.slab; sa.o_w1 = L.o_w1; sa.o_w2 = L.o_w2; sa.o_uw = L.o_uw; sa.o_vw = L.o_vw;
    sa.o_tauu = L.o_tauu; sa.o_tauv = L.o_tauv; sa.o_m = L.o_m; sa.o_j = L.o_j; sa.o_us = L.o_us; sa.o_vs = L.o_vs;
    sa.o_sig = L.o_sig; sa.o_vn = L.o_vn;
    sa.o_vcu = L.o_vcu; sa.o_vcv = L.o_vcv; sa.o_tbu = L.o_tbu; sa.o_tbv = L.o_tbv; sa.o_wbu = L.o_wbu; sa.o_wbv = L.o_wbv;
    sa.wcols = L.wcols;
    sa.o_mt = L.o_mt; sa.o_taum = L.o_taum; sa.o_lb = L.o_lb; sa.o_vcm = L.o_vcm; sa.o_tbm = L.o_tbm; sa.o_wbm = L.o_wbm;
    sa.pd_stack = reinterpret_cast<PanelDesc<T> *>(base + D.o_pds);
    sa.pd_core = reinterpret_cast<PanelDesc<T> *>(base + D.o_pdc);
    sa.qr_core = reinterpret_cast<QrProb<T> *>(base + D.o_qrc);
    sa.lq = reinterpret_cast<LqProb<T> *>(base + D.o_lq);
    sa.pc = reinterpret_cast<PrecondProb<T> *>(base + D.o_pc);
    sa.o_pj = L.o_pj; sa.o_pos = L.o_pos;
    sa.use_lq = 0;  // sorted + preconditioned stacks: Jacobi converges in ~6 sweeps on the core itself
    sa.kA_b = s.kA; sa.kB_b = s.kB; sa.kC_b = s.kC; sa.r_b = L.r_b;
    sa.rk_new = reinterpret_cast<int *>(base + D.o_rk);
    sa.info = d_info;
    sa.g1 = reinterpret_cast<GemmProb<T> *>(base + D.o_g1);
    sa.g2 = reinterpret_cast<GemmProb<T> *>(base + D.o_g2);
    sa.g3 = reinterpret_cast<GemmProb<T> *>(base + D.o_g3);
    sa.gv = reinterpret_cast<GemmProb<T> *>(base + D.o_gv);
    sa.gc = reinterpret_cast<GemmProb<T> *>(base + D.o_gc);
    sa.cp = reinterpret_cast<CopyProb<T> *>(base + D.o_cp);
    sa.qr = reinterpret_cast<QrProb<T> *>(base + D.o_qr);
    sa.rf = reinterpret_cast<ReflProb<T> *>(base + D.o_rf);
    sa.svd = reinterpret_cast<SvdProb<T> *>(base + D.o_svd);
    sa.rc = reinterpret_cast<RecompProb<T> *>(base + D.o_rc);
    sa.inc_enabled = inc_enabled ? 1 : 0;
    sa.o_gu = L.o_gu; sa.o_gu2 = L.o_gu2; sa.o_q2 = L.o_q2; sa.o_tu = L.o_tu;
    sa.gi = reinterpret_cast<GemmProb<T> *>(base + D.o_gi);
    sa.pd_inc = sa.pd_stack + 2 * (size_t) n;
    sa.inc_sj = reinterpret_cast<StripJob *>(base + D.o_isj);
    sa.nst_inc = D.nst_inc; sa.kp_b = L.kp_b;
    sa.inc_refresh = getenv("HCB_INC_REFRESH") ? std::max(1, atoi(getenv("HCB_INC_REFRESH"))) : 16;
    // incremental V side: pays when the tile is much taller than the stacked rank (the R-only QR runs on r x r instead of
    // n x r); HCB_VINC=0/1 overrides the default (on)
    const bool vinc_enabled = inc_enabled && !(getenv("HCB_VINC") && atoi(getenv("HCB_VINC")) == 0);
    sa.vinc_enabled = vinc_enabled ? 1 : 0;
    sa.o_sig0 = L.o_sig0; sa.o_hv = L.o_hv; sa.o_zv = L.o_zv; sa.o_q2v = L.o_q2v; sa.o_q2vt = L.o_q2vt; sa.o_rvp = L.o_rvp;
    sa.o_t1 = L.o_t1;
    sa.o_xu = L.o_xu; sa.o_xv = L.o_xv; sa.o_rn = L.o_rn; sa.o_t1b = L.o_t1b;
    sa.giv = reinterpret_cast<GemmProb<T> *>(base + D.o_giv);
    sa.gt1 = sa.giv + 4 * (size_t) n;
    sa.gx = sa.giv + 5 * (size_t) n;
    sa.pd_vcore = reinterpret_cast<PanelDesc<T> *>(base + D.o_pvc);
    // CholeskyQR2 of the new-column panels (fp64 incremental path, kp <= 64): HCB_NO_CHOLQR=1 keeps the Householder panels
    const bool cq_enabled = inc_enabled && L.kp_b <= CQ_KP && !(getenv("HCB_NO_CHOLQR") && atoi(getenv("HCB_NO_CHOLQR")) != 0);
    sa.cq_fail = reinterpret_cast<int *>(base + D.o_cqf);
    sa.cq_enabled = cq_enabled ? 1 : 0;
    sa.stats = ctx->d_stats;
    sa.err_flag = ctx->d_err;
    if (d_info && reset_info) HCB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t) * (size_t) n, ctx->stream));
    {
        PhaseScope ph(ctx, 0);
        k_setup_tlr<T><<<cdiv(n, 128), 128, 0, ctx->stream>>>(sa);
        HCB_LAUNCH_CHECK("k_setup_tlr");
    }

    // contraction phases (grids sized from rank BOUNDS; kernels read the true shapes from the descriptors)
    const int kmaxAB = std::max(s.kA, s.kB);
    int g1m = s.m, g1n = s.n, g2m = s.m, g2n = s.n;
    switch (s.mix) {
        case CCC: g1m = s.kA; g1n = s.kB; g2m = s.n; g2n = s.kA; break;
        case CCD: g1m = s.kA; g1n = s.kB; g2m = s.kA; g2n = s.n; break;
        case CDD: g1m = s.kA; g1n = s.n; break;
        case DCD: g1m = s.m; g1n = s.kB; break;
        case CDC: g1m = s.n; g1n = s.kA; break;
        case DCC: g1m = s.m; g1n = s.kB; break;
        default: break;
    }
    (void) kmaxAB;
    {
        PhaseScope ph(ctx, 1);
        HCB_TRY(launch_gemm<T>(ctx, sa.g1, n, g1m, g1n));
        if (s.mix == CCC) {
            // orthogonalise the product term's right factor (small Jacobi in shared memory, one CTA per tile)
            size_t want = ((size_t) s.kA * s.kB + (size_t) s.kA * s.kA + (size_t) s.kB) * sizeof(T);
            const size_t cap = ctx->smem_optin > 2048 ? ctx->smem_optin - 1024 : 0;
            if (want > cap) want = 0;  // does not fit: the kernel writes J = I
            want = align_up(want, 16);
            HCB_CUDA(cudaFuncSetAttribute(k_precond_product<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(want, 16)));
            // one warp per column pair of the round-robin (ka / 2 pairs per round)
            const int pc_threads = std::min(1024, std::max(256, 32 * ((s.kA + 3) / 4)));  // one warp per TWO column pairs: 2+ CTAs per SM, the 256 tiles of a batch run as one wave
            static const int pc_sweeps = getenv("HCB_PRECOND_SWEEPS") ? std::max(0, atoi(getenv("HCB_PRECOND_SWEEPS"))) : 8;
            k_precond_product<T><<<n, pc_threads, want, ctx->stream>>>(sa.pc, (int) (want / sizeof(T)), pc_sweeps);
            HCB_LAUNCH_CHECK("k_precond_product");
        }
        if (s.mix == CCC || s.mix == CCD || s.mix == CDD || s.mix == DCD || s.mix == DDC)
            HCB_TRY(launch_gemm<T>(ctx, sa.g2, n, g2m, g2n));
        if (s.mix == CCD) HCB_TRY(launch_gemm<T>(ctx, sa.g3, n, s.m, s.n));
        if (s.mix == CCC) HCB_TRY(launch_gemm<T>(ctx, sa.g3, n, s.m, s.kA));
    }

    if (s.mix == DDC) {
        dim3 grid(std::max(1, std::min(256, cdiv((long long) s.m * s.n, 256))), n);
        k_ddc_finalize<T><<<grid, 256, 0, ctx->stream>>>(sa.C, sa.ws, L.slab, L.o_w1, d_info);
        HCB_LAUNCH_CHECK("k_ddc_finalize");
        return HCB_OK;
    }
    if (!(s.mix == CCC || s.mix == CDC || s.mix == DCC)) return HCB_OK;  // dense C: done

    // recompression (Compressed.cpp:332-682)
    {
        PhaseScope ph(ctx, 2);
        HCB_TRY(launch_copy<T>(ctx, sa.cp, 4 * n, std::max(s.m, s.n), std::max(L.r_b, 1)));
        if (vinc_enabled) {
            k_vinc_sig0<T><<<n, 256, 0, ctx->stream>>>(sa.rc);
            HCB_LAUNCH_CHECK("k_vinc_sig0");
        }
        // sort the stack columns by decreasing V-stack norm (same permutation on both stacks)
        const size_t want = align_up((size_t) std::max(L.r_b, 1) * sizeof(T), 16);
        k_stack_order<T><<<n, 256, want, ctx->stream>>>(sa.rc, (int) (want / sizeof(T)));
        HCB_LAUNCH_CHECK("k_stack_order");
        dim3 grid(std::max(1, std::min(L.r_b, 64)), 2 * n);
        k_permute_stacks<T><<<grid, 256, 0, ctx->stream>>>(sa.rc);
        HCB_LAUNCH_CHECK("k_permute_stacks");
    }
    const int npan = 2 * n, nbt = L.nblk * npan;
    char *blk_store = base + D.o_bqr;
    {
        PhaseScope ph(ctx, 3);
        if (!blocked) HCB_TRY(launch_qr<T>(ctx, sa.qr, 2 * n));
        else {
            // block classical Gram-Schmidt of the new U columns against the orthonormal CU and of the new V columns against
            // W = CV^T diag(sigma)^-1 (Zv = S^-2 (CV Y), Y -= CV^T Zv): first pass of both sides, then the device decides per
            // tile side whether the second pass is needed (k_inc_gate: "twice is enough" criterion), then the second pass
            // (its GEMMs are no-ops where the gate switched them off)
            static const bool gate = !(getenv("HCB_GS_ALWAYS_TWICE") && atoi(getenv("HCB_GS_ALWAYS_TWICE")) != 0);
            dim3 gs(std::max(1, std::min(16, cdiv((long long) s.kC * L.kp_b, 256))), n);
            if (inc_enabled) {
                HCB_TRY(launch_gemm<T>(ctx, sa.gi + 0 * (size_t) n, n, s.kC, L.kp_b));
                HCB_TRY(launch_gemm<T>(ctx, sa.gi + 1 * (size_t) n, n, s.m, L.kp_b));
            }
            if (vinc_enabled) {
                HCB_TRY(launch_gemm<T>(ctx, sa.giv + 0 * (size_t) n, n, s.kC, L.kp_b));
                k_vinc_scale<T><<<gs, 256, 0, ctx->stream>>>(sa.rc, 0);
                HCB_LAUNCH_CHECK("k_vinc_scale");
                HCB_TRY(launch_gemm<T>(ctx, sa.giv + 1 * (size_t) n, n, s.n, L.kp_b));
            }
            if (inc_enabled && gate) {
                // (HCB_GS_FORCE_ONCE=1, tests only: drop the second pass even where the criterion asks for it)
                const int force_once = getenv("HCB_GS_FORCE_ONCE") && atoi(getenv("HCB_GS_FORCE_ONCE")) != 0;
                k_inc_gate<T><<<2 * n, 256, 0, ctx->stream>>>(sa.rc, sa.gi, sa.giv, n, force_once);
                HCB_LAUNCH_CHECK("k_inc_gate");
            }
            if (inc_enabled) {
                HCB_TRY(launch_gemm<T>(ctx, sa.gi + 2 * (size_t) n, n, s.kC, L.kp_b));
                HCB_TRY(launch_gemm<T>(ctx, sa.gi + 3 * (size_t) n, n, s.m, L.kp_b));
            }
            if (vinc_enabled) {
                HCB_TRY(launch_gemm<T>(ctx, sa.giv + 2 * (size_t) n, n, s.kC, L.kp_b));
                k_vinc_scale<T><<<gs, 256, 0, ctx->stream>>>(sa.rc, 1);
                HCB_LAUNCH_CHECK("k_vinc_scale");
                HCB_TRY(launch_gemm<T>(ctx, sa.giv + 3 * (size_t) n, n, s.n, L.kp_b));
            }
            if (cq_enabled) {
                if constexpr (std::is_same<T, double>::value) {
                    // CholeskyQR2 of the new columns (k_cholqr_pass, one fused kernel per pass): where it holds, Q2 / R2 are in
                    // place afterwards and the Householder panel + explicit-Q strips below find their descriptors switched off
                    for (int pass = 0; pass < 2; ++pass) {
                        k_cholqr_pass<T><<<2 * n, 256, 0, ctx->stream>>>(sa.rc, sa.pd_inc, sa.inc_sj, D.nst_inc, n, pass);
                        HCB_LAUNCH_CHECK("k_cholqr_pass");
                    }
                }
            }
            // one pass over 4n panels: the two stacks of every tile (inactive where incremental) + the P and Y panels
            HCB_TRY(run_blocked_qr<T>(ctx, sa.pd_stack, inc_enabled ? 4 * n : npan, std::max(s.m, s.n), L.r_b, blk_store,
                                      inc_enabled ? L.kp_b : 0, inc_enabled ? npan : 0));
            if (inc_enabled) {
                dim3 ge(std::max(1, std::min(64, cdiv((long long) std::max(s.m, s.n) * L.kp_b, 2048))), 2 * n);
                k_inc_eye<T><<<ge, 256, 0, ctx->stream>>>(sa.rc);
                HCB_LAUNCH_CHECK("k_inc_eye");
                HCB_TRY(launch_strips(ctx, sa.inc_sj, D.nst_inc * 2 * n, std::max(s.m, s.n)));   // explicit Q2 (U and V side)
            }
            if (vinc_enabled) {
                // R-only Householder QR of the r x r matrices RV * Pi (second pass: it needs R2v from the Y panels)
                dim3 gb(std::max(1, std::min(64, cdiv((long long) L.r_b * L.r_b, 1024))), n);
                k_vinc_build_rv<T><<<gb, 256, 0, ctx->stream>>>(sa.rc);
                HCB_LAUNCH_CHECK("k_vinc_build_rv");
                if constexpr (std::is_same<T, double>::value) {
                    // the graded factor R' by a Cholesky factorisation of the assembled, column-scaled Gram matrix (k_vcore_chol);
                    // tiles that fail its pivot test keep their Householder panel descriptor and are factored below
                    const bool vchol = !(getenv("HCB_NO_VCHOL") && atoi(getenv("HCB_NO_VCHOL")) != 0);   // (read per call: tests toggle it)
                    const size_t vsm = vcore_chol_smem(L.r_b);
                    if (vchol && L.kp_b <= CQ_KP && vsm + 1024 <= ctx->smem_optin) {
                        HCB_CUDA(cudaFuncSetAttribute(k_vcore_chol<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) vsm));
                        k_vcore_chol<T><<<n, 256, vsm, ctx->stream>>>(sa.rc, sa.pd_vcore, L.r_b);
                        HCB_LAUNCH_CHECK("k_vcore_chol");
                    }
                }
                HCB_TRY(run_blocked_qr<T>(ctx, sa.pd_vcore, n, L.r_b, L.r_b, base + D.o_bqr2));
            }
        }
    }
    {
        PhaseScope ph(ctx, 4);
        dim3 grid(std::max(1, std::min(64, cdiv((long long) L.pq_b * L.pq_b, 256))), n);
        if (!sa.use_lq) {
            // triangles out of the QR'd stacks, then core = RU RV^T as one batched (DMMA) GEMM
            dim3 gx(std::max(1, std::min(128, cdiv((long long) 2 * L.pq_b * L.r_b, 256))), n);
            k_extract_r<T><<<gx, 256, 0, ctx->stream>>>(sa.rc);
            HCB_LAUNCH_CHECK("k_extract_r");
            if (vinc_enabled) {  // tiles with both sides incremental: gather part of the core (the GEMM below adds the rest)
                dim3 gk(std::max(1, std::min(64, cdiv(L.pq_b, 32) * cdiv(L.pq_b, 32))), n);
                k_both_core<T><<<gk, dim3(32, 8), 0, ctx->stream>>>(sa.rc);
                HCB_LAUNCH_CHECK("k_both_core");
            }
            HCB_TRY(launch_gemm<T>(ctx, sa.gc, n, L.pq_b, L.pq_b));
        } else {
            k_core_build<T><<<grid, 256, 0, ctx->stream>>>(sa.rc);
            HCB_LAUNCH_CHECK("k_core_build");
        }
        if (sa.use_lq) {
            // LQ preconditioning: QR of the transposed core, L = R^T goes to the Jacobi kernel
            if (!blocked) HCB_TRY(launch_qr<T>(ctx, sa.qr_core, n));
            else HCB_TRY(run_blocked_qr<T>(ctx, sa.pd_core, n, L.pq_b, L.pq_b, blk_store));
            k_extract_l<T><<<grid, 256, 0, ctx->stream>>>(sa.lq);
            HCB_LAUNCH_CHECK("k_extract_l");
        }
    }
    {
        PhaseScope ph(ctx, 7);
        // accuracy-aware stop (default since round 2: full GPU suite green, 442 -> 426 ms per step; HCB_JACOBI_ACC_STOP=0
        // restores machine precision): the sweeps end once the remaining non-orthogonality is far below the compression
        // accuracy
        static const bool acc_stop = !(getenv("HCB_JACOBI_ACC_STOP") && atoi(getenv("HCB_JACOBI_ACC_STOP")) == 0);
        HCB_TRY(launch_svd<T>(ctx, sa.svd, n, L.pq_b, L.pq_b, acc_stop ? prm->accuracy : 0.0));
    }
    {
        PhaseScope ph(ctx, 8);
        HCB_TRY(launch_gemm<T>(ctx, sa.gv, n, L.pq_b, L.pq_b));
        if (vinc_enabled) {  // V S' = (RV Pi) ((RU Pi)^T Us) for the incremental tiles
            HCB_TRY(launch_gemm<T>(ctx, sa.gt1, n, L.pq_b, L.pq_b));
            HCB_TRY(launch_gemm<T>(ctx, sa.gx, n, L.pq_b, L.pq_b));
        }
        k_truncate<T><<<n, 256, 0, ctx->stream>>>(sa.rc, (T) prm->accuracy, prm->truncated_svd, (int) prm->fixed_rank);
        HCB_LAUNCH_CHECK("k_truncate");
    }
    if (!blocked) {
        PhaseScope ph(ctx, 5);
        HCB_TRY(launch_refl<T>(ctx, sa.rf, 2 * n, rk_bound));
    } else if (std::is_same<T, double>::value && strip_path_ok(ctx, std::max(s.m, s.n))) {
        // rebuild C := Q [X;0], strip-resident: every 32-column strip of C takes all reflector blocks in one kernel
        PhaseScope ph(ctx, 5);
        StripJob *sj = reinterpret_cast<StripJob *>(base + D.o_asj);
        const int nstrips = cdiv(rk_bound, NBQ);
        GemmProb<T> *gru = reinterpret_cast<GemmProb<T> *>(base + D.o_gru);
        k_setup_apply_strips<T><<<cdiv(nstrips * npan, 128), 128, 0, ctx->stream>>>(sa.rc, sj, nstrips, npan, gru);
        HCB_LAUNCH_CHECK("k_setup_apply_strips");
        HCB_TRY(launch_strips(ctx, sj, nstrips * npan, std::max(s.m, s.n)));
        if (inc_enabled) {  // CU' = [CU | Q2] * Us for the incremental tiles (two GEMMs into TU)
            HCB_TRY(launch_gemm<T>(ctx, gru, n, s.m, rk_bound));
            if (!std::is_same<T, double>::value) HCB_TRY(launch_gemm<T>(ctx, gru + n, n, s.m, rk_bound));
            if (vinc_enabled) {  // VN = [CV ; Q2v^T]^T * (S^-1-scaled V S')
                dim3 gt(std::max(1, std::min(64, cdiv(s.n, 32) * cdiv(L.kp_b, 32))), n);
                k_vinc_transpose_q2<T><<<gt, dim3(32, 8), 0, ctx->stream>>>(sa.rc);
                HCB_LAUNCH_CHECK("k_vinc_transpose_q2");
                HCB_TRY(launch_gemm<T>(ctx, gru + 2 * (size_t) n, n, s.n, rk_bound));
            }
        }
    } else {
        // blocked rebuild C := Q [X;0]: blocks last-to-first, three batched GEMMs per block, rank read on the device
        PhaseScope ph(ctx, 5);
        ApplyBlockArrays<T> aa{reinterpret_cast<GemmProb<T> *>(base + D.o_agw), reinterpret_cast<GemmProb<T> *>(base + D.o_agw2),
                               reinterpret_cast<GemmProb<T> *>(base + D.o_agup), L.nblk, npan};
        k_setup_apply_blocks<T><<<cdiv(nbt, 128), 128, 0, ctx->stream>>>(sa.rc, aa);
        HCB_LAUNCH_CHECK("k_setup_apply_blocks");
        const int mx = std::max(s.m, s.n);
        for (int b = L.nblk - 1; b >= 0; --b) {
            const size_t o = (size_t) b * npan;
            HCB_TRY(launch_gemm<T>(ctx, aa.gw + o, npan, NBQ, rk_bound));
            HCB_TRY(launch_gemm<T>(ctx, aa.gw2 + o, npan, NBQ, rk_bound));
            HCB_TRY(launch_gemm<T>(ctx, aa.gup + o, npan, mx - b * NBQ, rk_bound));
        }
    }
    {
        PhaseScope ph(ctx, 6);
        dim3 grid(std::max(1, std::min(256, cdiv(rk_bound, 32) * cdiv(s.n, 32))), n);
        k_finalize<T><<<grid, dim3(32, 8), 0, ctx->stream>>>(sa.rc);
        HCB_LAUNCH_CHECK("k_finalize");
    }
    return HCB_OK;
}

// info words of a partitioned (mixed) batch back to the caller's order: dst[perm[i]] = src[i]
__global__ void k_scatter_info(const int32_t *__restrict__ src, const int *__restrict__ perm, int32_t *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}

// Any batch: triples of DIFFERENT Dense/Compressed mixes (a Cholesky trailing update, a matrix with dense diagonal tiles)
// are partitioned by mix on the host -- stable, so equal-mix triples keep their order -- and every group runs as one
// homogeneous fused call; d_info comes back in the caller's order.  Homogeneous batches (the common case) pass straight
// through with no copy.
template<typename T>
int t_tlr_gemm_any(hcb_ctx *ctx, int64_t n64, const hcb_tile *A, int opA, const hcb_tile *B, int opB, const hcb_tile *C,
                   T alpha, T beta, const hcb_compress_params *prm, int32_t *d_info) {
    HCB_TRY(check_ctx(ctx));
    if (n64 <= 0) return HCB_OK;
    if (!A || !B || !C || !prm) return fail(HCB_EINVAL, "tlr_gemm_batched: null argument");
    if (n64 > (1 << 24)) return fail(HCB_EINVAL, "tlr_gemm_batched: batch too large");
    const int n = (int) n64;
    auto mix_of = [&](int t) {
        return (A[t].type == HCB_TILE_COMPRESSED ? 4 : 0) | (B[t].type == HCB_TILE_COMPRESSED ? 2 : 0) |
               (C[t].type == HCB_TILE_COMPRESSED ? 1 : 0);
    };
    bool same = true;
    for (int t = 1; t < n && same; ++t) same = mix_of(t) == mix_of(0);
    if (same) return t_tlr_gemm_batched<T>(ctx, n, A, opA, B, opB, C, alpha, beta, prm, d_info);
    std::vector<int> perm;
    perm.reserve(n);
    std::vector<hcb_tile> a, b, c;
    a.reserve(n); b.reserve(n); c.reserve(n);
    int group_begin[9] = {0};
    for (int mix = 0; mix < 8; ++mix) {
        group_begin[mix] = (int) perm.size();
        for (int t = 0; t < n; ++t)
            if (mix_of(t) == mix) { perm.push_back(t); a.push_back(A[t]); b.push_back(B[t]); c.push_back(C[t]); }
    }
    group_begin[8] = n;
    int32_t *tmp = nullptr;
    int *d_perm = nullptr;
    if (d_info) {  // group-ordered info words + the permutation, in the promotion arena's neighbour (grow-only)
        const size_t need = (size_t) n * (sizeof(int32_t) + sizeof(int)) + 256;
        if (need > ctx->info_tmp_bytes) {
            HCB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (ctx->info_tmp) HCB_CUDA(cudaFree(ctx->info_tmp));
            ctx->info_tmp = nullptr; ctx->info_tmp_bytes = 0;
            HCB_CUDA(cudaMalloc(&ctx->info_tmp, need * 2));
            ctx->info_tmp_bytes = need * 2;
        }
        tmp = reinterpret_cast<int32_t *>(ctx->info_tmp);
        d_perm = reinterpret_cast<int *>(tmp + n);
        void *st = nullptr;
        HCB_TRY(ring_upload(ctx, perm.data(), sizeof(int) * (size_t) n, &st));
        HCB_CUDA(cudaMemcpyAsync(d_perm, st, sizeof(int) * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    for (int mix = 0; mix < 8; ++mix) {
        const int g0 = group_begin[mix], cnt = group_begin[mix + 1] - g0;
        if (cnt <= 0) continue;
        HCB_TRY(t_tlr_gemm_batched<T>(ctx, cnt, a.data() + g0, opA, b.data() + g0, opB, c.data() + g0, alpha, beta, prm,
                                      tmp ? tmp + g0 : nullptr));
    }
    if (d_info) {
        k_scatter_info<<<cdiv(n, 256), 256, 0, ctx->stream>>>(tmp, d_perm, d_info, n);
        HCB_LAUNCH_CHECK("k_scatter_info");
    }
    return HCB_OK;
}

// FP32 batch on the FP64 path (see t_tlr_gemm_batched<float>).  Every distinct tile buffer of the batch gets an FP64
// shadow in the context's second arena (a tile that appears in many triples -- A(j, k) in every C(j, :) -- is converted
// once), element for element over its whole capacity, so offsets (V at m * max_rank, ld = rank) carry over unchanged.
int tlr_gemm_promoted(hcb_ctx *ctx, int n, const hcb_tile *A, int opA, const hcb_tile *B, int opB, const hcb_tile *C,
                      float alpha, float beta, const hcb_compress_params *prm, int32_t *d_info, bool reset_info) {
    auto elems = [](const hcb_tile &t) -> size_t {
        return t.type == HCB_TILE_COMPRESSED ? ((size_t) t.m + (size_t) t.n) * (size_t) t.max_rank : (size_t) t.ld * (size_t) t.n;
    };
    std::vector<hcb_tile> sh(3 * (size_t) n);
    std::vector<ConvProb<float, double>> up;
    std::vector<ConvProb<double, float>> down;
    std::unordered_map<const void *, size_t> where;  // tile buffer -> shadow offset (in doubles)
    size_t off = 0;
    for (int side = 0; side < 3; ++side) {
        const hcb_tile *src = side == 0 ? A : (side == 1 ? B : C);
        for (int t = 0; t < n; ++t) {
            const hcb_tile &x = src[t];
            auto it = where.find(x.d_data);
            size_t o;
            if (it == where.end()) {
                o = off;
                where.emplace(x.d_data, o);
                off += align_up(elems(x), 32);
                up.push_back({reinterpret_cast<const float *>(x.d_data), nullptr, elems(x)});
                up.back().dst = reinterpret_cast<double *>(o);  // patched with the arena base below
                if (side == 2) down.push_back({reinterpret_cast<const double *>(o), reinterpret_cast<float *>(x.d_data), elems(x)});
            } else {
                o = it->second;
                if (side == 2) return fail(HCB_EINVAL, "tlr_gemm_batched: a C tile appears twice in the batch (or aliases an operand)");
            }
            sh[(size_t) side * n + t] = x;
            sh[(size_t) side * n + t].d_data = reinterpret_cast<void *>(o);
        }
    }
    const size_t bytes = off * sizeof(double) + 256;
    if (bytes > ctx->ws2_bytes) {
        HCB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->ws2) HCB_CUDA(cudaFree(ctx->ws2));
        ctx->ws2 = nullptr; ctx->ws2_bytes = 0;
        const size_t want = align_up(bytes + bytes / 8, 1 << 20);
        cudaError_t e = cudaMalloc(&ctx->ws2, want);
        if (e != cudaSuccess) return fail(HCB_ENOMEM, std::string("fp32 promotion arena cudaMalloc: ") + cudaGetErrorString(e));
        ctx->ws2_bytes = want;
    }
    double *base = reinterpret_cast<double *>(ctx->ws2);
    for (auto &c : up) c.dst = base + reinterpret_cast<size_t>(c.dst);
    for (auto &c : down) c.src = base + reinterpret_cast<size_t>(c.src);
    for (auto &t : sh) t.d_data = base + reinterpret_cast<size_t>(t.d_data);
    size_t biggest = 0;
    for (auto &c : up) biggest = std::max(biggest, c.n);
    auto run = [&](auto &vec, auto kern) -> int {
        using P = typename std::remove_reference<decltype(vec[0])>::type;
        for (size_t o = 0; o < vec.size(); o += 4096) {
            const size_t cnt = std::min<size_t>(4096, vec.size() - o);
            void *d = nullptr;
            HCB_TRY(ring_upload(ctx, vec.data() + o, sizeof(P) * cnt, &d));
            dim3 grid((unsigned) std::max<size_t>(1, std::min<size_t>(64, (biggest + 2047) / 2048)), (unsigned) cnt);
            kern<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const P *>(d));
            HCB_LAUNCH_CHECK("k_convert_batched");
        }
        return HCB_OK;
    };
    HCB_TRY(run(up, k_convert_batched<float, double>));
    HCB_TRY(t_tlr_gemm_batched<double>(ctx, n, sh.data(), opA, sh.data() + n, opB, sh.data() + 2 * (size_t) n, (double) alpha,
                                       (double) beta, prm, d_info, reset_info));
    HCB_TRY(run(down, k_convert_batched<double, float>));
    return HCB_OK;
}

template<typename T>
int t_tlr_matmul(hcb_ctx *ctx, int64_t mt, int64_t nt, int64_t kt, const hcb_tile *A, const hcb_tile *B,
                 const hcb_tile *C, const int64_t *owned, int64_t n_owned, int64_t k_begin, int64_t k_end, T alpha,
                 T beta, const hcb_compress_params *prm, int32_t *d_info) {
    HCB_TRY(check_ctx(ctx));
    if (mt <= 0 || nt <= 0 || kt <= 0) return HCB_OK;
    if (k_begin < 0 || k_end > kt || k_begin > k_end) return fail(HCB_EINVAL, "tlr_matmul: bad k range");
    const int64_t total = owned ? n_owned : mt * nt;
    if (total <= 0) return HCB_OK;
    std::vector<hcb_tile> a(total), b(total), c(total);
    for (int64_t k = k_begin; k < k_end; ++k) {  // the k-sum of one C tile is sequential (each step recompresses)
        for (int64_t q = 0; q < total; ++q) {
            const int64_t lin = owned ? owned[q] : q;
            if (lin < 0 || lin >= mt * nt) return fail(HCB_EINVAL, "tlr_matmul: owned index out of range");
            const int64_t j = lin % mt, i = lin / mt;
            a[q] = A[j + k * mt];
            b[q] = B[k + i * kt];
            c[q] = C[lin];
        }
        // d_info flags are sticky over the k loop: zeroed once, OR-ed by every step (sweep count: maximum)
        HCB_TRY(t_tlr_gemm_batched<T>(ctx, total, a.data(), 0, b.data(), 0, c.data(), alpha, beta, prm, d_info, k == k_begin));
    }
    return HCB_OK;
}

// One k-step of the multi-tile product on this context's tiles: C(j, i) += alpha * Apan[j] * Bpan[i] for all j < mt,
// i < nt -- the local step of the distributed driver (hcorepp_b200/distributed.py): Apan / Bpan are the broadcast panels.
template<typename T>
int t_tlr_matmul_panel_step(hcb_ctx *ctx, int64_t mt, int64_t nt, const hcb_tile *Apan, const hcb_tile *Bpan,
                            const hcb_tile *C, T alpha, T beta, const hcb_compress_params *prm, int32_t *d_info, int first) {
    HCB_TRY(check_ctx(ctx));
    if (mt <= 0 || nt <= 0) return HCB_OK;
    if (!Apan || !Bpan || !C) return fail(HCB_EINVAL, "tlr_matmul_panel_step: null argument");
    const int64_t total = mt * nt;
    std::vector<hcb_tile> a(total), b(total);
    for (int64_t i = 0; i < nt; ++i)
        for (int64_t j = 0; j < mt; ++j) {
            a[j + i * mt] = Apan[j];
            b[j + i * mt] = Bpan[i];
        }
    return t_tlr_gemm_batched<T>(ctx, total, a.data(), 0, b.data(), 0, C, alpha, beta, prm, d_info, first != 0);
}

template<typename T>
int t_compress_full(hcb_ctx *ctx, int64_t n64, const T *const *dense, int64_t ld, const hcb_tile *out,
                    const hcb_compress_params *prm, int32_t *const *infos) {  // infos: per-tile device info words or NULL
    if (n64 <= 0) return HCB_OK;
    int m = 0, n = 0;
    for (int64_t t = 0; t < n64; ++t) {
        if (out[t].type != HCB_TILE_COMPRESSED || !out[t].d_rank || !out[t].d_data || !dense[t])
            return fail(HCB_EINVAL, "compress_batched: bad output tile");
        m = std::max(m, out[t].m);
        n = std::max(n, out[t].n);
    }
    const int a = std::max(m, n), b = std::min(m, n);
    const SvdJobLayout<T> L(a, b);
    const size_t eUs = align_up((size_t) a * b, 32), eVs = align_up((size_t) b * b, 32), eS = align_up((size_t) b, 32);
    const size_t slab = L.slab + eUs + eVs + eS;
    // chunk the batch so that the scratch stays below ~8 GiB
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n64, (int64_t) (((size_t) 8 << 30) / (slab * sizeof(T)))));
    for (int64_t c0 = 0; c0 < n64; c0 += chunk) {
        const int cnt = (int) std::min<int64_t>(chunk, n64 - c0);
        const size_t desc = svd_jobs_desc_bytes<T>(cnt, a, b) + align_up(sizeof(CompressProb<T>) * cnt, 256);
        HCB_TRY(ensure_ws(ctx, desc + slab * sizeof(T) * cnt + 512));
        char *base = reinterpret_cast<char *>(ctx->ws);
        CompressProb<T> *d_fp = reinterpret_cast<CompressProb<T> *>(base + svd_jobs_desc_bytes<T>(cnt, a, b));
        T *ws = reinterpret_cast<T *>(base + align_up(desc, 256));
        T *outs = ws + (size_t) cnt * L.slab;
        std::vector<SvdJob<T>> jobs(cnt);
        std::vector<CompressProb<T>> fp(cnt);
        for (int t = 0; t < cnt; ++t) {
            const hcb_tile &o = out[c0 + t];
            const int tm = o.m, tn = o.n, ta = std::max(tm, tn), tb = std::min(tm, tn);
            const bool tr = tm < tn;
            T *Us = outs + (size_t) t * (eUs + eVs + eS), *Vs = Us + eUs, *sg = Vs + eVs;
            int *uinfo = infos ? infos[c0 + t] : nullptr;
            jobs[t] = SvdJob<T>{dense[c0 + t], (int) ld, tr ? 1 : 0, ta, tb, Us, Vs, sg, uinfo};
            T *U = reinterpret_cast<T *>(o.d_data), *V = U + (size_t) tm * o.max_rank;
            fp[t] = CompressProb<T>{Us, Vs, sg, U, V, o.d_rank, nullptr, tm, tn, tb, ta, tr ? 1 : 0, o.max_rank, ta, tb, 0, uinfo};
        }
        HCB_TRY(stage_array(ctx, fp, d_fp));
        HCB_TRY(run_svd_jobs<T>(ctx, jobs, a, b, ws, base));
        k_compress_finalize<T><<<cnt, 256, 0, ctx->stream>>>(d_fp, (T) prm->accuracy, prm->truncated_svd,
                                                              (int) prm->fixed_rank);
        HCB_LAUNCH_CHECK("k_compress_finalize");
        if (c0 + chunk < n64) HCB_CUDA(cudaStreamSynchronize(ctx->stream));  // scratch is reused by the next chunk
    }
    return HCB_OK;
}

// Sketched compression (fp64, tiles up to 1024 rows): A ~ Q (Q^T A) with Q = orth(A Omega), k1 = 96 sketch columns;
// the small factor B = Q^T A (k1 x n) goes through the batched SVD pipeline, U = Q U_B, V = Sigma W^T.  A tile whose
// captured spectrum has not decayed two decades below the truncation threshold within k1 - 8 values is flagged on the
// device and re-done with the full SVD.  ~0.5 GFLOP of GEMM-shaped work per 1024^2 tile instead of a 40 GFLOP Jacobi.
constexpr int SKETCH_K = 96, SKETCH_K2 = 288, SKETCH_TAIL = 8;  // second, wider sketch for the tiles the first one rejects

template<typename T>
int t_compress_sketched(hcb_ctx *ctx, int cnt, const T *const *dense, int64_t ld, const hcb_tile *out,
                        const hcb_compress_params *prm, std::vector<int> &redo, int sketch_k = SKETCH_K,
                        int32_t *const *infos = nullptr) {
    // A ~ Qx Qx^T A, Qx = orth(A Omega) (m x k1).  B^T = A^T Qx = Qb Rb (n x k1 panel QR), Rb^T = Ub S Z^T (k1 x k1 Jacobi:
    // LEFT vectors Ub to high relative accuracy), so A ~ (Qx Ub) S (Qb Z)^T:  U = Qx Ub,  V^T = Qb [Z S ; 0].
    const int k1 = sketch_k, nblk = cdiv(k1, NBQ);
    int m = 0, n = 0;
    for (int t = 0; t < cnt; ++t) { m = std::max(m, out[t].m); n = std::max(n, out[t].n); }
    const SvdJobLayout<T> L(k1, k1);
    const size_t eMK = align_up((size_t) m * k1, 32), eNK = align_up((size_t) n * k1, 32), eKK = align_up((size_t) k1 * k1, 32),
                 eK = align_up((size_t) k1, 32), eTB = align_up((size_t) NBQ * NBQ * nblk, 32), eWB = align_up((size_t) 2 * NBQ * k1, 32);
    // per tile: Y VC Qx Uq (m x k1) | Bt VCb Vn (n x k1) | Mr Us Vs (k1 x k1) | tau taub sigma | TB TBb | WB WBb
    const size_t per = 4 * eMK + 3 * eNK + 3 * eKK + 3 * eK + 2 * eTB + 2 * eWB;
    const size_t nsj = (size_t) nblk * cnt;
    const size_t desc = svd_jobs_desc_bytes<T>(cnt, k1, k1) + align_up(sizeof(CompressProb<T>) * cnt, 256) +
                        3 * align_up(sizeof(GemmProb<T>) * cnt, 256) + 2 * align_up(sizeof(PanelDesc<T>) * cnt, 256) +
                        2 * qr_block_desc_bytes<T>(cnt, k1) + 2 * align_up(sizeof(StripJob) * nsj, 256) +
                        2 * align_up(sizeof(SketchGlue<T>) * cnt, 256) + align_up(sizeof(T *) * cnt, 256) +
                        2 * align_up(sizeof(int) * cnt, 256) + 1024;
    const size_t eOm = align_up((size_t) n * k1, 32);
    HCB_TRY(ensure_ws(ctx, desc + ((L.slab + per) * cnt + eOm) * sizeof(T) + 512));
    char *base = reinterpret_cast<char *>(ctx->ws);
    char *pd = base + svd_jobs_desc_bytes<T>(cnt, k1, k1);
    auto carve = [&](size_t bytes) { char *q = pd; pd += align_up(bytes, 256); return q; };
    auto *d_fp = reinterpret_cast<CompressProb<T> *>(carve(sizeof(CompressProb<T>) * cnt));
    auto *d_gy = reinterpret_cast<GemmProb<T> *>(carve(sizeof(GemmProb<T>) * cnt));
    auto *d_gb = reinterpret_cast<GemmProb<T> *>(carve(sizeof(GemmProb<T>) * cnt));
    auto *d_gu = reinterpret_cast<GemmProb<T> *>(carve(sizeof(GemmProb<T>) * cnt));
    auto *d_pan = reinterpret_cast<PanelDesc<T> *>(carve(sizeof(PanelDesc<T>) * cnt));
    auto *d_panb = reinterpret_cast<PanelDesc<T> *>(carve(sizeof(PanelDesc<T>) * cnt));
    char *d_qrb = carve(qr_block_desc_bytes<T>(cnt, k1));
    char *d_qrbb = carve(qr_block_desc_bytes<T>(cnt, k1));
    auto *d_sj = reinterpret_cast<StripJob *>(carve(sizeof(StripJob) * nsj));
    auto *d_sjb = reinterpret_cast<StripJob *>(carve(sizeof(StripJob) * nsj));
    auto *d_g0 = reinterpret_cast<SketchGlue<T> *>(carve(sizeof(SketchGlue<T>) * cnt));
    auto *d_g1 = reinterpret_cast<SketchGlue<T> *>(carve(sizeof(SketchGlue<T>) * cnt));
    auto *d_qx = reinterpret_cast<T **>(carve(sizeof(T *) * cnt));
    auto *d_ms = reinterpret_cast<int *>(carve(sizeof(int) * cnt));
    auto *d_flags = reinterpret_cast<int *>(carve(sizeof(int) * cnt));
    T *ws = reinterpret_cast<T *>(base + align_up(desc, 256));
    T *svd_ws = ws, *blk0 = ws + (size_t) cnt * L.slab, *Om = blk0 + per * cnt;
    k_fill_uniform<T><<<148, 256, 0, ctx->stream>>>(Om, (size_t) n * k1, 0x5EEDull);
    HCB_LAUNCH_CHECK("k_fill_uniform");

    std::vector<CompressProb<T>> fp(cnt);
    std::vector<GemmProb<T>> gy(cnt), gb(cnt), gu(cnt);
    std::vector<PanelDesc<T>> pan(cnt), panb(cnt);
    std::vector<StripJob> sj(nsj), sjb(nsj);
    std::vector<SketchGlue<T>> g0(cnt), g1(cnt);
    std::vector<SvdJob<T>> jobs(cnt);
    std::vector<T *> qx(cnt);
    std::vector<int> ms(cnt);
    for (int t = 0; t < cnt; ++t) {
        const hcb_tile &o = out[t];
        const int tm = o.m, tn = o.n;
        T *Y = blk0 + (size_t) t * per, *VC = Y + eMK, *Qx = VC + eMK, *Uq = Qx + eMK, *Bt = Uq + eMK, *VCb = Bt + eNK,
          *Vn = VCb + eNK, *Mr = Vn + eNK, *Us = Mr + eKK, *Vs = Us + eKK, *tau = Vs + eKK, *taub = tau + eK, *sg = taub + eK,
          *TB = sg + eK, *TBb = TB + eTB, *WB = TBb + eTB, *WBb = WB + eWB;
        gy[t] = GemmProb<T>{dense[t], Om, Y, tm, k1, tn, (int) ld, n, tm, 0, 0, T(1), T(0)};      // Y = A Omega
        pan[t] = PanelDesc<T>{Y, tau, VC, TB, WB, tm, k1, k1, 1};
        gb[t] = GemmProb<T>{dense[t], Qx, Bt, tn, k1, tm, (int) ld, tm, tn, 1, 0, T(1), T(0)};     // B^T = A^T Qx
        panb[t] = PanelDesc<T>{Bt, taub, VCb, TBb, WBb, tn, k1, k1, 1};
        for (int st = 0; st < nblk; ++st) {  // explicit Q [X; 0]: every 32-column strip takes the blocks last to first
            const int nc = std::min(NBQ, k1 - st * NBQ);
            sj[(size_t) t * nblk + st] = StripJob{reinterpret_cast<double *>(Qx) + (size_t) st * NBQ * tm,
                                                  reinterpret_cast<const double *>(VC), reinterpret_cast<const double *>(TB), tm, tm,
                                                  tm, nc, std::min(tm, k1), nblk - 1, nblk, -1, 0};
            sjb[(size_t) t * nblk + st] = StripJob{reinterpret_cast<double *>(Vn) + (size_t) st * NBQ * tn,
                                                   reinterpret_cast<const double *>(VCb), reinterpret_cast<const double *>(TBb), tn,
                                                   tn, tn, nc, std::min(tn, k1), nblk - 1, nblk, -1, 0};
        }
        g0[t] = SketchGlue<T>{Bt, Mr, k1, k1, tn};   // Mr = Rb^T
        jobs[t] = SvdJob<T>{Mr, k1, 0, k1, k1, Us, Vs, sg, infos ? infos[t] : nullptr};  // Rb^T = Ub S Z^T: Us = Ub, Vs = Z S
        g1[t] = SketchGlue<T>{Vs, Vn, tn, k1, k1};   // Vn = [Z S ; 0]
        gu[t] = GemmProb<T>{Qx, Us, Uq, tm, k1, k1, tm, k1, tm, 0, 0, T(1), T(0)};                 // Uq = Qx Ub
        T *U = reinterpret_cast<T *>(o.d_data), *V = U + (size_t) tm * o.max_rank;
        fp[t] = CompressProb<T>{Uq, Vn, sg, U, V, o.d_rank, d_flags + t, tm, tn, k1, tm, 0, o.max_rank, tm, tn, SKETCH_TAIL,
                                infos ? infos[t] : nullptr};
        qx[t] = Qx;
        ms[t] = tm;
    }
    HCB_TRY(stage_array(ctx, fp, d_fp));
    HCB_TRY(stage_array(ctx, gy, d_gy));
    HCB_TRY(stage_array(ctx, gb, d_gb));
    HCB_TRY(stage_array(ctx, gu, d_gu));
    HCB_TRY(stage_array(ctx, pan, d_pan));
    HCB_TRY(stage_array(ctx, panb, d_panb));
    HCB_TRY(stage_array(ctx, sj, d_sj));
    HCB_TRY(stage_array(ctx, sjb, d_sjb));
    HCB_TRY(stage_array(ctx, g0, d_g0));
    HCB_TRY(stage_array(ctx, g1, d_g1));
    HCB_TRY(stage_array(ctx, qx, d_qx));
    HCB_TRY(stage_array(ctx, ms, d_ms));
    HCB_TRY(launch_gemm<T>(ctx, d_gy, cnt, m, k1));                       // Y = A Omega
    HCB_TRY(run_blocked_qr<T>(ctx, d_pan, cnt, m, k1, d_qrb));            // Y = Qx R
    k_eye_batched<T><<<cnt, 256, 0, ctx->stream>>>(d_qx, d_ms, k1);
    HCB_LAUNCH_CHECK("k_eye_batched");
    HCB_TRY(launch_strips(ctx, d_sj, (int) nsj, m));                      // Qx explicit
    HCB_TRY(launch_gemm<T>(ctx, d_gb, cnt, n, k1));                       // B^T = A^T Qx
    HCB_TRY(run_blocked_qr<T>(ctx, d_panb, cnt, n, k1, d_qrbb));          // B^T = Qb Rb
    k_sketch_glue<T><<<cnt, 256, 0, ctx->stream>>>(d_g0, 0);
    HCB_LAUNCH_CHECK("k_sketch_glue");
    HCB_TRY(run_svd_jobs<T>(ctx, jobs, k1, k1, svd_ws, base));            // Rb^T = Ub S Z^T
    k_sketch_glue<T><<<cnt, 256, 0, ctx->stream>>>(d_g1, 1);
    HCB_LAUNCH_CHECK("k_sketch_glue");
    HCB_TRY(launch_strips(ctx, d_sjb, (int) nsj, n));                     // Vn = Qb [Z S ; 0]
    HCB_TRY(launch_gemm<T>(ctx, d_gu, cnt, m, k1));                       // Uq = Qx Ub
    k_compress_finalize<T><<<cnt, 256, 0, ctx->stream>>>(d_fp, (T) prm->accuracy, prm->truncated_svd, 0);
    HCB_LAUNCH_CHECK("k_compress_finalize");
    std::vector<int> flags(cnt);
    HCB_CUDA(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    HCB_CUDA(cudaStreamSynchronize(ctx->stream));  // (also: the scratch is reused by the fallback / next chunk)
    for (int t = 0; t < cnt; ++t)
        if (flags[t] != 0) redo.push_back(t);
    return HCB_OK;
}

template<typename T>
int t_compress_batched(hcb_ctx *ctx, int64_t n64, const T *const *dense, int64_t ld, const hcb_tile *out,
                       const hcb_compress_params *prm, int32_t *d_info) {
    HCB_TRY(check_ctx(ctx));
    if (n64 <= 0) return HCB_OK;
    if (!dense || !out || !prm) return fail(HCB_EINVAL, "compress_batched: null argument");
    int m = 0, n = 0, mn_min = 1 << 30;
    for (int64_t t = 0; t < n64; ++t) {
        if (out[t].type != HCB_TILE_COMPRESSED || !out[t].d_rank || !out[t].d_data || !dense[t])
            return fail(HCB_EINVAL, "compress_batched: bad output tile");
        m = std::max(m, out[t].m);
        n = std::max(n, out[t].n);
        mn_min = std::min(mn_min, std::min(out[t].m, out[t].n));
    }
    // sketch when it pays (both dimensions well above the sketch width) and the strip kernels apply
    const bool sketch = std::is_same<T, double>::value && prm->fixed_rank <= 0 && mn_min >= 3 * SKETCH_K &&
                        strip_path_ok(ctx, std::max(m, n)) && !getenv("HCB_COMPRESS_FULL_SVD");
    // d_info: zeroed here, then |= 1 (Jacobi not converged) / |= 2 (rank clipped at max_rank) per tile, bits 8.. = sweeps
    std::vector<int32_t *> ip;
    if (d_info) {
        HCB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t) * (size_t) n64, ctx->stream));
        ip.resize((size_t) n64);
        for (int64_t t = 0; t < n64; ++t) ip[t] = d_info + t;
    }
    int32_t *const *infos = d_info ? ip.data() : nullptr;
    if (!sketch) return t_compress_full<T>(ctx, n64, dense, ld, out, prm, infos);
    const int64_t chunk = 512;
    for (int64_t c0 = 0; c0 < n64; c0 += chunk) {
        const int cnt = (int) std::min<int64_t>(chunk, n64 - c0);
        std::vector<int> redo;
        HCB_TRY(t_compress_sketched<T>(ctx, cnt, dense + c0, ld, out + c0, prm, redo, SKETCH_K, infos ? infos + c0 : nullptr));
        if (!redo.empty() && mn_min >= 3 * SKETCH_K2) {
            // rank beyond the first sketch (e.g. neighbouring clusters of a covariance matrix: ranks 100-200 at nb = 1024):
            // a second, 288-column sketch before giving up on sketching
            std::vector<const T *> rd(redo.size());
            std::vector<hcb_tile> ro(redo.size());
            std::vector<int32_t *> ri(redo.size(), nullptr);
            for (size_t i = 0; i < redo.size(); ++i) {
                rd[i] = dense[c0 + redo[i]]; ro[i] = out[c0 + redo[i]];
                if (infos) ri[i] = infos[c0 + redo[i]];
            }
            std::vector<int> redo2, again;
            for (size_t b0 = 0; b0 < rd.size(); b0 += 128) {
                const int bc = (int) std::min<size_t>(128, rd.size() - b0);
                redo2.clear();
                HCB_TRY(t_compress_sketched<T>(ctx, bc, rd.data() + b0, ld, ro.data() + b0, prm, redo2, SKETCH_K2,
                                               infos ? ri.data() + b0 : nullptr));
                for (int r : redo2) again.push_back(redo[b0 + r]);
            }
            redo.swap(again);
        }
        if (!redo.empty()) {  // spectrum too flat for the sketches: full SVD for those tiles
            std::vector<const T *> rd(redo.size());
            std::vector<hcb_tile> ro(redo.size());
            std::vector<int32_t *> ri(redo.size(), nullptr);
            for (size_t i = 0; i < redo.size(); ++i) {
                rd[i] = dense[c0 + redo[i]]; ro[i] = out[c0 + redo[i]];
                if (infos) ri[i] = infos[c0 + redo[i]];
            }
            HCB_TRY(t_compress_full<T>(ctx, (int64_t) redo.size(), rd.data(), ld, ro.data(), prm, infos ? ri.data() : nullptr));
            HCB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    return HCB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// TLR Cholesky pieces: dense potrf / trsm / syrk of the kernel table + the batched tile forms
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
int t_potrf_lower(hcb_ctx *ctx, int64_t n64, T *A, int64_t lda, int32_t *d_info, T *D) {
    const int n = (int) n64;
    k_diag_upper<T><<<cdiv(n, 256), 256, 0, ctx->stream>>>(0, n, A, (int) lda, D);
    HCB_LAUNCH_CHECK("k_diag_upper");
    for (int j0 = 0; j0 < n; j0 += CH_NB) {
        const int jb = std::min(CH_NB, n - j0);
        if (j0 > 0) {  // left-looking update of block column j0 (all rows from j0 down) with the columns to its left
            GemmProb<T> g{A + j0, A + j0, A + j0 + (size_t) j0 * lda, n - j0, jb, j0, (int) lda, (int) lda, (int) lda, 0, 1, T(-1), T(1)};
            const GemmProb<T> *d;
            HCB_TRY(upload_one(ctx, g, &d));
            HCB_TRY(launch_gemm<T>(ctx, d, 1, n - j0, jb));
        }
        k_potrf_panel<T><<<1 + cdiv(std::max(0, n - j0 - jb), 256), 256, 0, ctx->stream>>>(A, n, (int) lda, j0, jb, d_info);
        HCB_LAUNCH_CHECK("k_potrf_panel");
    }
    k_diag_upper<T><<<cdiv(n, 256), 256, 0, ctx->stream>>>(1, n, A, (int) lda, D);
    HCB_LAUNCH_CHECK("k_diag_upper");
    return HCB_OK;
}

template<typename T>
int t_potrf(hcb_ctx *ctx, int uplo, int64_t n, T *A, int64_t lda, int32_t *d_info) {
    HCB_TRY(check_ctx(ctx));
    if (n <= 0) return HCB_OK;
    if (lda < n) return fail(HCB_EINVAL, "potrf: lda < n");
    const bool upper = (uplo == 'U' || uplo == 'u' || uplo == 1);
    const size_t eD = align_up((size_t) n * CH_NB, 32), eS = upper ? (size_t) n * n : 0;
    HCB_TRY(ensure_ws(ctx, (eD + 2 * eS) * sizeof(T) + 1024));  // (once, up front: growing the arena frees the old one)
    T *D = reinterpret_cast<T *>(ctx->ws), *S = D + eD;
    if (d_info) HCB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), ctx->stream));
    if (!upper) return t_potrf_lower<T>(ctx, n, A, lda, d_info, D);
    // A = U^T U with the upper triangle stored: factor the transposed triangle as a lower problem, transpose back
    HCB_TRY(t_copy<T>(ctx, A, (int) lda, S, (int) n, (int) n, (int) n, 1, T(1)));      // S = A^T (lower of S = upper of A)
    HCB_TRY(t_potrf_lower<T>(ctx, n, S, n, d_info, D));
    // write back only the upper triangle (incl. diagonal) of A := (lower of S)^T
    T *S2 = S + eS;
    HCB_TRY(t_copy<T>(ctx, S, (int) n, S2, (int) n, (int) n, (int) n, 1, T(1)));       // S2 = S^T
    return t_lacpy<T>(ctx, 'U', n, n, S2, n, A, lda);
}

template<typename T>
int t_trsm(hcb_ctx *ctx, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *A, int64_t lda,
           T *B, int64_t ldb) {
    HCB_TRY(check_ctx(ctx));
    if (m <= 0 || n <= 0) return HCB_OK;
    const int right = (side == 'R' || side == 'r' || side == 1), upper = (uplo == 'U' || uplo == 'u' || uplo == 1);
    const int tr = (trans != 0 && trans != 'N' && trans != 'n'), unit = (diag == 'U' || diag == 'u' || diag == 1);
    const int nrhs = (int) (right ? m : n);
    k_trsm_generic<T><<<cdiv(nrhs, 64), 64, 0, ctx->stream>>>(right, upper, tr, unit, (int) m, (int) n, alpha, A, (int) lda, B, (int) ldb);
    HCB_LAUNCH_CHECK("k_trsm_generic");
    return HCB_OK;
}

template<typename T>
int t_syrk(hcb_ctx *ctx, int uplo, int trans, int64_t n, int64_t k, T alpha, const T *A, int64_t lda, T beta, T *Cm, int64_t ldc) {
    HCB_TRY(check_ctx(ctx));
    if (n <= 0) return HCB_OK;
    const int upper = (uplo == 'U' || uplo == 'u' || uplo == 1), tr = (trans != 0 && trans != 'N' && trans != 'n');
    HCB_TRY(ensure_ws(ctx, (size_t) n * n * sizeof(T) + 1024));
    T *W = reinterpret_cast<T *>(ctx->ws);
    // W = op(A) op(A)^T : NoTrans -> A (n x k) A^T ; Trans -> A^T (n x k) A with A stored k x n
    GemmProb<T> g{A, A, W, (int) n, (int) n, (int) k, (int) lda, (int) lda, (int) n, tr, tr ? 0 : 1, T(1), T(0)};
    const GemmProb<T> *d;
    HCB_TRY(upload_one(ctx, g, &d));
    HCB_TRY(launch_gemm<T>(ctx, d, 1, (int) n, (int) n));
    dim3 block(32, 8), grid(cdiv(n, 32), cdiv(n, 8));
    k_syrk_combine<T><<<grid, block, 0, ctx->stream>>>(upper, (int) n, alpha, W, beta, Cm, (int) ldc);
    HCB_LAUNCH_CHECK("k_syrk_combine");
    return HCB_OK;
}

template<typename T>
int t_fill_triangle(hcb_ctx *ctx, int uplo, int64_t n, T *A, int64_t lda, T value) {
    HCB_TRY(check_ctx(ctx));
    if (n <= 0) return HCB_OK;
    dim3 block(32, 8), grid(cdiv(n, 32), cdiv(n, 8));
    k_fill_triangle<T><<<grid, block, 0, ctx->stream>>>((uplo == 'U' || uplo == 'u' || uplo == 1) ? 1 : 0, (int) n, A, (int) lda, value);
    HCB_LAUNCH_CHECK("k_fill_triangle");
    return HCB_OK;
}

template<typename T>
int t_symmetrize(hcb_ctx *ctx, int uplo, int64_t n, T *A, int64_t lda) {
    HCB_TRY(check_ctx(ctx));
    if (n <= 0) return HCB_OK;
    dim3 block(32, 8), grid(cdiv(n, 32), cdiv(n, 8));
    k_symmetrize<T><<<grid, block, 0, ctx->stream>>>((uplo == 'U' || uplo == 'u' || uplo == 1) ? 1 : 0, (int) n, A, (int) lda);
    HCB_LAUNCH_CHECK("k_symmetrize");
    return HCB_OK;
}

// V := V L^-T for a batch of compressed tiles X (the tile-Cholesky panel solve A(i,k) := A(i,k) L_kk^-T acts on the
// right factor only), L = lower Cholesky factors (one per tile; usually the same diagonal tile for a whole block column)
template<typename T>
int t_tlr_trsm_batched(hcb_ctx *ctx, int64_t n64, const hcb_tile *X, const T *const *dL, const int64_t *ldl) {
    HCB_TRY(check_ctx(ctx));
    if (n64 <= 0) return HCB_OK;
    if (!X || !dL || !ldl) return fail(HCB_EINVAL, "tlr_trsm_batched: null argument");
    const int n = (int) n64;
    int cols = 0, rkb = 0;
    std::vector<int> lds(n);
    for (int t = 0; t < n; ++t) {
        if (X[t].type != HCB_TILE_COMPRESSED || !X[t].d_rank || !X[t].d_data || !dL[t])
            return fail(HCB_EINVAL, "tlr_trsm_batched: X must be compressed tiles, L non-null");
        cols = std::max(cols, X[t].n);
        rkb = std::max(rkb, bound_of(X[t]));
        lds[t] = (int) ldl[t];
    }
    const int nsteps = cdiv(cols, CH_NB);
    const size_t bG = align_up(sizeof(GemmProb<T>) * (size_t) nsteps * n, 256), bT = align_up(sizeof(TrsmProb<T>) * n, 256),
                 bX = align_up(sizeof(hcb_tile) * n, 256), bL = align_up(sizeof(T *) * n, 256), bI = align_up(sizeof(int) * n, 256);
    HCB_TRY(ensure_ws(ctx, bG + bT + bX + bL + bI + 256));
    char *base = reinterpret_cast<char *>(ctx->ws);
    auto *d_g = reinterpret_cast<GemmProb<T> *>(base);
    auto *d_t = reinterpret_cast<TrsmProb<T> *>(base + bG);
    auto *d_x = reinterpret_cast<hcb_tile *>(base + bG + bT);
    auto *d_l = reinterpret_cast<const T **>(base + bG + bT + bX);
    auto *d_i = reinterpret_cast<int *>(base + bG + bT + bX + bL);
    std::vector<hcb_tile> xs(X, X + n);
    std::vector<const T *> ls(dL, dL + n);
    HCB_TRY(stage_array(ctx, xs, d_x));
    HCB_TRY(stage_array(ctx, ls, d_l));
    HCB_TRY(stage_array(ctx, lds, d_i));
    k_setup_trsm_tiles<T><<<cdiv(nsteps * n, 128), 128, 0, ctx->stream>>>(d_x, d_l, d_i, n, nsteps, d_g, d_t);
    HCB_LAUNCH_CHECK("k_setup_trsm_tiles");
    for (int step = 0; step < nsteps; ++step) {
        if (step > 0) HCB_TRY(launch_gemm<T>(ctx, d_g + (size_t) step * n, n, rkb, CH_NB));
        dim3 grid(std::max(1, cdiv(rkb, 128)), n);
        k_trsm_rlt_block<T><<<grid, 128, 0, ctx->stream>>>(d_t, step * CH_NB, CH_NB);
        HCB_LAUNCH_CHECK("k_trsm_rlt_block");
    }
    return HCB_OK;
}

// C[t] := beta C[t] + alpha A[t] A[t]^T for compressed A[t] = U V and dense C[t] (m x m): HCore<T>::Syrk with a compressed
// operand (HCore.cpp:484-575) without the triangle fill (the caller keeps C symmetric / reads one triangle)
template<typename T>
int t_tlr_syrk_batched(hcb_ctx *ctx, int64_t n64, const hcb_tile *A, T *const *dC, const int64_t *ldc, T alpha, T beta) {
    HCB_TRY(check_ctx(ctx));
    if (n64 <= 0) return HCB_OK;
    if (!A || !dC || !ldc) return fail(HCB_EINVAL, "tlr_syrk_batched: null argument");
    const int n = (int) n64;
    int m = 0, cols = 0, rkb = 0, cap = 0;
    std::vector<int> lds(n);
    for (int t = 0; t < n; ++t) {
        if (A[t].type != HCB_TILE_COMPRESSED || !A[t].d_rank || !A[t].d_data || !dC[t])
            return fail(HCB_EINVAL, "tlr_syrk_batched: A must be compressed tiles, C non-null");
        m = std::max(m, A[t].m); cols = std::max(cols, A[t].n);
        rkb = std::max(rkb, bound_of(A[t])); cap = std::max(cap, A[t].max_rank);
        lds[t] = (int) ldc[t];
    }
    const size_t slab = align_up((size_t) cap * cap, 32) + align_up((size_t) m * cap, 32);
    const size_t bG = align_up(sizeof(GemmProb<T>) * n, 256), bX = align_up(sizeof(hcb_tile) * n, 256),
                 bC = align_up(sizeof(T *) * n, 256), bI = align_up(sizeof(int) * n, 256);
    HCB_TRY(ensure_ws(ctx, 3 * bG + bX + bC + bI + slab * sizeof(T) * n + 512));
    char *base = reinterpret_cast<char *>(ctx->ws);
    auto *g1 = reinterpret_cast<GemmProb<T> *>(base), *g2 = reinterpret_cast<GemmProb<T> *>(base + bG),
         *g3 = reinterpret_cast<GemmProb<T> *>(base + 2 * bG);
    auto *d_x = reinterpret_cast<hcb_tile *>(base + 3 * bG);
    auto *d_c = reinterpret_cast<T **>(base + 3 * bG + bX);
    auto *d_i = reinterpret_cast<int *>(base + 3 * bG + bX + bC);
    T *ws = reinterpret_cast<T *>(base + align_up(3 * bG + bX + bC + bI, 256));
    std::vector<hcb_tile> xs(A, A + n);
    std::vector<T *> cs(dC, dC + n);
    HCB_TRY(stage_array(ctx, xs, d_x));
    HCB_TRY(stage_array(ctx, cs, d_c));
    HCB_TRY(stage_array(ctx, lds, d_i));
    k_setup_syrk_tiles<T><<<cdiv(n, 128), 128, 0, ctx->stream>>>(d_x, d_c, d_i, n, ws, slab, alpha, beta, g1, g2, g3);
    HCB_LAUNCH_CHECK("k_setup_syrk_tiles");
    HCB_TRY(launch_gemm<T>(ctx, g1, n, rkb, rkb));
    HCB_TRY(launch_gemm<T>(ctx, g2, n, m, rkb));
    return launch_gemm<T>(ctx, g3, n, m, m);
}

// Right-looking tile Cholesky A = L L^T of a symmetric positive definite matrix held as dense diagonal tiles + compressed
// tiles below the diagonal (the reference has the tile routines HCore<T>::Potrf / Trsm / Syrk / Gemm(aCholesky) but no
// driver):  for k:  L_kk = potrf(A_kk);  A_ik := A_ik L_kk^-T (i > k);  A_ii -= A_ik A_ik^T;
// A_ij -= A_ik A_jk^T (i > j > k, ONE batched recompressing call).  low = nt x nt column-major grid, entries i > j used.
template<typename T>
int t_tlr_potrf(hcb_ctx *ctx, int64_t nt, int64_t nb, T *const *diag, int64_t ldd, const hcb_tile *low,
                const hcb_compress_params *prm, int32_t *d_info, int32_t *d_potrf_info) {
    HCB_TRY(check_ctx(ctx));
    if (nt <= 0) return HCB_OK;
    if (!diag || !low || !prm || nb <= 0 || ldd < nb) return fail(HCB_EINVAL, "tlr_potrf: bad argument");
    if (d_info) HCB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t) * (size_t) (nt * nt), ctx->stream));
    if (d_potrf_info) HCB_CUDA(cudaMemsetAsync(d_potrf_info, 0, sizeof(int32_t) * (size_t) nt, ctx->stream));
    std::vector<hcb_tile> col, a, b, c;
    std::vector<const T *> ls;
    std::vector<T *> cs;
    std::vector<int64_t> lds;
    for (int64_t k = 0; k < nt; ++k) {
        HCB_TRY(ensure_ws(ctx, align_up((size_t) nb * CH_NB, 32) * sizeof(T) + 1024));
        HCB_TRY(t_potrf_lower<T>(ctx, nb, diag[k], ldd, d_potrf_info ? d_potrf_info + k : nullptr, reinterpret_cast<T *>(ctx->ws)));
        const int64_t cnt = nt - k - 1;
        if (cnt <= 0) break;
        col.assign(cnt, hcb_tile{});
        ls.assign(cnt, diag[k]);
        lds.assign(cnt, ldd);
        cs.resize(cnt);
        for (int64_t i = k + 1; i < nt; ++i) {
            col[i - k - 1] = low[i + k * nt];
            cs[i - k - 1] = diag[i];
        }
        HCB_TRY(t_tlr_trsm_batched<T>(ctx, cnt, col.data(), ls.data(), lds.data()));
        HCB_TRY(t_tlr_syrk_batched<T>(ctx, cnt, col.data(), cs.data(), lds.data(), T(-1), T(1)));
        a.clear(); b.clear(); c.clear();
        for (int64_t j = k + 1; j < nt; ++j)
            for (int64_t i = j + 1; i < nt; ++i) {
                a.push_back(low[i + k * nt]);
                b.push_back(low[j + k * nt]);
                c.push_back(low[i + j * nt]);
            }
        if (!a.empty()) {
            // d_info: sticky over the steps, indexed by the position of the C tile in the step's batch (diagnostics)
            HCB_TRY(t_tlr_gemm_batched<T>(ctx, (int64_t) a.size(), a.data(), 0, b.data(), 1, c.data(), T(-1), T(1), prm, d_info, false));
        }
    }
    return HCB_OK;
}

template<typename T>
size_t t_workspace(int64_t n_tiles, int64_t m, int64_t n, int64_t k, int64_t r_bound) {
    BatchShape s;
    s.m = (int) m; s.n = (int) n; s.k = (int) k; s.mix = CCC;
    // r_bound = bound on the stacked rank kc + ka of any tile of the batch; the split is not known here, so every
    // single rank is bounded by r_bound as well (an upper bound of what a call with these bounds makes the arena grow to)
    s.kA = s.kB = s.kC = (int) r_bound;
    s.r_force = (int) r_bound;
    s.maxrankC = (int) std::max<int64_t>(1, std::min(m, n) / 3);
    const Layout<T> L = make_layout<T>(s);
    const DescArrays<T> D((int) n_tiles, L.nblk, std::max(L.r_b, L.pq_b), std::max(1, std::min(L.pq_b, s.maxrankC)));
    return D.bytes + L.slab * sizeof(T) * (size_t) n_tiles + 512;
}

}  // namespace hcb

// =================================================================================================================
// extern "C"
// =================================================================================================================
using namespace hcb;

extern "C" {

const char *hcb_last_error(void) { return g_last_error.c_str(); }
const char *hcb_version(void) {
    return "hcore_b200 0.1 (sm_100a; batched TLR GEMM + recompression; no CPU fallback; built " __DATE__ " " __TIME__ ")";
}
uint64_t hcb_launch_count(void) { return g_launches.load(); }
void hcb_launch_count_reset(void) { g_launches.store(0); }

static int ctx_init(int device, cudaStream_t stream, bool own, hcb_ctx **out) {
    if (!out) return fail(HCB_EINVAL, "null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(HCB_ENODEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) +
                                       "): libhcore_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(HCB_EINVAL, "device index out of range");
    HCB_CUDA(cudaSetDevice(device));
    hcb_ctx *c = new hcb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    HCB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    HCB_CUDA(cudaMalloc((void **) &c->d_err, sizeof(int)));
    HCB_CUDA(cudaMemset(c->d_err, 0, sizeof(int)));
    HCB_CUDA(cudaMalloc((void **) &c->d_stats, 8 * sizeof(unsigned long long)));
    HCB_CUDA(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
    HCB_CUDA(cudaMallocHost((void **) &c->h_err, sizeof(int)));
    *c->h_err = 0;
    if (own) {
        HCB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    } else {
        c->stream = stream;
    }
    *out = c;
    return HCB_OK;
}

int hcb_ctx_create(int device, hcb_ctx **out) { return ctx_init(device, nullptr, true, out); }
int hcb_ctx_create_on_stream(int device, void *s, hcb_ctx **out) { return ctx_init(device, (cudaStream_t) s, false, out); }

int hcb_ctx_destroy(hcb_ctx *c) {
    if (!c) return HCB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->ws) cudaFree(c->ws);
    if (c->ws2) cudaFree(c->ws2);
    if (c->info_tmp) cudaFree(c->info_tmp);
    if (c->svd_sched) cudaFree(c->svd_sched);
    if (c->d_err) cudaFree(c->d_err);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->h_err) cudaFreeHost(c->h_err);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto e : c->ring.ev) if (e) cudaEventDestroy(e);
    if (c->ring.h) cudaFreeHost(c->ring.h);
    if (c->ring.d) cudaFree(c->ring.d);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return HCB_OK;
}

int hcb_ctx_phase_timing(hcb_ctx *c, int enable) {
    HCB_TRY(check_ctx(c));
    c->timing = enable != 0;
    return HCB_OK;
}
int hcb_ctx_phase_times(hcb_ctx *c, double *ms, uint64_t *launches) {
    HCB_TRY(check_ctx(c));
    HCB_CUDA(cudaStreamSynchronize(c->stream));
    for (auto &r : c->phase_recs) {
        float t = 0.f;
        HCB_CUDA(cudaEventElapsedTime(&t, r.beg, r.end));
        if (ms) ms[r.phase] += t;
        if (launches) launches[r.phase] += 1;
    }
    c->phase_recs.clear();
    c->ev_used = 0;
    return HCB_OK;
}
const char *hcb_phase_name(int phase) {
    static const char *names[HCB_N_PHASES] = {"setup", "contraction", "stack", "panel_qr", "core_lq", "apply_q",
                                              "finalize", "jacobi_svd", "vsigma_truncate"};
    return (phase >= 0 && phase < HCB_N_PHASES) ? names[phase] : "?";
}

int hcb_ctx_sync(hcb_ctx *c) {
    HCB_TRY(check_ctx(c));
    // the fused path reports a violated rank bound (tile left untouched) here even when the caller passed no d_info
    HCB_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    HCB_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_err) {
        const int e = *c->h_err;
        *c->h_err = 0;
        HCB_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
        if (e & 4)
            return fail(HCB_EBOUND, "a tile's rank exceeded its rank_bound in a fused call since the last sync: that "
                                    "tile's update was NOT applied (raise rank_bound, or leave it 0 = max_rank)");
    }
    return HCB_OK;
}
int hcb_ctx_stats(hcb_ctx *c, uint64_t *out8, int reset) {
    HCB_TRY(check_ctx(c));
    if (!out8) return fail(HCB_EINVAL, "ctx_stats: null output");
    HCB_CUDA(cudaStreamSynchronize(c->stream));
    HCB_CUDA(cudaMemcpy(out8, c->d_stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) HCB_CUDA(cudaMemset(c->d_stats, 0, 8 * sizeof(uint64_t)));
    return HCB_OK;
}
void *hcb_ctx_stream(hcb_ctx *c) { return c ? (void *) c->stream : nullptr; }
int hcb_ctx_device(hcb_ctx *c) { return c ? c->device : -1; }
int hcb_ctx_sm_count(hcb_ctx *c) { return c ? c->sm_count : 0; }
int hcb_ctx_reserve_workspace(hcb_ctx *c, size_t bytes) {
    HCB_TRY(check_ctx(c));
    return ensure_ws(c, bytes);
}
size_t hcb_ctx_workspace_bytes(hcb_ctx *c) { return c ? c->ws_bytes : 0; }

int hcb_malloc(hcb_ctx *c, size_t bytes, void **d_out) {
    HCB_TRY(check_ctx(c));
    if (!d_out) return fail(HCB_EINVAL, "null out pointer");
    cudaError_t e = cudaMalloc(d_out, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(HCB_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return HCB_OK;
}
int hcb_free(hcb_ctx *c, void *p) {
    HCB_TRY(check_ctx(c));
    if (p) HCB_CUDA(cudaFree(p));
    return HCB_OK;
}
int hcb_memcpy(hcb_ctx *c, void *dst, const void *src, size_t bytes, int kind) {
    HCB_TRY(check_ctx(c));
    static const cudaMemcpyKind kinds[5] = {cudaMemcpyHostToDevice, cudaMemcpyDeviceToDevice, cudaMemcpyDeviceToHost,
                                            cudaMemcpyHostToHost, cudaMemcpyDefault};
    if (kind < 0 || kind > 4) return fail(HCB_EINVAL, "memcpy: bad kind");
    if (bytes) HCB_CUDA(cudaMemcpyAsync(dst, src, bytes, kinds[kind], c->stream));
    return HCB_OK;
}
int hcb_memset(hcb_ctx *c, void *dst, int value, size_t bytes) {
    HCB_TRY(check_ctx(c));
    if (bytes) HCB_CUDA(cudaMemsetAsync(dst, value, bytes, c->stream));
    return HCB_OK;
}

#define HCB_DEFINE_KERNEL_TABLE(P, T)                                                                                 \
    int hcb_##P##gemm(hcb_ctx *c, int ta, int tb, int64_t m, int64_t n, int64_t k, T alpha, const T *A, int64_t lda,   \
                      const T *B, int64_t ldb, T beta, T *C, int64_t ldc) {                                           \
        return t_gemm<T>(c, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);                                   \
    }                                                                                                                 \
    int hcb_##P##multiply_by_alpha(hcb_ctx *c, T *arr, int64_t rows, int64_t cols, int64_t m, int64_t rank, T alpha) { \
        return t_multiply_by_alpha<T>(c, arr, rows, cols, m, rank, alpha);                                           \
    }                                                                                                                 \
    int hcb_##P##process_v(hcb_ctx *c, int64_t n, int64_t crank, int ungqr, int64_t vm, T beta, const T *CV,          \
                           int64_t ldcv, T *V, int64_t arank, const T *B, int cholesky) {                             \
        return t_process_v<T>(c, n, crank, ungqr, vm, beta, CV, ldcv, V, arank, B, cholesky);                        \
    }                                                                                                                 \
    int hcb_##P##new_rank(hcb_ctx *c, int trunc, const T *sig, int64_t size_s, T acc, int64_t *host_rank) {            \
        return t_new_rank<T>(c, trunc, sig, size_s, acc, host_rank);                                                 \
    }                                                                                                                 \
    int hcb_##P##new_rank_device(hcb_ctx *c, int trunc, const T *sig, int64_t size_s, T acc, int32_t *d_rank) {        \
        return t_new_rank_device<T>(c, trunc, sig, size_s, acc, d_rank);                                             \
    }                                                                                                                 \
    int hcb_##P##uvptr(hcb_ctx *c, int64_t rank, int64_t vm, T *UV, const T *Vnew) {                                   \
        return t_uvptr<T>(c, rank, vm, UV, Vnew);                                                                    \
    }                                                                                                                 \
    int hcb_##P##vtnew(hcb_ctx *c, int64_t rk, int ungqr, int64_t mn, const T *sig, T *VT, int64_t size_s,             \
                       int64_t vm) {                                                                                  \
        return t_vtnew<T>(c, rk, ungqr, mn, sig, VT, size_s, vm);                                                    \
    }                                                                                                                 \
    int hcb_##P##uvptr_conj(hcb_ctx *c, int64_t, int64_t, T *) { return check_ctx(c); }                                \
    int hcb_##P##fill_identity(hcb_ctx *c, int64_t n, T *A) { return t_fill_identity<T>(c, n, A); }                    \
    int hcb_##P##lacpy(hcb_ctx *c, int type, int64_t m, int64_t n, const T *A, int64_t lda, T *B, int64_t ldb) {       \
        return t_lacpy<T>(c, type, m, n, A, lda, B, ldb);                                                            \
    }                                                                                                                 \
    int hcb_##P##laset(hcb_ctx *c, int type, int64_t m, int64_t n, T off, T diag, T *A, int64_t lda) {                 \
        return t_laset<T>(c, type, m, n, off, diag, A, lda);                                                         \
    }                                                                                                                 \
    int hcb_##P##geqrf(hcb_ctx *c, int64_t m, int64_t n, T *A, int64_t lda, T *tau) {                                  \
        return t_geqrf<T>(c, m, n, A, lda, tau);                                                                     \
    }                                                                                                                 \
    int hcb_##P##ungqr(hcb_ctx *c, int64_t m, int64_t n, int64_t k, T *A, int64_t lda, const T *tau) {                 \
        return t_ungqr<T>(c, m, n, k, A, lda, tau);                                                                  \
    }                                                                                                                 \
    int hcb_##P##unmqr(hcb_ctx *c, int side, int trans, int64_t m, int64_t n, int64_t k, const T *A, int64_t lda,      \
                       const T *tau, T *C, int64_t ldc) {                                                             \
        return t_unmqr<T>(c, side, trans, m, n, k, A, lda, tau, C, ldc);                                             \
    }                                                                                                                 \
    int hcb_##P##svd(hcb_ctx *c, int64_t m, int64_t n, T *A, int64_t lda, T *S, T *U, int64_t ldu, T *VT,              \
                     int64_t ldvt) {                                                                                  \
        return t_svd<T>(c, m, n, A, lda, S, U, ldu, VT, ldvt);                                                       \
    }                                                                                                                 \
    int hcb_##P##trmm(hcb_ctx *c, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *A,  \
                      int64_t lda, T *B, int64_t ldb) {                                                               \
        return t_trmm<T>(c, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);                                   \
    }                                                                                                                 \
    int hcb_##P##tlr_gemm_batched(hcb_ctx *c, int64_t n, const hcb_tile *A, int opA, const hcb_tile *B, int opB,       \
                                  const hcb_tile *C, T alpha, T beta, const hcb_compress_params *p, int32_t *info) {  \
        return t_tlr_gemm_any<T>(c, n, A, opA, B, opB, C, alpha, beta, p, info);                                     \
    }                                                                                                                 \
    int hcb_##P##compress_batched(hcb_ctx *c, int64_t n, const T *const *dense, int64_t ld, const hcb_tile *out,       \
                                  const hcb_compress_params *p, int32_t *info) {                                      \
        return t_compress_batched<T>(c, n, dense, ld, out, p, info);                                                 \
    }                                                                                                                 \
    int hcb_##P##tlr_matmul(hcb_ctx *c, int64_t mt, int64_t nt, int64_t kt, const hcb_tile *A, const hcb_tile *B,      \
                            const hcb_tile *C, const int64_t *owned, int64_t n_owned, int64_t k_begin,                \
                            int64_t k_end, T alpha, T beta, const hcb_compress_params *p, int32_t *info) {            \
        return t_tlr_matmul<T>(c, mt, nt, kt, A, B, C, owned, n_owned, k_begin, k_end, alpha, beta, p, info);        \
    }                                                                                                                 \
    int hcb_##P##transpose(hcb_ctx *c, int64_t rows, int64_t cols, const T *A, int64_t lda, T *Out, int64_t ldo) {     \
        int rc_ = check_ctx(c);                                                                                       \
        if (rc_ != HCB_OK) return rc_;                                                                                \
        return t_copy<T>(c, A, (int) lda, Out, (int) ldo, (int) cols, (int) rows, 1, T(1));                          \
    }                                                                                                                 \
    int hcb_##P##potrf(hcb_ctx *c, int uplo, int64_t n, T *A, int64_t lda, int32_t *d_info) {                          \
        return t_potrf<T>(c, uplo, n, A, lda, d_info);                                                               \
    }                                                                                                                 \
    int hcb_##P##trsm(hcb_ctx *c, int side, int uplo, int trans, int diag, int64_t m, int64_t n, T alpha, const T *A,  \
                      int64_t lda, T *B, int64_t ldb) {                                                               \
        return t_trsm<T>(c, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);                                   \
    }                                                                                                                 \
    int hcb_##P##syrk(hcb_ctx *c, int uplo, int trans, int64_t n, int64_t k, T alpha, const T *A, int64_t lda, T beta,  \
                      T *Cm, int64_t ldc) {                                                                           \
        return t_syrk<T>(c, uplo, trans, n, k, alpha, A, lda, beta, Cm, ldc);                                        \
    }                                                                                                                 \
    int hcb_##P##fill_triangle(hcb_ctx *c, int uplo, int64_t n, T *A, int64_t lda, T value) {                          \
        return t_fill_triangle<T>(c, uplo, n, A, lda, value);                                                        \
    }                                                                                                                 \
    int hcb_##P##symmetrize(hcb_ctx *c, int uplo, int64_t n, T *A, int64_t lda) {                                      \
        return t_symmetrize<T>(c, uplo, n, A, lda);                                                                  \
    }                                                                                                                 \
    int hcb_##P##tlr_trsm_batched(hcb_ctx *c, int64_t n, const hcb_tile *X, const T *const *dL, const int64_t *ldl) {  \
        return t_tlr_trsm_batched<T>(c, n, X, dL, ldl);                                                              \
    }                                                                                                                 \
    int hcb_##P##tlr_syrk_batched(hcb_ctx *c, int64_t n, const hcb_tile *A, T *const *dC, const int64_t *ldc, T alpha,  \
                                  T beta) {                                                                           \
        return t_tlr_syrk_batched<T>(c, n, A, dC, ldc, alpha, beta);                                                 \
    }                                                                                                                 \
    int hcb_##P##tlr_potrf(hcb_ctx *c, int64_t nt, int64_t nb, T *const *diag, int64_t ldd, const hcb_tile *low,       \
                           const hcb_compress_params *p, int32_t *info, int32_t *potrf_info) {                        \
        return t_tlr_potrf<T>(c, nt, nb, diag, ldd, low, p, info, potrf_info);                                       \
    }                                                                                                                 \
    int hcb_##P##tlr_matmul_panel_step(hcb_ctx *c, int64_t mt, int64_t nt, const hcb_tile *Apan, const hcb_tile *Bpan, \
                                       const hcb_tile *C, T alpha, T beta, const hcb_compress_params *p, int32_t *info, \
                                       int first) {                                                                   \
        return t_tlr_matmul_panel_step<T>(c, mt, nt, Apan, Bpan, C, alpha, beta, p, info, first);                    \
    }                                                                                                                 \
    size_t hcb_##P##tlr_gemm_workspace(int64_t n_tiles, int64_t m, int64_t n, int64_t k, int64_t r_bound) {            \
        return t_workspace<T>(n_tiles, m, n, k, r_bound);                                                            \
    }

HCB_DEFINE_KERNEL_TABLE(d, double)
HCB_DEFINE_KERNEL_TABLE(s, float)

}  // extern "C"
