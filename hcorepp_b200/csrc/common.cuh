// common.cuh -- shared device/host helpers for libhcore_b200 (sm_100a only; no multi-backend dispatch).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>
#include <mutex>
#include "../../include/hcore_b200.h"

namespace hcb {

// ---------------------------------------------------------------------------------------------------------------
// error handling: C ABI returns codes; the text of the last failure is kept per thread
// ---------------------------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define HCB_CUDA(call)                                                                                             \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            return ::hcb::fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? HCB_ENODEVICE        \
                                                                                          : HCB_ECUDA,             \
                               std::string(#call) + ": " + cudaGetErrorString(e_));                                \
    } while (0)

#define HCB_LAUNCH_CHECK(name)                                                                                     \
    do {                                                                                                           \
        ::hcb::g_launches.fetch_add(1, std::memory_order_relaxed);                                                 \
        cudaError_t e_ = cudaGetLastError();                                                                       \
        if (e_ != cudaSuccess) return ::hcb::fail(HCB_ECUDA, std::string(name) + ": " + cudaGetErrorString(e_));   \
    } while (0)

#define HCB_TRY(expr)                                                                                              \
    do {                                                                                                           \
        int rc_ = (expr);                                                                                          \
        if (rc_ != HCB_OK) return rc_;                                                                             \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return (int) ((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------------------------
// device-side reductions
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole CTA; every thread gets the result. `red` = shared scratch of >= 33 elements.
template<typename T>
__device__ __forceinline__ T block_sum(T v, T *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        T t = (lane < nw) ? red[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// d_info word of the fused path: low byte = sticky flags (OR), bits 8..15 = Jacobi sweeps (MAXIMUM over the steps that
// shared the word, e.g. the k loop of tlr_matmul)
__device__ __forceinline__ void info_max_sweeps(int *info, int sweeps) {
    int old = *reinterpret_cast<volatile int *>(info);
    for (;;) {
        const int cur = (old >> 8) & 0xff;
        if (cur >= sweeps) return;
        const int want = (old & ~0xff00) | ((sweeps & 0xff) << 8);
        const int seen = atomicCAS(info, old, want);
        if (seen == old) return;
        old = seen;
    }
}

template<typename T> struct Eps;
template<> struct Eps<double> { static __host__ __device__ constexpr double v() { return 2.220446049250313e-16; } };
template<> struct Eps<float> { static __host__ __device__ constexpr float v() { return 1.1920929e-07f; } };

template<typename T> __device__ __forceinline__ T t_sqrt(T x);
template<> __device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
template<> __device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
template<typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }

// ---------------------------------------------------------------------------------------------------------------
// device-side problem descriptors (built ON THE DEVICE by k_setup from tile descriptors + device-resident ranks)
// ---------------------------------------------------------------------------------------------------------------
template<typename T>
struct GemmProb {  // C = alpha*op(A)*op(B) + beta*C ; m == 0 means "skip"
    const T *A;
    const T *B;
    T *C;
    int m, n, k;
    int lda, ldb, ldc;
    int ta, tb;
    T alpha, beta;
    // optional second A segment: op(A) = [A | A2] along k, split at k1 (ta == 0 only; A2 == nullptr: none)
    const T *A2;
    int k1, lda2;
};

template<typename T>
struct CopyProb {  // dst[i + j*ldd] = scale * (trans ? src[j + i*lds] : src[i + j*lds]),  i < rows, j < cols
    const T *src;
    T *dst;
    int rows, cols;
    int lds, ldd;
    int trans;
    T scale;
};

template<typename T>
struct QrProb {  // Householder QR in place (LAPACK layout): A (m x n, lda), tau[min(m,n)]
    T *A;
    T *tau;
    int m, n, lda;
};

template<typename T>
struct ReflProb {  // apply k reflectors stored in V (LAPACK layout, ldv) + tau to C (mc x nc, ldc)
    const T *V;
    const T *tau;
    T *C;
    int mv;          // reflector length (rows of V)
    int k;           // number of reflectors
    int ldv;
    int mc, nc, ldc;
    int side_right;  // 0: C := op(Q) C (columns of C independent) ; 1: C := C op(Q) (rows of C independent)
    int forward;     // 1: apply H_0 first ... H_{k-1} last ; 0: H_{k-1} first ... H_0 last
    const int *nc_dev;  // optional: device-resident column (or row) count overriding nc / mc (data-dependent rank)
};

template<typename T>
struct SvdProb {  // one-sided Jacobi on M (a x b, a >= b): M = Uout * diag(sigma) * V^T, sigma descending
    T *M;         // input (global), preserved
    T *J;         // work (global) a*b elements: rotated copy of M when the problem does not fit shared memory
    T *Uout;      // a x b, ld ldu : normalised left vectors (zero columns where sigma == 0)
    T *Vout;      // b x b, ld ldv : receives V diag(sigma) = M^T Uout from the follow-up GEMM
    T *sigma;     // b
    int *info;    // optional: |= 1 when not converged
    int a, b, ldm, ldu, ldv;
};

template<typename T>
struct LqProb {   // LQ preconditioning: L = R^T of the QR-factored transpose (see k_extract_l)
    const T *MT;  // QR-factored M^T (b x a, ld b): R in the upper trapezoid
    T *Lb;        // a x b (ld a) receives L = R^T
    int a, b;
};

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
struct ParamRing {  // pinned host staging + device mirror for small descriptor uploads
    // Four segments, one event each: a segment is re-used only after the copies issued from it on the previous lap have
    // completed (event wait, normally long satisfied) -- never a stream synchronise in steady state.
    static constexpr int NSEG = 4;
    char *h = nullptr;
    char *d = nullptr;
    size_t cap = 0, off = 0;
    int seg = 0;
    cudaEvent_t ev[NSEG] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_pending[NSEG] = {false, false, false, false};
};

}  // namespace hcb

struct hcb_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void *ws = nullptr;  // grow-only scratch arena
    size_t ws_bytes = 0;
    void *ws2 = nullptr;  // second grow-only arena: FP64 shadow copies of FP32 tiles (promoted fused path)
    size_t ws2_bytes = 0;
    void *info_tmp = nullptr;  // group-ordered info words + permutation of a mixed-mix batch (grow-only)
    size_t info_tmp_bytes = 0;
    int *svd_sched = nullptr;  // work counters of the persistent Jacobi kernel (2 + problems ints, grow-only)
    size_t svd_sched_n = 0;
    hcb::ParamRing ring;
    unsigned long long *d_stats = nullptr;  // device counters of the adaptive fast paths (hcb_ctx_stats), 8 entries
    int *d_err = nullptr;      // device-side sticky error word of the fused path (bit 2: a rank exceeded its bound)
    int *h_err = nullptr;      // pinned mirror read by hcb_ctx_sync
    // optional per-phase device timing of the fused path (CUDA events on this stream; see hcb_ctx_phase_timing)
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct PhaseRec { int phase; cudaEvent_t beg, end; };
    std::vector<PhaseRec> phase_recs;
};

namespace hcb {
// Copies `bytes` of host descriptors to the device through the pinned ring; returns the device address.
int ring_upload(hcb_ctx *ctx, const void *host, size_t bytes, void **d_out);
int ensure_ws(hcb_ctx *ctx, size_t bytes);
// Ampere-style asynchronous global -> shared copies (LDGSTS): 16-byte (L2 only) and 8-byte variants
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// CTA-scope mbarrier helpers (producer/consumer hand-off between warps without a block-wide barrier)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"((unsigned) __cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {  // release.cta
    asm volatile("mbarrier.arrive.shared.b64 _, [%0];\n" ::"r"((unsigned) __cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long *bar, unsigned parity) {  // acquire.cta; phase k <-> parity k & 1
    unsigned ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(ok)
                 : "r"((unsigned) __cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    // watchdog: a protocol error must abort the kernel (sticky launch failure), never hang the GPU
    for (unsigned spins = 0; !mbar_try(bar, parity); ++spins)
        if (spins > (1u << 22)) __trap();
}

}  // namespace hcb
