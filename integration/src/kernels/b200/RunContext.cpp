// src/kernels/b200/RunContext.cpp -- hcorepp::kernels::RunContext for the B200 backend.
//
// The class is the one the reference declares for its CUDA build (include/hcorepp/kernels/cuda/RunContext.hpp:15-53,
// compiled with -DUSE_CUDA); only its implementation file is exchanged (this one instead of
// src/kernels/cuda/RunContext.cpp).  What the new backend needs from a context is a stream -- libhcore_b200.so attaches
// its own hcb_ctx to that stream on first use (src/kernels/b200/kernels.cpp: ctx_of) -- plus the device info word
// HCoreKernels<T>::potrf reports through.  There is no cuSOLVER handle any more: the member stays null and
// GetCusolverDnHandle() returns it unchanged for source compatibility.
#include <hcorepp/kernels/cuda/RunContext.hpp>

namespace hcorepp {
namespace kernels {

namespace {
constexpr int kDevice = 0;
constexpr int64_t kQueueBatch = 2048;  // argument of blas::Queue (unused by the new kernels)
}  // namespace

// A root context owns its queue (= stream) and the info word; forked contexts share both and own nothing.
RunContext::RunContext()
    : mWorkBufferSize(0), mpWorkBuffer(nullptr), mpInfo(nullptr), mpQueue(std::make_shared<blas::Queue>(kDevice, kQueueBatch)),
      mCuSolverHandle(nullptr), mCuSolverOwner(true), mWorkSpaceOwner(false) {
    cudaMalloc(reinterpret_cast<void **>(&mpInfo), sizeof(int));
}

RunContext::RunContext(const RunContext &aParent)
    : mWorkBufferSize(0), mpWorkBuffer(nullptr), mpInfo(aParent.mpInfo), mpQueue(aParent.mpQueue), mCuSolverHandle(nullptr),
      mCuSolverOwner(false), mWorkSpaceOwner(false) {}

RunContext::~RunContext() {
    if (mWorkSpaceOwner && mpWorkBuffer) cudaFree(mpWorkBuffer);
    if (mCuSolverOwner) {  // root context: drain the stream before the info word goes away
        mpQueue->sync();
        cudaFree(mpInfo);
    }
}

RunContext RunContext::ForkChildContext() { return RunContext(*this); }

void RunContext::Sync() const { mpQueue->sync(); }

cudaStream_t RunContext::GetStream() const { return mpQueue->stream(); }

blas::Queue &RunContext::GetBLASQueue() const { return *mpQueue; }

cusolverDnHandle_t RunContext::GetCusolverDnHandle() const { return mCuSolverHandle; }

int *RunContext::GetInfoPointer() const { return mpInfo; }

// Grow-only scratch some reference call sites ask for when they are handed no workspace (CudaKernels.cu:544-554).
// The new kernels never need it (their scratch is the library's context arena); kept for source compatibility.
void *RunContext::RequestWorkBuffer(size_t aBufferSize) const {
    if (aBufferSize <= mWorkBufferSize) return mpWorkBuffer;
    if (mpWorkBuffer) cudaFree(mpWorkBuffer);
    cudaMalloc(&mpWorkBuffer, aBufferSize);
    mWorkBufferSize = aBufferSize;
    mWorkSpaceOwner = true;
    return mpWorkBuffer;
}

}  // namespace kernels
}  // namespace hcorepp
