// src/kernels/b200/kernels.cpp -- the file a maintainer of ecrc/hcorepp adds next to src/kernels/cuda/ to run the
// UNMODIFIED operator layer (src/api/HCore.cpp, src/operators/concrete/{Compressed,Dense}.cpp, ...) on libhcore_b200.so.
// It replaces src/kernels/cuda/kernels.cpp + CudaKernels.cu (cuBLAS via the BLAS++ queue, cuSOLVER, element-wise SIMT
// kernels): every member of hcorepp::kernels::HCoreKernels<T> (include/hcorepp/kernels/kernels.hpp:27-129) forwards to
// the C-ABI symbol of include/hcore_b200.h that replaces it.  Built with the reference's own -DUSE_CUDA headers
// (cuda/RunContext.hpp, cuda/memory.hpp stay as they are: they only need a stream); no cuBLAS / cuSOLVER is linked.
// tests/dropin/Makefile compiles exactly this arrangement and tests/dropin/dropin_test.cpp runs the reference's
// tile-at-a-time flow through it on the GPU.
#include <hcorepp/kernels/kernels.hpp>
#include <hcorepp/operators/helpers/CompressionParameters.hpp>
#include <hcore_b200.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

namespace hcorepp {
namespace kernels {

namespace {
// one hcb_ctx per CUDA stream of a RunContext (the reference's CUDA RunContext has no room for a new member without
// touching its header; a maintainer would add `hcb_ctx *mpB200` there instead of this table)
hcb_ctx *ctx_of(const RunContext &aContext) {
    static std::map<void *, hcb_ctx *> table;
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    void *stream = (void *) aContext.GetStream();
    auto it = table.find(stream);
    if (it != table.end()) return it->second;
    int device = 0;
    cudaGetDevice(&device);
    hcb_ctx *c = nullptr;
    if (hcb_ctx_create_on_stream(device, stream, &c) != HCB_OK)
        throw std::runtime_error(std::string("libhcore_b200: ") + hcb_last_error());
    table[stream] = c;
    return c;
}
inline void check(int rc, const char *what) {
    if (rc != HCB_OK) throw std::runtime_error(std::string(what) + ": " + hcb_last_error());
}
inline int op(blas::Op o) { return o == blas::Op::NoTrans ? 0 : 1; }

template<typename T> struct abi;
#define HCB_ADAPTER_ABI(P, T)                                                                                            \
    template<> struct abi<T> {                                                                                           \
        static constexpr auto gemm = hcb_##P##gemm; static constexpr auto multiply_by_alpha = hcb_##P##multiply_by_alpha;  \
        static constexpr auto process_v = hcb_##P##process_v; static constexpr auto new_rank = hcb_##P##new_rank;          \
        static constexpr auto uvptr = hcb_##P##uvptr; static constexpr auto vtnew = hcb_##P##vtnew;                       \
        static constexpr auto uvptr_conj = hcb_##P##uvptr_conj; static constexpr auto fill_identity = hcb_##P##fill_identity; \
        static constexpr auto lacpy = hcb_##P##lacpy; static constexpr auto laset = hcb_##P##laset;                       \
        static constexpr auto geqrf = hcb_##P##geqrf; static constexpr auto ungqr = hcb_##P##ungqr;                       \
        static constexpr auto unmqr = hcb_##P##unmqr; static constexpr auto svd = hcb_##P##svd;                           \
        static constexpr auto trmm = hcb_##P##trmm; static constexpr auto potrf = hcb_##P##potrf;                         \
        static constexpr auto trsm = hcb_##P##trsm; static constexpr auto syrk = hcb_##P##syrk;                           \
        static constexpr auto fill_triangle = hcb_##P##fill_triangle; static constexpr auto symmetrize = hcb_##P##symmetrize; \
        static constexpr auto transpose = hcb_##P##transpose;                                                            \
    };
HCB_ADAPTER_ABI(d, double)
HCB_ADAPTER_ABI(s, float)
#undef HCB_ADAPTER_ABI
}  // namespace

template<typename T>
void HCoreKernels<T>::Gemm(blas::Layout aLayout, blas::Op aTransA, blas::Op aTransB, size_t aM, size_t aN, size_t aK, T &aAlpha,
                           T const *apA, size_t aLdA, T const *apB, size_t aLdB, T &aBeta, T *apC, size_t aLdC,
                           const RunContext &aContext) {   // was blas::gemm(..., queue)  cuda/kernels.cpp:18-24
    if (aLayout == blas::Layout::ColMajor)
        check(abi<T>::gemm(ctx_of(aContext), op(aTransA), op(aTransB), aM, aN, aK, aAlpha, apA, aLdA, apB, aLdB, aBeta, apC, aLdC), "Gemm");
    else  // row-major C = op(A) op(B)  <=>  column-major C^T = op(B)^T op(A)^T
        check(abi<T>::gemm(ctx_of(aContext), op(aTransB), op(aTransA), aN, aM, aK, aAlpha, apB, aLdB, apA, aLdA, aBeta, apC, aLdC), "Gemm");
}

template<typename T>
void HCoreKernels<T>::MultiplyByAlpha(T *apArray, size_t aRows, size_t aCols, size_t aM, size_t aRank, T &aAlpha,
                                      const RunContext &aContext) {
    check(abi<T>::multiply_by_alpha(ctx_of(aContext), apArray, aRows, aCols, aM, aRank, aAlpha), "MultiplyByAlpha");
}

template<typename T>
void HCoreKernels<T>::ProcessVpointer(size_t aN, size_t aCRank, bool aGetUngqr, size_t Vm, T &aBeta, T *apCV, size_t aLdcV, T *V,
                                      size_t aArank, const T *apBdata, const RunContext &aContext, bool aCholesky) {
    check(abi<T>::process_v(ctx_of(aContext), aN, aCRank, aGetUngqr, Vm, aBeta, apCV, aLdcV, V, aArank, apBdata, aCholesky), "ProcessVpointer");
}

template<typename T>
void HCoreKernels<T>::CalculateNewRank(size_t &aNewRank, bool aTruncatedSvd, blas::real_type<T> *apSigma, size_t sizeS,
                                       blas::real_type<T> accuracy, const RunContext &aContext) {
    int64_t r = 0;  // host result: one sync per tile, like CudaKernels.cu:656-697 (the fused path keeps the rank on the device)
    check(abi<T>::new_rank(ctx_of(aContext), aTruncatedSvd, apSigma, sizeS, accuracy, &r), "CalculateNewRank");
    aNewRank = (size_t) r;
}

template<typename T>
void HCoreKernels<T>::CalculateUVptr(size_t aRank, size_t aVm, T *UVptr, const T *Vnew, const RunContext &aContext) {
    check(abi<T>::uvptr(ctx_of(aContext), aRank, aVm, UVptr, Vnew), "CalculateUVptr");
}

template<typename T>
void HCoreKernels<T>::CalculateVTnew(size_t aRkNew, bool aUngqr, size_t aMinVmVn, blas::real_type<T> *apSigma, T *apVTnew,
                                     size_t aSizeS, size_t aVm, const RunContext &aContext) {
    check(abi<T>::vtnew(ctx_of(aContext), aRkNew, aUngqr, aMinVmVn, apSigma, apVTnew, aSizeS, aVm), "CalculateVTnew");
}

template<typename T>
void HCoreKernels<T>::CalculateUVptrConj(size_t aRank, size_t aVm, T *UVptr, const RunContext &aContext) {
    check(abi<T>::uvptr_conj(ctx_of(aContext), aRank, aVm, UVptr), "CalculateUVptrConj");
}

template<typename T>
void HCoreKernels<T>::FillIdentityMatrix(size_t aNumOfElements, T *apMatrix, const RunContext &aContext) {
    check(abi<T>::fill_identity(ctx_of(aContext), aNumOfElements, apMatrix), "FillIdentityMatrix");
}

template<typename T>
void HCoreKernels<T>::LaCpy(common::MatrixType aType, size_t aM, size_t aRank, T *apCU, size_t aLD, T *apU, size_t aUm,
                            const RunContext &aContext) {
    check(abi<T>::lacpy(ctx_of(aContext), (int) aType, aM, aRank, apCU, aLD, apU, aUm), "LaCpy");
}

template<typename T>
void HCoreKernels<T>::Geqrf(size_t aM, size_t aN, T *apA, size_t aLdA, T *apTau, T *, size_t, size_t, const RunContext &aContext) {
    check(abi<T>::geqrf(ctx_of(aContext), aM, aN, apA, aLdA, apTau), "Geqrf");   // was cusolverDnXgeqrf  CudaKernels.cu:534-561
}

template<typename T>
void HCoreKernels<T>::Laset(common::MatrixType aMatrixType, size_t aM, size_t aN, T aOffdiag, T aDiag, T *apA, size_t aLdA,
                            const RunContext &aContext) {
    check(abi<T>::laset(ctx_of(aContext), (int) aMatrixType, aM, aN, aOffdiag, aDiag, apA, aLdA), "Laset");
}

template<typename T>
void HCoreKernels<T>::Trmm(blas::Layout, blas::Side aSide, blas::Uplo aUplo, blas::Op aTrans, blas::Diag aDiag, size_t aM, size_t aN,
                           T aAlpha, T const *apA, size_t aLdA, T *apB, size_t aLdB, const RunContext &aContext) {
    check(abi<T>::trmm(ctx_of(aContext), (int) aSide, (int) aUplo, (int) aTrans, (int) aDiag, aM, aN, aAlpha, apA, aLdA, apB, aLdB), "Trmm");
}

template<typename T>
void HCoreKernels<T>::SVD(common::Job, common::Job, size_t aM, size_t aN, T *apA, size_t aLdA, T *apS, T *apU, size_t aLdU, T *apVT,
                          size_t aLdVt, common::CompressionType, T *, size_t, size_t, const RunContext &aContext) {
    check(abi<T>::svd(ctx_of(aContext), aM, aN, apA, aLdA, apS, apU, aLdU, apVT, aLdVt), "SVD");  // was cusolverDnXgesvd  CudaKernels.cu:699-730
}

template<typename T>
void HCoreKernels<T>::Unmqr(common::SideMode aSide, common::BlasOperation aTrans, size_t aM, size_t aN, size_t aK, T const *apA,
                            size_t aLdA, T const *apTau, T *apC, size_t aLdC, T *, size_t, const RunContext &aContext) {
    check(abi<T>::unmqr(ctx_of(aContext), (int) aSide, aTrans == common::OP_NoTRANS ? 0 : 1, aM, aN, aK, apA, aLdA, apTau, apC, aLdC), "Unmqr");
}

template<typename T>
void HCoreKernels<T>::ungqr(size_t aM, size_t aN, size_t aK, T *apA, size_t aLdA, T *apTau, T *, size_t, const RunContext &aContext) {
    check(abi<T>::ungqr(ctx_of(aContext), aM, aN, aK, apA, aLdA, apTau), "ungqr");
}

template<typename T>
size_t HCoreKernels<T>::CalculateGemmWorkspaceSize(size_t, size_t, size_t, size_t, size_t, const operators::CompressionParameters &,
                                                   size_t &aHostSize, const RunContext &) {
    aHostSize = 0;  // scratch lives in the library's context arena; callers need not provide any (cuda/kernels.cpp:137-265)
    return 0;
}

template<typename T>
int HCoreKernels<T>::potrf(blas::Uplo aUplo, T *, size_t, size_t, size_t aMatrixOrder, T *apMatrix, size_t aLeadingDim, blas::Layout,
                           const kernels::RunContext &aContext) {
    int *d_info = aContext.GetInfoPointer();
    check(abi<T>::potrf(ctx_of(aContext), (int) aUplo, aMatrixOrder, apMatrix, aLeadingDim, d_info), "potrf");
    int info = 0;
    cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, aContext.GetStream());
    aContext.Sync();
    return info;
}

template<typename T>
void HCoreKernels<T>::FillMatrixTriangle(blas::Uplo aUplo, size_t aRows, size_t aCols, T *apMatrix, blas::Layout, size_t aValue,
                                         const kernels::RunContext &aContext) {
    if (aRows != aCols) return;
    check(abi<T>::fill_triangle(ctx_of(aContext), (int) aUplo, aRows, apMatrix, aRows, (T) aValue), "FillMatrixTriangle");
}

template<typename T>
void HCoreKernels<T>::trsm(blas::Layout, blas::Side aSide, blas::Uplo aUplo, blas::Op aTrans, blas::Diag aDiag, size_t aRows, size_t aCols,
                           T aAlpha, T const *apMatrixA, size_t aLeadingDimA, T *apMatrixB, size_t aLeadingDimB,
                           const kernels::RunContext &aContext) {
    check(abi<T>::trsm(ctx_of(aContext), (int) aSide, (int) aUplo, op(aTrans), (int) aDiag, aRows, aCols, aAlpha, apMatrixA, aLeadingDimA,
                       apMatrixB, aLeadingDimB), "trsm");
}

template<typename T>
void HCoreKernels<T>::syrk(blas::Layout, blas::Uplo aUplo, blas::Op aTrans, size_t aRows, size_t aCols, T aAlpha, T const *apMatrixA,
                           size_t aLeadingDimA, T aBeta, T *apMatrixB, size_t aLeadingDimB, const kernels::RunContext &aContext) {
    check(abi<T>::syrk(ctx_of(aContext), (int) aUplo, op(aTrans), aRows, aCols, aAlpha, apMatrixA, aLeadingDimA, aBeta, apMatrixB,
                       aLeadingDimB), "syrk");
}

template<typename T>
void HCoreKernels<T>::Symmetrize(blas::Layout, T *apMatrixA, size_t aRows, size_t aCols, blas::Uplo aUplo, const RunContext &aContext) {
    if (aRows != aCols) return;
    check(abi<T>::symmetrize(ctx_of(aContext), (int) aUplo, aRows, apMatrixA, aRows), "Symmetrize");
}

template<typename T>
void HCoreKernels<T>::transpose(blas::Layout aLayout, size_t aRows, size_t aCols, const T *aA, size_t aLeadingDimA, T *aOut,
                                size_t aLeadingDimOut, const kernels::RunContext &aContext) {
    if (aA == nullptr || aOut == nullptr) return;
    if (aLayout == blas::Layout::RowMajor) { const size_t t = aRows; aRows = aCols; aCols = t; }
    check(abi<T>::transpose(ctx_of(aContext), aRows, aCols, aA, aLeadingDimA, aOut, aLeadingDimOut), "transpose");
}

template<typename T>
size_t HCoreKernels<T>::CalculatePotrfWorkspaceSize(T *, blas::Uplo, size_t, size_t, size_t &aHostSize, const RunContext &) {
    aHostSize = 0;
    return 0;
}

HCOREPP_INSTANTIATE_CLASS(HCoreKernels)

}  // namespace kernels
}  // namespace hcorepp
