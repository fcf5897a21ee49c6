// hcorepp.hpp -- C++ host layer that mirrors the reference's public API for the TLR-GEMM path on top of the C ABI
// (include/hcore_b200.h).  Same namespaces, class names, argument order and exception behaviour as ecrc/hcorepp, so
// that code written against
//     hcorepp::api::HCore<T>::Gemm / CalculateMemoryPoolSize        include/hcorepp/api/HCore.hpp:41-63
//     hcorepp::operators::{Tile, DenseTile, CompressedTile}          include/hcorepp/operators/**
//     hcorepp::operators::CompressionParameters                      .../helpers/CompressionParameters.hpp:44-46
//     hcorepp::kernels::{RunContext, ContextManager, HCoreKernels}   include/hcorepp/kernels/**
//     hcorepp::memory::{AllocateArray, DestroyArray, Memcpy, Memset}  include/hcorepp/kernels/memory.hpp:46-121
//     hcorepp::dataunits::{DataHolder, MemoryUnit}                   include/hcorepp/data-units/**
// builds against this header and libhcore_b200.so instead of the reference's cuBLAS/cuSOLVER backend.
// Header-only, no CUDA headers needed (everything crosses the C ABI).  Real types only (float, double), like the
// reference (Definitions.hpp:10-13).  There is no CPU fallback: without a CUDA device constructors throw.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../hcore_b200.h"

// The reference's signatures use these BLAS++ enums (blas/util.hh, v2023.01.00); values are BLAS++'s public ones.
#ifndef HCOREPP_B200_NO_BLAS_ENUMS
namespace blas {
enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
enum class Uplo : char { Upper = 'U', Lower = 'L', General = 'G' };
enum class Diag : char { NonUnit = 'N', Unit = 'U' };
enum class Side : char { Left = 'L', Right = 'R' };
template<typename T> using real_type = T;
}  // namespace blas
#endif

namespace hcorepp {

namespace detail {
inline void check(int rc, const char *where) {
    if (rc != HCB_OK) throw std::runtime_error(std::string(where) + ": " + hcb_last_error());
}
template<typename T> struct abi;  // maps T to the hcb_{d,s}* symbols
#define HCOREPP_B200_ABI(P, T)                                                                                          \
    template<> struct abi<T> {                                                                                          \
        static constexpr auto gemm = hcb_##P##gemm;                                                                     \
        static constexpr auto multiply_by_alpha = hcb_##P##multiply_by_alpha;                                           \
        static constexpr auto process_v = hcb_##P##process_v;                                                           \
        static constexpr auto new_rank = hcb_##P##new_rank;                                                             \
        static constexpr auto uvptr = hcb_##P##uvptr;                                                                   \
        static constexpr auto vtnew = hcb_##P##vtnew;                                                                   \
        static constexpr auto uvptr_conj = hcb_##P##uvptr_conj;                                                         \
        static constexpr auto fill_identity = hcb_##P##fill_identity;                                                   \
        static constexpr auto lacpy = hcb_##P##lacpy;                                                                   \
        static constexpr auto laset = hcb_##P##laset;                                                                   \
        static constexpr auto geqrf = hcb_##P##geqrf;                                                                   \
        static constexpr auto ungqr = hcb_##P##ungqr;                                                                   \
        static constexpr auto unmqr = hcb_##P##unmqr;                                                                   \
        static constexpr auto svd = hcb_##P##svd;                                                                       \
        static constexpr auto trmm = hcb_##P##trmm;                                                                     \
        static constexpr auto tlr_gemm_batched = hcb_##P##tlr_gemm_batched;                                             \
        static constexpr auto compress_batched = hcb_##P##compress_batched;                                             \
        static constexpr auto tlr_matmul = hcb_##P##tlr_matmul;                                                         \
        static constexpr auto tlr_gemm_workspace = hcb_##P##tlr_gemm_workspace;                                         \
        static constexpr auto potrf = hcb_##P##potrf;                                                                   \
        static constexpr auto trsm = hcb_##P##trsm;                                                                     \
        static constexpr auto syrk = hcb_##P##syrk;                                                                     \
        static constexpr auto fill_triangle = hcb_##P##fill_triangle;                                                   \
        static constexpr auto symmetrize = hcb_##P##symmetrize;                                                         \
        static constexpr auto tlr_trsm_batched = hcb_##P##tlr_trsm_batched;                                             \
        static constexpr auto tlr_syrk_batched = hcb_##P##tlr_syrk_batched;                                             \
        static constexpr auto tlr_potrf = hcb_##P##tlr_potrf;                                                           \
    };
HCOREPP_B200_ABI(d, double)
HCOREPP_B200_ABI(s, float)
#undef HCOREPP_B200_ABI
}  // namespace detail

namespace common {  // include/hcorepp/common/Definitions.hpp:21-104
enum SideMode { SIDE_LEFT = 0, SIDE_RIGHT = 1 };
enum BlasOperation { OP_NoTRANS = 0, OP_TRANS = 1, OP_C = 2, OP_HERMITAN = 2, OP_CONJG = 3 };
enum class Job { NoVec = 'N', Vec = 'V', UpdateVec = 'U', AllVec = 'A', SomeVec = 'S', OverwriteVec = 'O' };
enum class MatrixType { General = 'G', Lower = 'L', Upper = 'U' };
enum CompressionType { LAPACK_GESVD, LAPACK_GESDD };
enum class MemoryHandlerStrategy { ONDEMAND, POOL };
}  // namespace common

// ---------------------------------------------------------------------------------------------------------------------
namespace kernels {

/// CUDA run context (cuda/RunContext.hpp:15-53): one device + one stream; every kernel is enqueued on GetStream().
class RunContext {
public:
    explicit RunContext(int aDevice = 0) { detail::check(hcb_ctx_create(aDevice, &mpCtx), "RunContext"); }
    RunContext(const RunContext &) = delete;
    RunContext &operator=(const RunContext &) = delete;
    RunContext(RunContext &&o) noexcept : mpCtx(o.mpCtx) { o.mpCtx = nullptr; }
    ~RunContext() { if (mpCtx) hcb_ctx_destroy(mpCtx); }
    void Sync() const { detail::check(hcb_ctx_sync(mpCtx), "RunContext::Sync"); }
    void *GetStream() const { return hcb_ctx_stream(mpCtx); }
    [[nodiscard]] bool SupportsOMP() const { return false; }
    RunContext ForkChildContext() const { return RunContext(hcb_ctx_device(mpCtx)); }
    hcb_ctx *Handle() const { return mpCtx; }
private:
    hcb_ctx *mpCtx = nullptr;
};

/// ContextManager (ContextManager.hpp:18-74): lazily created contexts, index 0 is the main one.
class ContextManager {
public:
    static ContextManager &GetInstance() { static ContextManager m; return m; }
    RunContext &GetContext(size_t aIdx = 0) {
        if (aIdx >= MaxContexts) throw std::runtime_error("Trying to fetch invalid Context Idx");
        if (mContexts.size() <= aIdx) mContexts.resize(aIdx + 1);
        if (!mContexts[aIdx]) mContexts[aIdx].reset(new RunContext(0));
        return *mContexts[aIdx];
    }
    void SyncMainContext() { GetContext(0).Sync(); }
    void SyncAll() { for (auto &c : mContexts) if (c) c->Sync(); }
    static void DestroyInstance() { GetInstance().mContexts.clear(); }
    static constexpr size_t MaxContexts = 50;  // MAX_NUM_STREAMS, ContextManager.hpp:10
private:
    std::vector<std::unique_ptr<RunContext>> mContexts;
};

}  // namespace kernels

// ---------------------------------------------------------------------------------------------------------------------
namespace memory {  // include/hcorepp/kernels/memory.hpp:22-121
enum class MemoryTransfer { HOST_TO_DEVICE, DEVICE_TO_DEVICE, DEVICE_TO_HOST, HOST_TO_HOST, AUTOMATIC };

template<typename T>
T *AllocateArray(size_t aNumElements, const kernels::RunContext &aContext) {
    void *p = nullptr;
    detail::check(hcb_malloc(aContext.Handle(), aNumElements * sizeof(T), &p), "AllocateArray");
    return static_cast<T *>(p);
}
template<typename T>
void DestroyArray(T *apArray, const kernels::RunContext &aContext) {
    if (apArray) detail::check(hcb_free(aContext.Handle(), apArray), "DestroyArray");
}
template<typename T>
void Memcpy(T *apDestination, const T *apSrcDataArray, size_t aNumOfElements, const kernels::RunContext &aContext,
            MemoryTransfer aTransferType = MemoryTransfer::DEVICE_TO_DEVICE, bool aBlocking = false) {
    detail::check(hcb_memcpy(aContext.Handle(), apDestination, apSrcDataArray, aNumOfElements * sizeof(T),
                             (int) aTransferType), "Memcpy");
    if (aBlocking) aContext.Sync();
}
template<typename T>
void Memset(T *apDestination, char aValue, size_t aNumOfElements, const kernels::RunContext &aContext) {
    detail::check(hcb_memset(aContext.Handle(), apDestination, aValue, aNumOfElements * sizeof(T)), "Memset");
}
}  // namespace memory

// ---------------------------------------------------------------------------------------------------------------------
namespace operators {

/// CompressionParameters.hpp:44-46
class CompressionParameters {
public:
    CompressionParameters(double aAccuracy = 1e-4, bool aUseTrmm = false, bool aUseUngqr = true, bool aTruncatedSvd = false,
                          size_t aFixedRank = 0, common::CompressionType aOpType = common::CompressionType::LAPACK_GESDD)
        : mUseTrmm(aUseTrmm), mUseUngqr(aUseUngqr), mTruncatedSvd(aTruncatedSvd), mFixedRank(aFixedRank), mOpType(aOpType),
          mAccuracy(aAccuracy) {}
    bool GetTrmm() const { return mUseTrmm; }
    bool GetUngqr() const { return mUseUngqr; }
    bool GetTruncatedSvd() const { return mTruncatedSvd; }
    size_t GetFixedRank() const { return mFixedRank; }
    common::CompressionType GetOperationType() const { return mOpType; }
    double GetAccuracy() const { return mAccuracy; }
    hcb_compress_params ToC() const {
        return hcb_compress_params{mAccuracy, mUseTrmm, mUseUngqr, mTruncatedSvd, (int64_t) mFixedRank,
                                   mOpType == common::LAPACK_GESVD ? 0 : 1, 0};
    }
private:
    bool mUseTrmm, mUseUngqr, mTruncatedSvd;
    size_t mFixedRank;
    common::CompressionType mOpType;
    double mAccuracy;
};
}  // namespace operators

// ---------------------------------------------------------------------------------------------------------------------
namespace dataunits {

/// DataHolder.hpp:29-147: (rows, cols, ld, device ptr, owns?) -- the owning constructor allocates and copies / zero-fills.
template<typename T>
class DataHolder {
public:
    DataHolder(size_t aRows, size_t aCols, size_t aLeadingDim, T *apData, const kernels::RunContext &aContext,
               bool aMemoryOwnership = true)
        : mRows(aRows), mCols(aCols), mLd(aLeadingDim), mOwns(aMemoryOwnership), mContext(aContext) {
        if (!mOwns) { mpData = apData; return; }
        mpData = memory::AllocateArray<T>(aRows * aCols, aContext);
        if (apData) {  // host or device source (cudaMemcpyDefault): compact the ld on the way in
            for (size_t j = 0; j < aCols; ++j)
                memory::Memcpy<T>(mpData + j * aRows, apData + j * aLeadingDim, aRows, aContext, memory::MemoryTransfer::AUTOMATIC);
            aContext.Sync();
            mLd = aRows;
        } else {
            memory::Memset<T>(mpData, 0, aRows * aCols, aContext);
            mLd = aRows;
        }
    }
    DataHolder(const DataHolder &) = delete;
    ~DataHolder() { if (mOwns && mpData) hcb_free(mContext.Handle(), mpData); }
    T *GetData() const { return mpData; }
    size_t GetNumOfRows() const { return mRows; }
    size_t GetNumOfCols() const { return mCols; }
    size_t GetLeadingDim() const { return mLd; }
    void Resize(size_t aRows, size_t aCols, size_t aLeadingDim) {
        if (mOwns && aRows * aCols > mRows * mCols) {
            T *p = memory::AllocateArray<T>(aRows * aCols, mContext);
            memory::Memset<T>(p, 0, aRows * aCols, mContext);
            memory::Memcpy<T>(p, mpData, mRows * mCols, mContext);
            mContext.Sync();
            hcb_free(mContext.Handle(), mpData);
            mpData = p;
        }
        mRows = aRows; mCols = aCols; mLd = aLeadingDim;
    }
private:
    T *mpData = nullptr;
    size_t mRows, mCols, mLd;
    bool mOwns;
    const kernels::RunContext &mContext;
};

/// MemoryUnit (pool/MemoryHandler.hpp:21-157).  The fused path keeps its scratch in the context's grow-only arena, so
/// the unit only records what was asked of it; it never memsets (the reference's Reset() memsets the whole pool).
template<typename T>
class MemoryUnit {
public:
    explicit MemoryUnit(const kernels::RunContext &aContext, size_t aPoolSize = 0) : mContext(aContext) {
        if (aPoolSize) Initialize(aPoolSize);
    }
    [[nodiscard]] bool IsInitialized() const { return mInitialized; }
    void Initialize(size_t aSize) {
        detail::check(hcb_ctx_reserve_workspace(mContext.Handle(), aSize * sizeof(T)), "MemoryUnit::Initialize");
        mInitialized = true;
    }
    void BufferMemSet(char, size_t, size_t = 0) {}
    void FreeAllocations() { mInitialized = false; }
    void Reset() {}
    common::MemoryHandlerStrategy GetStrategy() { return common::MemoryHandlerStrategy::POOL; }
private:
    const kernels::RunContext &mContext;
    bool mInitialized = false;
};

template<typename T>
class MemoryHandler {  // pool/MemoryHandler.hpp:171-196: one unit per context
public:
    static MemoryHandler &GetInstance() { static MemoryHandler h; return h; }
    MemoryUnit<T> &GetMemoryUnit(size_t aIdx = 0) {
        if (mUnits.size() <= aIdx) mUnits.resize(aIdx + 1);
        if (!mUnits[aIdx]) mUnits[aIdx].reset(new MemoryUnit<T>(kernels::ContextManager::GetInstance().GetContext(aIdx)));
        return *mUnits[aIdx];
    }
    void FreeAllocations() { for (auto &u : mUnits) if (u) u->FreeAllocations(); }
    static void DestroyInstance() { GetInstance().mUnits.clear(); }
private:
    std::vector<std::unique_ptr<MemoryUnit<T>>> mUnits;
};
}  // namespace dataunits

// ---------------------------------------------------------------------------------------------------------------------
namespace operators {

enum TileType { DENSE, COMPRESSED };

struct TileMetadata {  // Tile.hpp:30-52: the (metadata, buffer) wire format
    size_t mNumOfRows, mNumOfCols, mMatrixRank, mMaxRank, mLeadingDimension;
    blas::Layout mLayout;
    TileType mType;
    TileMetadata(size_t r, size_t c, size_t rank, size_t maxRank, size_t ld, blas::Layout layout, TileType type)
        : mNumOfRows(r), mNumOfCols(c), mMatrixRank(rank), mMaxRank(maxRank), mLeadingDimension(ld), mLayout(layout), mType(type) {}
};

template<typename T>
class Tile {  // Tile.hpp:124-277
public:
    virtual ~Tile() = default;
    blas::Layout GetLayout() const { return mLayout; }
    size_t GetNumOfRows() const { return mNumOfRows; }
    size_t GetNumOfCols() const { return mNumOfCols; }
    size_t GetLeadingDim() const { return mLeadingDim; }
    virtual size_t GetTileRank() const { return mRank; }
    std::reference_wrapper<dataunits::DataHolder<T>> GetDataHolder() const { return *mpDataArray; }
    virtual T *GetTileSubMatrix(size_t aIndex) const = 0;
    virtual size_t GetTileStride(size_t aIndex) const = 0;
    virtual bool isDense() const = 0;
    virtual bool isCompressed() const = 0;
    virtual int64_t GetNumOfSubMatrices() const = 0;
    virtual TileType GetTileType() = 0;
    virtual hcb_tile Descriptor() const = 0;  // what crosses the C ABI
    /// Tile.hpp:214,231 -- the (metadata, buffer) wire format: UnPackTile hands out a NEW TileMetadata (the caller deletes
    /// it) and a BORROWED data pointer; PackTile adopts a buffer without taking ownership (Compressed.cpp:769-805).
    virtual std::pair<TileMetadata *, T *> UnPackTile(const kernels::RunContext &aContext) = 0;
    virtual void PackTile(TileMetadata aMetadata, T *aDataArray, const kernels::RunContext &aContext) = 0;
protected:
    blas::Layout mLayout = blas::Layout::ColMajor;
    size_t mLeadingDim = 0, mNumOfRows = 0, mNumOfCols = 0, mRank = 0, mMaxRank = 0;
    dataunits::DataHolder<T> *mpDataArray = nullptr;
};

template<typename T>
class DenseTile : public Tile<T> {  // Dense.hpp:56-91
public:
    DenseTile(size_t aNumOfRows, size_t aNumOfCols, T *aPdata, size_t aLeadingDim, blas::Layout aLayout,
              const kernels::RunContext &aContext, bool aMemoryOwnership = true) {
        // RowMajor (Dense.cpp:69-96 passes the tile's layout to the GEMM): an m x n row-major buffer IS the column-major
        // n x m transpose, which is how it is held; HCore::Gemm then computes C^T = op(B)^T op(A)^T on those views.
        const bool rm = aLayout == blas::Layout::RowMajor;
        this->mLayout = aLayout; this->mNumOfRows = aNumOfRows; this->mNumOfCols = aNumOfCols; this->mRank = 0;
        this->mpDataArray = new dataunits::DataHolder<T>(rm ? aNumOfCols : aNumOfRows, rm ? aNumOfRows : aNumOfCols, aLeadingDim, aPdata,
                                                         aContext, aMemoryOwnership);
        this->mLeadingDim = this->mpDataArray->GetLeadingDim();
    }
    DenseTile(size_t m, size_t n, T *d, size_t ld, const kernels::RunContext &c) : DenseTile(m, n, d, ld, blas::Layout::ColMajor, c) {}
    DenseTile() = default;  // for TilePacker::PackTile (Dense.hpp default constructor)
    ~DenseTile() override { delete this->mpDataArray; }
    T *GetTileSubMatrix(size_t aIndex) const override {
        if (aIndex != 0) throw std::invalid_argument("GetTileSubMatrix ::Index out of range, should be 0 in case of dense tile.\n");
        return this->mpDataArray->GetData();
    }
    size_t GetTileStride(size_t aIndex) const override {
        if (aIndex != 0) throw std::invalid_argument("DenseTile::GetTileStride:: Index out of range, should be 0 in case of dense tile.\n");
        return this->mpDataArray->GetLeadingDim();
    }
    bool isDense() const override { return true; }
    bool isCompressed() const override { return false; }
    int64_t GetNumOfSubMatrices() const override { return 1; }
    TileType GetTileType() override { return DENSE; }
    hcb_tile Descriptor() const override {
        const bool rm = this->mLayout == blas::Layout::RowMajor;  // the column-major view of a row-major tile is its transpose
        return hcb_tile{HCB_TILE_DENSE, (int32_t) (rm ? this->mNumOfCols : this->mNumOfRows),
                        (int32_t) (rm ? this->mNumOfRows : this->mNumOfCols),
                        (int32_t) this->mpDataArray->GetLeadingDim(), 0, 0, nullptr, this->mpDataArray->GetData()};
    }
    std::pair<TileMetadata *, T *> UnPackTile(const kernels::RunContext &) override {
        return {new TileMetadata(this->mNumOfRows, this->mNumOfCols, 0, 0, this->mLeadingDim, this->mLayout, DENSE),
                this->mpDataArray->GetData()};
    }
    /// Dense.cpp:137-143: adopt a DEVICE buffer (no ownership) described by the metadata
    void PackTile(TileMetadata aMetadata, T *aDataArray, const kernels::RunContext &aContext) override {
        delete this->mpDataArray;
        this->mLayout = aMetadata.mLayout; this->mLeadingDim = aMetadata.mLeadingDimension;
        this->mNumOfRows = aMetadata.mNumOfRows; this->mNumOfCols = aMetadata.mNumOfCols; this->mRank = aMetadata.mMatrixRank;
        const bool rm = this->mLayout == blas::Layout::RowMajor;
        this->mpDataArray = new dataunits::DataHolder<T>(rm ? this->mNumOfCols : this->mNumOfRows, rm ? this->mNumOfRows : this->mNumOfCols,
                                                         this->mLeadingDim, aDataArray, aContext, false);
    }
};

#define HCOREPP_B200_MAX_RANK_RATIO 3  // Compressed.hpp:14

template<typename T>
class CompressedTile : public Tile<T> {  // Compressed.hpp:67-151 ; buffer = [U (m x maxRank) | V (maxRank x n)]
public:
    /// (m, n, U, V, ld, rank, layout, ctx): maxRank = rank  (Compressed.cpp:20-47)
    CompressedTile(size_t aNumOfRows, size_t aNumOfCols, T *apDataU, T *apDataV, size_t aLeadingDim, size_t aRank,
                   blas::Layout aLayout, const kernels::RunContext &aContext) : mpContext(&aContext) {
        Init(aNumOfRows, aNumOfCols, aLeadingDim, aRank, std::max<size_t>(aRank, 1), aLayout);
        memory::Memcpy<T>(GetUMatrix(), apDataU, aNumOfRows * aRank, aContext, memory::MemoryTransfer::AUTOMATIC);
        memory::Memcpy<T>(GetVMatrix(), apDataV, aRank * aNumOfCols, aContext, memory::MemoryTransfer::AUTOMATIC);
        aContext.Sync();
    }
    /// (m, n, UV, ld, rank[, layout], ctx): UV = [U | V] packed, maxRank = rank  (Compressed.cpp:49-73)
    CompressedTile(size_t aNumOfRows, size_t aNumOfCols, T *apData, size_t aLeadingDim, size_t aRank, blas::Layout aLayout,
                   const kernels::RunContext &aContext) : mpContext(&aContext) {
        Init(aNumOfRows, aNumOfCols, aLeadingDim, aRank, std::max<size_t>(aRank, 1), aLayout);
        if (apData) {
            memory::Memcpy<T>(GetUMatrix(), apData, aNumOfRows * aRank, aContext, memory::MemoryTransfer::AUTOMATIC);
            memory::Memcpy<T>(GetVMatrix(), apData + aNumOfRows * aRank, aRank * aNumOfCols, aContext, memory::MemoryTransfer::AUTOMATIC);
            aContext.Sync();
        }
    }
    CompressedTile(size_t m, size_t n, T *d, size_t ld, size_t rank, const kernels::RunContext &c)
        : CompressedTile(m, n, d, ld, rank, blas::Layout::ColMajor, c) {}
    /// compressing constructor: SVD + truncation on the device, maxRank = max(min(m,n)/3, 1)  (Compressed.cpp:75-146)
    CompressedTile(size_t aNumOfRows, size_t aNumOfCols, T *apData, size_t aLeadingDim, const CompressionParameters &aParameters,
                   blas::Layout aLayout, const kernels::RunContext &aContext) : mpContext(&aContext) {
        const size_t maxRank = std::max<size_t>(std::min(aNumOfRows, aNumOfCols) / HCOREPP_B200_MAX_RANK_RATIO, 1);
        Init(aNumOfRows, aNumOfCols, aLeadingDim, maxRank, maxRank, aLayout);
        if (apData) {
            dataunits::DataHolder<T> dense(aNumOfRows, aNumOfCols, aLeadingDim, apData, aContext);  // host or device source
            const T *ptrs[1] = {dense.GetData()};
            hcb_tile out = Descriptor();
            hcb_compress_params p = aParameters.ToC();
            detail::check(detail::abi<T>::compress_batched(aContext.Handle(), 1, ptrs, (int64_t) aNumOfRows, &out, &p, nullptr),
                          "CompressedTile(compress)");
            aContext.Sync();
        }
    }
    CompressedTile() = default;  // for TilePacker::PackTile (Compressed.hpp default constructor)
    CompressedTile(const CompressedTile &) = delete;
    ~CompressedTile() override {
        delete this->mpDataArray;
        if (mpRank && mpContext) hcb_free(mpContext->Handle(), mpRank);
    }
    T *GetUMatrix() const { return this->mpDataArray->GetData(); }
    T *GetVMatrix() const { return this->mpDataArray->GetData() + this->mNumOfRows * this->mMaxRank; }  // Compressed.cpp:180-185
    T *GetTileSubMatrix(size_t aIndex) const override {
        if (aIndex > 1) throw std::invalid_argument("CompressedTile::GetTileSubMatrix:: Index out of range, should be 0 or 1 in case of compressed tile.\n");
        return aIndex == 0 ? GetUMatrix() : GetVMatrix();
    }
    size_t GetTileStride(size_t aIndex) const override {
        if (aIndex > 1) throw std::invalid_argument("CompressedTile::GetTileStride::Index out of range, should be 0 or 1 in case of compressed tile.\n");
        return aIndex == 0 ? GetULeadingDim() : GetVLeadingDim();
    }
    /// The device owns the rank; the host getter refreshes it (synchronises the context, like the reference's
    /// CalculateNewRank does per GEMM, CudaKernels.cu:656-697 -- here only when somebody asks).
    size_t GetTileRank() const override {
        int32_t r = 0;
        detail::check(hcb_memcpy(mpContext->Handle(), &r, mpRank, sizeof(r), 2), "GetTileRank");
        mpContext->Sync();
        return (size_t) r;
    }
    /// Per-tile fixed rank of the replay drivers (par_fixed_rank_streams_main.cpp:465-477,540-541; Compressed.cpp:510-515):
    /// > 0 makes every recompression of THIS tile keep exactly that rank, 0 restores truncation by accuracy.
    void SetFixedRank(size_t aRank) { mFixedRank = (int32_t) aRank; }
    /// The library keeps a state word next to the rank (hcb_tile.d_state: "U has orthonormal columns"); whoever writes the
    /// factors through GetUMatrix() / GetVMatrix() directly must call this.
    void InvalidateState() {
        const int32_t z = 0;
        detail::check(hcb_memcpy(mpContext->Handle(), mpRank + 1, &z, sizeof(z), 0), "InvalidateState");
    }
    size_t GetMaxRank() const { return this->mMaxRank; }
    size_t GetULeadingDim() const { return this->mNumOfRows; }
    size_t GetVLeadingDim() const { return GetTileRank(); }
    bool isDense() const override { return false; }
    bool isCompressed() const override { return true; }
    int64_t GetNumOfSubMatrices() const override { return 2; }
    TileType GetTileType() override { return COMPRESSED; }
    hcb_tile Descriptor() const override {
        return hcb_tile{HCB_TILE_COMPRESSED, (int32_t) this->mNumOfRows, (int32_t) this->mNumOfCols, 0, (int32_t) this->mMaxRank, 0,
                        mpRank, this->mpDataArray->GetData(), mpRank + 1, mFixedRank, 0};
    }
    /// Dense*Dense -> Compressed makes the tile full rank (HCore.cpp:291-298): grow the buffer like DataHolder::Resize.
    void EnsureCapacity(size_t aMaxRank) {
        if (aMaxRank <= this->mMaxRank) return;
        const size_t m = this->mNumOfRows, n = this->mNumOfCols, rk = GetTileRank();
        auto *fresh = new dataunits::DataHolder<T>(m * aMaxRank + aMaxRank * n, 1, m * aMaxRank + aMaxRank * n, nullptr, *mpContext);
        memory::Memcpy<T>(fresh->GetData(), GetUMatrix(), m * rk, *mpContext);
        memory::Memcpy<T>(fresh->GetData() + m * aMaxRank, GetVMatrix(), rk * n, *mpContext);
        mpContext->Sync();
        delete this->mpDataArray;
        this->mpDataArray = fresh;
        this->mMaxRank = aMaxRank;
    }
    std::pair<TileMetadata *, T *> UnPackTile(const kernels::RunContext &) override {
        return {new TileMetadata(this->mNumOfRows, this->mNumOfCols, GetTileRank(), this->mMaxRank, this->mLeadingDim,
                                 this->mLayout, COMPRESSED), this->mpDataArray->GetData()};
    }
    /// Compressed.cpp:780-805: adopt a DEVICE buffer [U (m x maxRank) | V (maxRank x n), ld = rank] without ownership; the
    /// rank of the metadata goes to the device-resident rank word, the state word starts at "unknown".
    void PackTile(TileMetadata aMetadata, T *aDataArray, const kernels::RunContext &aContext) override {
        if (aMetadata.mLayout != blas::Layout::ColMajor) throw std::invalid_argument("CompressedTile: only ColMajor tiles are supported");
        delete this->mpDataArray;
        if (mpRank && mpContext) hcb_free(mpContext->Handle(), mpRank);
        mpContext = &aContext;
        this->mLayout = aMetadata.mLayout; this->mRank = aMetadata.mMatrixRank; this->mLeadingDim = aMetadata.mLeadingDimension;
        this->mNumOfRows = aMetadata.mNumOfRows; this->mNumOfCols = aMetadata.mNumOfCols; this->mMaxRank = aMetadata.mMaxRank;
        const size_t elems = this->mNumOfRows * this->mMaxRank + this->mMaxRank * this->mNumOfCols;
        this->mpDataArray = new dataunits::DataHolder<T>(elems, 1, elems, aDataArray, aContext, false);
        AllocRankWord((int32_t) aMetadata.mMatrixRank);
    }
private:
    void AllocRankWord(int32_t aRank) {
        void *p = nullptr;
        detail::check(hcb_malloc(mpContext->Handle(), 2 * sizeof(int32_t), &p), "CompressedTile");
        mpRank = static_cast<int32_t *>(p);
        const int32_t init[2] = {aRank, 0};  // {rank, state: unknown}
        detail::check(hcb_memcpy(mpContext->Handle(), mpRank, init, sizeof(init), 0), "CompressedTile");
        mpContext->Sync();
    }
    void Init(size_t m, size_t n, size_t ld, size_t rank, size_t maxRank, blas::Layout layout) {
        if (layout != blas::Layout::ColMajor) throw std::invalid_argument("CompressedTile: only ColMajor tiles are supported");
        this->mLayout = layout; this->mNumOfRows = m; this->mNumOfCols = n; this->mLeadingDim = ld;
        this->mRank = rank; this->mMaxRank = maxRank;
        const size_t elems = m * maxRank + maxRank * n;
        this->mpDataArray = new dataunits::DataHolder<T>(elems, 1, elems, nullptr, *mpContext);
        AllocRankWord((int32_t) rank);
    }
    const kernels::RunContext *mpContext = nullptr;
    int32_t *mpRank = nullptr;  // device-resident {rank, state}
    int32_t mFixedRank = 0;
};

/// TilePacker (include/hcorepp/operators/interface/TilePacker.hpp:18-60, src/operators/TilePacker.cpp:4-26)
template<typename T>
class TilePacker {
public:
    virtual ~TilePacker() = default;
    static std::pair<TileMetadata *, T *> UnPackTile(Tile<T> &aTile, const kernels::RunContext &aContext) {
        return aTile.UnPackTile(aContext);
    }
    static Tile<T> *PackTile(TileMetadata aMetadata, T *apDataArray, const kernels::RunContext &aContext) {
        Tile<T> *tile;
        if (aMetadata.mType == DENSE) tile = new DenseTile<T>();
        else tile = new CompressedTile<T>();
        tile->PackTile(aMetadata, apDataArray, aContext);
        return tile;
    }
};
}  // namespace operators

// ---------------------------------------------------------------------------------------------------------------------
namespace kernels {
/// The kernel table (kernels.hpp:27-129), forwarding 1:1 to the C ABI. Pointers are device pointers.
template<typename T>
class HCoreKernels {
    using A = detail::abi<T>;
    static int op(blas::Op o) { return o == blas::Op::NoTrans ? 0 : 1; }
public:
    static void Gemm(blas::Layout aLayout, blas::Op aTransA, blas::Op aTransB, size_t aM, size_t aN, size_t aK, T &aAlpha,
                     T const *apA, size_t aLdA, T const *apB, size_t aLdB, T &aBeta, T *apC, size_t aLdC, const RunContext &c) {
        if (aLayout == blas::Layout::ColMajor)
            detail::check(A::gemm(c.Handle(), op(aTransA), op(aTransB), aM, aN, aK, aAlpha, apA, aLdA, apB, aLdB, aBeta, apC, aLdC), "Gemm");
        else  // C^T = op(B)^T op(A)^T
            detail::check(A::gemm(c.Handle(), op(aTransB), op(aTransA), aN, aM, aK, aAlpha, apB, aLdB, apA, aLdA, aBeta, apC, aLdC), "Gemm");
    }
    static void MultiplyByAlpha(T *apArray, size_t aRows, size_t aCols, size_t aM, size_t aRank, T &aAlpha, const RunContext &c) {
        detail::check(A::multiply_by_alpha(c.Handle(), apArray, aRows, aCols, aM, aRank, aAlpha), "MultiplyByAlpha");
    }
    static void ProcessVpointer(size_t aN, size_t aCRank, bool aGetUngqr, size_t Vm, T &aBeta, T *apCV, size_t aLdcV, T *V,
                                size_t aArank, const T *apBdata, const RunContext &c, bool aCholesky = false) {
        detail::check(A::process_v(c.Handle(), aN, aCRank, aGetUngqr, Vm, aBeta, apCV, aLdcV, V, aArank, apBdata, aCholesky), "ProcessVpointer");
    }
    static void CalculateNewRank(size_t &aNewRank, bool aTruncatedSvd, T *apSigma, size_t sizeS, T accuracy, const RunContext &c) {
        int64_t r = 0;
        detail::check(A::new_rank(c.Handle(), aTruncatedSvd, apSigma, sizeS, accuracy, &r), "CalculateNewRank");
        aNewRank = (size_t) r;
    }
    static void CalculateUVptr(size_t aRank, size_t aVm, T *UVptr, const T *Vnew, const RunContext &c) {
        detail::check(A::uvptr(c.Handle(), aRank, aVm, UVptr, Vnew), "CalculateUVptr");
    }
    static void CalculateVTnew(size_t aRkNew, bool aUngqr, size_t aMinVmVn, T *apSigma, T *apVTnew, size_t aSizeS, size_t aVm, const RunContext &c) {
        detail::check(A::vtnew(c.Handle(), aRkNew, aUngqr, aMinVmVn, apSigma, apVTnew, aSizeS, aVm), "CalculateVTnew");
    }
    static void CalculateUVptrConj(size_t aRank, size_t aVm, T *UVptr, const RunContext &c) {
        detail::check(A::uvptr_conj(c.Handle(), aRank, aVm, UVptr), "CalculateUVptrConj");
    }
    static void FillIdentityMatrix(size_t aNumOfElements, T *apMatrix, const RunContext &c) {
        detail::check(A::fill_identity(c.Handle(), aNumOfElements, apMatrix), "FillIdentityMatrix");
    }
    static void LaCpy(common::MatrixType aType, size_t aM, size_t aRank, T *apCU, size_t aLD, T *apU, size_t aUm, const RunContext &c) {
        detail::check(A::lacpy(c.Handle(), (int) aType, aM, aRank, apCU, aLD, apU, aUm), "LaCpy");
    }
    static void Geqrf(size_t aM, size_t aN, T *apA, size_t aLdA, T *apTau, T *, size_t, size_t, const RunContext &c) {
        detail::check(A::geqrf(c.Handle(), aM, aN, apA, aLdA, apTau), "Geqrf");
    }
    static void Laset(common::MatrixType aMatrixType, size_t aM, size_t aN, T aOffdiag, T aDiag, T *apA, size_t aLdA, const RunContext &c) {
        detail::check(A::laset(c.Handle(), (int) aMatrixType, aM, aN, aOffdiag, aDiag, apA, aLdA), "Laset");
    }
    static void Trmm(blas::Layout, blas::Side aSide, blas::Uplo aUplo, blas::Op aTrans, blas::Diag aDiag, size_t aM, size_t aN, T aAlpha,
                     T const *apA, size_t aLdA, T *apB, size_t aLdB, const RunContext &c) {
        detail::check(A::trmm(c.Handle(), (int) aSide, (int) aUplo, (int) aTrans, (int) aDiag, aM, aN, aAlpha, apA, aLdA, apB, aLdB), "Trmm");
    }
    static void SVD(common::Job, common::Job, size_t aM, size_t aN, T *apA, size_t aLdA, T *apS, T *apU, size_t aLdU, T *apVT,
                    size_t aLdVt, common::CompressionType, T *, size_t, size_t, const RunContext &c) {
        detail::check(A::svd(c.Handle(), aM, aN, apA, aLdA, apS, apU, aLdU, apVT, aLdVt), "SVD");
    }
    static void Unmqr(common::SideMode aSide, common::BlasOperation aTrans, size_t aM, size_t aN, size_t aK, T const *apA, size_t aLdA,
                      T const *apTau, T *apC, size_t aLdC, T *, size_t, const RunContext &c) {
        detail::check(A::unmqr(c.Handle(), (int) aSide, aTrans == common::OP_NoTRANS ? 0 : 1, aM, aN, aK, apA, aLdA, apTau, apC, aLdC), "Unmqr");
    }
    static void ungqr(size_t aM, size_t aN, size_t aK, T *apA, size_t aLdA, T *apTau, T *, size_t, const RunContext &c) {
        detail::check(A::ungqr(c.Handle(), aM, aN, aK, apA, aLdA, apTau), "ungqr");
    }
    /// kernels.hpp:103-129 -- the Cholesky pieces of the table (src/kernels/omp/kernels.cpp:234-303)
    static int potrf(blas::Uplo aUplo, T *, size_t, size_t, size_t aMatrixOrder, T *apMatrix, size_t aLeadingDim, blas::Layout,
                     const RunContext &c) {
        int32_t *d_info = nullptr;
        detail::check(hcb_malloc(c.Handle(), sizeof(int32_t), (void **) &d_info), "potrf");
        detail::check(A::potrf(c.Handle(), (int) aUplo, aMatrixOrder, apMatrix, aLeadingDim, d_info), "potrf");
        int32_t info = 0;
        detail::check(hcb_memcpy(c.Handle(), &info, d_info, sizeof(info), 2), "potrf");
        c.Sync();
        hcb_free(c.Handle(), d_info);
        return info;
    }
    static void trsm(blas::Layout, blas::Side aSide, blas::Uplo aUplo, blas::Op aTrans, blas::Diag aDiag, size_t aRows, size_t aCols,
                     T aAlpha, const T *apMatrixA, size_t aLeadingDimA, T *apMatrixB, size_t aLeadingDimB, const RunContext &c) {
        detail::check(A::trsm(c.Handle(), (int) aSide, (int) aUplo, op(aTrans), (int) aDiag, aRows, aCols, aAlpha, apMatrixA,
                              aLeadingDimA, apMatrixB, aLeadingDimB), "trsm");
    }
    static void syrk(blas::Layout, blas::Uplo aUplo, blas::Op aTrans, size_t aRows, size_t aCols, T aAlpha, const T *apMatrixA,
                     size_t aLeadingDimA, T aBeta, T *apMatrixB, size_t aLeadingDimB, const RunContext &c) {
        detail::check(A::syrk(c.Handle(), (int) aUplo, op(aTrans), aRows, aCols, aAlpha, apMatrixA, aLeadingDimA, aBeta, apMatrixB,
                              aLeadingDimB), "syrk");
    }
    static void FillMatrixTriangle(blas::Uplo aUplo, size_t aRows, size_t aCols, T *apMatrix, blas::Layout, size_t aValue,
                                   const RunContext &c) {
        if (aRows != aCols) return;
        detail::check(A::fill_triangle(c.Handle(), (int) aUplo, aRows, apMatrix, aRows, (T) aValue), "FillMatrixTriangle");
    }
    static void Symmetrize(blas::Layout, T *apMatrixA, size_t aRows, size_t aCols, blas::Uplo aUplo, const RunContext &c) {
        if (aRows != aCols) return;
        detail::check(A::symmetrize(c.Handle(), (int) aUplo, aRows, apMatrixA, aRows), "Symmetrize");
    }
    static size_t CalculatePotrfWorkspaceSize(T *, blas::Uplo, size_t, size_t, size_t &aHostSize, const RunContext &) {
        aHostSize = 0;
        return 0;
    }
    static size_t CalculateGemmWorkspaceSize(size_t, size_t, size_t, size_t, size_t, const operators::CompressionParameters &,
                                             size_t &aHostSize, const RunContext &) {
        aHostSize = 0;  // scratch lives in the context's arena; callers need not provide any
        return 0;
    }
};
}  // namespace kernels

// ---------------------------------------------------------------------------------------------------------------------
namespace api {
template<typename T>
class HCore {
public:
    /// HCore<T>::Gemm (HCore.hpp:41-46): C = alpha*op(A)*op(B) + beta*C on Dense / Compressed tiles, recompression
    /// included; one fused device call, nothing synchronises.  aFlops receives the reference's dense-equivalent model
    /// for the contraction (2mnk).
    static void Gemm(T aAlpha, operators::Tile<T> const &aA, blas::Op const &aAOp, operators::Tile<T> const &aB,
                     blas::Op const &aBOp, T aBeta, operators::Tile<T> &aC, const kernels::RunContext &aContext, size_t &aFlops,
                     dataunits::MemoryUnit<T> &aMemoryUnit, const operators::CompressionParameters &aSVDArguments = {1e-9},
                     bool aCholesky = false) {
        (void) aMemoryUnit;
        // aCholesky (HCore.cpp:160-174): the reference's Cholesky variant is C += alpha * A * B^T on compressed tiles whose V
        // factors it keeps as n x k.  Here every tile keeps the one layout of this library (V = rank x n), so the variant is the
        // plain transposed product -- it only checks what the reference's branch requires.
        if (aCholesky && !(aA.isCompressed() && aB.isCompressed() && aAOp == blas::Op::NoTrans && aBOp == blas::Op::Trans))
            throw std::runtime_error("HCore::Gemm: the Cholesky variant takes compressed A, B with (NoTrans, Trans)");
        if (aA.isDense() && aB.isDense() && aC.isCompressed())
            static_cast<operators::CompressedTile<T> &>(aC).EnsureCapacity(std::min(aC.GetNumOfRows(), aC.GetNumOfCols()));
        const hcb_tile a = aA.Descriptor(), b = aB.Descriptor(), c = aC.Descriptor();
        const hcb_compress_params p = aSVDArguments.ToC();
        const bool rm = aC.GetLayout() == blas::Layout::RowMajor;
        if (rm != (aA.GetLayout() == blas::Layout::RowMajor) || rm != (aB.GetLayout() == blas::Layout::RowMajor))
            throw std::invalid_argument("HCore::Gemm: A, B and C must share one layout");
        if (rm) {
            // Row-major tiles (dense only, like the reference: Dense.cpp:69-96 hands the layout to the GEMM; compressed tiles are
            // ColMajor, Compressed.cpp:29-34): the descriptors are the column-major TRANSPOSED views, and
            // C^T = alpha * op(B)^T op(A)^T + beta * C^T = alpha * op_B(B^T-view) * op_A(A^T-view) + beta * C^T-view.
            if (!(aA.isDense() && aB.isDense() && aC.isDense())) throw std::invalid_argument("HCore::Gemm: RowMajor is a dense-tile layout");
            detail::check(detail::abi<T>::tlr_gemm_batched(aContext.Handle(), 1, &b, aBOp == blas::Op::NoTrans ? 0 : 1, &a,
                                                           aAOp == blas::Op::NoTrans ? 0 : 1, &c, aAlpha, aBeta, &p, nullptr), "HCore::Gemm");
        } else
        detail::check(detail::abi<T>::tlr_gemm_batched(aContext.Handle(), 1, &a, aAOp == blas::Op::NoTrans ? 0 : 1, &b,
                                                       aBOp == blas::Op::NoTrans ? 0 : 1, &c, aAlpha, aBeta, &p, nullptr), "HCore::Gemm");
        aFlops += 2 * aC.GetNumOfRows() * aC.GetNumOfCols() * (aAOp == blas::Op::NoTrans ? aA.GetNumOfCols() : aA.GetNumOfRows());
    }
    /// Batched form used by multi-tile drivers: many independent (A,B,C) triples of one operand mix per call.
    static void GemmBatched(T aAlpha, const std::vector<const operators::Tile<T> *> &aA, blas::Op aAOp,
                            const std::vector<const operators::Tile<T> *> &aB, blas::Op aBOp, T aBeta,
                            const std::vector<operators::Tile<T> *> &aC, const kernels::RunContext &aContext,
                            const operators::CompressionParameters &aSVDArguments = {1e-9}) {
        const size_t n = aC.size();
        if (aA.size() != n || aB.size() != n) throw std::invalid_argument("HCore::GemmBatched: size mismatch");
        std::vector<hcb_tile> a(n), b(n), c(n);
        for (size_t i = 0; i < n; ++i) { a[i] = aA[i]->Descriptor(); b[i] = aB[i]->Descriptor(); c[i] = aC[i]->Descriptor(); }
        const hcb_compress_params p = aSVDArguments.ToC();
        detail::check(detail::abi<T>::tlr_gemm_batched(aContext.Handle(), (int64_t) n, a.data(), aAOp == blas::Op::NoTrans ? 0 : 1,
                                                       b.data(), aBOp == blas::Op::NoTrans ? 0 : 1, c.data(), aAlpha, aBeta, &p, nullptr),
                      "HCore::GemmBatched");
    }
    /// HCore<T>::Potrf (HCore.cpp:586-621): dense tiles only, in place; the other triangle is left as it was.
    static void Potrf(operators::Tile<T> &aA, const blas::Uplo aUplo, const kernels::RunContext &aContext, size_t &aFlops,
                      dataunits::MemoryUnit<T> &) {
        if (aA.GetNumOfSubMatrices() != 1) throw std::runtime_error(" Potrf works only with dense tiles");
        auto &dh = aA.GetDataHolder().get();
        kernels::HCoreKernels<T>::potrf(aUplo, nullptr, 0, 0, aA.GetNumOfRows(), dh.GetData(), dh.GetLeadingDim(), aA.GetLayout(), aContext);
        aFlops += aA.GetNumOfRows() * aA.GetNumOfRows() * aA.GetNumOfRows() / 3;
    }
    /// HCore<T>::Trsm (HCore.cpp:624-647): A dense triangular, B compressed; blas::trsm with m = rows(B), n = rank(B) on B's V
    /// buffer with the tile's leading dimension (the reference's Cholesky convention stores V as n x k); U is not touched.
    static void Trsm(blas::Side aSide, blas::Uplo aUplo, blas::Op aTrans, blas::Diag aDiag, T aAlpha, operators::Tile<T> &aA,
                     operators::Tile<T> &aB, const kernels::RunContext &aContext, size_t &, dataunits::MemoryUnit<T> &) {
        if (aB.GetNumOfSubMatrices() != 2) throw std::runtime_error(" TRSM: Tile B must be compressed ");
        auto &b = static_cast<operators::CompressedTile<T> &>(aB);
        auto &dh = aA.GetDataHolder().get();
        kernels::HCoreKernels<T>::trsm(aB.GetLayout(), aSide, aUplo, aTrans, aDiag, aB.GetNumOfRows(), b.GetTileRank(), aAlpha,
                                       dh.GetData(), dh.GetLeadingDim(), b.GetVMatrix(), aB.GetNumOfRows(), aContext);
    }
    /// HCore<T>::Syrk (HCore.cpp:484-583).  Dense A: blas::syrk with n = rows(C), k = cols(C) on the `uplo` triangle.
    /// Compressed A (= U V in this library's layout): C := beta C - (U (V V^T)) U^T and the other strict triangle zeroed --
    /// the reference's compressed branch (alpha is hard-coded to -1 there, HCore.cpp:546).
    static void Syrk(T aAlpha, const operators::Tile<T> &aA, const blas::Op &aAOp, const blas::Uplo aUplo, T aBeta,
                     operators::Tile<T> &aC, const kernels::RunContext &aContext, size_t &, dataunits::MemoryUnit<T> &) {
        auto &cd = aC.GetDataHolder().get();
        if (aA.GetNumOfSubMatrices() == 1) {
            auto &ad = aA.GetDataHolder().get();
            kernels::HCoreKernels<T>::syrk(aC.GetLayout(), aUplo, aAOp, aC.GetNumOfRows(), aC.GetNumOfCols(), aAlpha, ad.GetData(),
                                           ad.GetLeadingDim(), aBeta, cd.GetData(), cd.GetLeadingDim(), aContext);
            return;
        }
        if (!aC.isDense()) throw std::runtime_error("HCore::Syrk: compressed A needs a dense C");
        const hcb_tile a = aA.Descriptor();
        T *cptr[1] = {cd.GetData()};
        const int64_t ldc[1] = {(int64_t) cd.GetLeadingDim()};
        detail::check(detail::abi<T>::tlr_syrk_batched(aContext.Handle(), 1, &a, cptr, ldc, T(-1), aBeta), "HCore::Syrk");
        kernels::HCoreKernels<T>::FillMatrixTriangle(aUplo == blas::Uplo::Lower ? blas::Uplo::Upper : blas::Uplo::Lower, aC.GetNumOfRows(),
                                                     aC.GetNumOfCols(), cd.GetData(), aC.GetLayout(), 0, aContext);
    }
    /// HCore.hpp:59-63 -- in ELEMENTS, like the reference; what one fused call needs for these tiles.
    static size_t CalculateMemoryPoolSize(const operators::Tile<T> &aA, const operators::Tile<T> &aB, const operators::Tile<T> &aC,
                                          operators::CompressionParameters, const kernels::RunContext &) {
        const size_t ka = aA.isCompressed() ? static_cast<const operators::CompressedTile<T> &>(aA).GetMaxRank() : 0;
        const size_t kc = aC.isCompressed() ? static_cast<const operators::CompressedTile<T> &>(aC).GetMaxRank() : 0;
        const size_t k = aA.GetNumOfCols();
        return detail::abi<T>::tlr_gemm_workspace(1, aC.GetNumOfRows(), aC.GetNumOfCols(), k, ka + kc) / sizeof(T) + 1;
    }
};
}  // namespace api

namespace helpers {  // include/hcorepp/helpers/{RawMatrix,TileMatrix}.hpp -- the multi-tile container and the driver loop

/// Host column-major matrix (RawMatrix.hpp:27-160, the part the drivers use; generators / LAPACK helpers stay in the
/// reference's test harness).
template<typename T>
class RawMatrix {
public:
    RawMatrix(size_t aM, size_t aN) : mM(aM), mN(aN), mData(aM * aN, T(0)) {}
    RawMatrix(size_t aM, size_t aN, const T *apData) : mM(aM), mN(aN), mData(apData, apData + aM * aN) {}
    T *GetData() { return mData.data(); }
    const T *GetData() const { return mData.data(); }
    size_t GetM() const { return mM; }
    size_t GetN() const { return mN; }
    /// Frobenius norm
    double Norm() const {
        double s = 0;
        for (const T &v : mData) s += (double) v * (double) v;
        return std::sqrt(s);
    }
    /// this -= aReference (RawMatrix.hpp ReferenceDifference)
    void ReferenceDifference(const RawMatrix<T> &aReference) {
        if (aReference.mM != mM || aReference.mN != mN) throw std::runtime_error("Reference Matrix Is Not The Same Size");
        for (size_t i = 0; i < mData.size(); ++i) mData[i] -= aReference.mData[i];
    }
private:
    size_t mM, mN;
    std::vector<T> mData;
};

/// TileMatrix.hpp:27-196: mt x nt grid of tiles (tile (row, col) at [col][row]); dense or compressed on construction.
template<typename T>
class TileMatrix {
public:
    TileMatrix(const RawMatrix<T> &aRawMatrix, size_t aRowTileSize, size_t aColumnTileSize, kernels::RunContext &aContext)
        : TileMatrix(aRawMatrix, aRowTileSize, aColumnTileSize, nullptr, aContext) {}
    TileMatrix(const RawMatrix<T> &aRawMatrix, size_t aRowTileSize, size_t aColumnTileSize,
               const operators::CompressionParameters &aParameters, kernels::RunContext &aContext)
        : TileMatrix(aRawMatrix, aRowTileSize, aColumnTileSize, &aParameters, aContext) {}
    TileMatrix(const TileMatrix &) = delete;
    ~TileMatrix() {
        for (auto &col : mMatrixTiles)
            for (auto *t : col) delete t;
    }
    operators::Tile<T> *GetTile(size_t aRowIndex, size_t aColIndex) { return mMatrixTiles[aColIndex][aRowIndex]; }
    size_t GetRowTileCount() const { return mRowTileCount; }
    size_t GetColTileCount() const { return mColTileCount; }
    size_t GetRowTileSize() const { return mRowTileSize; }
    size_t GetColTileSize() const { return mColTileSize; }
    size_t GetM() const { return mM; }
    size_t GetN() const { return mN; }
    /// U*V per tile, assembled on the host (TileMatrix.cpp:187-252)
    RawMatrix<T> ToRawMatrix(kernels::RunContext &aContext) {
        RawMatrix<T> out(mM, mN);
        for (size_t c = 0; c < mColTileCount; ++c)
            for (size_t r = 0; r < mRowTileCount; ++r) {
                operators::Tile<T> *t = mMatrixTiles[c][r];
                const size_t tm = t->GetNumOfRows(), tn = t->GetNumOfCols();
                std::vector<T> d(tm * tn, T(0));
                if (t->isDense()) {
                    memory::Memcpy<T>(d.data(), t->GetTileSubMatrix(0), tm * tn, aContext, memory::MemoryTransfer::DEVICE_TO_HOST);
                    aContext.Sync();
                } else {
                    auto *ct = static_cast<operators::CompressedTile<T> *>(t);
                    const size_t rk = ct->GetTileRank();
                    std::vector<T> U(tm * rk), V(rk * tn);
                    memory::Memcpy<T>(U.data(), ct->GetUMatrix(), tm * rk, aContext, memory::MemoryTransfer::DEVICE_TO_HOST);
                    memory::Memcpy<T>(V.data(), ct->GetVMatrix(), rk * tn, aContext, memory::MemoryTransfer::DEVICE_TO_HOST);
                    aContext.Sync();
                    for (size_t j = 0; j < tn; ++j)
                        for (size_t l = 0; l < rk; ++l)
                            for (size_t i = 0; i < tm; ++i) d[i + j * tm] += U[i + l * tm] * V[l + j * rk];
                }
                for (size_t j = 0; j < tn; ++j)
                    std::memcpy(out.GetData() + (r * mRowTileSize) + (c * mColTileSize + j) * mM, d.data() + j * tm, tm * sizeof(T));
            }
        return out;
    }
    /// bytes held by the tiles (TileMatrix.cpp:254-266: compressed tiles count (m + n) * rank elements)
    size_t GetMemoryFootprint() {
        size_t bytes = 0;
        for (auto &col : mMatrixTiles)
            for (auto *t : col)
                bytes += sizeof(T) * (t->isDense() ? t->GetNumOfRows() * t->GetNumOfCols()
                                                   : (t->GetNumOfRows() + t->GetNumOfCols()) * t->GetTileRank());
        return bytes;
    }
private:
    TileMatrix(const RawMatrix<T> &aRaw, size_t aTm, size_t aTn, const operators::CompressionParameters *apParams,
               kernels::RunContext &aContext)
        : mRowTileSize(aTm), mColTileSize(aTn), mM(aRaw.GetM()), mN(aRaw.GetN()) {
        mRowTileCount = (mM + aTm - 1) / aTm;
        mColTileCount = (mN + aTn - 1) / aTn;
        mMatrixTiles.resize(mColTileCount);
        // compressed: ONE device copy of the whole matrix and ONE batched compressing call over all tiles (the reference
        // compresses tile by tile, TileMatrix.cpp:60-90; its per-tile constructor is still available on CompressedTile)
        T *dRaw = nullptr;
        std::vector<const T *> ptrs;
        std::vector<hcb_tile> descs;
        if (apParams) {
            dRaw = memory::AllocateArray<T>(mM * mN, aContext);
            memory::Memcpy<T>(dRaw, aRaw.GetData(), mM * mN, aContext, memory::MemoryTransfer::HOST_TO_DEVICE);
        }
        for (size_t c = 0; c < mColTileCount; ++c) {
            mMatrixTiles[c].resize(mRowTileCount, nullptr);
            for (size_t r = 0; r < mRowTileCount; ++r) {
                const size_t tm = std::min(aTm, mM - r * aTm), tn = std::min(aTn, mN - c * aTn);
                T *src = const_cast<T *>(aRaw.GetData()) + r * aTm + c * aTn * mM;  // sub-matrix view, ld = M
                if (apParams) {
                    auto *t = new operators::CompressedTile<T>(tm, tn, nullptr, mM, *apParams, blas::Layout::ColMajor, aContext);
                    mMatrixTiles[c][r] = t;
                    ptrs.push_back(dRaw + r * aTm + c * aTn * mM);
                    descs.push_back(t->Descriptor());
                } else {
                    mMatrixTiles[c][r] = new operators::DenseTile<T>(tm, tn, src, mM, blas::Layout::ColMajor, aContext);
                }
            }
        }
        if (apParams) {
            const hcb_compress_params p = apParams->ToC();
            detail::check(detail::abi<T>::compress_batched(aContext.Handle(), (int64_t) descs.size(), ptrs.data(), (int64_t) mM, descs.data(),
                                                           &p, nullptr), "TileMatrix(compress)");
            aContext.Sync();
            hcb_free(aContext.Handle(), dRaw);
        }
    }
    std::vector<std::vector<operators::Tile<T> *>> mMatrixTiles;
    size_t mRowTileCount = 0, mColTileCount = 0, mRowTileSize, mColTileSize, mM, mN;
};

/// The drivers' triple loop (examples/matrix_multiplication/omp_main.cpp:112-126): C(j,i) = alpha * sum_k A(j,k) B(k,i) +
/// beta * C(j,i), here ONE batched device call per k over all C tiles (hcb_?tlr_matmul), nothing synchronises.
template<typename T>
void TileMatrixMultiplication(TileMatrix<T> &aA, TileMatrix<T> &aB, TileMatrix<T> &aC, T aAlpha, T aBeta,
                              const operators::CompressionParameters &aParameters, const kernels::RunContext &aContext) {
    const size_t mt = aC.GetRowTileCount(), nt = aC.GetColTileCount(), kt = aA.GetColTileCount();
    if (aA.GetRowTileCount() != mt || aB.GetColTileCount() != nt || aB.GetRowTileCount() != kt)
        throw std::invalid_argument("TileMatrixMultiplication: tile grids do not conform");
    std::vector<hcb_tile> a(mt * kt), b(kt * nt), c(mt * nt);
    for (size_t k = 0; k < kt; ++k)
        for (size_t j = 0; j < mt; ++j) a[j + k * mt] = aA.GetTile(j, k)->Descriptor();
    for (size_t i = 0; i < nt; ++i)
        for (size_t k = 0; k < kt; ++k) b[k + i * kt] = aB.GetTile(k, i)->Descriptor();
    for (size_t i = 0; i < nt; ++i)
        for (size_t j = 0; j < mt; ++j) c[j + i * mt] = aC.GetTile(j, i)->Descriptor();
    const hcb_compress_params p = aParameters.ToC();
    detail::check(detail::abi<T>::tlr_matmul(aContext.Handle(), (int64_t) mt, (int64_t) nt, (int64_t) kt, a.data(), b.data(), c.data(),
                                             nullptr, 0, 0, (int64_t) kt, aAlpha, aBeta, &p, nullptr),
                  "TileMatrixMultiplication");
}

}  // namespace helpers
}  // namespace hcorepp
