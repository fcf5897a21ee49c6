// tests/dropin/dropin_test.cpp -- the reference's OWN classes (compiled unmodified, -DUSE_CUDA) running on
// libhcore_b200.so through integration/src/kernels/b200/kernels.cpp.  Test infrastructure.
//
// usage: dropin_test <input.bin> <output.bin>
//   input  (little-endian): int64 header {nb, ka, n_steps}, double accuracy, then per step the factors of A_k and B_k:
//            AU (nb x ka), AV (ka x nb), BU (nb x ka), BV (ka x nb), column-major doubles
//   output: int64 rank after every step (n_steps), then the final C as a dense nb x nb column-major matrix
// The run is the reference's tile-at-a-time flow: CompressedTile constructors, HCore<double>::Gemm(A_k, B_k, C) with
// recompression for every k -- i.e. Compressed.cpp:208-694 calling Geqrf / ungqr / SVD / CalculateNewRank / ... of the
// kernel table, each of which is now a libhcore_b200 kernel.  Also replays TestGemm.cpp's CCC known answer and a Potrf.
#include <hcorepp/api/HCore.hpp>
#include <hcorepp/kernels/ContextManager.hpp>
#include <hcorepp/kernels/memory.hpp>
#include <hcorepp/operators/concrete/Compressed.hpp>
#include <hcorepp/operators/concrete/Dense.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace hcorepp;
using namespace hcorepp::operators;

static std::vector<double> to_host(const double *d, size_t n, const kernels::RunContext &ctx) {
    std::vector<double> h(n);
    memory::Memcpy<double>(h.data(), d, n, ctx, memory::MemoryTransfer::DEVICE_TO_HOST, true);
    return h;
}
static double *to_device(const std::vector<double> &h, const kernels::RunContext &ctx) {
    double *d = memory::AllocateArray<double>(h.size(), ctx);
    memory::Memcpy<double>(d, h.data(), h.size(), ctx, memory::MemoryTransfer::HOST_TO_DEVICE, true);
    return d;
}
static std::vector<double> dense_of(CompressedTile<double> &t, const kernels::RunContext &ctx) {
    const size_t m = t.GetNumOfRows(), n = t.GetNumOfCols(), rk = t.GetTileRank();
    auto U = to_host(t.GetUMatrix(), m * rk, ctx), V = to_host(t.GetVMatrix(), rk * n, ctx);
    std::vector<double> out(m * n, 0.0);
    for (size_t j = 0; j < n; ++j)
        for (size_t l = 0; l < rk; ++l)
            for (size_t i = 0; i < m; ++i) out[i + j * m] += U[i + l * m] * V[l + j * rk];
    return out;
}

int main(int argc, char **argv) {
    auto &ctx = kernels::ContextManager::GetInstance().GetContext();
    dataunits::MemoryUnit<double> unit(ctx);
    int failures = 0;
    size_t flops = 0;
    {   // TestGemm.cpp CCC-style known answer: A = U_a V_a (3x2 rank 1 ...) through the reference's classes
        std::vector<double> au = {1, 2, 3}, av = {1, 1}, bu = {1, 2}, bv = {2, 3};      // A = au*av (3x2), B = bu*bv (2x2)
        double *dau = to_device(au, ctx), *dav = to_device(av, ctx), *dbu = to_device(bu, ctx), *dbv = to_device(bv, ctx);
        CompressedTile<double> A(3, 2, dau, dav, 3, 1, blas::Layout::ColMajor, ctx);
        CompressedTile<double> B(2, 2, dbu, dbv, 2, 1, blas::Layout::ColMajor, ctx);
        std::vector<double> z(3 + 2, 0.0);
        double *dz = to_device(z, ctx);
        CompressedTile<double> Cc(3, 2, dz, 3, 1, blas::Layout::ColMajor, ctx);
        api::HCore<double>::Gemm(1.0, A, blas::Op::NoTrans, B, blas::Op::NoTrans, 1.0, Cc, ctx, flops, unit,
                                 CompressionParameters(1e-12));
        ctx.Sync();
        auto got = dense_of(Cc, ctx);   // A*B = [1;2;3]*(1*1+1*2)*[2 3] = [1;2;3]*3*[2 3]
        const double want[6] = {6, 12, 18, 9, 18, 27};
        bool ok = Cc.GetTileRank() == 1;
        for (int i = 0; i < 6; ++i) ok = ok && std::fabs(got[i] - want[i]) < 1e-9;
        std::printf("reference CompressedTile + HCore::Gemm on libhcore_b200: %s (rank %zu)\n", ok ? "PASS" : "FAIL", Cc.GetTileRank());
        failures += !ok;
    }
    {   // HCore<T>::Potrf through the adapter
        std::vector<double> a = {4, 2, 2, 2, 5, 3, 2, 3, 6};
        double *da = to_device(a, ctx);
        DenseTile<double> A(3, 3, da, 3, blas::Layout::ColMajor, ctx);
        api::HCore<double>::Potrf(A, blas::Uplo::Lower, ctx, flops, unit);
        ctx.Sync();
        auto l = to_host(A.GetTileSubMatrix(0), 9, ctx);
        bool ok = std::fabs(l[0] - 2) < 1e-12 && std::fabs(l[1] - 1) < 1e-12 && std::fabs(l[2] - 1) < 1e-12 && std::fabs(l[4] - 2) < 1e-12 &&
                  std::fabs(l[5] - 1) < 1e-12 && std::fabs(l[8] - 2) < 1e-12;
        std::printf("reference DenseTile + HCore::Potrf on libhcore_b200: %s\n", ok ? "PASS" : "FAIL");
        failures += !ok;
    }
    if (argc >= 3) {
        FILE *f = std::fopen(argv[1], "rb");
        if (!f) { std::printf("cannot open %s\n", argv[1]); return 90; }
        int64_t hdr[3];
        double acc;
        if (std::fread(hdr, sizeof(int64_t), 3, f) != 3 || std::fread(&acc, sizeof(double), 1, f) != 1) return 91;
        const size_t nb = hdr[0], ka = hdr[1], steps = hdr[2];
        std::vector<double> z(2 * nb, 0.0);
        const size_t max_rank = nb / 3;
        // C0: the rank-1 zero tile the compressing constructor gives (capacity nb/3): built through the reference's
        // (UV, ld, rank) constructor on a buffer sized for max_rank, then its rank set to 1
        std::vector<double> cbuf(2 * nb * max_rank, 0.0);
        double *dc = to_device(cbuf, ctx);
        CompressedTile<double> Cc(nb, nb, dc, nb, max_rank, blas::Layout::ColMajor, ctx);
        Cc.ReadjustTileRank(1, ctx);
        std::vector<int64_t> ranks;
        std::vector<double> au(nb * ka), av(ka * nb), bu(nb * ka), bv(ka * nb);
        for (size_t k = 0; k < steps; ++k) {
            if (std::fread(au.data(), 8, au.size(), f) != au.size() || std::fread(av.data(), 8, av.size(), f) != av.size() ||
                std::fread(bu.data(), 8, bu.size(), f) != bu.size() || std::fread(bv.data(), 8, bv.size(), f) != bv.size()) return 92;
            double *dau = to_device(au, ctx), *dav = to_device(av, ctx), *dbu = to_device(bu, ctx), *dbv = to_device(bv, ctx);
            CompressedTile<double> A(nb, nb, dau, dav, nb, ka, blas::Layout::ColMajor, ctx);
            CompressedTile<double> B(nb, nb, dbu, dbv, nb, ka, blas::Layout::ColMajor, ctx);
            api::HCore<double>::Gemm(1.0, A, blas::Op::NoTrans, B, blas::Op::NoTrans, 1.0, Cc, ctx, flops, unit, CompressionParameters(acc));
            ctx.Sync();
            ranks.push_back((int64_t) Cc.GetTileRank());
            memory::DestroyArray(dau, ctx); memory::DestroyArray(dav, ctx); memory::DestroyArray(dbu, ctx); memory::DestroyArray(dbv, ctx);
        }
        std::fclose(f);
        auto dense = dense_of(Cc, ctx);
        FILE *o = std::fopen(argv[2], "wb");
        std::fwrite(ranks.data(), sizeof(int64_t), ranks.size(), o);
        std::fwrite(dense.data(), sizeof(double), dense.size(), o);
        std::fclose(o);
        std::printf("k-sum of %zu steps through the reference's tile-at-a-time flow: ranks", steps);
        for (auto r : ranks) std::printf(" %lld", (long long) r);
        std::printf("\n");
    }
    std::printf("%d failure(s)\n", failures);
    return failures;
}
