// tests/cpp/test_api.cpp -- the reference's own operator/API tests, replayed through the C++ host layer
// (include/hcorepp_b200/hcorepp.hpp -> C ABI -> CUDA).  Known answers transcribed from
//   tests/operators/TestCompressedTile.cpp:124-260, tests/api/TestGemm.cpp:27-1046, tests/operators/TestDenseTile.cpp.
// Prints one line per case; exit code = number of failures.  Needs a CUDA device (no CPU fallback).
#include <hcorepp_b200/hcorepp.hpp>

#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

using namespace hcorepp;
using namespace hcorepp::operators;
using hcorepp::api::HCore;
using hcorepp::kernels::RunContext;

template<typename T> using Mat = std::vector<std::vector<T>>;  // semantic matrix: rows of columns

template<typename T>
static std::vector<T> colmajor(const Mat<double> &m) {
    const size_t r = m.size(), c = m[0].size();
    std::vector<T> out(r * c);
    for (size_t i = 0; i < r; ++i)
        for (size_t j = 0; j < c; ++j) out[i + j * r] = (T) m[i][j];
    return out;
}

template<typename T>
static std::vector<T> to_host(const T *d, size_t n, const RunContext &ctx) {
    std::vector<T> h(n);
    memory::Memcpy<T>(h.data(), d, n, ctx, memory::MemoryTransfer::DEVICE_TO_HOST);
    ctx.Sync();
    return h;
}

// product of a tile as a dense column-major host matrix
template<typename T>
static std::vector<T> dense_of(Tile<T> &t, const RunContext &ctx) {
    const size_t m = t.GetNumOfRows(), n = t.GetNumOfCols();
    if (t.isDense()) return to_host<T>(t.GetTileSubMatrix(0), m * n, ctx);
    auto &c = static_cast<CompressedTile<T> &>(t);
    const size_t rk = c.GetTileRank();
    auto U = to_host<T>(c.GetUMatrix(), m * rk, ctx), V = to_host<T>(c.GetVMatrix(), rk * n, ctx);
    std::vector<T> out(m * n, 0);
    for (size_t j = 0; j < n; ++j)
        for (size_t l = 0; l < rk; ++l)
            for (size_t i = 0; i < m; ++i) out[i + j * m] += U[i + l * m] * V[l + j * rk];
    return out;
}

template<typename T>
static bool approx(const std::vector<T> &got, const Mat<double> &want, double tol = 1e-2) {  // Catch Approx().epsilon(1e-2)
    const size_t r = want.size(), c = want[0].size();
    for (size_t i = 0; i < r; ++i)
        for (size_t j = 0; j < c; ++j) {
            const double g = got[i + j * r], w = want[i][j];
            if (std::fabs(g - w) > tol * std::max(std::fabs(g), std::fabs(w)) + 1e-5) return false;
        }
    return true;
}

static int failures = 0;
static void report(const char *name, const char *type, bool ok) {
    std::printf("%-34s %-6s %s\n", name, type, ok ? "PASS" : "FAIL");
    if (!ok) ++failures;
}

template<typename T>
static void run(const char *type) {
    RunContext &ctx = kernels::ContextManager::GetInstance().GetContext();
    auto &unit = dataunits::MemoryHandler<T>::GetInstance().GetMemoryUnit();
    const CompressionParameters eps_params(std::numeric_limits<T>::epsilon());
    size_t flops = 0;
    auto dense = [&](const Mat<double> &m) {
        auto h = colmajor<T>(m);
        return new DenseTile<T>(m.size(), m[0].size(), h.data(), m.size(), blas::Layout::ColMajor, ctx);
    };
    auto comp = [&](const Mat<double> &u, const Mat<double> &v) {
        auto hu = colmajor<T>(u), hv = colmajor<T>(v);
        return new CompressedTile<T>(u.size(), v[0].size(), hu.data(), hv.data(), u.size(), u[0].size(), blas::Layout::ColMajor, ctx);
    };
    auto zeros_c = [&](size_t m, size_t n, size_t rank) {
        std::vector<T> z((m + n) * rank, 0);
        return new CompressedTile<T>(m, n, z.data(), m, rank, ctx);
    };
    auto zeros_d = [&](size_t m, size_t n) { return new DenseTile<T>(m, n, nullptr, m, blas::Layout::ColMajor, ctx); };

    {  // TestGemm.cpp:27 -- DDD
        auto *A = dense({{1, 2, 3}, {4, 5, 6}, {7, 8, 9}}), *B = dense({{2, 4, 6}, {8, 10, 12}, {14, 16, 18}}), *C = zeros_d(3, 3);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit);
        report("TestGemm 1 (DDD)", type, approx(dense_of(*C, ctx), {{60, 72, 84}, {132, 162, 192}, {204, 252, 300}}));
        delete A; delete B; delete C;
    }
    {  // RowMajor dense tiles (Dense.cpp:69-96): the same product with every buffer row-major, ragged shapes, op(B) = B^T
        const Mat<double> a = {{1, 2, 3}, {4, 5, 6}}, b = {{1, 0, 2}, {0, 1, 1}, {3, 1, 0}, {2, 2, 2}};  // A 2x3, B 4x3: C = A B^T (2x4)
        auto rowmajor = [](const Mat<double> &m) { std::vector<T> o; for (auto &r : m) for (double v : r) o.push_back((T) v); return o; };
        auto ha = rowmajor(a), hb = rowmajor(b);
        DenseTile<T> A(2, 3, ha.data(), 3, blas::Layout::RowMajor, ctx), B(4, 3, hb.data(), 3, blas::Layout::RowMajor, ctx);
        std::vector<T> hc = rowmajor({{1, 1, 1, 1}, {2, 2, 2, 2}});
        DenseTile<T> C(2, 4, hc.data(), 4, blas::Layout::RowMajor, ctx);
        HCore<T>::Gemm(2, A, blas::Op::NoTrans, B, blas::Op::Trans, 1, C, ctx, flops, unit);
        auto got = to_host<T>(C.GetTileSubMatrix(0), 8, ctx);  // row-major 2 x 4
        const double want[8] = {2 * 7 + 1, 2 * 5 + 1, 2 * 5 + 1, 2 * 12 + 1, 2 * 16 + 2, 2 * 11 + 2, 2 * 17 + 2, 2 * 30 + 2};
        bool ok = true;
        for (int i = 0; i < 8; ++i) ok = ok && std::fabs(got[i] - want[i]) < 1e-4;
        bool mixed_throws = false;
        try { auto *Cc = zeros_d(2, 4); try { HCore<T>::Gemm(1, A, blas::Op::NoTrans, B, blas::Op::Trans, 1, *Cc, ctx, flops, unit); } catch (const std::invalid_argument &) { mixed_throws = true; } delete Cc; } catch (...) {}
        report("RowMajor DDD (A B^T, ragged)", type, ok && mixed_throws);
    }
    {  // TestGemm.cpp:206 -- CDD
        auto *A = comp({{1}, {4}, {7}}, {{10, 11, 12}}); auto *B = dense({{2, 4}, {8, 10}, {14, 16}}); auto *C = zeros_d(3, 2);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit);
        report("TestGemm 4 (CDD)", type, approx(dense_of(*C, ctx), {{276, 342}, {1104, 1368}, {1932, 2394}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:289 -- DCD
        auto *A = dense({{2, 1, 4}, {8, 5, 10}, {14, 10, 16}}); auto *B = comp({{1}, {4}, {7}}, {{10, 11, 12}}); auto *C = zeros_d(3, 3);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit);
        report("TestGemm 5 (DCD)", type, approx(dense_of(*C, ctx), {{340, 374, 408}, {980, 1078, 1176}, {1660, 1826, 1992}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:370 -- CCD
        auto *A = comp({{2, 4}, {8, 10}, {14, 16}}, {{1, 5, 6, 7}, {3, 4, 8, 9}});
        auto *B = comp({{2, 10, 18}, {4, 12, 20}, {6, 14, 22}, {8, 16, 24}}, {{5, 25}, {10, 30}, {15, 35}});
        auto *C = zeros_d(3, 2);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit);
        report("TestGemm 6 (CCD)", type, approx(dense_of(*C, ctx), {{66760, 178840}, {195400, 523480}, {324040, 868120}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:468 -- CDC, C rank 1 (capacity 1): result is the best rank-1 approximation
        auto *A = comp({{2, 8}, {4, 10}, {6, 12}}, {{5, 15, 25}, {10, 20, 30}}); auto *B = dense({{1, 9}, {5, 10}, {7, 12}});
        auto *C = zeros_c(3, 2, 1);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        report("TestGemm 7 (CDC)", type, approx(dense_of(*C, ctx), {{3070, 6190}, {4220, 8480}, {5370, 10770}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:605 -- DCC
        auto *A = dense({{2, 8, 14, 20}, {4, 10, 16, 22}, {6, 12, 18, 24}});
        auto *B = comp({{1, 9}, {3, 11}, {5, 13}, {7, 15}}, {{5, 15}, {10, 20}});
        auto *C = zeros_c(3, 2, 1);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        report("TestGemm 8 (DCC)", type, approx(dense_of(*C, ctx), {{7060, 15300}, {8180, 17700}, {9300, 20100}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:748 -- CCC
        auto *A = comp({{1, 10, 20}, {2, 11, 21}, {3, 12, 22}, {4, 13, 23}, {5, 14, 24}}, {{2, 8, 14, 20}, {4, 10, 16, 22}, {6, 12, 18, 24}});
        auto *B = comp({{1, 9}, {3, 11}, {5, 13}, {7, 15}}, {{5, 15}, {10, 20}});
        auto *C = zeros_c(5, 2, 1);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        report("TestGemm 9 (CCC)", type,
               approx(dense_of(*C, ctx), {{274860, 594300}, {299400, 647400}, {323940, 700500}, {348480, 753600}, {373020, 806700}}));
        delete A; delete B; delete C;
    }
    {  // TestGemm.cpp:917 -- DDC: C becomes full rank (U = A*B, V = I)
        auto *A = dense({{1, 10, 20}, {2, 11, 21}, {3, 12, 22}, {4, 13, 23}, {5, 14, 24}});
        auto *B = dense({{2, 8, 14, 20}, {4, 10, 16, 22}, {6, 12, 18, 24}});
        auto *C = zeros_c(5, 4, 4);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        const bool ok = approx(dense_of(*C, ctx), {{162, 348, 534, 720}, {174, 378, 582, 786}, {186, 408, 630, 852}, {198, 438, 678, 918}, {210, 468, 726, 984}});
        report("TestGemm 10 (DDC)", type, ok && C->GetTileRank() == 4);
        delete A; delete B; delete C;
    }
    {  // TestCompressedTile.cpp:124-260: C(3x2, rank 2 zeros) += A(3x3) * B(3x2) with acc = eps, as a DCC call (BU = I)
        auto *A = dense({{1, 4, 7}, {2, 5, 8}, {3, 6, 9}});
        auto *B = comp({{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, {{2, 8}, {4, 10}, {6, 12}});
        auto *C = zeros_c(3, 2, 2);
        HCore<T>::Gemm(1, *A, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        report("CompressedTile::Gemm known answer", type, approx(dense_of(*C, ctx), {{60, 132}, {72, 162}, {84, 192}}) && C->GetTileRank() == 2);
        // layout contract (TestCompressedTile.cpp:29-121): V sits at m*maxRank, ldU = m, ldV = rank
        report("CompressedTile layout", type, C->GetVMatrix() == C->GetUMatrix() + 3 * 2 && C->GetULeadingDim() == 3 && C->GetTileStride(1) == 2);
        delete A; delete B; delete C;
    }
    {  // compressing constructor (Compressed.cpp:75-146): rank-2 matrix, acc 1e-6 -> rank 2, maxRank = min/3
        const size_t n = 12;
        Mat<double> m(n, std::vector<double>(n));
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) m[i][j] = std::sin(0.3 * i) * std::cos(0.2 * j) + 0.5 * (i + 1.0) * (j + 2.0) / (n * n);
        auto h = colmajor<T>(m);
        CompressedTile<T> Ct(n, n, h.data(), n, CompressionParameters(1e-5), blas::Layout::ColMajor, ctx);
        report("CompressedTile compress ctor", type, Ct.GetTileRank() == 2 && Ct.GetMaxRank() == 4 && approx(dense_of(Ct, ctx), m, 1e-3));
    }
    {  // error behaviour (Dense.cpp:36-42, Compressed.cpp:188-206)
        auto *A = dense({{1, 2}, {3, 4}});
        bool threw = false;
        try { A->GetTileSubMatrix(1); } catch (const std::invalid_argument &) { threw = true; }
        report("invalid sub-matrix index throws", type, threw);
        delete A;
    }
    {  // multi-tile container + driver loop (TileMatrix.hpp, omp_main.cpp:112-126): 3 x 2 x 2 tiles of 48, rank-5 tiles
        const size_t nb = 48, mt = 3, nt = 2, kt = 2, rk = 5;
        auto lowrank = [&](size_t M, size_t N, size_t tm, size_t tn, unsigned seed) {
            helpers::RawMatrix<T> raw(M, N);
            unsigned s = seed;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double) (s >> 8) / (1 << 24) - 0.5; };
            for (size_t bc = 0; bc < N / tn; ++bc)
                for (size_t br = 0; br < M / tm; ++br) {
                    std::vector<double> U(tm * rk), V(rk * tn);
                    for (auto &v : U) v = rnd();
                    for (auto &v : V) v = rnd();
                    for (size_t j = 0; j < tn; ++j)
                        for (size_t i = 0; i < tm; ++i) {
                            double acc = 0;
                            for (size_t l = 0; l < rk; ++l) acc += U[i + l * tm] * V[l + j * rk] * std::pow(0.5, (double) l);
                            raw.GetData()[(br * tm + i) + (bc * tn + j) * M] = (T) acc;
                        }
                }
            return raw;
        };
        helpers::RawMatrix<T> rA = lowrank(mt * nb, kt * nb, nb, nb, 11), rB = lowrank(kt * nb, nt * nb, nb, nb, 23);
        helpers::RawMatrix<T> rC(mt * nb, nt * nb);
        const double acc = sizeof(T) == 8 ? 1e-8 : 1e-4;
        CompressionParameters prm(acc);
        RunContext &mctx = const_cast<RunContext &>(ctx);
        helpers::TileMatrix<T> tA(rA, nb, nb, prm, mctx), tB(rB, nb, nb, prm, mctx), tC(rC, nb, nb, prm, mctx);
        helpers::TileMatrixMultiplication<T>(tA, tB, tC, (T) 1, (T) 1, prm, ctx);
        helpers::RawMatrix<T> got = tC.ToRawMatrix(mctx);
        helpers::RawMatrix<T> want(mt * nb, nt * nb);
        for (size_t j = 0; j < nt * nb; ++j)
            for (size_t l = 0; l < kt * nb; ++l)
                for (size_t i = 0; i < mt * nb; ++i)
                    want.GetData()[i + j * mt * nb] += rA.GetData()[i + l * mt * nb] * rB.GetData()[l + j * kt * nb];
        const double nrm = want.Norm();
        got.ReferenceDifference(want);
        bool ranks_ok = true;
        for (size_t i = 0; i < nt; ++i)
            for (size_t j = 0; j < mt; ++j) ranks_ok = ranks_ok && tC.GetTile(j, i)->GetTileRank() <= 2 * rk;
        report("TileMatrix + driver loop", type, got.Norm() <= 10 * acc * nrm && ranks_ok && tA.GetTile(0, 0)->GetTileRank() == rk &&
                                                     tC.GetMemoryFootprint() > 0);
    }
    {  // wire format (Tile.hpp:30-52, Compressed.cpp:769-805, TilePacker.cpp:4-26): UnPackTile -> (metadata, buffer) -> PackTile
        auto *A = dense({{1, 2}, {3, 4}, {5, 6}});
        const std::vector<T> u = colmajor<T>({{1, 0}, {0, 1}, {1, 1}}), v = colmajor<T>({{1, 2, 3, 4}, {0, 1, 0, 1}});
        std::vector<T> uh = u, vh = v;
        CompressedTile<T> Ct(3, 4, uh.data(), vh.data(), 3, 2, blas::Layout::ColMajor, ctx);
        auto pc = TilePacker<T>::UnPackTile(Ct, ctx);
        auto pd = TilePacker<T>::UnPackTile(*A, ctx);
        bool meta_ok = pc.first->mType == COMPRESSED && pc.first->mMatrixRank == 2 && pc.first->mMaxRank == 2 && pc.first->mNumOfRows == 3 &&
                       pc.first->mNumOfCols == 4 && pd.first->mType == DENSE && pd.first->mNumOfRows == 3 && pd.first->mNumOfCols == 2;
        Tile<T> *rc = TilePacker<T>::PackTile(*pc.first, pc.second, ctx);   // views over the same device buffers
        Tile<T> *rd = TilePacker<T>::PackTile(*pd.first, pd.second, ctx);
        bool same = rc->isCompressed() && rd->isDense() && rc->GetTileRank() == 2 && dense_of(*rc, ctx) == dense_of(Ct, ctx) &&
                    dense_of(*rd, ctx) == dense_of(*A, ctx) && rc->GetTileSubMatrix(0) == Ct.GetUMatrix();
        // the re-packed view is a full citizen: use it as an operand
        auto *C = zeros_c(3, 2, 2);
        auto *B = dense({{1, 0}, {0, 1}, {1, 1}, {2, 0}});
        HCore<T>::Gemm(1, *rc, blas::Op::NoTrans, *B, blas::Op::NoTrans, 1, *C, ctx, flops, unit, eps_params);
        report("PackTile / UnPackTile / TilePacker", type, meta_ok && same && approx(dense_of(*C, ctx), {{12, 5}, {2, 1}, {14, 6}}));
        delete pc.first; delete pd.first; delete rc; delete rd; delete A; delete B; delete C;
    }
    {  // per-tile fixed rank (par_fixed_rank_streams_main.cpp:465-477,540-541; Compressed.cpp:510-515)
        const size_t n = 24;
        Mat<double> ma(n, std::vector<double>(n)), mb(n, std::vector<double>(n));
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) {
                ma[i][j] = std::sin(0.31 * i + 0.1) * std::cos(0.17 * j) + 0.3 * std::cos(0.05 * i * j);
                mb[i][j] = std::cos(0.23 * i) * std::sin(0.19 * j + 0.4) + 0.2 * std::sin(0.07 * i * j);
            }
        auto ha = colmajor<T>(ma), hb = colmajor<T>(mb);
        const double acc = sizeof(T) == 8 ? 1e-9 : 1e-4;
        CompressedTile<T> A(n, n, ha.data(), n, CompressionParameters(acc), blas::Layout::ColMajor, ctx);
        CompressedTile<T> B(n, n, hb.data(), n, CompressionParameters(acc), blas::Layout::ColMajor, ctx);
        auto *C1 = zeros_c(n, n, 8);
        HCore<T>::Gemm(1, A, blas::Op::NoTrans, B, blas::Op::NoTrans, 1, *C1, ctx, flops, unit, CompressionParameters(acc));
        const size_t free_rank = C1->GetTileRank();
        auto *C2 = zeros_c(n, n, 8);
        static_cast<CompressedTile<T> *>(C2)->SetFixedRank(3);
        HCore<T>::Gemm(1, A, blas::Op::NoTrans, B, blas::Op::NoTrans, 1, *C2, ctx, flops, unit, CompressionParameters(acc));
        report("per-tile fixed rank", type, free_rank > 3 && C2->GetTileRank() == 3);
        delete C1; delete C2;
    }
    {  // Cholesky pieces (HCore.cpp:482-647): Potrf on a dense tile, Syrk, Trsm on the V buffer, the aCholesky product
        auto *A = dense({{4, 99, 99}, {2, 5, 99}, {2, 3, 6}});   // lower triangle of an SPD matrix, junk above
        HCore<T>::Potrf(*A, blas::Uplo::Lower, ctx, flops, unit);
        const double l22 = std::sqrt(6.0 - 1.0 - 1.0);
        bool potrf_ok = approx(dense_of(*A, ctx), {{2, 99, 99}, {1, 2, 99}, {1, 1, l22}}, 1e-5);
        bool threw = false;
        auto *Cc = zeros_c(3, 3, 1);
        try { HCore<T>::Potrf(*Cc, blas::Uplo::Lower, ctx, flops, unit); } catch (const std::runtime_error &) { threw = true; }
        // Syrk, dense A: C := -A A^T + C on the lower triangle only
        auto *S = dense({{1, 2}, {3, 4}});
        auto *Cs = dense({{10, 7}, {10, 10}});
        HCore<T>::Syrk(-1, *S, blas::Op::NoTrans, blas::Uplo::Lower, 1, *Cs, ctx, flops, unit);
        bool syrk_ok = approx(dense_of(*Cs, ctx), {{5, 7}, {-1, -15}}, 1e-5);
        // Trsm: L X = V on the V buffer of a compressed tile (viewed m x rank)
        const std::vector<T> u = colmajor<T>({{1}, {1}, {1}}), v = colmajor<T>({{2, 4, 6}});
        std::vector<T> uh = u, vh = v;
        CompressedTile<T> Bt(3, 3, uh.data(), vh.data(), 3, 1, blas::Layout::ColMajor, ctx);
        auto *L = dense({{2, 0, 0}, {1, 1, 0}, {0, 1, 2}});
        HCore<T>::Trsm(blas::Side::Left, blas::Uplo::Lower, blas::Op::NoTrans, blas::Diag::NonUnit, 1, *L, Bt, ctx, flops, unit);
        auto vb = to_host<T>(Bt.GetVMatrix(), 3, ctx);   // x = L^-1 (2,4,6)^T = (1, 3, 1.5)
        bool trsm_ok = std::fabs(vb[0] - 1) < 1e-5 && std::fabs(vb[1] - 3) < 1e-5 && std::fabs(vb[2] - 1.5) < 1e-5;
        // aCholesky product: C += alpha A B^T on compressed tiles
        std::vector<T> u2 = colmajor<T>({{1}, {0}, {2}}), v2 = colmajor<T>({{1, 1, 0}});
        CompressedTile<T> X(3, 3, u2.data(), v2.data(), 3, 1, blas::Layout::ColMajor, ctx);
        auto *Cx = zeros_c(3, 3, 3);
        HCore<T>::Gemm(-1, X, blas::Op::NoTrans, X, blas::Op::Trans, 1, *Cx, ctx, flops, unit, eps_params, true);
        bool chol_gemm_ok = approx(dense_of(*Cx, ctx), {{-2, 0, -4}, {0, 0, 0}, {-4, 0, -8}}, 1e-5);
        report("Potrf / Syrk / Trsm / aCholesky Gemm", type, potrf_ok && threw && syrk_ok && trsm_ok && chol_gemm_ok);
        delete A; delete Cc; delete S; delete Cs; delete L; delete Cx;
    }
    (void) flops;
}

int main() {
    try {
        run<double>("double");
        run<float>("float");
    } catch (const std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 100;
    }
    std::printf("%d failure(s)\n", failures);
    return failures;
}
