// examples/matrix_multiplication/main.cpp -- the reference's headline driver (examples/matrix_multiplication/omp_main.cpp:
// same command line, same flow, same CSV lines) on top of the C++ mirror of its API (include/hcorepp_b200/hcorepp.hpp ->
// C ABI -> CUDA):
//     b200-hcorepp-matrix [nTiles = 2] [accuracy list = "1e-1,1e-4,1e-8"] [tileDim = 512] [perTileGen = 0]
//     HCOREPP_VERBOSE=ON prints the CSV header (omp_main.cpp:173-202).
// Flow (omp_main.cpp:219-418): generate A, B with the LATMS spectrum law, C = 0; reference dense GEMM on the device;
// dense tile flow; per accuracy: compressed tile matrices (compression on the device), the tile GEMM (ONE batched device
// call per k), reconstruction, error against the dense reference, normalised like the reference
// (error / ((|A| + |B| + |C0|) * accuracy * min(M, N)) must stay below 10), memory footprint in KB, times in ms.
// The matrix generator is this file's own (the reference's wraps LAPACK dlatms on the host, outside the hot path): singular
// values from the reference's law (LatmsGenerator.cpp:36-53), orthogonal factors = products of random Householder reflectors.
#include <hcorepp_b200/hcorepp.hpp>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <sstream>
#include <string>
#include <vector>

using namespace hcorepp;
using namespace hcorepp::helpers;
using hcorepp::kernels::RunContext;
using hcorepp::operators::CompressionParameters;

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// sigma_i of the reference generator for an n x n matrix in double (LatmsGenerator.cpp:36-53)
static std::vector<double> spectrum(size_t n) {
    const double eps = std::numeric_limits<double>::epsilon(), sep = 10 * eps;
    std::vector<double> s(n);
    for (size_t i = 0; i < n; ++i) {
        if (i < 80) s[i] = std::pow(sep, (double) i / 80.0);
        else s[i] = sep * std::pow(eps / sep, (double) (i - 80) / (double) (n > 81 ? n - 81 : 1));
    }
    return s;
}

// A (n x n, ld) = H_1 .. H_k diag(sigma) G_k .. G_1 with random Householder reflectors H, G (an orthogonally mixed matrix with
// exactly the singular values sigma)
static void generate(double *A, size_t n, size_t ld, std::mt19937_64 &rng, int reflectors = 24) {
    const std::vector<double> sig = spectrum(n);
    for (size_t j = 0; j < n; ++j)
        for (size_t i = 0; i < n; ++i) A[i + j * ld] = (i == j) ? sig[i] : 0.0;
    std::normal_distribution<double> g(0.0, 1.0);
    std::vector<double> v(n), w(n);
    for (int side = 0; side < 2; ++side)
        for (int r = 0; r < reflectors; ++r) {
            double nv = 0;
            for (auto &x : v) { x = g(rng); nv += x * x; }
            const double sc = 2.0 / nv;
            if (side == 0) {  // A := (I - sc v v^T) A
                for (size_t j = 0; j < n; ++j) {
                    double d = 0;
                    for (size_t i = 0; i < n; ++i) d += v[i] * A[i + j * ld];
                    d *= sc;
                    for (size_t i = 0; i < n; ++i) A[i + j * ld] -= d * v[i];
                }
            } else {          // A := A (I - sc v v^T)
                std::fill(w.begin(), w.end(), 0.0);
                for (size_t j = 0; j < n; ++j)
                    for (size_t i = 0; i < n; ++i) w[i] += A[i + j * ld] * v[j];
                for (size_t j = 0; j < n; ++j)
                    for (size_t i = 0; i < n; ++i) A[i + j * ld] -= sc * w[i] * v[j];
            }
        }
}

static RawMatrix<double> make_matrix(size_t tiles, size_t tileDim, bool per_tile, std::mt19937_64 &rng) {
    const size_t n = tiles * tileDim;
    RawMatrix<double> M(n, n);
    if (!per_tile) generate(M.GetData(), n, n, rng);
    else
        for (size_t c = 0; c < tiles; ++c)
            for (size_t r = 0; r < tiles; ++r) generate(M.GetData() + r * tileDim + c * tileDim * n, tileDim, n, rng);
    return M;
}

int main(int argc, char **argv) {
    int tileDim = 512, nTiles = 2, perTileGen = 0;
    std::vector<double> accuracies = {1e-1, 1e-4, 1e-8};
    if (argc > 1) nTiles = atoi(argv[1]);
    if (argc > 2) {
        accuracies.clear();
        std::stringstream ss(argv[2]);
        for (double v; ss >> v;) {
            accuracies.push_back(v);
            if (ss.peek() == ',') ss.ignore();
        }
    }
    if (argc > 3) tileDim = atoi(argv[3]);
    if (argc > 4) perTileGen = atoi(argv[4]);
    const char *verbose = std::getenv("HCOREPP_VERBOSE");
    bool wantHeader = verbose && std::string(verbose) == "ON";
    try {
        RunContext &context = kernels::ContextManager::GetInstance().GetContext();
        double alpha = 1, beta = 1;
        const size_t n = (size_t) nTiles * tileDim;
        std::mt19937_64 rng(1);
        double t0 = now_ms();
        RawMatrix<double> hostA = make_matrix(nTiles, tileDim, perTileGen > 0, rng);
        RawMatrix<double> hostB = make_matrix(nTiles, tileDim, perTileGen > 0, rng);
        RawMatrix<double> refC(n, n), zeroC(n, n);
        const double msGenerate = now_ms() - t0;
        // reference solution: one dense GEMM on the device (omp_main.cpp:258-289)
        double msRefGemm;
        {
            double *a = memory::AllocateArray<double>(n * n, context), *b = memory::AllocateArray<double>(n * n, context),
                   *c = memory::AllocateArray<double>(n * n, context);
            memory::Memcpy<double>(a, hostA.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            memory::Memcpy<double>(b, hostB.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            memory::Memcpy<double>(c, refC.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            context.Sync();
            t0 = now_ms();
            kernels::HCoreKernels<double>::Gemm(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, n, n, n, alpha, a, n, b, n,
                                                beta, c, n, context);
            context.Sync();
            msRefGemm = now_ms() - t0;
            memory::Memcpy<double>(refC.GetData(), c, n * n, context, memory::MemoryTransfer::DEVICE_TO_HOST);
            context.Sync();
            hcb_free(context.Handle(), a); hcb_free(context.Handle(), b); hcb_free(context.Handle(), c);
        }
        const size_t flopsRef = 2 * n * n * n, kbRef = 3 * n * n * sizeof(double) / 1024;
        const double normA = hostA.Norm(), normB = hostB.Norm(), normC0 = zeroC.Norm();
        const size_t nTileGemms = (size_t) nTiles * nTiles * nTiles;
        const size_t flopsTiles = nTileGemms * 2 * (size_t) tileDim * tileDim * tileDim;   // what HCore::Gemm adds to aFlops
        int nFailed = 0;
        // dense flow (omp_main.cpp:305-335)
        double msDenseBuild, msDenseGemm, denseErr, denseErrScaled;
        size_t kbDense;
        {
            CompressionParameters none(1e-9);
            t0 = now_ms();
            TileMatrix<double> a(hostA, tileDim, tileDim, context), b(hostB, tileDim, tileDim, context),
                c(zeroC, tileDim, tileDim, context);
            context.Sync();
            msDenseBuild = now_ms() - t0;
            t0 = now_ms();
            TileMatrixMultiplication<double>(a, b, c, alpha, beta, none, context);
            context.Sync();
            msDenseGemm = now_ms() - t0;
            RawMatrix<double> got = c.ToRawMatrix(context);
            got.ReferenceDifference(refC);
            denseErr = got.Norm();
            denseErrScaled = denseErr / ((normA + normB + normC0) * std::numeric_limits<double>::epsilon() * (double) n);
            if (denseErrScaled >= 10) { std::printf("Example didn't pass, dense HCore++ error > 10 \n"); ++nFailed; }
            kbDense = (a.GetMemoryFootprint() + b.GetMemoryFootprint() + c.GetMemoryFootprint()) / 1024;
        }
        bool headerPending = true;
        for (double accuracy : accuracies) {
            CompressionParameters prm(accuracy);
            for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up, like the reference (omp_main.cpp:341-348)
                t0 = now_ms();
                TileMatrix<double> a(hostA, tileDim, tileDim, prm, context), b(hostB, tileDim, tileDim, prm, context),
                    c(zeroC, tileDim, tileDim, prm, context);
                context.Sync();
                const double msBuild = now_ms() - t0;
                t0 = now_ms();
                TileMatrixMultiplication<double>(a, b, c, alpha, beta, prm, context);
                context.Sync();
                const double msGemm = now_ms() - t0;
                if (pass == 0) continue;
                RawMatrix<double> got = c.ToRawMatrix(context);
                got.ReferenceDifference(refC);
                const double err = got.Norm(), errScaled = err / ((normA + normB + normC0) * accuracy * (double) n);
                if (errScaled >= 10) { std::printf("Example didn't pass, compressed HCore++ error > 10 \n"); ++nFailed; }
                const size_t kb = (a.GetMemoryFootprint() + b.GetMemoryFootprint() + c.GetMemoryFootprint()) / 1024;
                if (headerPending) {
                    if (wantHeader)
                        std::printf("tile_count, tileDim, matrix_size, type, error, error_normalized, memory(KB), creation(ms), gemm_time(ms), flops\n");
                    std::printf("%d, %d, %d, ref, 0, 0, %zu, %f, %f, %zu\n", nTiles, tileDim, nTiles * tileDim, kbRef,
                                msGenerate, msRefGemm, flopsRef);
                    std::printf("%d, %d, %d, dense, %e, %e, %zu, %f, %f, %zu\n", nTiles, tileDim, nTiles * tileDim,
                                denseErr, denseErrScaled, kbDense, msDenseBuild, msDenseGemm, flopsTiles);
                    headerPending = false;
                }
                std::printf("%d, %d, %d, %2.1e, %e, %e, %zu, %f, %f, %zu\n", nTiles, tileDim, nTiles * tileDim, accuracy,
                            err, errScaled, kb, msBuild, msGemm, flopsTiles);
            }
        }
        return nFailed;
    } catch (const std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 100;
    }
}
