// examples/matrix_multiplication/main.cpp -- the reference's headline driver (examples/matrix_multiplication/omp_main.cpp:
// same command line, same flow, same CSV lines) on top of the C++ mirror of its API (include/hcorepp_b200/hcorepp.hpp ->
// C ABI -> CUDA):
//     b200-hcorepp-matrix [matrix_tiles = 2] [accuracy list = "1e-1,1e-4,1e-8"] [tile_size = 512] [per_tile_generation = 0]
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

static RawMatrix<double> make_matrix(size_t tiles, size_t tile_size, bool per_tile, std::mt19937_64 &rng) {
    const size_t n = tiles * tile_size;
    RawMatrix<double> M(n, n);
    if (!per_tile) generate(M.GetData(), n, n, rng);
    else
        for (size_t c = 0; c < tiles; ++c)
            for (size_t r = 0; r < tiles; ++r) generate(M.GetData() + r * tile_size + c * tile_size * n, tile_size, n, rng);
    return M;
}

int main(int argc, char **argv) {
    int tile_size = 512, matrix_tiles = 2, per_tile_generation = 0;
    std::vector<double> accuracy_list = {1e-1, 1e-4, 1e-8};
    if (argc > 1) matrix_tiles = atoi(argv[1]);
    if (argc > 2) {
        accuracy_list.clear();
        std::stringstream ss(argv[2]);
        for (double v; ss >> v;) {
            accuracy_list.push_back(v);
            if (ss.peek() == ',') ss.ignore();
        }
    }
    if (argc > 3) tile_size = atoi(argv[3]);
    if (argc > 4) per_tile_generation = atoi(argv[4]);
    const char *verbose = std::getenv("HCOREPP_VERBOSE");
    bool print_header = verbose && std::string(verbose) == "ON";
    try {
        RunContext &context = kernels::ContextManager::GetInstance().GetContext();
        double alpha = 1, beta = 1;
        const size_t n = (size_t) matrix_tiles * tile_size;
        std::mt19937_64 rng(1);
        double t0 = now_ms();
        RawMatrix<double> full_a = make_matrix(matrix_tiles, tile_size, per_tile_generation > 0, rng);
        RawMatrix<double> full_b = make_matrix(matrix_tiles, tile_size, per_tile_generation > 0, rng);
        RawMatrix<double> full_c(n, n), initial_c(n, n);
        const double t_generation = now_ms() - t0;
        // reference solution: one dense GEMM on the device (omp_main.cpp:258-289)
        double t_ref;
        {
            double *a = memory::AllocateArray<double>(n * n, context), *b = memory::AllocateArray<double>(n * n, context),
                   *c = memory::AllocateArray<double>(n * n, context);
            memory::Memcpy<double>(a, full_a.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            memory::Memcpy<double>(b, full_b.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            memory::Memcpy<double>(c, full_c.GetData(), n * n, context, memory::MemoryTransfer::HOST_TO_DEVICE);
            context.Sync();
            t0 = now_ms();
            kernels::HCoreKernels<double>::Gemm(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, n, n, n, alpha, a, n, b, n,
                                                beta, c, n, context);
            context.Sync();
            t_ref = now_ms() - t0;
            memory::Memcpy<double>(full_c.GetData(), c, n * n, context, memory::MemoryTransfer::DEVICE_TO_HOST);
            context.Sync();
            hcb_free(context.Handle(), a); hcb_free(context.Handle(), b); hcb_free(context.Handle(), c);
        }
        const size_t ref_flops = 2 * n * n * n, ref_kb = 3 * n * n * sizeof(double) / 1024;
        const double a_norm = full_a.Norm(), b_norm = full_b.Norm(), c_init_norm = initial_c.Norm();
        const size_t tile_gemms = (size_t) matrix_tiles * matrix_tiles * matrix_tiles;
        const size_t tile_flops = tile_gemms * 2 * (size_t) tile_size * tile_size * tile_size;   // what HCore::Gemm adds to aFlops
        int failures = 0;
        // dense flow (omp_main.cpp:305-335)
        double t_dense_creation, t_dense_gemm, dense_error, dense_error_normalized;
        size_t dense_kb;
        {
            CompressionParameters none(1e-9);
            t0 = now_ms();
            TileMatrix<double> a(full_a, tile_size, tile_size, context), b(full_b, tile_size, tile_size, context),
                c(initial_c, tile_size, tile_size, context);
            context.Sync();
            t_dense_creation = now_ms() - t0;
            t0 = now_ms();
            TileMatrixMultiplication<double>(a, b, c, alpha, beta, none, context);
            context.Sync();
            t_dense_gemm = now_ms() - t0;
            RawMatrix<double> got = c.ToRawMatrix(context);
            got.ReferenceDifference(full_c);
            dense_error = got.Norm();
            dense_error_normalized = dense_error / ((a_norm + b_norm + c_init_norm) * std::numeric_limits<double>::epsilon() * (double) n);
            if (dense_error_normalized >= 10) { std::printf("Example didn't pass, dense HCore++ error > 10 \n"); ++failures; }
            dense_kb = (a.GetMemoryFootprint() + b.GetMemoryFootprint() + c.GetMemoryFootprint()) / 1024;
        }
        bool first_print = true;
        for (double accuracy : accuracy_list) {
            CompressionParameters prm(accuracy);
            for (int pass = 0; pass < 2; ++pass) {   // pass 0 = warm-up, like the reference (omp_main.cpp:341-348)
                t0 = now_ms();
                TileMatrix<double> a(full_a, tile_size, tile_size, prm, context), b(full_b, tile_size, tile_size, prm, context),
                    c(initial_c, tile_size, tile_size, prm, context);
                context.Sync();
                const double t_creation = now_ms() - t0;
                t0 = now_ms();
                TileMatrixMultiplication<double>(a, b, c, alpha, beta, prm, context);
                context.Sync();
                const double t_gemm = now_ms() - t0;
                if (pass == 0) continue;
                RawMatrix<double> got = c.ToRawMatrix(context);
                got.ReferenceDifference(full_c);
                const double err = got.Norm(), err_n = err / ((a_norm + b_norm + c_init_norm) * accuracy * (double) n);
                if (err_n >= 10) { std::printf("Example didn't pass, compressed HCore++ error > 10 \n"); ++failures; }
                const size_t kb = (a.GetMemoryFootprint() + b.GetMemoryFootprint() + c.GetMemoryFootprint()) / 1024;
                if (first_print) {
                    if (print_header)
                        std::printf("tile_count, tile_size, matrix_size, type, error, error_normalized, memory(KB), creation(ms), gemm_time(ms), flops\n");
                    std::printf("%d, %d, %d, ref, 0, 0, %zu, %f, %f, %zu\n", matrix_tiles, tile_size, matrix_tiles * tile_size, ref_kb,
                                t_generation, t_ref, ref_flops);
                    std::printf("%d, %d, %d, dense, %e, %e, %zu, %f, %f, %zu\n", matrix_tiles, tile_size, matrix_tiles * tile_size,
                                dense_error, dense_error_normalized, dense_kb, t_dense_creation, t_dense_gemm, tile_flops);
                    first_print = false;
                }
                std::printf("%d, %d, %d, %2.1e, %e, %e, %zu, %f, %f, %zu\n", matrix_tiles, tile_size, matrix_tiles * tile_size, accuracy,
                            err, err_n, kb, t_creation, t_gemm, tile_flops);
            }
        }
        return failures;
    } catch (const std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 100;
    }
}
