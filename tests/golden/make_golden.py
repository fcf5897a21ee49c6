"""Generate tests/golden/ref_*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref/libhcorepp_ref.so, built from the
unmodified sources under /root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin oracle/tlr_oracle.py (tests/test_oracle.py) and are also replayed against the CUDA path
(tests/test_gpu_parity.py) on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MIXES = ["DDD", "DDC", "DCD", "DCC", "CDD", "CDC", "CCD", "CCC"]


def single_tile_mixes(n, dtype, acc, alpha=3.5, beta=2.5):
    """tests/api/AdvancedGemmTest.cpp:31-120 inputs (generate_dense_matrix x3 from one seed, compress_dense_matrix)."""
    seed = np.array([0, 0, 0, 1], dtype=np.int64)
    A = ref.generate_dense(n, n, dtype, seed)
    B = ref.generate_dense(n, n, dtype, seed)
    Cm = ref.generate_dense(n, n, dtype, seed)
    AU, AV = ref.compress_dense(A, acc)
    BU, BV = ref.compress_dense(B, acc)
    CU, CV = ref.compress_dense(Cm, acc)
    out = dict(A=A, B=B, C=Cm, AU=AU, AV=AV, BU=BU, BV=BV, CU=CU, CV=CV, acc=acc, alpha=alpha, beta=beta)
    p = ref.Params(acc)
    for mix in MIXES:
        mk = lambda kind, D, U, V: ref.RefTile.dense(D) if kind == "D" else ref.RefTile.from_uv(U, V)
        ta, tb, tc = mk(mix[0], A, AU, AV), mk(mix[1], B, BU, BV), mk(mix[2], Cm, CU, CV)
        flops = ref.gemm(alpha, ta, False, tb, False, beta, tc, p)
        out[f"{mix}_out"] = tc.to_dense()
        out[f"{mix}_rank"] = tc.info()["rank"]
        out[f"{mix}_flops"] = flops
        if mix[2] == "C":
            u, v = tc.read()
            out[f"{mix}_U"], out[f"{mix}_V"] = u, v
    return out


def multi_tile(nb, T, acc, dtype=np.float64):
    """examples/matrix_multiplication/omp_main.cpp flow with the per-tile LATMS generator: compress A,B; C0 = 0."""
    A, B = ref.latms_law(nb * T, nb * T, dtype, tile_size=nb, reps=2)
    p = ref.Params(acc)
    tile = lambda M, j, i: M[j * nb:(j + 1) * nb, i * nb:(i + 1) * nb]
    Z = np.zeros((nb, nb), dtype=dtype)
    rA = [[ref.RefTile.compress(tile(A, j, k), p) for k in range(T)] for j in range(T)]
    rB = [[ref.RefTile.compress(tile(B, j, k), p) for k in range(T)] for j in range(T)]
    rC = [[ref.RefTile.compress(Z, p) for _ in range(T)] for _ in range(T)]
    out = dict(nb=nb, T=T, acc=acc)
    for name, g in (("A", rA), ("B", rB), ("C0", rC)):
        out[f"{name}_ranks"] = np.array([[t.info()["rank"] for t in r] for r in g])
        for j in range(T):
            for i in range(T):
                u, v = g[j][i].read()
                out[f"{name}_U_{j}_{i}"], out[f"{name}_V_{j}_{i}"] = u, v
    _, flops = ref.matmul(rA, rB, rC, 1.0, 1.0, p, nthreads=1)
    out["flops"] = flops
    out["C_ranks"] = np.array([[t.info()["rank"] for t in r] for r in rC])
    out["C_dense"] = np.block([[t.to_dense() for t in r] for r in rC])
    return out


def kernels(dtype):
    rng = np.random.default_rng(7)
    m, n = 40, 7
    a = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
    qr = a.copy(order="F")
    tau = np.zeros(n, dtype=dtype)
    ref.fn("k_geqrf", dtype)(m, n, ref.ptr(qr), m, ref.ptr(tau))
    q = qr.copy(order="F")
    ref.fn("k_ungqr", dtype)(m, n, n, ref.ptr(q), m, ref.ptr(tau))
    core = np.asfortranarray(rng.standard_normal((9, 9)).astype(dtype) * (10.0 ** -np.arange(9))[None, :].astype(dtype))
    work = core.copy(order="F")
    s = np.zeros(9, dtype=dtype)
    u = np.zeros((9, 9), dtype=dtype, order="F")
    vt = np.zeros((9, 9), dtype=dtype, order="F")
    ref.fn("k_svd", dtype)(9, 9, ref.ptr(work), 9, ref.ptr(s), ref.ptr(u), 9, ref.ptr(vt), 9, 1)
    return dict(geqrf_in=a, geqrf_qr=qr, geqrf_tau=tau, ungqr_q=q, svd_in=core, svd_s=s, svd_u=u, svd_vt=vt)


def compress_case(nb, acc, dtype):
    (A,) = ref.latms_law(nb, nb, dtype, tile_size=0, reps=1)
    t = ref.RefTile.compress(A, ref.Params(acc))
    u, v = t.read()
    return dict(A=A, U=u, V=v, rank=t.info()["rank"], max_rank=t.info()["max_rank"], acc=acc)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(OUT, "ref_mixes_f64.npz"), **single_tile_mixes(64, np.float64, 1e-4))
    np.savez_compressed(os.path.join(OUT, "ref_mixes_f32.npz"), **single_tile_mixes(64, np.float32, 1e-4))
    np.savez_compressed(os.path.join(OUT, "ref_multitile_f64.npz"), **multi_tile(64, 3, 1e-6))
    np.savez_compressed(os.path.join(OUT, "ref_kernels_f64.npz"), **kernels(np.float64))
    np.savez_compressed(os.path.join(OUT, "ref_kernels_f32.npz"), **kernels(np.float32))
    np.savez_compressed(os.path.join(OUT, "ref_compress_f64.npz"), **compress_case(120, 1e-8, np.float64))
    np.savez_compressed(os.path.join(OUT, "ref_compress_f32.npz"), **compress_case(120, 1e-4, np.float32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
