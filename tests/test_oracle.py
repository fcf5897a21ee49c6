"""CPU tests: pin oracle/tlr_oracle.py (the checker) against
  (i)   the reference's own known-answer vectors (tests/golden/reference_vectors.json),
  (ii)  fixtures produced by running the reference itself (tests/golden/ref_*.npz, see make_golden.py),
  (iii) the compiled reference live (oracle/_ref) when it is present.
"""
import json
import os

import numpy as np
import pytest

from oracle import tlr_oracle as O
from oracle import ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))
K = VEC["kernels"]
F = lambda x, dt=np.float64: np.asfortranarray(np.array(x, dtype=dt))
flat = lambda x, shape, dt=np.float64: np.array(x, dtype=dt).reshape(shape, order="F").copy(order="F")
DTYPES = [np.float64, np.float32]


def approx(a, b, tol=1e-2):
    """Catch Approx().epsilon(tol): |a-b| <= tol * max(|a|,|b|) (+ tiny absolute slack for exact zeros)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b)) + 1e-5)


# ---------------------------------------------------------------- (i) reference known answers
@pytest.mark.parametrize("dt", DTYPES)
def test_kernel_known_answers(dt):
    g = K["gemm"]
    assert np.array_equal(O.k_gemm(False, False, dt(1), F(g["A"], dt), F(g["B"], dt), dt(1), F(g["C0"], dt)), F(g["C"], dt))
    g = K["multiply_by_alpha"]
    a = np.array(g["flat_in"], dtype=dt)
    O.k_multiply_by_alpha(a, g["rows"], g["cols"], g["m"], g["rank"], dt(g["alpha"]))
    assert np.array_equal(a, np.array(g["flat_out"], dtype=dt))
    g = K["process_v"]
    V = O.k_process_v(g["n"], g["crank"], g["ungqr"], g["vm"], dt(g["beta"]), np.array(g["cv_flat"], dtype=dt),
                      g["ldcv"], g["arank"], np.array(g["b_flat"], dtype=dt))
    assert np.array_equal(V.reshape(-1, order="F"), np.array(g["v_flat"], dtype=dt))
    for key in ("new_rank_abs", "new_rank_rel"):
        g = K[key]
        assert O.k_new_rank(g["truncated"], np.array(g["sigma"], dtype=dt), len(g["sigma"]), dt(g["accuracy"])) == g["rank"]
    g = K["uvptr"]
    out = O.k_uvptr(g["rank"], g["vm"], flat(g["vnew_flat"], (g["vm"], g["rank"]), dt))
    assert np.array_equal(out.reshape(-1, order="F"), np.array(g["uv_flat"], dtype=dt))
    for key in ("vtnew_noungqr", "vtnew_ungqr"):
        g = K[key]
        vt = flat(g["vt_flat"], (g["size_s"], g["vm"]), dt)
        O.k_vtnew(g["rk"], g["ungqr"], min(g["vm"], g["vn"]), np.array(g["sigma"], dtype=dt), vt, g["size_s"], g["vm"])
        assert np.array_equal(vt.reshape(-1, order="F"), np.array(g["out_flat"], dtype=dt))
    n = K["fill_identity"]["n"]
    assert np.array_equal(O.k_fill_identity(n, np.zeros((n, n), dtype=dt, order="F")), np.eye(n, dtype=dt))
    g = K["lacpy"]
    for kind in "GUL":
        b = O.k_lacpy(kind, g["m"], g["n"], flat(g["a_flat"], (4, 4), dt), np.zeros((4, 4), dtype=dt, order="F"))
        assert np.array_equal(b.reshape(-1, order="F"), np.array(g[kind], dtype=dt))
    g = K["laset"]
    for kind in "GUL":
        a = O.k_laset(kind, 4, 4, dt(g["offdiag"]), dt(g["diag"]), np.zeros((4, 4), dtype=dt, order="F"))
        assert np.array_equal(a.reshape(-1, order="F"), np.array(g[kind], dtype=dt))


@pytest.mark.parametrize("dt", DTYPES)
def test_geqrf_known_answer(dt):
    g = K["geqrf"]
    a = flat(g["a_flat"], (g["m"], g["n"]), dt)
    for impl in (O.k_geqrf, O.householder_qr_numpy):
        qr, tau = impl(a)
        assert approx(qr.reshape(-1, order="F"), g["qr_flat"], g["tol_rel"]), impl.__name__
        assert approx(tau, g["tau"], g["tol_rel"])


@pytest.mark.parametrize("dt", DTYPES)
def test_svd_reconstructs(dt):
    a = F(K["svd"]["A"], dt)
    for svd in ("gesvd", "gesdd"):
        u, s, vt = O.k_svd(a, svd)
        assert approx((u * s) @ vt, a)
    u, s, vt = O.jacobi_svd_numpy(a)
    assert approx((u * s) @ vt, a)
    assert np.allclose(s, np.linalg.svd(a.astype(np.float64), compute_uv=False), atol=1e-6)


def _tile(case, name, dt):
    if name in case:
        return O.DenseTile(F(case[name], dt))
    return O.CompressedTile.from_uv(F(case[name + "U"], dt), F(case[name + "V"], dt))


@pytest.mark.parametrize("dt", DTYPES)
def test_compressed_tile_gemm_known_answer(dt):
    g = VEC["compressed_tile_gemm"]
    A, B = F(g["A"], dt), F(g["B"], dt)
    C = O.CompressedTile.from_uv(np.zeros((3, g["c_rank"]), dt), np.zeros((g["c_rank"], 2), dt))
    O.compressed_tile_gemm(C, dt(1), A, 3, 3, B, dt(1), O.CompressionParameters(float(np.finfo(dt).eps)), [0])
    assert approx(C.to_dense(), g["C"], g["tol_rel"])


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", VEC["hcore_gemm"]["cases"], ids=lambda c: c["name"])
def test_hcore_gemm_known_answers(case, dt):
    A, B = _tile(case, "A", dt), _tile(case, "B", dt)
    m, n = A.m, B.n
    if case["C0"] == "zeros":
        C = O.DenseTile(np.zeros((m, n), dt, order="F"))
    else:
        r = case["c_rank"]
        C = O.CompressedTile.from_uv(np.zeros((m, r), dt), np.zeros((r, n), dt))
    p = O.CompressionParameters(float(np.finfo(dt).eps)) if case.get("acc") == "eps" else O.CompressionParameters()
    O.hcore_gemm(dt(case["alpha"]), A, False, B, False, dt(case["beta"]), C, p)
    assert approx(C.to_dense(), case["C"], VEC["hcore_gemm"]["tol_rel"])


def test_rank_law():
    for acc, rk in VEC["examples"]["rank_law"]["double"].items():
        s = O.latms_spectrum(512, np.float64)
        assert O.k_new_rank(False, s, 512, float(acc)) == rk


# ---------------------------------------------------------------- (ii) fixtures produced by the reference itself
MIXES = ["DDD", "DDC", "DCD", "DCC", "CDD", "CDC", "CCD", "CCC"]


@pytest.mark.parametrize("dt,name", [(np.float64, "f64"), (np.float32, "f32")])
@pytest.mark.parametrize("mix", MIXES)
def test_mixes_vs_reference_fixture(mix, dt, name):
    z = np.load(os.path.join(GOLDEN, f"ref_mixes_{name}.npz"))
    mk = lambda kind, D, U, V: O.DenseTile(F(z[D], dt)) if kind == "D" else O.CompressedTile.from_uv(z[U], z[V])
    a, b, c = mk(mix[0], "A", "AU", "AV"), mk(mix[1], "B", "BU", "BV"), mk(mix[2], "C", "CU", "CV")
    flops = O.hcore_gemm(dt(z["alpha"]), a, False, b, False, dt(z["beta"]), c, O.CompressionParameters(float(z["acc"])))
    assert flops == int(z[f"{mix}_flops"])
    assert (0 if c.dense else c.rank) == int(z[f"{mix}_rank"])
    ref_out = z[f"{mix}_out"]
    tol = 1e-12 if dt == np.float64 else 2e-5
    assert np.linalg.norm(c.to_dense() - ref_out) <= tol * np.linalg.norm(ref_out)


def test_multitile_vs_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "ref_multitile_f64.npz"))
    T = int(z["T"])
    grid = lambda name: [[O.CompressedTile(F(z[f"{name}_U_{j}_{i}"]), F(z[f"{name}_V_{j}_{i}"]), int(z["nb"]) // 3)
                          for i in range(T)] for j in range(T)]
    A, B, C = grid("A"), grid("B"), grid("C0")
    flops, _ = O.tile_matmul(A, B, C, 1.0, 1.0, O.CompressionParameters(float(z["acc"])))
    assert flops == int(z["flops"])
    assert np.array_equal(np.array([[t.rank for t in r] for r in C]), z["C_ranks"])
    Cd = np.block([[t.to_dense() for t in r] for r in C])
    assert np.linalg.norm(Cd - z["C_dense"]) <= 1e-12 * np.linalg.norm(z["C_dense"])


@pytest.mark.parametrize("dt,name", [(np.float64, "f64"), (np.float32, "f32")])
def test_compress_vs_reference_fixture(dt, name):
    z = np.load(os.path.join(GOLDEN, f"ref_compress_{name}.npz"))
    t = O.CompressedTile.compress(z["A"], O.CompressionParameters(float(z["acc"])))
    assert t.rank == int(z["rank"]) and t.max_rank == int(z["max_rank"])
    ref_d = z["U"] @ z["V"]
    assert np.linalg.norm(t.to_dense() - ref_d) <= (1e-12 if dt == np.float64 else 2e-5) * np.linalg.norm(ref_d)
    # the generator law itself: leading singular values of the reference-generated matrix follow latms_spectrum
    s = np.linalg.svd(z["A"].astype(np.float64), compute_uv=False)
    assert np.allclose(s[:20], O.latms_spectrum(z["A"].shape[0], dt)[:20].astype(np.float64), rtol=1e-3 if dt == np.float32 else 1e-9)


@pytest.mark.parametrize("dt,name", [(np.float64, "f64"), (np.float32, "f32")])
def test_kernels_vs_reference_fixture(dt, name):
    z = np.load(os.path.join(GOLDEN, f"ref_kernels_{name}.npz"))
    tol = 1e-12 if dt == np.float64 else 1e-5
    qr, tau = O.k_geqrf(z["geqrf_in"])
    assert np.allclose(qr, z["geqrf_qr"], atol=tol) and np.allclose(tau, z["geqrf_tau"], atol=tol)
    qr2, tau2 = O.householder_qr_numpy(z["geqrf_in"])  # the unblocked restatement agrees with LAPACK's blocked code
    assert np.allclose(qr2, z["geqrf_qr"], atol=tol * 10) and np.allclose(tau2, z["geqrf_tau"], atol=tol * 10)
    q = O.k_ungqr(40, 7, 7, qr, tau)
    assert np.allclose(q, z["ungqr_q"], atol=tol)
    u, s, vt = O.k_svd(z["svd_in"])
    assert np.allclose(s, z["svd_s"], atol=tol)
    _, sj, _ = O.jacobi_svd_numpy(z["svd_in"])
    assert np.allclose(sj, z["svd_s"].astype(np.float64), atol=1e-12 if dt == np.float64 else 1e-6)


# ---------------------------------------------------------------- (iii) live against the compiled reference
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (make -C oracle ref)")


@needs_ref
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [33, 100])
def test_live_all_mixes(n, dt):
    seed = np.array([0, 0, 0, 1], dtype=np.int64)
    A, B, Cm = (R.generate_dense(n, n, dt, seed) for _ in range(3))
    acc = 1e-4
    uv = [R.compress_dense(M, acc) for M in (A, B, Cm)]
    for mix in MIXES:
        mats = [A, B, Cm]
        rt = [R.RefTile.dense(mats[i]) if mix[i] == "D" else R.RefTile.from_uv(*uv[i]) for i in range(3)]
        ot = [O.DenseTile(F(mats[i], dt)) if mix[i] == "D" else O.CompressedTile.from_uv(*uv[i]) for i in range(3)]
        f1 = R.gemm(3.5, rt[0], False, rt[1], False, 2.5, rt[2], R.Params(acc))
        f2 = O.hcore_gemm(3.5, ot[0], False, ot[1], False, 2.5, ot[2], O.CompressionParameters(acc))
        assert f1 == f2, mix
        d1, d2 = rt[2].to_dense(), ot[2].to_dense()
        assert np.linalg.norm(d1 - d2) <= (1e-12 if dt == np.float64 else 2e-5) * np.linalg.norm(d1), mix
        assert rt[2].info()["rank"] == (0 if ot[2].dense else ot[2].rank)


@needs_ref
def test_live_multitile_matches_and_example_line():
    """Small replay of examples/matrix_multiplication/omp_main.cpp (per-tile generator) through both paths."""
    nb, T, acc = 144, 2, 1e-8  # maxRank = 48 >= the 44 the spectrum law needs at 1e-8
    A, B = R.latms_law(nb * T, nb * T, np.float64, tile_size=nb, reps=2)
    tile = lambda M, j, i: M[j * nb:(j + 1) * nb, i * nb:(i + 1) * nb]
    Z = np.zeros((nb, nb))
    p, po = R.Params(acc), O.CompressionParameters(acc)
    rA = [[R.RefTile.compress(tile(A, j, k), p) for k in range(T)] for j in range(T)]
    rB = [[R.RefTile.compress(tile(B, j, k), p) for k in range(T)] for j in range(T)]
    rC = [[R.RefTile.compress(Z, p) for _ in range(T)] for _ in range(T)]
    oA = [[O.CompressedTile.compress(tile(A, j, k), po) for k in range(T)] for j in range(T)]
    oB = [[O.CompressedTile.compress(tile(B, j, k), po) for k in range(T)] for j in range(T)]
    oC = [[O.CompressedTile.compress(Z, po) for _ in range(T)] for _ in range(T)]
    assert [[t.info()["rank"] for t in r] for r in rA] == [[t.rank for t in r] for r in oA]
    _, f1 = R.matmul(rA, rB, rC, 1.0, 1.0, p, nthreads=2)
    f2, _ = O.tile_matmul(oA, oB, oC, 1.0, 1.0, po)
    assert f1 == f2
    assert [[t.info()["rank"] for t in r] for r in rC] == [[t.rank for t in r] for r in oC]
    Cr = np.block([[t.to_dense() for t in r] for r in rC])
    Co = np.block([[t.to_dense() for t in r] for r in oC])
    assert np.linalg.norm(Cr - Co) <= 1e-12 * np.linalg.norm(Cr)
    assert np.linalg.norm(Cr - A @ B) <= 10 * acc * np.linalg.norm(A @ B)


# ------------------------------------------------------------------------------------------------------------------
# TLR Cholesky pieces of the reference (NOT built on the GPU side yet): the reference has no driver
# and no enabled tests for them, so what it actually does is recorded here against the compiled reference.
# ------------------------------------------------------------------------------------------------------------------
@needs_ref
def test_reference_cholesky_pieces_semantics():
    from oracle import ref as R
    rng = np.random.default_rng(0)
    n, k = 64, 8
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    # Potrf (HCore.cpp:586-621): dense tiles only, LAPACK potrf in place (the other triangle is left alone)
    t = R.RefTile.dense(np.asfortranarray(S.copy()))
    R.potrf(t, "L")
    L = t.to_dense()
    assert np.allclose(np.tril(L), np.linalg.cholesky(S)) and np.allclose(np.triu(L, 1), np.triu(S, 1))
    # Syrk with a dense A (HCore.cpp:576-582): k is taken from C's column count, i.e. A must be a square tile
    A = rng.standard_normal((n, n))
    C0 = rng.standard_normal((n, n))
    C0 = C0 + C0.T
    ta, tc = R.RefTile.dense(np.asfortranarray(A.copy())), R.RefTile.dense(np.asfortranarray(C0.copy()))
    R.syrk(-1.0, ta, False, "L", 1.0, tc)
    assert np.allclose(np.tril(tc.to_dense()), np.tril(C0 - A @ A.T))
    # Trsm (HCore.cpp:624-647): B must be compressed; the solve runs on B's V buffer viewed as (rows(B) x rank) with the
    # tile's leading dimension; with side = Right only the leading rank x rank block of A is used.  U is untouched.
    Lm = np.tril(rng.standard_normal((n, n))) + n * np.eye(n)
    U, V = rng.standard_normal((n, k)), rng.standard_normal((k, n))
    tA = R.RefTile.dense(np.asfortranarray(Lm.copy()))
    tB = R.RefTile.from_uv(np.asfortranarray(U.copy()), np.asfortranarray(V.copy()))
    R.trsm("R", "L", True, "N", 1.0, tA, tB)
    U2, V2 = tB.read()
    assert np.allclose(U2, U)
    Vn = V.reshape(-1, order="F").reshape((n, k), order="F")
    got = V2.reshape(-1, order="F").reshape((n, k), order="F")
    assert np.allclose(got, Vn @ np.linalg.inv(Lm[:k, :k]).T)
    with pytest.raises(RuntimeError):
        R.potrf(tB, "L")  # "Potrf works only with dense tiles"


@needs_ref
def test_cholesky_pieces_restatement_matches_reference():
    """oracle restatement of Potrf / Syrk(dense) / Trsm == the compiled reference (groundwork for SURVEY 8f row 1)."""
    from oracle import ref as R
    rng = np.random.default_rng(5)
    n, k = 48, 6
    M = rng.standard_normal((n, n))
    S = M @ M.T + n * np.eye(n)
    t, o = R.RefTile.dense(np.asfortranarray(S.copy())), O.DenseTile(np.asfortranarray(S.copy()))
    R.potrf(t, "L")
    O.hcore_potrf(o, "L")
    assert np.allclose(t.to_dense(), o.data, rtol=1e-12, atol=1e-12)
    A = rng.standard_normal((n, n))
    C0 = rng.standard_normal((n, n))
    for uplo in ("L", "U"):
        ta, tc = R.RefTile.dense(np.asfortranarray(A.copy())), R.RefTile.dense(np.asfortranarray(C0.copy()))
        oc = O.DenseTile(np.asfortranarray(C0.copy()))
        R.syrk(-1.0, ta, False, uplo, 0.5, tc)
        O.hcore_syrk_dense(-1.0, O.DenseTile(np.asfortranarray(A.copy())), False, uplo, 0.5, oc)
        assert np.allclose(tc.to_dense(), oc.data, rtol=1e-12, atol=1e-12), uplo
    Lm = np.tril(rng.standard_normal((n, n))) + n * np.eye(n)
    U, V = rng.standard_normal((n, k)), rng.standard_normal((k, n))
    for side, trans in (("R", True), ("R", False), ("L", False), ("L", True)):
        tA = R.RefTile.dense(np.asfortranarray(Lm.copy()))
        tB = R.RefTile.from_uv(np.asfortranarray(U.copy()), np.asfortranarray(V.copy()))
        oB = O.CompressedTile.from_uv(U.copy(), V.copy())
        R.trsm(side, "L", trans, "N", 2.0, tA, tB)
        O.hcore_trsm(side, "L", trans, "N", 2.0, O.DenseTile(np.asfortranarray(Lm.copy())), oB)
        U2, V2 = tB.read()
        assert np.allclose(U2, oB.U) and np.allclose(V2, oB.V, rtol=1e-10, atol=1e-12), (side, trans)


def test_tile_cholesky_driver_composition_vs_compiled_reference():
    """The Cholesky driver loop of the oracle (the reference has tile routines but no driver) with its
    recompressing update done (a) by the numpy restatement and (b) by the COMPILED reference's HCore::Gemm (opB = Trans):
    same factor, same ranks; and the factor reproduces the covariance matrix to the compression accuracy."""
    from oracle import ref as R
    nt, nb, acc = 4, 96, 1e-6
    pts, tile = O.covariance_tiles(nt, nb, ell=0.2, nugget=1e-2, seed=3)
    p = O.CompressionParameters(acc)

    def build():
        return ([tile(k, k) for k in range(nt)],
                {(i, j): O.CompressedTile.compress(tile(i, j), p) for j in range(nt) for i in range(j + 1, nt)})
    d1, l1 = O.tile_cholesky(*build(), p)

    def ref_gemm(a, b, c):
        # the compiled reference returns NaN for a TRANSPOSED compressed operand outside its aCholesky mode (HCore.cpp:57-110
        # reads the factors with the wrong shapes), so B^T is handed over as the explicit tile (BV^T)(BU^T), NoTrans
        ta, tb = R.RefTile.from_uv(a.U, a.V), R.RefTile.from_uv(np.asfortranarray(b.V.T), np.asfortranarray(b.U.T))
        tc = R.RefTile.from_uv_cap(c.U, c.V, c.max_rank)
        R.gemm(-1.0, ta, False, tb, False, 1.0, tc, R.Params(acc))
        u, v = tc.read()
        c.U, c.V = np.asfortranarray(u), np.asfortranarray(v)
    d2, l2 = O.tile_cholesky(*build(), p, gemm=ref_gemm)
    A = np.block([[tile(i, j) for j in range(nt)] for i in range(nt)])
    for dd, ll in ((d1, l1), (d2, l2)):
        L = np.zeros_like(A)
        for k in range(nt):
            L[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = np.tril(dd[k])
        for (i, j), t in ll.items():
            L[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = t.to_dense()
        assert np.linalg.norm(A - L @ L.T) / np.linalg.norm(A) <= 10 * acc
    for key in l1:
        assert l1[key].rank == l2[key].rank
        assert np.linalg.norm(l1[key].to_dense() - l2[key].to_dense()) <= 10 * acc * max(1.0, np.linalg.norm(l2[key].to_dense()))
    assert np.all(np.linalg.eigvalsh(A) > 0)
