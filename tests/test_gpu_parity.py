"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI of libhcore_b200.so
(directly, or via the thin hcorepp_b200.api mirror) and is compared with the oracle (oracle/tlr_oracle.py), the
reference's own known-answer vectors, and fixtures produced by running the reference (tests/golden/).

Tolerances (BASELINE.json north_star): ||C_gpu - C_ref||_F / ||C_ref||_F <= 10 * accuracy on reconstructed products,
output ranks within +/-1 of the reference; element-wise index/copy kernels are compared exactly.
/root/reference is NOT used here (it does not exist on the GPU box).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tlr_oracle as O  # noqa: E402  (checker only)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))
K = VEC["kernels"]
DT = [np.float64, np.float32]
TT = {np.float64: torch.float64, np.float32: torch.float32}
F = lambda x, dt=np.float64: np.asfortranarray(np.array(x, dtype=dt))


@pytest.fixture(scope="module")
def hc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import hcorepp_b200 as h
    return h


@pytest.fixture(scope="module")
def ctx(hc):
    return hc.RunContext(0)


def dev(a):
    """numpy (any shape, interpreted column-major via Fortran flattening) -> flat device tensor."""
    a = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).cuda()


def host(t, shape):
    return t.cpu().numpy().reshape(shape, order="F")


def fn(name, dt):
    from hcorepp_b200 import _capi
    return getattr(_capi.lib, f"hcb_{'d' if dt == np.float64 else 's'}{name}")


def ok(rc):
    from hcorepp_b200 import _capi
    _capi.check(rc)


def ct(dt):
    return C.c_double if dt == np.float64 else C.c_float


def relerr(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300)


def approx(a, b, tol=1e-2):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(a), np.abs(b)) + 1e-5)


# ------------------------------------------------------------------------------------------------ kernel table
@pytest.mark.parametrize("dt", DT)
def test_kernel_table_known_answers(ctx, dt):
    c = ct(dt)
    g = K["gemm"]
    A, B, C0 = dev(F(g["A"], dt)), dev(F(g["B"], dt)), dev(F(g["C0"], dt))
    ok(fn("gemm", dt)(ctx.h, 0, 0, 3, 5, 4, c(1), A.data_ptr(), 3, B.data_ptr(), 4, c(1), C0.data_ptr(), 3))
    assert np.array_equal(host(C0, (3, 5)), F(g["C"], dt))

    g = K["multiply_by_alpha"]
    a = dev(np.array(g["flat_in"], dtype=dt))
    ok(fn("multiply_by_alpha", dt)(ctx.h, a.data_ptr(), g["rows"], g["cols"], g["m"], g["rank"], c(g["alpha"])))
    assert np.array_equal(a.cpu().numpy(), np.array(g["flat_out"], dtype=dt))

    g = K["process_v"]
    cv, b, v = dev(np.array(g["cv_flat"], dtype=dt)), dev(np.array(g["b_flat"], dtype=dt)), dev(np.zeros(10, dtype=dt))
    ok(fn("process_v", dt)(ctx.h, g["n"], g["crank"], 0, g["vm"], c(g["beta"]), cv.data_ptr(), g["ldcv"], v.data_ptr(),
                           g["arank"], b.data_ptr(), 0))
    assert np.array_equal(v.cpu().numpy(), np.array(g["v_flat"], dtype=dt))

    for key in ("new_rank_abs", "new_rank_rel"):
        g = K[key]
        s = dev(np.array(g["sigma"], dtype=dt))
        r = C.c_int64(0)
        ok(fn("new_rank", dt)(ctx.h, int(g["truncated"]), s.data_ptr(), len(g["sigma"]), c(g["accuracy"]), C.byref(r)))
        assert r.value == g["rank"]
        dr = torch.zeros(1, dtype=torch.int32, device="cuda")
        ok(fn("new_rank_device", dt)(ctx.h, int(g["truncated"]), s.data_ptr(), len(g["sigma"]), c(g["accuracy"]), dr.data_ptr()))
        assert int(dr.item()) == g["rank"]

    g = K["uvptr"]
    vn, uv = dev(np.array(g["vnew_flat"], dtype=dt)), dev(np.zeros(6, dtype=dt))
    ok(fn("uvptr", dt)(ctx.h, g["rank"], g["vm"], uv.data_ptr(), vn.data_ptr()))
    assert np.array_equal(uv.cpu().numpy(), np.array(g["uv_flat"], dtype=dt))

    for key in ("vtnew_noungqr", "vtnew_ungqr"):
        g = K[key]
        vt, s = dev(np.array(g["vt_flat"], dtype=dt)), dev(np.array(g["sigma"], dtype=dt))
        ok(fn("vtnew", dt)(ctx.h, g["rk"], int(g["ungqr"]), min(g["vm"], g["vn"]), s.data_ptr(), vt.data_ptr(), g["size_s"], g["vm"]))
        assert np.array_equal(vt.cpu().numpy(), np.array(g["out_flat"], dtype=dt))

    a = dev(np.zeros(9, dtype=dt))
    ok(fn("fill_identity", dt)(ctx.h, 3, a.data_ptr()))
    assert np.array_equal(host(a, (3, 3)), np.eye(3, dtype=dt))

    g = K["lacpy"]
    for kind in "GUL":
        a, b = dev(np.array(g["a_flat"], dtype=dt)), dev(np.zeros(16, dtype=dt))
        ok(fn("lacpy", dt)(ctx.h, ord(kind), 4, 4, a.data_ptr(), 4, b.data_ptr(), 4))
        assert np.array_equal(b.cpu().numpy(), np.array(g[kind], dtype=dt))
    g = K["laset"]
    for kind in "GUL":
        a = dev(np.zeros(16, dtype=dt))
        ok(fn("laset", dt)(ctx.h, ord(kind), 4, 4, c(g["offdiag"]), c(g["diag"]), a.data_ptr(), 4))
        assert np.array_equal(a.cpu().numpy(), np.array(g[kind], dtype=dt))

    for key in ("trmm", "trmm_unit"):
        g = K[key]
        a, b = dev(np.array(g["a_flat"], dtype=dt)), dev(np.array(g["b_flat"], dtype=dt))
        ok(fn("trmm", dt)(ctx.h, ord(g["side"]), ord(g["uplo"]), ord(g["trans"]), ord(g["diag"]), 4, 4, c(g["alpha"]),
                          a.data_ptr(), 4, b.data_ptr(), 4))
        assert approx(b.cpu().numpy(), g["out_flat"], g["tol_rel"])

    g = K["geqrf"]
    a, tau = dev(np.array(g["a_flat"], dtype=dt)), dev(np.zeros(2, dtype=dt))
    ok(fn("geqrf", dt)(ctx.h, g["m"], g["n"], a.data_ptr(), g["m"], tau.data_ptr()))
    assert approx(a.cpu().numpy(), g["qr_flat"], g["tol_rel"]) and approx(tau.cpu().numpy(), g["tau"], g["tol_rel"])

    A = F(K["svd"]["A"], dt)
    a, s, u, vt = dev(A), dev(np.zeros(3, dt)), dev(np.zeros(9, dt)), dev(np.zeros(9, dt))
    ok(fn("svd", dt)(ctx.h, 3, 3, a.data_ptr(), 3, s.data_ptr(), u.data_ptr(), 3, vt.data_ptr(), 3))
    U, S, VT = host(u, (3, 3)), s.cpu().numpy(), host(vt, (3, 3))
    assert approx((U * S) @ VT, A)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(40, 7), (300, 45), (1024, 71), (5, 9), (33, 33), (1, 1)])
def test_geqrf_ungqr_unmqr_vs_lapack(ctx, dt, shape):
    m, n = shape
    rng = np.random.default_rng(m * 1000 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))
    k = min(m, n)
    a, tau = dev(A), dev(np.zeros(k, dt))
    ok(fn("geqrf", dt)(ctx.h, m, n, a.data_ptr(), m, tau.data_ptr()))
    qr_g, tau_g = host(a, (m, n)), tau.cpu().numpy()
    qr_o, tau_o = O.k_geqrf(A)
    tol = 1e-11 if dt == np.float64 else 2e-4
    assert relerr(np.triu(qr_g), np.triu(qr_o)) < tol          # R, including LAPACK's sign convention
    assert relerr(qr_g, qr_o) < tol and relerr(tau_g, tau_o) < tol
    if m >= n:
        q = dev(qr_g)
        ok(fn("ungqr", dt)(ctx.h, m, n, n, q.data_ptr(), m, tau.data_ptr()))
        Q = host(q, (m, n))
        assert relerr(Q, O.k_ungqr(m, n, n, qr_o, tau_o)) < tol
        assert relerr(Q.T @ Q, np.eye(n)) < tol
        assert relerr(Q @ np.triu(qr_g[:n, :]), A) < tol
    # unmqr, all four (side, trans) combinations against LAPACK
    ncols = 6
    for side, trans in (("L", "N"), ("L", "T"), ("R", "N"), ("R", "T")):
        Cm = np.asfortranarray(rng.standard_normal((m, ncols) if side == "L" else (ncols, m)).astype(dt))
        cdev = dev(Cm)
        ok(fn("unmqr", dt)(ctx.h, 0 if side == "L" else 1, 0 if trans == "N" else 1, Cm.shape[0], Cm.shape[1], k,
                           a.data_ptr(), m, tau.data_ptr(), cdev.data_ptr(), Cm.shape[0]))
        want = O.k_unmqr(side, trans, qr_o[:, :k], tau_o, Cm)
        assert relerr(host(cdev, Cm.shape), want) < tol, (side, trans)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(9, 9), (71, 71), (130, 130), (64, 20), (20, 64), (7, 1), (1, 5), (200, 200),
                                   (357, 357), (700, 300), (300, 421), (600, 600), (768, 400), (769, 390)])
def test_svd_vs_lapack(ctx, dt, shape):
    m, n = shape
    rng = np.random.default_rng(m * 7 + n)
    k = min(m, n)
    # graded spectrum, like the recompression cores: singular values from 1 down to ~eps
    A = rng.standard_normal((m, k)) @ np.diag(np.logspace(0, -14 if dt == np.float64 else -6, k)) @ rng.standard_normal((k, n))
    A = np.asfortranarray(A.astype(dt))
    a, s, u, vt = dev(A), dev(np.zeros(k, dt)), dev(np.zeros(m * k, dt)), dev(np.zeros(k * n, dt))
    ok(fn("svd", dt)(ctx.h, m, n, a.data_ptr(), m, s.data_ptr(), u.data_ptr(), m, vt.data_ptr(), k))
    U, S, VT = host(u, (m, k)), s.cpu().numpy(), host(vt, (k, n))
    _, s_ref, _ = O.k_svd(A)
    eps = np.finfo(dt).eps
    assert np.all(np.diff(S) <= 0)                                            # sorted
    assert np.max(np.abs(S - s_ref)) <= 4 * max(m, n) * eps * s_ref[0]        # absolute agreement with gesdd
    assert relerr((U * S) @ VT, A) < 200 * eps                                # reconstructs A
    big = S > 1e3 * eps * S[0]
    # The Jacobi factor (left vectors of the taller of A / A^T) is orthonormal to rounding; the other factor is
    # recovered as X = A^T U / sigma (what the recompression consumes is X*sigma, accurate to eps*sigma_0
    # absolutely), so its vector i is orthonormal to ~ eps*sigma_0/sigma_i: weigh that check by sigma.
    direct, derived = (U[:, big], VT[big].T) if m >= n else (VT[big].T, U[:, big])
    assert np.max(np.abs(direct.T @ direct - np.eye(big.sum()))) < 1e3 * eps
    W = derived * (S[big] / S[0])[None, :]
    assert np.max(np.abs(W.T @ W - np.diag((S[big] / S[0]) ** 2))) < 1e3 * eps
    lead = S > 1e-3 * S[0]
    dl = VT[lead].T if m >= n else U[:, lead]
    assert np.max(np.abs(dl.T @ dl - np.eye(lead.sum()))) < 1e6 * eps


# ------------------------------------------------------------------------------------------------ tile level
def mk_tile(hc, ctx, kind, D, U, V, dt, max_rank=None):
    if kind == "D":
        return hc.DenseTile(np.asarray(D, dtype=dt), ctx)
    return hc.CompressedTile.from_uv(np.asarray(U, dtype=dt), np.asarray(V, dtype=dt), ctx, max_rank=max_rank)


@pytest.mark.parametrize("dt", DT)
def test_compressed_tile_gemm_known_answer(hc, ctx, dt):
    g = VEC["compressed_tile_gemm"]  # tests/operators/TestCompressedTile.cpp:124-260 (C rank 2 zeros += A*B, DDC-free path)
    A, B = F(g["A"], dt), F(g["B"], dt)
    # CompressedTile::Gemm(alpha, A-as-U_AB, B-as-V_AB): exactly a DCC-shaped call with BU = I: use A dense, B = (I, B)
    Ct = hc.CompressedTile.from_uv(np.zeros((3, 2), dt), np.zeros((2, 2), dt), ctx)
    Bt = hc.CompressedTile.from_uv(np.eye(3, dtype=dt), B, ctx)
    hc.HCore.Gemm(1.0, hc.DenseTile(A, ctx), False, Bt, False, 1.0, Ct, ctx, hc.CompressionParameters(float(np.finfo(dt).eps)))
    assert approx(Ct.to_dense(), g["C"], g["tol_rel"])
    assert Ct.GetTileRank() == 2


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("case", VEC["hcore_gemm"]["cases"], ids=lambda c: c["name"])
def test_hcore_gemm_known_answers(hc, ctx, case, dt):
    def tile(name):
        if name in case:
            return hc.DenseTile(F(case[name], dt), ctx)
        return hc.CompressedTile.from_uv(F(case[name + "U"], dt), F(case[name + "V"], dt), ctx)
    A, B = tile("A"), tile("B")
    m, n = A.m, B.n
    if case["C0"] == "zeros":
        Ct = hc.DenseTile(np.zeros((m, n), dt), ctx)
    else:
        r = case["c_rank"]
        Ct = hc.CompressedTile.from_uv(np.zeros((m, r), dt), np.zeros((r, n), dt), ctx)
    p = hc.CompressionParameters(float(np.finfo(dt).eps)) if case.get("acc") == "eps" else hc.CompressionParameters()
    hc.HCore.Gemm(case["alpha"], A, False, B, False, case["beta"], Ct, ctx, p)
    assert approx(Ct.to_dense(), case["C"], VEC["hcore_gemm"]["tol_rel"])


MIXES = ["DDD", "DDC", "DCD", "DCC", "CDD", "CDC", "CCD", "CCC"]


@pytest.mark.parametrize("dt,name", [(np.float64, "f64"), (np.float32, "f32")])
@pytest.mark.parametrize("mix", MIXES)
def test_mixes_vs_reference_fixture(hc, ctx, mix, dt, name):
    """tests/api/AdvancedGemmTest.cpp inputs; outputs of the reference itself (tests/golden/ref_mixes_*.npz)."""
    z = np.load(os.path.join(GOLDEN, f"ref_mixes_{name}.npz"))
    acc = float(z["acc"])
    a = mk_tile(hc, ctx, mix[0], z["A"], z["AU"], z["AV"], dt)
    b = mk_tile(hc, ctx, mix[1], z["B"], z["BU"], z["BV"], dt)
    c = mk_tile(hc, ctx, mix[2], z["C"], z["CU"], z["CV"], dt)
    hc.HCore.Gemm(float(z["alpha"]), a, False, b, False, float(z["beta"]), c, ctx, hc.CompressionParameters(acc))
    out, ref_out = c.to_dense(), z[f"{mix}_out"]
    assert relerr(out, ref_out) <= 10 * acc
    if mix[2] == "C":
        assert abs(c.GetTileRank() - int(z[f"{mix}_rank"])) <= 1
    if mix in ("DDD", "DCD", "CDD", "CCD"):  # pure contractions: rounding-level agreement
        assert relerr(out, ref_out) <= (1e-13 if dt == np.float64 else 1e-5)
    # the reference's own acceptance bound (AdvancedGemmTest.cpp:245-256)
    A, B, C0 = z["A"].astype(np.float64), z["B"].astype(np.float64), z["C"].astype(np.float64)
    ninf = lambda M: np.abs(M).sum(axis=1).max()
    cref = float(z["alpha"]) * A @ B + float(z["beta"]) * C0
    if mix != "DDC":
        bound = np.sqrt(A.shape[1] + 2) * abs(float(z["alpha"])) * ninf(A) * ninf(B) + 2 * abs(float(z["beta"])) * ninf(C0)
        err_ref = ninf(ref_out - cref) / bound
        assert ninf(out - cref) / bound < max(3 * acc, 1.5 * err_ref)


def test_multitile_vs_reference_fixture(hc, ctx):
    z = np.load(os.path.join(GOLDEN, "ref_multitile_f64.npz"))
    T, nb, acc = int(z["T"]), int(z["nb"]), float(z["acc"])
    dt = np.float64

    def grid(name):
        tm = hc.TileMatrix(T, T, nb, nb, torch.float64, ctx, compressed=True)
        for j in range(T):
            for i in range(T):
                U, V = z[f"{name}_U_{j}_{i}"], z[f"{name}_V_{j}_{i}"]
                t = tm.GetTile(j, i)
                rk = U.shape[1]
                t.buf[: nb * rk] = torch.from_numpy(np.ascontiguousarray(U.T).reshape(-1)).cuda()
                t.buf[nb * tm.max_rank: nb * tm.max_rank + rk * nb] = torch.from_numpy(np.ascontiguousarray(V.T).reshape(-1)).cuda()
                t.rank.fill_(rk)
        return tm
    A, B, Cm = grid("A"), grid("B"), grid("C0")
    info = torch.zeros(T * T, dtype=torch.int32, device="cuda")
    hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, hc.CompressionParameters(acc), info=info)
    ctx.Sync()
    assert int((info & 0xff).max().item()) in (0, 2)  # 2 = clipped at maxRank, which the reference does silently too
    assert np.max(np.abs(Cm.rank_table() - z["C_ranks"])) <= 1
    assert relerr(Cm.ToRawMatrix(), z["C_dense"]) <= 10 * acc


@pytest.mark.parametrize("dt,name", [(np.float64, "f64"), (np.float32, "f32")])
def test_compress_vs_reference_fixture(hc, ctx, dt, name):
    z = np.load(os.path.join(GOLDEN, f"ref_compress_{name}.npz"))
    acc = float(z["acc"])
    t = hc.CompressedTile.compress(z["A"], hc.CompressionParameters(acc), ctx)
    assert t.max_rank == int(z["max_rank"])
    assert abs(t.GetTileRank() - int(z["rank"])) <= 1
    assert relerr(t.to_dense(), z["U"] @ z["V"]) <= 10 * acc
    U = t.GetUMatrix().astype(np.float64)
    assert np.max(np.abs(U.T @ U - np.eye(U.shape[1]))) < (1e-10 if dt == np.float64 else 1e-3)  # U orthonormal, S in V


@pytest.mark.parametrize("shape", [(512, 512), (1024, 700), (600, 1024), (1024, 1024)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_sketched_compression_vs_oracle(hc, ctx, shape):
    """Large fp64 tiles take the sketched constructor (range finder + small SVD): same rank and factors as the
    reference's full SVD (Compressed.cpp:75-146) on the generator's spectrum law, U orthonormal."""
    m, n = shape
    acc = 1e-8
    rng = np.random.default_rng(m * 7 + n)
    k = min(m, n)
    qu, _ = np.linalg.qr(rng.standard_normal((m, k)))
    qv, _ = np.linalg.qr(rng.standard_normal((n, k)))
    A = (qu * O.latms_spectrum(k)) @ qv.T
    t = hc.CompressedTile.compress(A, hc.CompressionParameters(acc), ctx)
    o = O.CompressedTile.compress(A, O.CompressionParameters(acc))
    assert t.max_rank == o.max_rank
    assert abs(t.GetTileRank() - o.rank) <= 1, (t.GetTileRank(), o.rank)
    assert relerr(t.to_dense(), o.to_dense()) <= 10 * acc
    U = t.GetUMatrix()
    assert np.max(np.abs(U.T @ U - np.eye(U.shape[1]))) < 1e-10


def test_sketched_compression_falls_back_on_flat_spectrum(hc, ctx):
    """A spectrum that has not decayed within the sketch width is detected on the device and the tile is redone with
    the full SVD; a batch mixing both kinds gives each tile what it gives alone (the reference clamps at maxRank)."""
    acc, nb = 1e-8, 512
    rng = np.random.default_rng(99)

    def tile(sig):
        qu, _ = np.linalg.qr(rng.standard_normal((nb, nb)))
        qv, _ = np.linalg.qr(rng.standard_normal((nb, nb)))
        return (qu * sig) @ qv.T
    flat = tile(0.95 ** np.arange(nb))            # rank at 1e-8 is ~360 > maxRank = 170: clamped
    steep = tile(O.latms_spectrum(nb))            # rank 44
    raw = np.concatenate([flat, steep, flat * 0.5], axis=1)
    tm = hc.TileMatrix.from_dense(raw, nb, nb, ctx, hc.CompressionParameters(acc))
    for i, A in enumerate([flat, steep, flat * 0.5]):
        o = O.CompressedTile.compress(A, O.CompressionParameters(acc))
        g = tm.GetTile(0, i)
        assert abs(g.GetTileRank() - o.rank) <= 1, (i, g.GetTileRank(), o.rank)
        # clamped tiles: both keep the leading maxRank triplets; compare the reconstructions
        assert relerr(g.to_dense(), o.to_dense()) <= (1e-6 if i != 1 else 10 * acc)
    # the constructor's d_info is written -- bit 2 where the rank was clipped at maxRank (silent in the
    # reference, Compressed.cpp:117-119), clean for the tile that fits, Jacobi converged everywhere
    ci = tm.compress_info.cpu().numpy()
    assert [int(v) & 0xff for v in ci] == [2, 0, 2], ci


# ------------------------------------------------------------------------------------------------ live vs the oracle
def oracle_tile(kind, D, UV, dt):
    return O.DenseTile(np.asfortranarray(D.astype(dt))) if kind == "D" else O.CompressedTile.from_uv(UV[0].astype(dt), UV[1].astype(dt))


def lowrank(rng, m, n, k, dt, decay=0.5):
    U, _ = np.linalg.qr(rng.standard_normal((m, k)))
    V, _ = np.linalg.qr(rng.standard_normal((n, k)))
    s = decay ** np.arange(k)
    return U.astype(dt), (s[:, None] * V.T).astype(dt)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("dims", [(96, 80, 112), (3, 2, 3), (17, 33, 5), (256, 256, 256), (1, 1, 1), (130, 64, 31)])
def test_all_mixes_ragged_vs_oracle(hc, ctx, dims, dt):
    m, n, k = dims
    rng = np.random.default_rng(m + 10 * n + 100 * k)
    acc = 1e-6 if dt == np.float64 else 1e-3
    ka, kb, kc = max(1, min(m, k) // 3), max(1, min(k, n) // 3), max(1, min(m, n) // 4)
    AUV, BUV, CUV = lowrank(rng, m, k, ka, dt), lowrank(rng, k, n, kb, dt), lowrank(rng, m, n, kc, dt)
    Ad, Bd, Cd = AUV[0] @ AUV[1], BUV[0] @ BUV[1], CUV[0] @ CUV[1]
    cap = max(1, min(m, n))  # generous capacity so that truncation, not clipping, decides the rank
    for mix in MIXES:
        a = mk_tile(hc, ctx, mix[0], Ad, *AUV, dt)
        b = mk_tile(hc, ctx, mix[1], Bd, *BUV, dt)
        c = mk_tile(hc, ctx, mix[2], Cd, *CUV, dt, max_rank=cap)
        oa, ob = oracle_tile(mix[0], Ad, AUV, dt), oracle_tile(mix[1], Bd, BUV, dt)
        oc = oracle_tile(mix[2], Cd, CUV, dt)
        if mix[2] == "C":
            oc.max_rank = cap
        hc.HCore.Gemm(1.5, a, False, b, False, -0.5, c, ctx, hc.CompressionParameters(acc))
        O.hcore_gemm(dt(1.5), oa, False, ob, False, dt(-0.5), oc, O.CompressionParameters(acc))
        ref = oc.to_dense()
        if mix == "DDC":
            # full-rank result held exactly in one factor; for m < n the reference's own index arithmetic is wrong
            # (HCore.cpp:296 "not handled correctly"), so the check is against the dense truth for every shape
            ref = 1.5 * Ad.astype(np.float64) @ Bd.astype(np.float64) - 0.5 * Cd.astype(np.float64)
        scale = max(np.linalg.norm(ref), 1e-30)
        assert np.linalg.norm(c.to_dense().astype(np.float64) - ref) <= 10 * acc * max(scale, 1.0), mix
        if mix[2] == "C" and mix != "DDC":
            assert abs(c.GetTileRank() - oc.rank) <= 1, (mix, c.GetTileRank(), oc.rank)


@pytest.mark.parametrize("dt", DT)
def test_transposed_operands_dense_output(hc, ctx, dt):
    """op = Trans on dense and compressed operands with a dense C (Dense.cpp:46-108 handles ops; HCore.cpp:73-81,101-109)."""
    rng = np.random.default_rng(5)
    m, n, k = 70, 50, 90
    AUV, BUV = lowrank(rng, k, m, 9, dt), lowrank(rng, n, k, 7, dt)  # stored transposed: A is k x m, B is n x k
    Ad, Bd = AUV[0] @ AUV[1], BUV[0] @ BUV[1]
    C0 = rng.standard_normal((m, n)).astype(dt)
    want = 2.0 * Ad.T.astype(np.float64) @ Bd.T + 0.5 * C0
    for mix in ("DDD", "CDD", "DCD", "CCD"):
        a = mk_tile(hc, ctx, mix[0], Ad, *AUV, dt)
        b = mk_tile(hc, ctx, mix[1], Bd, *BUV, dt)
        c = hc.DenseTile(C0, ctx)
        hc.HCore.Gemm(2.0, a, True, b, True, 0.5, c, ctx)
        assert relerr(c.to_dense(), want) < (1e-12 if dt == np.float64 else 1e-4), mix


def test_rank_truncation_straddling_accuracy(hc, ctx):
    """Singular values placed just above / below the absolute threshold: ranks must agree with the oracle +/-1."""
    rng = np.random.default_rng(11)
    nb, acc = 128, 1e-6
    U, _ = np.linalg.qr(rng.standard_normal((nb, 12)))
    V, _ = np.linalg.qr(rng.standard_normal((nb, 12)))
    s = np.array([1, .5, .1, 1e-2, 1e-3, 1e-4, 1e-5, 2e-6, 1.2e-6, 0.9e-6, 5e-7, 1e-8])
    A = (U * s) @ V.T
    t = hc.CompressedTile.compress(A, hc.CompressionParameters(acc), ctx)
    o = O.CompressedTile.compress(A, O.CompressionParameters(acc))
    assert abs(t.GetTileRank() - o.rank) <= 1 and o.rank == 9
    # fixed-rank mode (Compressed.cpp:103-109, 510-515)
    t = hc.CompressedTile.compress(A, hc.CompressionParameters(acc, fixed_rank=5), ctx)
    assert t.GetTileRank() == 5
    # relative threshold (truncatedSvd, omp/kernels.cpp:87-95)
    t = hc.CompressedTile.compress(A * 100, hc.CompressionParameters(acc, truncated_svd=True), ctx)
    o = O.CompressedTile.compress(A * 100, O.CompressionParameters(acc, truncated_svd=True))
    assert abs(t.GetTileRank() - o.rank) <= 1


def test_batched_equals_one_by_one_and_info(hc, ctx):
    """One batched call over many triples == the same triples one at a time; rank-bound violations are reported."""
    rng = np.random.default_rng(3)
    nb, n = 64, 9
    mk = lambda k: lowrank(rng, nb, nb, k, np.float64)
    As, Bs, C1, C2 = [], [], [], []
    for t in range(n):
        As.append(hc.CompressedTile.from_uv(*mk(5 + t % 3), ctx))
        Bs.append(hc.CompressedTile.from_uv(*mk(4 + t % 2), ctx))
        cu, cv = mk(3)
        C1.append(hc.CompressedTile.from_uv(cu, cv, ctx, max_rank=21))
        C2.append(hc.CompressedTile.from_uv(cu, cv, ctx, max_rank=21))
    p = hc.CompressionParameters(1e-8)
    info = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    hc.gemm_batched(1.0, As, False, Bs, False, 1.0, C1, ctx, p, info=info)
    for t in range(n):
        hc.HCore.Gemm(1.0, As[t], False, Bs[t], False, 1.0, C2[t], ctx, p)
    ctx.Sync()
    assert int((info & 0xff).max().item()) == 0
    assert 1 <= int((info >> 8).max().item()) <= 30  # Jacobi sweeps used (diagnostics)
    for t in range(n):
        assert C1[t].GetTileRank() == C2[t].GetTileRank()
        assert relerr(C1[t].to_dense(), C2[t].to_dense()) < 1e-12
    # a rank bound that is too small must be detected on the device and leave the tile untouched
    before = C1[0].to_dense()
    As[0].rank_bound = 2
    info1 = torch.zeros(1, dtype=torch.int32, device="cuda")
    hc.gemm_batched(1.0, As[:1], False, Bs[:1], False, 1.0, C1[:1], ctx, p, info=info1)
    from hcorepp_b200._capi import HcbError
    with pytest.raises(HcbError, match="rank_bound"):  # sticky context error: heard even without an info buffer
        ctx.Sync()
    ctx.Sync()                                          # ... once
    assert int(info1.item()) & 0xff == 4 and np.array_equal(C1[0].to_dense(), before)
    hc.gemm_batched(1.0, As[:1], False, Bs[:1], False, 1.0, C1[:1], ctx, p)   # same violation, no info buffer
    with pytest.raises(HcbError, match="rank_bound"):
        ctx.Sync()
    assert np.array_equal(C1[0].to_dense(), before)


def test_mixed_mix_batch_equals_one_by_one(hc, ctx):
    """one hcb_?tlr_gemm_batched call may hold triples of DIFFERENT Dense/Compressed mixes -- they are
    partitioned by mix inside the library; results, ranks and the order of d_info must be those of one-by-one calls."""
    rng = np.random.default_rng(31)
    nb, dt = 96, np.float64
    order = ["CCC", "DDD", "CDC", "CCD", "DCC", "CCC", "CDD", "DCD", "CCC", "DCC"]
    p = hc.CompressionParameters(1e-8)

    def build():
        As, Bs, Cs = [], [], []
        r2 = np.random.default_rng(77)
        for mix in order:
            def op(kind, k):
                if kind == "C":
                    return hc.CompressedTile.from_uv(*lowrank(r2, nb, nb, k, dt), ctx, max_rank=nb // 3)
                return hc.DenseTile(np.asfortranarray(r2.standard_normal((nb, nb)) / nb), ctx)
            As.append(op(mix[0], 7)); Bs.append(op(mix[1], 6)); Cs.append(op(mix[2], 5))
        return As, Bs, Cs
    A1, B1, C1 = build()
    A2, B2, C2 = build()
    info = torch.full((len(order),), -1, dtype=torch.int32, device="cuda")
    hc.gemm_batched(1.0, A1, False, B1, False, 1.0, C1, ctx, p, info=info)
    for a, b, c in zip(A2, B2, C2):
        hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, p)
    ctx.Sync()
    h = info.cpu().numpy()
    assert np.all((h & 0xff) == 0), h
    for t, mix in enumerate(order):
        assert relerr(C1[t].to_dense(), C2[t].to_dense()) < 1e-12, (t, mix)
        if mix[2] == "C":
            assert C1[t].GetTileRank() == C2[t].GetTileRank(), (t, mix)
            assert (h[t] >> 8) >= 1, (t, mix, h)        # a recompressing triple reports its Jacobi sweeps at ITS index
        else:
            assert (h[t] >> 8) == 0, (t, mix, h)        # dense-output triples run no Jacobi
    del rng


def test_incremental_u_side_ksum_vs_oracle(hc, ctx):
    """Round 2: a C tile that carries the state bit "U orthonormal" takes the incremental U path (block Gram-Schmidt of
    the new columns against CU + a kp-column panel QR + GEMM rebuild).  An 8-step k-sum against the oracle (<= 10*acc,
    ranks +/-1), against the full-QR path (HCB_NO_INCREMENTAL=1), U stays orthonormal, info flags are sticky over k."""
    nb, k, acc, kt = 512, 30, 1e-8, 8
    p, po = hc.CompressionParameters(acc), O.CompressionParameters(acc)
    tiles = [(O.synth_compressed_tile(nb, k, 7000 + i), O.synth_compressed_tile(nb, k, 8000 + i)) for i in range(kt)]
    oC = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
    oranks = []
    for a, b in tiles:
        O.hcore_gemm(1.0, a, False, b, False, 1.0, oC, po)
        oranks.append(oC.rank)
    ref = oC.to_dense()
    res = {}
    for mode in ("incremental", "full"):
        if mode == "full":
            os.environ["HCB_NO_INCREMENTAL"] = "1"
        try:
            Ct = hc.CompressedTile(nb, nb, nb // 3, torch.float64, ctx)
            ranks, states = [], []
            for a, b in tiles:
                A = hc.CompressedTile.from_uv(a.U, a.V, ctx)
                B = hc.CompressedTile.from_uv(b.U, b.V, ctx)
                hc.HCore.Gemm(1.0, A, False, B, False, 1.0, Ct, ctx, p)
                ranks.append(Ct.GetTileRank())
                states.append(int(Ct.state.item()))
            res[mode] = (Ct.to_dense(), ranks, states, Ct.factors()[0])
        finally:
            os.environ.pop("HCB_NO_INCREMENTAL", None)
    for mode, (d, ranks, states, U) in res.items():
        assert relerr(d, ref) <= 10 * acc, mode
        assert max(abs(a - b) for a, b in zip(ranks, oranks)) <= 1, (mode, ranks, oranks)
        assert all(s & 3 == 3 for s in states), (mode, states)   # U orthonormal, rows of V orthogonal
        assert [s >> 8 for s in states] == ([0] + list(range(1, kt)) if mode == "incremental" else [0] * kt), (mode, states)
        # accuracy-aware Jacobi stop: the left vectors are orthogonal to ~0.05 * accuracy per update (adds up over the
        # incremental updates until the periodic full re-factorisation)
        orth = torch.linalg.norm(U.t() @ U - torch.eye(U.shape[1], dtype=torch.float64, device="cuda")).item()
        assert orth < 0.5 * acc, (mode, orth)
    assert relerr(res["incremental"][0], res["full"][0]) <= 1e-9
    # a tile written behind the library's back: invalidate() clears the bit and the next update takes the full path
    Ct = hc.CompressedTile.from_uv(tiles[0][0].U * 3.0, tiles[0][0].V, ctx, max_rank=nb // 3)  # U NOT orthonormal
    assert int(Ct.state.item()) == 0
    A, B = hc.CompressedTile.from_uv(tiles[1][0].U, tiles[1][0].V, ctx), hc.CompressedTile.from_uv(tiles[1][1].U, tiles[1][1].V, ctx)
    o = O.CompressedTile.from_uv(tiles[0][0].U * 3.0, tiles[0][0].V)
    o.max_rank = nb // 3
    hc.HCore.Gemm(1.0, A, False, B, False, 1.0, Ct, ctx, p)
    O.hcore_gemm(1.0, tiles[1][0], False, tiles[1][1], False, 1.0, o, po)
    assert relerr(Ct.to_dense(), o.to_dense()) <= 10 * acc and abs(Ct.GetTileRank() - o.rank) <= 1


def test_selective_reorthogonalisation_saturated_tile(hc, ctx):
    """Round 2: the second block Gram-Schmidt pass of the incremental recompression is switched off per tile side when
    every new column kept more than half of its squared norm (k_inc_gate).  Here the updates are built so that it must
    STAY ON: A_k = C_U M_k (+ a 1e-8 component outside), B_k = N_k C_V -- every new direction lies almost entirely in the
    spans the tile already has (a saturated tile of a smooth kernel).  At accuracy 1e-12 the outside components survive the
    truncation, and the once-projected remainder is ~1e-8 of its column: ONE Gram-Schmidt pass would leave it orthogonal
    to C_U only to eps / 1e-8 ~ 1e-8, the second pass brings that to rounding level.  Checked: parity with the oracle, U
    orthonormal to 1e-12 (with the second pass forced off -- HCB_GS_FORCE_ONCE, measured 1.4e-11 against 4e-14 -- the same
    k-sum loses more than two orders of magnitude, which is asserted too), and agreement with the full Householder path."""
    nb, kc, ka, acc, kt = 512, 60, 24, 1e-12, 4
    rng = np.random.default_rng(4242)
    p, po = hc.CompressionParameters(acc), O.CompressionParameters(acc)
    c0 = O.synth_compressed_tile(nb, kc, 9100)                  # U orthonormal, V = S W^T
    W = (c0.V / np.linalg.norm(c0.V, axis=1)[:, None]).T        # nb x kc orthonormal
    steps = []
    for i in range(kt):
        qa, _ = np.linalg.qr(rng.standard_normal((nb, ka)))
        qb, _ = np.linalg.qr(rng.standard_normal((nb, ka)))
        M, N = rng.standard_normal((kc, ka)) / np.sqrt(kc), rng.standard_normal((kc, ka)) / np.sqrt(kc)
        sa = 0.5 ** np.arange(ka)
        AU = np.asfortranarray(c0.U @ M + 1e-8 * qa)            # columns inside span(CU) up to 1e-8
        AV = np.asfortranarray((sa[:, None] * qb.T))            # ka x nb
        BU = np.asfortranarray(qb)                              # B = qb (N^T W^T + 1e-8 qa^T): rows inside span(W)
        BV = np.asfortranarray(N.T @ W.T + 1e-8 * qa.T)
        steps.append((O.CompressedTile.from_uv(AU, AV), O.CompressedTile.from_uv(BU, BV)))
    oC = O.CompressedTile.from_uv(c0.U.copy(), c0.V.copy())
    oC.max_rank = nb // 3
    res = {}
    for mode in ("incremental", "full", "forced_once"):
        if mode == "full":
            os.environ["HCB_NO_INCREMENTAL"] = "1"
        if mode == "forced_once":
            os.environ["HCB_GS_FORCE_ONCE"] = "1"
        try:
            Ct = hc.CompressedTile.from_uv(c0.U, c0.V, ctx, max_rank=nb // 3)
            Ct.state.fill_(3)                                   # the tile IS in SVD form (built that way above)
            for a, b in steps:
                A, B = hc.CompressedTile.from_uv(a.U, a.V, ctx), hc.CompressedTile.from_uv(b.U, b.V, ctx)
                hc.HCore.Gemm(1.0, A, False, B, False, 1.0, Ct, ctx, p)
            U = Ct.factors()[0]
            orth = torch.linalg.norm(U.t() @ U - torch.eye(U.shape[1], dtype=torch.float64, device="cuda")).item()
            res[mode] = (Ct.to_dense(), Ct.GetTileRank(), orth, int(Ct.state.item()))
        finally:
            os.environ.pop("HCB_NO_INCREMENTAL", None)
            os.environ.pop("HCB_GS_FORCE_ONCE", None)
    for a, b in steps:
        O.hcore_gemm(1.0, a, False, b, False, 1.0, oC, po)
    ref = oC.to_dense()
    # the case really exercises the second pass: with it forced off the left factor loses orthogonality by orders of magnitude
    forced = res.pop("forced_once")
    assert forced[2] > 100 * res["incremental"][2] and forced[2] > 5e-12, (forced[2], res["incremental"][2])
    for mode, (d, rk, orth, st) in res.items():
        assert relerr(d, ref) <= 10 * acc, (mode, relerr(d, ref))
        assert abs(rk - oC.rank) <= 1, (mode, rk, oC.rank)
        assert orth < 1e-12, (mode, orth)
    assert res["incremental"][3] >> 8 == kt                      # every update took the incremental path
    assert relerr(res["incremental"][0], res["full"][0]) <= 1e-9


def test_cholqr_fast_path_fallback_and_deflation(hc):
    """Round 2: the kp new columns of an incremental update are factored by CholeskyQR2 (k_cholqr_pass) where every pivot of
    the scaled Gram matrix passes the safety test; otherwise that tile side falls back to the Householder panels.
    (a) generic updates take the fast path (counters), (b) an A tile whose U columns are nearly dependent (a user-built tile,
    condition ~1e9) must FALL BACK and still match the oracle, (c) a rank-deficient update is deflated or falls back, (d) the
    result equals the all-Householder path (HCB_NO_CHOLQR=1) to rounding."""
    nb, kc, ka, acc = 512, 50, 20, 1e-8
    rng = np.random.default_rng(515)
    c2 = hc.RunContext(0)
    p, po = hc.CompressionParameters(acc), O.CompressionParameters(acc)
    c0 = O.synth_compressed_tile(nb, kc, 9300)
    good = (O.synth_compressed_tile(nb, ka, 9301), O.synth_compressed_tile(nb, ka, 9302))
    # nearly dependent left columns: u_1 = u_0 + 1e-9 w
    qa, _ = np.linalg.qr(rng.standard_normal((nb, ka)))
    qa[:, 1] = qa[:, 0] + 1e-9 * qa[:, 1]
    bad_a = O.CompressedTile.from_uv(np.asfortranarray(qa), np.asfortranarray(good[0].V.copy()))
    # a repeated and a zero column in the right factor's row space (deflation on the V side)
    bv = good[1].V.copy()
    bv[3, :] = 0.0
    defl_b = O.CompressedTile.from_uv(np.asfortranarray(good[1].U.copy()), np.asfortranarray(bv))
    seq = [good, (bad_a, good[1]), (good[0], defl_b)]
    outs = {}
    for mode in ("cholqr", "householder", "no_vchol"):
        if mode == "householder":
            os.environ["HCB_NO_CHOLQR"] = "1"
            os.environ["HCB_NO_VCHOL"] = "1"
        if mode == "no_vchol":
            os.environ["HCB_NO_VCHOL"] = "1"
        try:
            Ct = hc.CompressedTile.from_uv(c0.U, c0.V, c2, max_rank=nb // 3)
            Ct.state.fill_(3)
            c2.stats(reset=True)
            st = []
            for a, b in seq:
                A, B = hc.CompressedTile.from_uv(a.U, a.V, c2), hc.CompressedTile.from_uv(b.U, b.V, c2)
                hc.HCore.Gemm(1.0, A, False, B, False, 1.0, Ct, c2, p)
                st.append(c2.stats(reset=True))
            U = Ct.factors()[0]
            orth = torch.linalg.norm(U.t() @ U - torch.eye(U.shape[1], dtype=torch.float64, device="cuda")).item()
            outs[mode] = (Ct.to_dense(), Ct.GetTileRank(), st, orth)
        finally:
            os.environ.pop("HCB_NO_CHOLQR", None)
            os.environ.pop("HCB_NO_VCHOL", None)
    oC = O.CompressedTile.from_uv(c0.U.copy(), c0.V.copy())
    oC.max_rank = nb // 3
    for a, b in seq:
        O.hcore_gemm(1.0, a, False, b, False, 1.0, oC, po)
    ref = oC.to_dense()
    for mode, (d, rk, st, orth) in outs.items():
        assert relerr(d, ref) <= 10 * acc, (mode, relerr(d, ref))
        assert abs(rk - oC.rank) <= 1, (mode, rk, oC.rank)
        assert orth < 0.5 * acc, (mode, orth)
    st = outs["cholqr"][2]
    assert st[0]["cholqr_panels"] == 2 and st[0]["cholqr_fallback_pass0"] == 0, st[0]          # generic: both sides fast
    assert st[1]["cholqr_fallback_pass0"] + st[1]["cholqr_fallback_pass1"] >= 1, st[1]          # dependent U columns: fallback
    # rank-deficient right factor (zero row of B.V): the degenerate new direction is either deflated or sends the side to the
    # Householder path -- never through an unsafe CholeskyQR
    assert st[2]["deflated_columns"] + st[2]["cholqr_fallback_pass0"] + st[2]["cholqr_fallback_pass1"] >= 1, st[2]
    assert all(v["cholqr_panels"] == 0 and v["vcore_cholesky"] == 0 for v in outs["householder"][2])
    assert relerr(outs["cholqr"][0], outs["householder"][0]) <= 1e-9
    # the graded r x r factor of the V side: by Cholesky of the assembled scaled Gram matrix (k_vcore_chol) in the default mode,
    # by the Householder R-only QR with HCB_NO_VCHOL=1 -- same result to rounding
    assert all(v["vcore_cholesky"] + v["vcore_fallback"] == 1 for v in st) and st[0]["vcore_cholesky"] == 1, st
    assert all(v["vcore_cholesky"] == 0 for v in outs["no_vchol"][2])
    assert relerr(outs["cholqr"][0], outs["no_vchol"][0]) <= 1e-9


def test_matmul_info_is_sticky_over_k(hc, ctx):
    """hcb_?tlr_matmul passes one d_info to every k-step -- flags are OR-ed, the sweep count is the maximum."""
    nb, T, k, acc = 128, 2, 12, 1e-8
    A = hc.TileMatrix(T, T, nb, nb, torch.float64, ctx, compressed=True)
    B = hc.TileMatrix(T, T, nb, nb, torch.float64, ctx, compressed=True)
    Cm = hc.TileMatrix.zeros_compressed(T, T, nb, nb, torch.float64, ctx)
    for tm, seed in ((A, 1), (B, 2)):
        for j in range(T):
            for i in range(T):
                t = O.synth_compressed_tile(nb, k, seed * 100 + j * 10 + i)
                g = tm.GetTile(j, i)
                g.buf[: nb * k] = dev(t.U)
                g.buf[nb * tm.max_rank: nb * tm.max_rank + k * nb] = dev(t.V)
                g.rank.fill_(k)
    # the bound is violated in the FIRST k-step only: the A(:, 0) tiles claim a bound below their rank
    A.descs[0].rank_bound = 3
    A.descs[1].rank_bound = 3
    info = torch.full((T * T,), -1, dtype=torch.int32, device="cuda")
    hc.tile_matrix_multiplication(A, B, Cm, 1.0, 1.0, ctx, hc.CompressionParameters(acc), info=info)
    from hcorepp_b200._capi import HcbError
    with pytest.raises(HcbError):
        ctx.Sync()
    h = info.cpu().numpy()
    assert np.all((h & 0xff) == 4)            # every C tile lost its k = 0 update: the flag survived the k = 1 step
    assert np.all(((h >> 8) & 0xff) >= 1)     # ... whose Jacobi sweeps are reported as the maximum over k


def test_full_size_tile_properties(hc, ctx):
    """BASELINE config sizes (nb = 1024, ranks 44, acc 1e-8): one C tile through a 3-step k-sweep, against the oracle
    (<= 10*acc, ranks +/-1) and through size-independent properties -- U orthonormal, the reference's own normalised
    error check against the dense product (omp_main.cpp:366-372), and idempotent recompression."""
    torch.manual_seed(0)
    nb, k, acc, kt = 1024, 44, 1e-8, 3
    s = torch.from_numpy(O.latms_spectrum(nb, np.float64)[:k].copy()).cuda()

    def synth():
        qu, _ = torch.linalg.qr(torch.randn(nb, k, dtype=torch.float64, device="cuda"))
        qv, _ = torch.linalg.qr(torch.randn(nb, k, dtype=torch.float64, device="cuda"))
        return qu, s[:, None] * qv.t()
    p, po = hc.CompressionParameters(acc), O.CompressionParameters(acc)
    Ct = hc.CompressedTile(nb, nb, nb // 3, torch.float64, ctx)
    oC = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
    dense = torch.zeros(nb, nb, dtype=torch.float64, device="cuda")
    ranks, oranks = [], []
    a_inf = b_inf = 0.0
    for _ in range(kt):
        (au, av), (bu, bv) = synth(), synth()
        aun, avn, bun, bvn = (x.cpu().numpy() for x in (au, av, bu, bv))
        A = hc.CompressedTile.from_uv(aun, avn, ctx, max_rank=nb // 3)
        B = hc.CompressedTile.from_uv(bun, bvn, ctx, max_rank=nb // 3)
        hc.HCore.Gemm(1.0, A, False, B, False, 1.0, Ct, ctx, p)
        O.hcore_gemm(1.0, O.CompressedTile.from_uv(aun, avn), False, O.CompressedTile.from_uv(bun, bvn), False, 1.0, oC, po)
        dense += (au @ av) @ (bu @ bv)
        a_inf += (au @ av).abs().sum(dim=1).max().item()
        b_inf += (bu @ bv).abs().sum(dim=1).max().item()
        ranks.append(Ct.GetTileRank())
        oranks.append(oC.rank)
    assert max(abs(a - b) for a, b in zip(ranks, oranks)) <= 1, (ranks, oranks)
    U, V = Ct.factors()
    ref = torch.from_numpy(oC.to_dense()).cuda()
    assert (torch.linalg.norm(U @ V - ref) / torch.linalg.norm(ref)).item() <= 10 * acc
    orth = torch.linalg.norm(U.t() @ U - torch.eye(U.shape[1], dtype=torch.float64, device="cuda")).item()
    assert orth < 1e-10
    err_inf = (U @ V - dense).abs().sum(dim=1).max().item()
    assert err_inf / ((a_inf / kt + b_inf / kt) * acc * nb) < 10  # the example driver's own pass criterion
    before = (U @ V).clone()
    Z = hc.CompressedTile(nb, nb, nb // 3, torch.float64, ctx)  # rank-1 zero tile
    hc.HCore.Gemm(1.0, Z, False, Z, False, 1.0, Ct, ctx, p)
    U2, V2 = Ct.factors()
    assert (torch.linalg.norm(U2 @ V2 - before) / torch.linalg.norm(before)).item() <= 10 * acc
    assert abs(Ct.GetTileRank() - ranks[-1]) <= 1



@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(1000, 900, 500, 50, 50, 60), (257, 300, 128, 40, 40, 50), (999, 513, 256, 45, 45, 40),
                                  (130, 700, 96, 30, 40, 45), (1024, 1024, 1024, 44, 44, 240),
                                  (2048, 2048, 2048, 64, 64, 100), (2048, 1500, 700, 40, 50, 90)],
                         ids=lambda d: "x".join(map(str, d)))
def test_blocked_recompression_ragged_vs_oracle(hc, ctx, dims):
    """Stacked rank > 64: the compact-WY path -- register panel QR clusters, strip-resident reflector clusters of 1 / 2 / 4
    CTAs, odd row counts (8-byte cp.async path), wide V stacks, the register-block Jacobi -- against the oracle."""
    m, n, k, ka, kb, kc = dims
    dt, acc = np.float64, 1e-8
    rng = np.random.default_rng(sum(dims))
    AUV, BUV, CUV = lowrank(rng, m, k, ka, dt, 0.9), lowrank(rng, k, n, kb, dt, 0.9), lowrank(rng, m, n, kc, dt, 0.97)
    cap = min(m, n)
    a = mk_tile(hc, ctx, "C", None, *AUV, dt)
    b = mk_tile(hc, ctx, "C", None, *BUV, dt)
    c = mk_tile(hc, ctx, "C", None, *CUV, dt, max_rank=cap)
    c.rank_bound = min(cap, kc + 2 * ka + 8)  # tight scratch bound (as the drivers set it): register-block Jacobi for r <= 384
    oa, ob, oc = (oracle_tile("C", None, x, dt) for x in (AUV, BUV, CUV))
    oc.max_rank = cap
    for _ in range(2):  # second pass: C now carries the rank of the first result (orthonormal U, graded V)
        hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, hc.CompressionParameters(acc))
        O.hcore_gemm(dt(1.0), oa, False, ob, False, dt(1.0), oc, O.CompressionParameters(acc))
        ref = oc.to_dense()
        assert np.linalg.norm(c.to_dense() - ref) <= 10 * acc * max(np.linalg.norm(ref), 1.0)
        assert abs(c.GetTileRank() - oc.rank) <= 1, (c.GetTileRank(), oc.rank)


@pytest.mark.gpu
def test_blocked_batch_of_different_shapes(hc, ctx):
    """One batched call over tiles of different shapes and ranks (descriptors are per tile, grids from the bounds) must give
    what the tiles give one by one."""
    dt, acc = np.float64, 1e-8
    shapes = [(1000, 900, 500, 50, 50, 60), (513, 257, 128, 40, 40, 50), (96, 80, 64, 10, 12, 9)]
    rng = np.random.default_rng(7)
    tiles, solo = [], []
    for (m, n, k, ka, kb, kc) in shapes:
        AUV, BUV, CUV = lowrank(rng, m, k, ka, dt, 0.9), lowrank(rng, k, n, kb, dt, 0.9), lowrank(rng, m, n, kc, dt, 0.97)
        mk = lambda: (mk_tile(hc, ctx, "C", None, *AUV, dt), mk_tile(hc, ctx, "C", None, *BUV, dt),
                      mk_tile(hc, ctx, "C", None, *CUV, dt, max_rank=min(m, n)))
        tiles.append(mk())
        solo.append(mk())
    prm = hc.CompressionParameters(acc)
    hc.gemm_batched(1.0, [t[0] for t in tiles], False, [t[1] for t in tiles], False, 1.0, [t[2] for t in tiles], ctx, prm)
    for (a, b, c), (_, _, cb) in zip(solo, tiles):
        hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, prm)
        ref = c.to_dense()
        assert np.linalg.norm(cb.to_dense() - ref) <= 1e-12 * max(np.linalg.norm(ref), 1.0)
        assert cb.GetTileRank() == c.GetTileRank()


def test_cpp_host_layer_replays_reference_tests():
    """include/hcorepp_b200/hcorepp.hpp (the C++ mirror of the reference API) over the C ABI: the reference's own
    operator/API known answers (tests/cpp/test_api.cpp), built by __graft_entry__.build() with plain g++."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "test_api")
    assert os.path.exists(exe), "tests/cpp/test_api not built (run __graft_entry__.build()): the C++ replay must not be skipped"
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "0 failure(s)" in r.stdout


# ------------------------------------------------------------------------------------------------ TLR Cholesky (8f row 1)
def test_example_driver_matrix_multiplication():
    """The reference's headline example (examples/matrix_multiplication/omp_main.cpp: same command line, flow and CSV
    lines) built on the C++ mirror: `b200-hcorepp-matrix 4 1e-4,1e-8 512` = BASELINE configs[0] shape.  The example's own
    pass criterion (normalised error < 10, omp_main.cpp:325-327,379-381) must hold for the dense and both compressed runs,
    compressed storage must be smaller than dense, and tighter accuracy must not give a larger error."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples", "matrix_multiplication", "b200-hcorepp-matrix")
    assert os.path.exists(exe), "examples/matrix_multiplication/b200-hcorepp-matrix not built (run __graft_entry__.build())"
    env = dict(os.environ, HCOREPP_VERBOSE="ON")
    r = subprocess.run([exe, "4", "1e-4,1e-8", "512"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l]
    assert lines[0].startswith("tile_count, tile_size, matrix_size, type, error, error_normalized, memory(KB)")
    rows = [[c.strip() for c in l.split(",")] for l in lines[1:]]
    kinds = [row[3] for row in rows]
    assert kinds[:2] == ["ref", "dense"] and len(rows) == 4, kinds
    dense, c4, c8 = rows[1], rows[2], rows[3]
    assert float(dense[5]) < 10 and float(c4[5]) < 10 and float(c8[5]) < 10, (dense, c4, c8)
    assert int(c4[6]) < int(c8[6]) < int(dense[6]), (c4[6], c8[6], dense[6])     # memory(KB): looser accuracy, smaller tiles
    assert float(c8[4]) <= float(c4[4])                                          # absolute error


def _spd(rng, n, dt=np.float64):
    a = rng.standard_normal((n, n))
    return np.asfortranarray((a @ a.T / n + np.eye(n)).astype(dt))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 31, 64, 300, 1024])
def test_potrf_vs_oracle(hc, ctx, n, dt):
    """hcb_?potrf == HCore<T>::Potrf (HCore.cpp:586-621 -> lapack::potrf): in place, the other triangle untouched, both uplo."""
    rng = np.random.default_rng(n)
    A = _spd(rng, n, dt)
    tol = 1e-11 if dt == np.float64 else 2e-4
    for uplo in ("L", "U"):
        t = hc.DenseTile(A.copy(), ctx)
        assert hc.HCore.Potrf(t, uplo, ctx) == 0
        o = O.DenseTile(np.asfortranarray(A.copy()))
        O.hcore_potrf(o, uplo)
        assert relerr(t.to_numpy(), o.data) < tol * max(1, n / 64), (uplo, n)
    with pytest.raises(RuntimeError, match="dense"):
        hc.HCore.Potrf(hc.CompressedTile.from_uv(np.ones((4, 1)), np.ones((1, 4)), ctx), "L", ctx)
    if n >= 31:  # LAPACK's info: 1-based index of the first non-positive pivot
        bad = A.copy()
        bad[17, 17] = -5.0
        assert hc.HCore.Potrf(hc.DenseTile(bad, ctx), "L", ctx) == 18


@pytest.mark.parametrize("dt", DT)
def test_trsm_syrk_triangle_helpers_vs_oracle(hc, ctx, dt):
    """hcb_?trsm (all side / uplo / trans / diag combinations), hcb_?syrk, fill_triangle, symmetrize and the HCore::Trsm /
    Syrk mirrors against the oracle (scipy / numpy restatements pinned against the compiled reference in test_oracle.py)."""
    import scipy.linalg as sl
    rng = np.random.default_rng(5)
    m, n = 70, 45
    tol = 1e-11 if dt == np.float64 else 5e-4
    for side in "LR":
        na = m if side == "L" else n
        Tm = (0.1 * rng.standard_normal((na, na)) + 4 * np.eye(na)).astype(dt)  # (well conditioned with a unit diagonal too)
        for uplo in "LU":
            for trans in (0, 1):
                for diag in "NU":
                    B = rng.standard_normal((m, n)).astype(dt)
                    dA, dB = dev(Tm), dev(B)
                    ok(fn("trsm", dt)(ctx.h, ord(side), ord(uplo), trans, ord(diag), m, n, ct(dt)(1.5), dA.data_ptr(), na,
                                      dB.data_ptr(), m))
                    tri = np.tril(Tm) if uplo == "L" else np.triu(Tm)
                    if diag == "U":
                        tri = tri - np.diag(np.diag(tri)) + np.eye(na)
                    op = (tri.T if trans else tri).astype(np.float64)
                    ref = 1.5 * (np.linalg.solve(op, B.astype(np.float64)) if side == "L"
                                 else np.linalg.solve(op.T, B.astype(np.float64).T).T)
                    assert relerr(host(dB, (m, n)), ref) < tol, (side, uplo, trans, diag)
    # syrk: only the uplo triangle is touched
    nn, kk = 60, 60
    A, Cm = rng.standard_normal((nn, kk)).astype(dt), rng.standard_normal((nn, nn)).astype(dt)
    for uplo in "LU":
        for ta in (False, True):
            tc = hc.DenseTile(Cm.copy(), ctx)
            hc.HCore.Syrk(-1.0, hc.DenseTile(A.copy(), ctx), ta, uplo, 0.5, tc, ctx)
            oc = O.DenseTile(np.asfortranarray(Cm.copy()))
            O.hcore_syrk_dense(dt(-1.0), O.DenseTile(np.asfortranarray(A.copy())), ta, uplo, dt(0.5), oc)
            assert relerr(tc.to_numpy(), oc.data) < tol, (uplo, ta)
    # FillMatrixTriangle / Symmetrize (omp/kernels.cpp:245-262, 283-303)
    for uplo in "LU":
        d = dev(Cm)
        ok(fn("fill_triangle", dt)(ctx.h, ord(uplo), nn, d.data_ptr(), nn, ct(dt)(0.0)))
        ref = np.tril(Cm) if uplo == "U" else np.triu(Cm)
        assert np.array_equal(host(d, (nn, nn)), ref)
        d = dev(Cm)
        ok(fn("symmetrize", dt)(ctx.h, ord(uplo), nn, d.data_ptr(), nn))
        tri = np.triu(Cm) if uplo == "U" else np.tril(Cm)
        assert np.array_equal(host(d, (nn, nn)), tri + tri.T - np.diag(np.diag(Cm)))
    # HCore::Trsm mirror: reference semantics (solve on B's V buffer viewed m x rank)
    Lm = np.asfortranarray((np.tril(rng.standard_normal((m, m))) + 4 * np.eye(m)).astype(dt))
    U, V = lowrank(rng, m, m, 9, dt)
    for side, trans in (("L", False), ("L", True)):
        tB = hc.CompressedTile.from_uv(U, V, ctx)
        hc.HCore.Trsm(side, "L", trans, "N", 2.0, hc.DenseTile(Lm.copy(), ctx), tB, ctx)
        oB = O.CompressedTile.from_uv(U, V)
        O.hcore_trsm(side, "L", trans, "N", dt(2.0), O.DenseTile(np.asfortranarray(Lm.copy())), oB)
        assert relerr(tB.GetVMatrix(), oB.V) < tol and relerr(tB.GetUMatrix(), oB.U) == 0.0
    with pytest.raises(RuntimeError, match="compressed"):
        hc.HCore.Trsm("L", "L", False, "N", 1.0, hc.DenseTile(Lm, ctx), hc.DenseTile(Lm, ctx), ctx)


@pytest.mark.parametrize("nt,nb,acc", [(4, 256, 1e-8), (5, 128, 1e-5), (3, 512, 1e-8)])
def test_tlr_cholesky_vs_oracle(hc, ctx, nt, nb, acc):
    """Tile Cholesky driver (hcb_dtlr_potrf) on a synthetic Matern covariance matrix against the oracle's loop over the
    reference's tile routines (oracle.tile_cholesky): factor L within 10*acc (relative Frobenius), tile ranks +/-1, and the
    size-independent property ||A - L L^T|| / ||A|| <= 10*acc."""
    pts, tile = O.covariance_tiles(nt, nb, ell=0.1, nugget=1e-2, seed=nt)
    p, po = hc.CompressionParameters(acc), O.CompressionParameters(acc)
    S = hc.SymTileMatrix.from_tiles(tile, nt, nb, torch.float64, ctx, p)
    diag = [tile(k, k) for k in range(nt)]
    low = {(i, j): O.CompressedTile.compress(tile(i, j), po) for j in range(nt) for i in range(j + 1, nt)}
    for (i, j), t in low.items():  # same compressed inputs on both sides (the compressing constructor has its own test)
        assert abs(S.low.GetTile(i, j).GetTileRank() - t.rank) <= 1
    pinfo = torch.full((nt,), -1, dtype=torch.int32, device="cuda")
    hc.tlr_cholesky(S, ctx, p, potrf_info=pinfo)
    ctx.Sync()
    assert int(pinfo.abs().max().item()) == 0
    O.tile_cholesky(diag, low, po)
    Lg = S.lower_factor_dense()
    Lo = np.zeros_like(Lg)
    for k in range(nt):
        Lo[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = np.tril(diag[k])
    for (i, j), t in low.items():
        Lo[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = t.to_dense()
        assert abs(S.low.GetTile(i, j).GetTileRank() - t.rank) <= 1, (i, j)
    assert relerr(Lg, Lo) <= 10 * acc
    A = np.block([[tile(i, j) for j in range(nt)] for i in range(nt)])
    assert relerr(Lg @ Lg.T, A) <= 10 * acc


def test_sketched_compression_second_sketch_for_ranks_beyond_96(hc, ctx):
    """Ranks between the first sketch (96 columns) and the second (288): the tile is rejected by the first range finder on
    the device and accepted by the wider one -- same rank and product as the reference's full SVD (oracle)."""
    rng = np.random.default_rng(11)
    nb, acc = 1024, 1e-8
    U, _ = np.linalg.qr(rng.standard_normal((nb, 260)))
    V, _ = np.linalg.qr(rng.standard_normal((nb, 260)))
    A = np.asfortranarray((U * (0.9 ** np.arange(260))) @ V.T)     # sigma_i = 0.9^i: rank 175 at 1e-8
    t = hc.CompressedTile.compress(A, hc.CompressionParameters(acc), ctx)
    o = O.CompressedTile.compress(A, O.CompressionParameters(acc))
    assert 120 < o.rank < 280 and abs(t.GetTileRank() - o.rank) <= 1
    assert relerr(t.to_dense(), o.to_dense()) <= 10 * acc


def test_per_tile_fixed_rank_replay_and_manifest(hc, ctx, tmp_path):
    """SURVEY.md 8f rows 2-3.  Fixed-rank replay (par_fixed_rank_streams_main.cpp:465-477,540-541): a first product learns
    the rank of every C tile, the replay pins each tile to ITS rank (hcb_tile.fixed_rank) and must reproduce ranks and
    products; against the oracle with the same per-tile fixed ranks.  Tile manifest: save -> load round trip is exact."""
    nb, T, k, acc = 128, 2, 10, 1e-6
    def fill(tm, og, name):
        for j in range(T):
            for i in range(T):
                t = O.synth_compressed_tile(nb, k + 2 * i + j, seed=ord(name) * 100 + j * 10 + i)   # different ranks per tile
                og[j][i] = t
                g = tm.GetTile(j, i)
                rk = t.U.shape[1]
                g.buf[: nb * rk] = dev(t.U)
                g.buf[nb * tm.max_rank: nb * tm.max_rank + rk * nb] = dev(t.V)
                g.rank.fill_(rk)
    A = hc.TileMatrix(T, T, nb, nb, torch.float64, ctx, compressed=True)
    B = hc.TileMatrix(T, T, nb, nb, torch.float64, ctx, compressed=True)
    oA, oB = [[None] * T for _ in range(T)], [[None] * T for _ in range(T)]
    fill(A, oA, "A"); fill(B, oB, "B")
    p = hc.CompressionParameters(acc)
    C1 = hc.TileMatrix.zeros_compressed(T, T, nb, nb, torch.float64, ctx)
    hc.tile_matrix_multiplication(A, B, C1, 1.0, 1.0, ctx, p)
    ctx.Sync()
    learned = C1.rank_table()
    forced = np.maximum(learned - np.array([[0, 3], [5, 1]]), 1)      # replay with per-tile ranks (some below the free rank)
    C2 = hc.TileMatrix.zeros_compressed(T, T, nb, nb, torch.float64, ctx)
    C2.set_fixed_ranks(forced)
    hc.tile_matrix_multiplication(A, B, C2, 1.0, 1.0, ctx, p)
    ctx.Sync()
    assert np.array_equal(C2.rank_table(), forced)
    for j in range(T):
        for i in range(T):
            oc = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
            pf = O.CompressionParameters(acc, fixed_rank=int(forced[j][i]))
            for kk in range(T):
                O.hcore_gemm(1.0, oA[j][kk], False, oB[kk][i], False, 1.0, oc, pf)
            assert oc.rank == forced[j][i]
            assert relerr(C2.GetTile(j, i).to_dense(), oc.to_dense()) <= 10 * acc
    C2.set_fixed_ranks(None)
    # manifest round trip
    path = str(tmp_path / "c1")
    C1.save(path)
    L = hc.TileMatrix.load(path, ctx)
    assert np.array_equal(L.rank_table(), learned)
    assert np.array_equal(L.ToRawMatrix(), C1.ToRawMatrix())
    import json
    man = json.load(open(path + ".json"))
    assert man["tiles"][1]["metadata"]["mType"] == "COMPRESSED" and man["tiles"][1]["metadata"]["mMatrixRank"] == int(learned[1][0])
    D = hc.TileMatrix.from_dense(np.arange(64.0).reshape(8, 8), 4, 4, ctx)
    D.save(str(tmp_path / "d"))
    assert np.array_equal(hc.TileMatrix.load(str(tmp_path / "d"), ctx).ToRawMatrix(), D.ToRawMatrix())


# ------------------------------------------------------------------------------------------------ round-2 test holes
def _ref_or_skip():
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    return R


def test_config0_replay_whole_matrix_generator(hc, ctx):
    """BASELINE.json configs[0] replayed EXACTLY: examples/matrix_multiplication/omp_main.cpp with 4 x 4 tiles of 512,
    accuracy 1e-8, the reference's WHOLE-MATRIX LATMS generator (A then B drawn from one seed state, omp_main.cpp:232-242),
    C0 = 0 -- on the GPU through TileMatrix(compressing) + tile_matrix_multiplication, side by side with the compiled
    reference on the same bytes.  The example prints error 1.067069e-07 and a C footprint of 10624 KB for this case."""
    R = _ref_or_skip()
    nb, T, acc = 512, 4, 1e-8
    A, B = R.latms_law(nb * T, nb * T, np.float64, tile_size=0, reps=2)
    p, pr = hc.CompressionParameters(acc), R.Params(acc)
    tA = hc.TileMatrix.from_dense(A, nb, nb, ctx, p)
    tB = hc.TileMatrix.from_dense(B, nb, nb, ctx, p)
    tC = hc.TileMatrix.from_dense(np.zeros_like(A), nb, nb, ctx, p)
    tile = lambda M, j, i: np.asfortranarray(M[j * nb:(j + 1) * nb, i * nb:(i + 1) * nb])
    rA = [[R.RefTile.compress(tile(A, j, k), pr) for k in range(T)] for j in range(T)]
    rB = [[R.RefTile.compress(tile(B, j, k), pr) for k in range(T)] for j in range(T)]
    rC = [[R.RefTile.compress(np.zeros((nb, nb), order="F"), pr) for _ in range(T)] for _ in range(T)]
    rank = lambda g: np.array([[t.info()["rank"] for t in r] for r in g])
    assert np.max(np.abs(tA.rank_table() - rank(rA))) <= 1 and np.max(np.abs(tB.rank_table() - rank(rB))) <= 1
    # "memory(KB)": the footprints TileMatrix records AT CONSTRUCTION (TileMatrix.cpp:176-184) -- C0's rank-1 zero tiles included
    fp_ref = int(rank(rA).sum() + rank(rB).sum() + rank(rC).sum()) * (nb + nb) * 8
    fp_gpu = tA.GetMemoryFootprint() + tB.GetMemoryFootprint() + tC.GetMemoryFootprint()
    hc.tile_matrix_multiplication(tA, tB, tC, 1.0, 1.0, ctx, p)
    ctx.Sync()
    R.matmul(rA, rB, rC, 1.0, 1.0, pr, nthreads=4)
    assert np.max(np.abs(tC.rank_table() - rank(rC))) <= 1, (tC.rank_table(), rank(rC))
    Cr = np.block([[t.to_dense() for t in r] for r in rC])
    Cg = tC.ToRawMatrix()
    assert relerr(Cg, Cr) <= 10 * acc
    # the example's printed "error" is RawMatrix::Norm() of the difference to the dense product = the INFINITY norm
    # (RawMatrix.cpp:45-48), its "memory" the footprint of A + B + C in KB (omp_main.cpp:366-377)
    dense = A @ B
    inf = lambda X: float(np.abs(X - dense).sum(axis=1).max())
    err_ref, err_gpu = inf(Cr), inf(Cg)
    assert abs(err_ref - 1.067069e-07) <= 0.02e-07 and abs(err_gpu - err_ref) <= 0.1 * err_ref, (err_ref, err_gpu)
    a_inf, b_inf = np.abs(A).sum(axis=1).max(), np.abs(B).sum(axis=1).max()
    assert err_gpu / ((a_inf + b_inf) * acc * nb * T) < 10        # the example's own pass criterion
    assert fp_ref == 10624 * 1024                                  # ... and its printed footprint
    assert abs(fp_gpu - fp_ref) <= 2 * T * T * 2 * nb * 8          # (A / B ranks within +/-1 per tile)


@pytest.mark.parametrize("mix", ["CCC", "CDC", "DCC"])
@pytest.mark.parametrize("ops", [(True, False), (False, True), (True, True)])
def test_transposed_operands_compressed_output(hc, ctx, mix, ops):
    """op = Trans with a COMPRESSED C (only dense C was covered).  The reference itself cannot serve as the
    checker here -- outside its aCholesky branch HCore::Gemm reads a transposed compressed operand with the wrong shapes
    (NaN from the compiled reference, tests/test_oracle.py) -- so the check is against the dense truth: the recompressed
    result within 10*acc and its rank = the numerical rank of the truth at that accuracy (+/-1)."""
    rng = np.random.default_rng(hash((mix, ops)) % 1000)
    dt, acc = np.float64, 1e-8
    m, n, k = 160, 140, 96
    ta, tb = ops
    AUV = lowrank(rng, k, m, 20, dt) if ta else lowrank(rng, m, k, 20, dt)
    BUV = lowrank(rng, n, k, 18, dt) if tb else lowrank(rng, k, n, 18, dt)
    CUV = lowrank(rng, m, n, 25, dt)
    Ad, Bd, Cd = AUV[0] @ AUV[1], BUV[0] @ BUV[1], CUV[0] @ CUV[1]
    a = mk_tile(hc, ctx, mix[0], Ad, *AUV, dt)
    b = mk_tile(hc, ctx, mix[1], Bd, *BUV, dt)
    c = mk_tile(hc, ctx, "C", None, *CUV, dt, max_rank=min(m, n))
    for _ in range(2):  # second pass: C carries the orthonormal-U state -> incremental path with transposed operands
        hc.HCore.Gemm(0.75, a, ta, b, tb, -1.25, c, ctx, hc.CompressionParameters(acc))
        Cd = 0.75 * (Ad.T if ta else Ad) @ (Bd.T if tb else Bd) - 1.25 * Cd
        assert relerr(c.to_dense(), Cd) <= 10 * acc
        s = np.linalg.svd(Cd, compute_uv=False)
        assert abs(c.GetTileRank() - int(O.k_new_rank(False, s, len(s), acc))) <= 1


@pytest.mark.parametrize("dims", [(1000, 900, 500, 50, 50, 60), (257, 300, 128, 40, 40, 50), (1024, 1024, 1024, 44, 44, 120)],
                         ids=lambda d: "x".join(map(str, d)))
def test_blocked_recompression_fp32(hc, ctx, dims):
    """fp32 on the blocked path (stacked rank > 64: GEMM-based blocked QR + shared-memory / register Jacobi) at accuracy
    1e-4 against the fp32 oracle (all blocked tests were f64)."""
    m, n, k, ka, kb, kc = dims
    dt, acc = np.float32, 1e-4
    rng = np.random.default_rng(sum(dims) + 1)
    AUV, BUV, CUV = lowrank(rng, m, k, ka, dt, 0.9), lowrank(rng, k, n, kb, dt, 0.9), lowrank(rng, m, n, kc, dt, 0.97)
    cap = min(m, n)
    a, b = mk_tile(hc, ctx, "C", None, *AUV, dt), mk_tile(hc, ctx, "C", None, *BUV, dt)
    c = mk_tile(hc, ctx, "C", None, *CUV, dt, max_rank=cap)
    oa, ob, oc = (oracle_tile("C", None, x, dt) for x in (AUV, BUV, CUV))
    oc.max_rank = cap
    for _ in range(2):
        hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, hc.CompressionParameters(acc))
        O.hcore_gemm(dt(1.0), oa, False, ob, False, dt(1.0), oc, O.CompressionParameters(acc))
        ref = oc.to_dense().astype(np.float64)
        assert np.linalg.norm(c.to_dense().astype(np.float64) - ref) <= 10 * acc * max(np.linalg.norm(ref), 1.0)
        # fp32 singular values near the threshold carry ~1e-6 * sigma_0 of noise on both sides: ranks within a few
        assert abs(c.GetTileRank() - oc.rank) <= max(1, int(0.03 * oc.rank)), (c.GetTileRank(), oc.rank)


def test_fp32_promoted_path_vs_native_and_oracle(hc, ctx, monkeypatch):
    """Round 2: FP32 tiles on the blocked path run on the FP64 machinery through FP64 shadows.  The same
    4-step k-sum at nb = 512, accuracy 1e-4, three ways: promoted (default), native FP32 kernels (HCB_FP32_NATIVE=1), FP32
    oracle.  Both GPU paths must meet the 10 * accuracy contract; a shared operand (A == B buffer) and a dense operand
    (DCC) go through the shadow de-duplication."""
    nb, ka, acc, dt = 512, 30, 1e-4, np.float32
    rng = np.random.default_rng(77)
    steps = [(lowrank(rng, nb, nb, ka, dt, 0.8), lowrank(rng, nb, nb, ka, dt, 0.8)) for _ in range(4)]
    C0 = lowrank(rng, nb, nb, 60, dt, 0.93)
    outs = {}
    for mode in ("promoted", "native"):
        if mode == "native":
            monkeypatch.setenv("HCB_FP32_NATIVE", "1")
        else:
            monkeypatch.delenv("HCB_FP32_NATIVE", raising=False)
        c = mk_tile(hc, ctx, "C", None, *C0, dt, max_rank=nb // 2)
        for AUV, BUV in steps:
            a, b = mk_tile(hc, ctx, "C", None, *AUV, dt), mk_tile(hc, ctx, "C", None, *BUV, dt)
            hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, hc.CompressionParameters(acc))
        ctx.Sync()
        outs[mode] = (c.to_dense().astype(np.float64), c.GetTileRank())
    monkeypatch.delenv("HCB_FP32_NATIVE", raising=False)
    oc = oracle_tile("C", None, C0, dt)
    oc.max_rank = nb // 2
    for AUV, BUV in steps:
        O.hcore_gemm(dt(1.0), oracle_tile("C", None, AUV, dt), False, oracle_tile("C", None, BUV, dt), False, dt(1.0), oc,
                     O.CompressionParameters(acc))
    ref = oc.to_dense().astype(np.float64)
    for mode, (d, rk) in outs.items():
        assert np.linalg.norm(d - ref) <= 10 * acc * np.linalg.norm(ref), mode
        assert abs(rk - oc.rank) <= max(1, int(0.03 * oc.rank)), (mode, rk, oc.rank)
    # X * X^T with ONE buffer for both operands, and a dense left operand (DCC): shadows are shared / dense tiles converted
    X = mk_tile(hc, ctx, "C", None, *steps[0][0], dt)
    c2 = mk_tile(hc, ctx, "C", None, *C0, dt, max_rank=nb // 2)
    hc.HCore.Gemm(1.0, X, False, X, True, 1.0, c2, ctx, hc.CompressionParameters(acc))
    xd = (steps[0][0][0].astype(np.float64) @ steps[0][0][1].astype(np.float64))
    want = C0[0].astype(np.float64) @ C0[1].astype(np.float64) + xd @ xd.T
    assert np.linalg.norm(c2.to_dense().astype(np.float64) - want) <= 10 * acc * np.linalg.norm(want)
    Dn = rng.standard_normal((nb, nb)).astype(dt) / np.sqrt(nb)
    dA = mk_tile(hc, ctx, "D", Dn, None, None, dt)
    c3 = mk_tile(hc, ctx, "C", None, *C0, dt, max_rank=nb // 2)
    hc.HCore.Gemm(1.0, dA, False, X, False, 1.0, c3, ctx, hc.CompressionParameters(acc))
    want3 = C0[0].astype(np.float64) @ C0[1].astype(np.float64) + Dn.astype(np.float64) @ xd
    assert np.linalg.norm(c3.to_dense().astype(np.float64) - want3) <= 10 * acc * np.linalg.norm(want3)


@pytest.mark.parametrize("nb,rank,mixes", [(1024, 128, MIXES), (1024, 256, MIXES), (2048, 256, ["CCC", "CDC", "DCC", "CCD"])],
                         ids=["nb1024-r128", "nb1024-r256", "nb2048-r256"])
def test_single_tile_sweep_large_ranks(hc, ctx, nb, rank, mixes):
    """BASELINE.json configs[1]: the single-tile HCore::Gemm sweep at its upper end -- tile 1024 / 2048, operand ranks 128
    and 256 (stacked rank up to 512: the one-column-per-warp register Jacobi, 8-CTA strip clusters) -- against the oracle."""
    dt, acc = np.float64, 1e-8
    rng = np.random.default_rng(nb + rank)
    AUV, BUV, CUV = lowrank(rng, nb, nb, rank, dt, 0.93), lowrank(rng, nb, nb, rank, dt, 0.93), lowrank(rng, nb, nb, rank, dt, 0.95)
    Ad, Bd, Cd = AUV[0] @ AUV[1], BUV[0] @ BUV[1], CUV[0] @ CUV[1]
    for mix in mixes:
        a = mk_tile(hc, ctx, mix[0], Ad, *AUV, dt)
        b = mk_tile(hc, ctx, mix[1], Bd, *BUV, dt)
        cap = nb if mix == "DDC" else nb // 3 + rank
        c = mk_tile(hc, ctx, mix[2], Cd, *CUV, dt, max_rank=cap)
        oa, ob, oc = oracle_tile(mix[0], Ad, AUV, dt), oracle_tile(mix[1], Bd, BUV, dt), oracle_tile(mix[2], Cd, CUV, dt)
        if mix[2] == "C":
            oc.max_rank = cap
        hc.HCore.Gemm(1.0, a, False, b, False, 1.0, c, ctx, hc.CompressionParameters(acc))
        O.hcore_gemm(1.0, oa, False, ob, False, 1.0, oc, O.CompressionParameters(acc))
        ref = oc.to_dense() if mix != "DDC" else Ad @ Bd + Cd
        assert relerr(c.to_dense(), ref) <= 10 * acc, mix
        if mix[2] == "C" and mix != "DDC":
            assert abs(c.GetTileRank() - oc.rank) <= 1, (mix, c.GetTileRank(), oc.rank)


def test_workspace_formula_covers_the_arena(hc):
    """a2 / a11: hcb_?tlr_gemm_workspace (the replacement of HCore::CalculateMemoryPoolSize, HCore.cpp:417-480) must be an
    upper bound of what a fused call actually makes the context's arena grow to (plus the arena's own 12.5 % + 1 MiB slack)."""
    from hcorepp_b200 import _capi
    rng = np.random.default_rng(2)
    ctx2 = hc.RunContext(0)          # fresh, empty arena
    nb, n_tiles, ka, kc = 512, 6, 30, 70
    As = [hc.CompressedTile.from_uv(*lowrank(rng, nb, nb, ka, np.float64), ctx2) for _ in range(n_tiles)]
    Bs = [hc.CompressedTile.from_uv(*lowrank(rng, nb, nb, ka, np.float64), ctx2) for _ in range(n_tiles)]
    Cs = [hc.CompressedTile.from_uv(*lowrank(rng, nb, nb, kc, np.float64), ctx2, max_rank=nb // 3) for _ in range(n_tiles)]
    for t in Cs:
        t.rank_bound = kc + ka
    assert _capi.lib.hcb_ctx_workspace_bytes(ctx2.h) == 0
    hc.gemm_batched(1.0, As, False, Bs, False, 1.0, Cs, ctx2, hc.CompressionParameters(1e-8))
    ctx2.Sync()
    used = _capi.lib.hcb_ctx_workspace_bytes(ctx2.h)
    formula = _capi.lib.hcb_dtlr_gemm_workspace(n_tiles, nb, nb, nb, (kc + ka) + ka)   # stacked bound = C's rank bound + ka
    assert 0 < used <= formula * 1.125 + (2 << 20), (used, formula)


def test_dropin_reference_operator_layer_on_libhcore_b200(tmp_path):
    """the reference's UNMODIFIED src/api/HCore.cpp + src/operators/concrete/{Compressed,Dense}.cpp
    (compiled with -DUSE_CUDA by tests/dropin/Makefile) linked against integration/src/kernels/b200/kernels.cpp, i.e. the
    kernel table HCoreKernels<T> forwarded to libhcore_b200.so.  The binary replays a TestGemm-style known answer, a Potrf,
    and a seeded 5-step k-sum through the reference's own tile-at-a-time flow (Geqrf / ungqr / SVD / CalculateNewRank / ...
    per tile); the result is compared with the oracle on the same bytes: ranks +/-1, relative Frobenius <= 10*acc."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin", "dropin_test")
    assert os.path.exists(exe), "tests/dropin/dropin_test not built (run __graft_entry__.build() where /root/reference exists)"
    nb, ka, steps, acc = 192, 16, 5, 1e-8
    tiles = [(O.synth_compressed_tile(nb, ka, 500 + k), O.synth_compressed_tile(nb, ka, 600 + k)) for k in range(steps)]
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([nb, ka, steps], dtype=np.int64).tobytes())
        f.write(np.array([acc], dtype=np.float64).tobytes())
        for a, b in tiles:
            for x in (a.U, a.V, b.U, b.V):
                f.write(np.asfortranarray(x, dtype=np.float64).tobytes(order="F"))
    r = subprocess.run([exe, inp, out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "0 failure(s)" in r.stdout, r.stdout[-3000:]
    raw = np.fromfile(out, dtype=np.uint8)
    ranks = raw[: 8 * steps].view(np.int64)
    dense = raw[8 * steps:].view(np.float64).reshape((nb, nb), order="F")
    oC = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
    oranks = []
    for a, b in tiles:
        O.hcore_gemm(1.0, a, False, b, False, 1.0, oC, O.CompressionParameters(acc))
        oranks.append(oC.rank)
    assert max(abs(int(x) - y) for x, y in zip(ranks, oranks)) <= 1, (ranks, oranks)
    assert relerr(dense, oC.to_dense()) <= 10 * acc
