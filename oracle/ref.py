"""ctypes binding to oracle/_ref/libhcorepp_ref.so -- the UNMODIFIED reference CPU path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
hcorepp_b200/.  The library is built by `make -C oracle ref` (see oracle/Makefile) from the sources where they lie
under /root/reference; on the GPU box only the prebuilt .so exists.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhcorepp_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
        # The scipy-wheel OpenBLAS is pthread-threaded; inside the reference's OMP regions that only prints
        # "Detect OpenMP Loop" warnings and oversubscribes (SURVEY.md 8c caveat iii) -> serial BLAS by default.
        _lib.hcref_set_blas_threads(1)
    return _lib


_P = {np.dtype(np.float64): "d", np.dtype(np.float32): "s"}
_CT = {np.dtype(np.float64): C.c_double, np.dtype(np.float32): C.c_float}
i64 = C.c_int64
vp = C.c_void_p


def _declare(L):
    for p, ct in (("d", C.c_double), ("s", C.c_float)):
        f = lambda n: getattr(L, f"hcref_{p}{n}")
        f("tile_dense").restype = vp
        f("tile_dense").argtypes = [i64, i64, vp, i64]
        f("tile_uv").restype = vp
        f("tile_uv").argtypes = [i64, i64, vp, vp, i64]
        f("tile_uv_cap").restype = vp
        f("tile_uv_cap").argtypes = [i64, i64, vp, vp, i64, i64]
        f("tile_compress").restype = vp
        f("tile_compress").argtypes = [i64, i64, vp, i64, C.c_double, C.c_int, C.c_int, C.c_int, i64, C.c_int]
        f("tile_info").argtypes = [vp, vp]
        f("tile_read").argtypes = [vp, vp, vp]
        f("tile_free").argtypes = [vp]
        f("gemm").restype = C.c_int
        f("gemm").argtypes = [ct, vp, C.c_int, vp, C.c_int, ct, vp, C.c_double, C.c_int, C.c_int, C.c_int, i64,
                              C.c_int, vp]
        f("potrf").restype = C.c_int
        f("potrf").argtypes = [vp, C.c_int]
        f("trsm").restype = C.c_int
        f("trsm").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ct, vp, vp]
        f("syrk").restype = C.c_int
        f("syrk").argtypes = [ct, vp, C.c_int, C.c_int, ct, vp]
        f("matmul").restype = C.c_int
        f("matmul").argtypes = [i64, i64, i64, vp, vp, vp, ct, ct, C.c_double, C.c_int, C.c_int, C.c_int, i64,
                                C.c_int, C.c_int, vp, vp]
        f("latms_law").argtypes = [i64, i64, i64, vp, vp, i64, C.c_int]
        f("generate_dense").argtypes = [i64, i64, vp, i64, vp]
        f("compress_dense").restype = i64
        f("compress_dense").argtypes = [i64, i64, vp, i64, C.c_double, vp]
        f("k_gemm").argtypes = [C.c_int, C.c_int, i64, i64, i64, ct, vp, i64, vp, i64, ct, vp, i64]
        f("k_multiply_by_alpha").argtypes = [vp, i64, i64, i64, i64, ct]
        f("k_process_v").argtypes = [i64, i64, C.c_int, i64, ct, vp, i64, vp, i64, vp, C.c_int]
        f("k_new_rank").restype = i64
        f("k_new_rank").argtypes = [C.c_int, vp, i64, ct]
        f("k_uvptr").argtypes = [i64, i64, vp, vp]
        f("k_vtnew").argtypes = [i64, C.c_int, i64, vp, vp, i64, i64]
        f("k_fill_identity").argtypes = [i64, vp]
        f("k_lacpy").argtypes = [C.c_int, i64, i64, vp, i64, vp, i64]
        f("k_laset").argtypes = [C.c_int, i64, i64, ct, ct, vp, i64]
        f("k_geqrf").argtypes = [i64, i64, vp, i64, vp]
        f("k_ungqr").argtypes = [i64, i64, i64, vp, i64, vp]
        f("k_unmqr").argtypes = [C.c_int, C.c_int, i64, i64, i64, vp, i64, vp, vp, i64]
        f("k_svd").argtypes = [i64, i64, vp, i64, vp, vp, i64, vp, i64, C.c_int]
        f("k_trmm").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, i64, i64, ct, vp, i64, vp, i64]
    L.hcref_set_blas_threads.argtypes = [C.c_int]
    L.hcref_max_threads.restype = C.c_int


def fn(name: str, dtype):
    return getattr(lib(), f"hcref_{_P[np.dtype(dtype)]}{name}")


def ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


def fcol(a, dtype=None) -> np.ndarray:
    """Column-major contiguous copy (the reference is column-major everywhere on this path)."""
    return np.asfortranarray(np.array(a, dtype=dtype, copy=True))


@dataclass
class Params:
    """Mirror of CompressionParameters (include/hcorepp/operators/helpers/CompressionParameters.hpp:44-46)."""
    accuracy: float = 1e-4
    use_trmm: bool = False
    use_ungqr: bool = True
    truncated_svd: bool = False
    fixed_rank: int = 0
    svd: int = 1  # 0 = LAPACK_GESVD, 1 = LAPACK_GESDD (default)

    def cargs(self):
        return (C.c_double(self.accuracy), int(self.use_trmm), int(self.use_ungqr), int(self.truncated_svd),
                i64(self.fixed_rank), int(self.svd))


class RefTile:
    """Owning handle to a reference DenseTile / CompressedTile."""

    def __init__(self, handle, dtype):
        self.h = handle
        self.dtype = np.dtype(dtype)

    @classmethod
    def dense(cls, a):
        a = fcol(a)
        return cls(fn("tile_dense", a.dtype)(a.shape[0], a.shape[1], ptr(a), a.shape[0]), a.dtype)

    @classmethod
    def from_uv(cls, U, V):
        """CompressedTile(m, n, U, V, ld, rank, ...) -- maxRank = rank (Compressed.cpp:20-47)."""
        U = fcol(U)
        V = fcol(V, U.dtype)
        assert U.shape[1] == V.shape[0]
        return cls(fn("tile_uv", U.dtype)(U.shape[0], V.shape[1], ptr(U), ptr(V), U.shape[1]), U.dtype)

    @classmethod
    def from_uv_cap(cls, U, V, max_rank):
        """Same state as a compress-constructed tile (capacity max_rank, V ld = rank) without running the SVD."""
        U = fcol(U)
        V = fcol(V, U.dtype)
        return cls(fn("tile_uv_cap", U.dtype)(U.shape[0], V.shape[1], ptr(U), ptr(V), U.shape[1], max_rank), U.dtype)

    @classmethod
    def compress(cls, a, p: Params):
        """Compressing constructor -- SVD + truncate, maxRank = min(m,n)/3 (Compressed.cpp:75-146)."""
        a = fcol(a)
        return cls(fn("tile_compress", a.dtype)(a.shape[0], a.shape[1], ptr(a), a.shape[0], *p.cargs()), a.dtype)

    def info(self):
        out = np.zeros(6, dtype=np.int64)
        fn("tile_info", self.dtype)(self.h, ptr(out))
        return dict(m=int(out[0]), n=int(out[1]), rank=int(out[2]), dense=bool(out[3]), max_rank=int(out[4]),
                    ld=int(out[5]))

    def read(self):
        """Dense tile -> (A,) ; compressed tile -> (U (m x rank), V (rank x n))."""
        i = self.info()
        if i["dense"]:
            a = np.zeros((i["m"], i["n"]), dtype=self.dtype, order="F")
            fn("tile_read", self.dtype)(self.h, ptr(a), None)
            return (a,)
        u = np.zeros((i["m"], i["rank"]), dtype=self.dtype, order="F")
        v = np.zeros((i["rank"], i["n"]), dtype=self.dtype, order="F")
        fn("tile_read", self.dtype)(self.h, ptr(u), ptr(v))
        return u, v

    def to_dense(self):
        r = self.read()
        return r[0] if len(r) == 1 else r[0] @ r[1]

    def free(self):
        if self.h:
            fn("tile_free", self.dtype)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def gemm(alpha, A: RefTile, opA: bool, B: RefTile, opB: bool, beta, Ct: RefTile, p: Params) -> int:
    """hcorepp::api::HCore<T>::Gemm (src/api/HCore.cpp:22-344). Returns the reference's flop counter."""
    ct = _CT[Ct.dtype]
    flops = i64(0)
    rc = fn("gemm", Ct.dtype)(ct(alpha), A.h, int(opA), B.h, int(opB), ct(beta), Ct.h, *p.cargs(), C.byref(flops))
    if rc != 0:
        raise RuntimeError("reference HCore::Gemm threw")
    return int(flops.value)


# ---- TLR Cholesky pieces: no driver and no enabled tests in the reference; pinned here only
def potrf(A: RefTile, uplo: str = "L"):
    """HCore<T>::Potrf (HCore.cpp:586-621): dense tiles only, in place."""
    if fn("potrf", A.dtype)(A.h, ord(uplo)) != 0:
        raise RuntimeError("reference HCore::Potrf threw")


def trsm(side: str, uplo: str, trans: bool, diag: str, alpha, A: RefTile, B: RefTile):
    """HCore<T>::Trsm (HCore.cpp:624-647): A dense triangular, B compressed; the solve is applied to B's V buffer."""
    if fn("trsm", B.dtype)(ord(side), ord(uplo), int(trans), ord(diag), _CT[B.dtype](alpha), A.h, B.h) != 0:
        raise RuntimeError("reference HCore::Trsm threw")


def syrk(alpha, A: RefTile, opA: bool, uplo: str, beta, Ct: RefTile):
    """HCore<T>::Syrk (HCore.cpp:484-583)."""
    if fn("syrk", Ct.dtype)(_CT[Ct.dtype](alpha), A.h, int(opA), ord(uplo), _CT[Ct.dtype](beta), Ct.h) != 0:
        raise RuntimeError("reference HCore::Syrk threw")


def matmul(A, B, Cg, alpha, beta, p: Params, nthreads: int = 0):
    """Tile loop of examples/matrix_multiplication/omp_main.cpp:112-126.

    A, B, Cg: 2-D lists (grid[row][col]) of RefTile. Returns (seconds, flops)."""
    mt, kt = len(A), len(A[0])
    nt = len(B[0])
    assert len(B) == kt and len(Cg) == mt and len(Cg[0]) == nt
    dtype = Cg[0][0].dtype
    arr = lambda g, r, c: (vp * (r * c))(*[g[j][i].h for i in range(c) for j in range(r)])
    a_, b_, c_ = arr(A, mt, kt), arr(B, kt, nt), arr(Cg, mt, nt)
    ct = _CT[dtype]
    sec = C.c_double(0)
    flops = i64(0)
    rc = fn("matmul", dtype)(mt, nt, kt, a_, b_, c_, ct(alpha), ct(beta), *p.cargs(), int(nthreads),
                             C.byref(sec), C.byref(flops))
    if rc != 0:
        raise RuntimeError("reference tile matmul threw")
    return sec.value, int(flops.value)


def latms_law(m, n, dtype=np.float64, tile_size=0, seed=(0, 0, 0, 1), reps=1):
    """LatmsGenerator / TileLatmsGenerator (src/helpers/generators/*.cpp), `reps` consecutive draws."""
    seed = np.array(seed, dtype=np.int64)
    out = np.zeros((reps, n, m), dtype=dtype)  # each draw is m x n column-major
    fn("latms_law", dtype)(m, n, tile_size, ptr(seed), ptr(out), m, reps)
    return [np.asfortranarray(out[r].T) for r in range(reps)]


def generate_dense(m, n, dtype=np.float64, seed=None):
    """matrixhelpers::generate_dense_matrix (src/helpers/MatrixHelpers.cpp:21-36): sigma_i = 10^-i; seed in/out."""
    if seed is None:
        seed = np.array([0, 0, 0, 1], dtype=np.int64)
    a = np.zeros((m, n), dtype=dtype, order="F")
    fn("generate_dense", dtype)(m, n, ptr(a), m, ptr(seed))
    return a


def compress_dense(a, acc):
    """matrixhelpers::compress_dense_matrix (src/helpers/MatrixHelpers.cpp:48-100) -> (U, V)."""
    a = fcol(a)
    m, n = a.shape
    uv = np.zeros((m + n) * min(m, n), dtype=a.dtype)
    rk = int(fn("compress_dense", a.dtype)(m, n, ptr(a), m, C.c_double(acc), ptr(uv)))
    U = uv[: m * rk].reshape((m, rk), order="F").copy(order="F")
    V = uv[m * rk: (m + n) * rk].reshape((rk, n), order="F").copy(order="F")
    return U, V
