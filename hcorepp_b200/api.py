"""Host-side mirror of the reference's operator interface for the TLR-GEMM path, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference (file:line relative to the ecrc/hcorepp tree):
  RunContext              include/hcorepp/kernels/cuda/RunContext.hpp:15-53
  CompressionParameters   include/hcorepp/operators/helpers/CompressionParameters.hpp:44-46
  DenseTile               include/hcorepp/operators/concrete/Dense.hpp:56-91
  CompressedTile          include/hcorepp/operators/concrete/Compressed.hpp:67-151
  HCore.Gemm              include/hcorepp/api/HCore.hpp:41-46   (src/api/HCore.cpp:22-344)
  TileMatrix              include/hcorepp/helpers/TileMatrix.hpp:22-189
  tile_matrix_multiplication   examples/matrix_multiplication/omp_main.cpp:70-154

PyTorch is used for device memory and streams only (plumbing); all arithmetic happens in libhcore_b200.so.
Nothing here imports the test-only checker package and there is no CPU path: every call ends in a CUDA kernel or raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi
from ._capi import lib, check, hcb_tile, hcb_compress_params, TILE_DENSE, TILE_COMPRESSED

MAX_RANK_RATIO = 3  # include/hcorepp/operators/concrete/Compressed.hpp:14

_PFX = {torch.float64: "d", torch.float32: "s"}
_CT = {torch.float64: C.c_double, torch.float32: C.c_float}
_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}


def _fn(name, dtype):
    return getattr(lib, f"hcb_{_PFX[dtype]}{name}")


class RunContext:
    """CUDA run context: device + the stream every kernel is enqueued on (torch's current stream)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        if not torch.cuda.is_available():
            # let the library report it: there is no CPU fallback
            check(lib.hcb_ctx_create(device, C.byref(h)))
        torch.cuda.set_device(device)
        self.device = torch.device("cuda", device)
        self.torch_stream = torch.cuda.current_stream(device)
        check(lib.hcb_ctx_create_on_stream(device, C.c_void_p(self.torch_stream.cuda_stream), C.byref(h)))
        self.h = h

    def Sync(self):
        check(lib.hcb_ctx_sync(self.h))

    def SupportsOMP(self) -> bool:
        return False

    def stats(self, reset: bool = False) -> dict:
        """Counters of the adaptive fast paths (include/hcore_b200.h: hcb_ctx_stats)."""
        out = (C.c_uint64 * 8)()
        check(lib.hcb_ctx_stats(self.h, out, 1 if reset else 0))
        return {"cholqr_panels": int(out[0]), "cholqr_fallback_pass0": int(out[1]), "cholqr_fallback_pass1": int(out[2]),
                "gs_second_pass_skipped": int(out[3]), "gs_second_pass_run": int(out[4]), "deflated_columns": int(out[5]), "vcore_cholesky": int(out[6]), "vcore_fallback": int(out[7])}

    def reserve_workspace(self, nbytes: int):
        check(lib.hcb_ctx_reserve_workspace(self.h, nbytes))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.hcb_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass


@dataclass
class CompressionParameters:
    accuracy: float = 1e-4
    use_trmm: bool = False
    use_ungqr: bool = True
    truncated_svd: bool = False
    fixed_rank: int = 0
    svd: int = 1  # LAPACK_GESDD in the reference; the device path is one-sided Jacobi either way

    def c(self) -> hcb_compress_params:
        return hcb_compress_params(self.accuracy, int(self.use_trmm), int(self.use_ungqr), int(self.truncated_svd),
                                   int(self.fixed_rank), int(self.svd), 0)


def _to_device_colmajor(a, ctx: RunContext) -> torch.Tensor:
    """m x n array -> device tensor of shape (n, m), C-contiguous == column-major m x n with ld = m."""
    if isinstance(a, torch.Tensor):
        t = a.detach().to(ctx.device)
        return t.t().contiguous()
    a = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(ctx.device)


class DenseTile:
    """DenseTile(m, n, data, ld, ColMajor, ctx) -- one device buffer, column-major."""

    def __init__(self, data, ctx: RunContext):
        self.ctx = ctx
        self.t = _to_device_colmajor(data, ctx)  # (n, m)
        self.n, self.m = self.t.shape
        self.dtype = self.t.dtype

    dense = True

    def isDense(self):
        return True

    def isCompressed(self):
        return False

    def GetNumOfRows(self):
        return self.m

    def GetNumOfCols(self):
        return self.n

    def desc(self) -> hcb_tile:
        return hcb_tile(TILE_DENSE, self.m, self.n, self.m, 0, 0, None, self.t.data_ptr(), None, 0, 0)

    def to_numpy(self) -> np.ndarray:
        return np.asfortranarray(self.t.t().cpu().numpy())

    to_dense = to_numpy


class CompressedTile:
    """U (m x rank, ld m) at offset 0, V (rank x n, ld rank) at offset m*max_rank; the rank lives on the device."""

    dense = False

    def __init__(self, m, n, max_rank, dtype, ctx: RunContext, buf=None, rank=None, rank_bound=0, state=None,
                 fixed_rank=0):
        self.ctx, self.m, self.n, self.max_rank, self.dtype = ctx, int(m), int(n), int(max_rank), dtype
        self.rank_bound = int(rank_bound)
        self.fixed_rank = int(fixed_rank)
        self.buf = buf if buf is not None else torch.zeros(self.m * self.max_rank + self.max_rank * self.n,
                                                          dtype=dtype, device=ctx.device)
        self.rank = rank if rank is not None else torch.ones(1, dtype=torch.int32, device=ctx.device)
        # device state word kept by the library next to the rank (bit 0: "U has orthonormal columns"); a new / externally
        # written tile starts at 0 = unknown
        self.state = state if state is not None else torch.zeros(1, dtype=torch.int32, device=ctx.device)

    # -- constructors mirroring Compressed.hpp:67-151
    @classmethod
    def from_uv(cls, U, V, ctx: RunContext, max_rank=None):
        """CompressedTile(m, n, U, V, ld, rank, ...): maxRank = rank unless given (Compressed.cpp:28)."""
        U, V = np.asarray(U), np.asarray(V)
        m, rk = U.shape
        n = V.shape[1]
        assert V.shape[0] == rk
        t = cls(m, n, max_rank or max(rk, 1), _NP2T[U.dtype], ctx)
        if rk > 0:
            t.buf[: m * rk] = torch.from_numpy(np.ascontiguousarray(U.T).reshape(-1)).to(ctx.device)
            t.buf[m * t.max_rank: m * t.max_rank + rk * n] = torch.from_numpy(
                np.ascontiguousarray(V.astype(U.dtype).T).reshape(-1)).to(ctx.device)
        t.rank.fill_(rk)
        return t

    @classmethod
    def compress(cls, dense, params: CompressionParameters, ctx: RunContext):
        """Compressing constructor (Compressed.cpp:75-146): maxRank = max(min(m,n)/3, 1)."""
        d = DenseTile(dense, ctx)
        t = cls(d.m, d.n, max(min(d.m, d.n) // MAX_RANK_RATIO, 1), d.dtype, ctx)
        ptrs = (C.c_void_p * 1)(d.t.data_ptr())
        descs = (hcb_tile * 1)(t.desc())
        prm = params.c()
        check(_fn("compress_batched", d.dtype)(ctx.h, 1, ptrs, d.m, descs, C.byref(prm), None))
        ctx.Sync()
        return t

    def isDense(self):
        return False

    def isCompressed(self):
        return True

    def GetNumOfRows(self):
        return self.m

    def GetNumOfCols(self):
        return self.n

    def GetTileRank(self) -> int:
        """Host getter: refreshes from the device-resident rank (synchronises, like RunContext::Sync())."""
        return int(self.rank.item())

    def GetULeadingDim(self):
        return self.m

    def GetVLeadingDim(self):
        return self.GetTileRank()

    def desc(self) -> hcb_tile:
        return hcb_tile(TILE_COMPRESSED, self.m, self.n, 0, self.max_rank, self.rank_bound, self.rank.data_ptr(),
                        self.buf.data_ptr(), self.state.data_ptr(), self.fixed_rank, 0)

    def invalidate(self):
        """Call after writing the factors through `buf` directly: the library may no longer assume an orthonormal U."""
        self.state.zero_()

    def factors(self):
        rk = self.GetTileRank()
        U = self.buf[: self.m * rk].reshape(rk, self.m).t()
        V = self.buf[self.m * self.max_rank: self.m * self.max_rank + rk * self.n].reshape(self.n, rk).t()
        return U, V

    def GetUMatrix(self) -> np.ndarray:
        return np.asfortranarray(self.factors()[0].cpu().numpy())

    def GetVMatrix(self) -> np.ndarray:
        return np.asfortranarray(self.factors()[1].cpu().numpy())

    def to_dense(self) -> np.ndarray:
        """U * V through the library's own GEMM entry (hcb_?gemm on the packed buffer: U ld m, V ld rank)."""
        rk = self.GetTileRank()
        out = torch.empty(self.n, self.m, dtype=self.dtype, device=self.buf.device)   # column-major m x n
        p = "d" if self.dtype == torch.float64 else "s"
        es = self.buf.element_size()
        torch.cuda.synchronize(self.buf.device)  # (buf may have been written on another torch stream)
        check(getattr(lib, f"hcb_{p}gemm")(self.ctx.h, 0, 0, self.m, self.n, rk, 1.0, self.buf.data_ptr(), self.m,
                                           self.buf.data_ptr() + es * self.m * self.max_rank, rk, 0.0, out.data_ptr(), self.m))
        torch.cuda.synchronize(self.buf.device)
        return np.asfortranarray(out.cpu().numpy().T)


def gemm_batched(alpha, A, opA, B, opB, beta, Cs, ctx: RunContext, params: CompressionParameters | None = None,
                 info: torch.Tensor | None = None):
    """One fused launch sequence for many tile triples: C[t] = alpha*op(A[t])*op(B[t]) + beta*C[t] (+ recompression)."""
    n = len(Cs)
    assert len(A) == n and len(B) == n
    params = params or CompressionParameters(1e-9)
    dtype = Cs[0].dtype
    da = (hcb_tile * n)(*[t.desc() for t in A])
    db = (hcb_tile * n)(*[t.desc() for t in B])
    dc = (hcb_tile * n)(*[t.desc() for t in Cs])
    prm = params.c()
    ct = _CT[dtype]
    check(_fn("tlr_gemm_batched", dtype)(ctx.h, n, da, int(bool(opA)), db, int(bool(opB)), dc, ct(alpha), ct(beta),
                                         C.byref(prm), None if info is None else info.data_ptr()))


class HCore:
    """hcorepp::api::HCore<T> -- static tile routines (only Gemm is on the hot path)."""

    @staticmethod
    def Gemm(alpha, A, opA, B, opB, beta, Ct, ctx: RunContext, params: CompressionParameters | None = None):
        """HCore<T>::Gemm(alpha, A, opA, B, opB, beta, C, context, flops, memoryUnit, params) (HCore.hpp:41-46).

        Dense*Dense -> Compressed leaves C full rank (U = alpha*A*B + beta*C, V = I; HCore.cpp:272-299): like the
        reference (DataHolder::Resize) the tile buffer is re-allocated when its capacity is too small."""
        if (A.dense and B.dense and not Ct.dense) and Ct.max_rank < min(Ct.m, Ct.n):
            rk = Ct.GetTileRank()
            U, V = Ct.factors()
            newcap = min(Ct.m, Ct.n)
            buf = torch.zeros(Ct.m * newcap + newcap * Ct.n, dtype=Ct.dtype, device=ctx.device)
            buf[: Ct.m * rk] = U.t().reshape(-1)
            buf[Ct.m * newcap: Ct.m * newcap + rk * Ct.n] = V.t().reshape(-1)
            Ct.buf, Ct.max_rank = buf, newcap
        gemm_batched(alpha, [A], opA, [B], opB, beta, [Ct], ctx, params)


def _chr(x):
    return ord(x) if isinstance(x, str) else int(x)


def _potrf(A: "DenseTile", uplo, ctx: RunContext):
    info = torch.zeros(1, dtype=torch.int32, device=ctx.device)
    check(_fn("potrf", A.dtype)(ctx.h, _chr(uplo), A.m, A.t.data_ptr(), A.m, info.data_ptr()))
    return info


def _hcore_potrf(A, uplo, ctx: RunContext):
    """HCore<T>::Potrf (HCore.cpp:586-621): dense tiles only, LAPACK potrf semantics in place (the other triangle is left
    as it was).  Returns LAPACK's info (0 = success) -- read from the device, so this synchronises."""
    if not A.dense:
        raise RuntimeError(" Potrf works only with dense tiles")
    return int(_potrf(A, uplo, ctx).item())


def _hcore_trsm(side, uplo, trans, diag, alpha, A, B, ctx: RunContext):
    """HCore<T>::Trsm (HCore.cpp:624-647): A dense triangular, B compressed; blas::trsm with m = rows(B), n = rank(B) on
    B's V buffer viewed as an (m x rank) matrix -- the reference's Cholesky convention stores V as n x k; U is not touched."""
    if B.dense:
        raise RuntimeError(" TRSM: Tile B must be compressed ")
    rk = B.GetTileRank()
    vptr = B.buf.data_ptr() + B.buf.element_size() * B.m * B.max_rank
    check(_fn("trsm", B.dtype)(ctx.h, _chr(side), _chr(uplo), int(bool(trans)), _chr(diag), B.m, rk, _CT[B.dtype](alpha),
                               A.t.data_ptr(), A.m, vptr, B.m))


def _hcore_syrk(alpha, A, opA, uplo, beta, Ct, ctx: RunContext):
    """HCore<T>::Syrk (HCore.cpp:484-583) with a dense A: blas::syrk with n = rows(C), k = cols(C); only the `uplo` triangle
    of C is referenced.  (A compressed A goes through tlr_syrk_batched / the Cholesky driver in this library's V = rank x n
    convention; the reference's own compressed branch is unpinned)"""
    if not A.dense or not Ct.dense:
        raise RuntimeError(" Syrk: dense tiles only in the reference-convention mirror (see tlr_cholesky for compressed A)")
    check(_fn("syrk", Ct.dtype)(ctx.h, _chr(uplo), int(bool(opA)), Ct.m, Ct.n, _CT[Ct.dtype](alpha), A.t.data_ptr(), A.m,
                                _CT[Ct.dtype](beta), Ct.t.data_ptr(), Ct.m))


class TileMatrix:
    """helpers::TileMatrix<T>: an mt x nt grid of tiles in ONE pooled device buffer (+ a device rank table and a
    pre-built descriptor array), so that a whole k-step is a single batched call with no per-tile host work."""

    def __init__(self, mt, nt, tm, tn, dtype, ctx: RunContext, compressed=True, max_rank=None, rank_bound=0):
        self.mt, self.nt, self.tm, self.tn, self.dtype, self.ctx = mt, nt, tm, tn, dtype, ctx
        self.compressed = compressed
        self.max_rank = (max_rank or max(min(tm, tn) // MAX_RANK_RATIO, 1)) if compressed else 0
        self.rank_bound = rank_bound
        self.tile_elems = (tm * self.max_rank + self.max_rank * tn) if compressed else tm * tn
        self.buf = torch.zeros(mt * nt * self.tile_elems, dtype=dtype, device=ctx.device)
        self.ranks = torch.ones(mt * nt, dtype=torch.int32, device=ctx.device)
        self.state = torch.zeros(mt * nt, dtype=torch.int32, device=ctx.device)  # per-tile state words (see hcb_tile.d_state)
        self._build_descs()

    def _build_descs(self):
        n = self.mt * self.nt
        esz = self.buf.element_size()
        base, rbase, sbase = self.buf.data_ptr(), self.ranks.data_ptr(), self.state.data_ptr()
        self.descs = (hcb_tile * n)()
        for lin in range(n):  # column-major grid: lin = j + i*mt  (TileMatrix.hpp:78-80)
            d = self.descs[lin]
            d.type = TILE_COMPRESSED if self.compressed else TILE_DENSE
            d.m, d.n, d.ld = self.tm, self.tn, self.tm
            d.max_rank, d.rank_bound = self.max_rank, self.rank_bound
            d.d_rank = rbase + 4 * lin if self.compressed else None
            d.d_state = sbase + 4 * lin if self.compressed else None
            d.fixed_rank = 0
            d.d_data = base + esz * self.tile_elems * lin

    def set_rank_bound(self, bound: int):
        self.rank_bound = int(bound)
        for d in self.descs:
            d.rank_bound = self.rank_bound

    def set_fixed_ranks(self, table):
        """Per-tile fixed ranks for the replay drivers (par_fixed_rank_streams_main.cpp:465-477): table[row][col] > 0
        makes the recompression of C(row, col) keep exactly that rank, 0 / None restores truncation by accuracy."""
        for col in range(self.nt):
            for row in range(self.mt):
                self.descs[self.lin(row, col)].fixed_rank = 0 if table is None else int(table[row][col])

    def lin(self, row, col):
        return row + col * self.mt

    def tile_buf(self, row, col):
        o = self.lin(row, col) * self.tile_elems
        return self.buf[o: o + self.tile_elems]

    def GetTile(self, row, col):
        """Tile<T>* GetTile(row, col) -- a view object sharing the pooled storage."""
        lin = self.lin(row, col)
        if self.compressed:
            return CompressedTile(self.tm, self.tn, self.max_rank, self.dtype, self.ctx, buf=self.tile_buf(row, col),
                                  rank=self.ranks[lin: lin + 1], rank_bound=self.rank_bound,
                                  state=self.state[lin: lin + 1])
        t = DenseTile.__new__(DenseTile)
        t.ctx, t.t, t.m, t.n, t.dtype = self.ctx, self.tile_buf(row, col).view(self.tn, self.tm), self.tm, self.tn, self.dtype
        return t

    # -- constructors (TileMatrix.hpp:40-41, 62-63)
    @classmethod
    def from_dense(cls, raw, tm, tn, ctx: RunContext, params: CompressionParameters | None = None):
        """TileMatrix(RawMatrix, rowTile, colTile[, params], ctx): dense tiles, or compressed (one SVD per tile,
        here ONE batched device call instead of the reference's serial loop, TileMatrix.cpp:150-171)."""
        raw_t = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw))
        raw_t = raw_t.to(ctx.device)
        M, N = raw_t.shape
        assert M % tm == 0 and N % tn == 0, "ragged edge tiles: build them tile by tile with CompressedTile/DenseTile"
        mt, nt = M // tm, N // tn
        dtype = raw_t.dtype
        # (mt, tm, nt, tn) -> (nt, mt, tn, tm): tile (j,i) column-major at lin = j + i*mt
        tiles = raw_t.view(mt, tm, nt, tn).permute(2, 0, 3, 1).contiguous()
        if params is None:
            tmx = cls(mt, nt, tm, tn, dtype, ctx, compressed=False)
            tmx.buf.copy_(tiles.reshape(-1))
            return tmx
        tmx = cls(mt, nt, tm, tn, dtype, ctx, compressed=True)
        n = mt * nt
        esz = tiles.element_size()
        ptrs = (C.c_void_p * n)(*[tiles.data_ptr() + esz * tm * tn * lin for lin in range(n)])
        prm = params.c()
        # per-tile info words of the compressing constructor (1: Jacobi not converged, 2: rank clipped at maxRank)
        tmx.compress_info = torch.zeros(n, dtype=torch.int32, device=ctx.device)
        check(_fn("compress_batched", dtype)(ctx.h, n, ptrs, tm, tmx.descs, C.byref(prm), C.c_void_p(tmx.compress_info.data_ptr())))
        ctx.Sync()
        return tmx

    @classmethod
    def zeros_compressed(cls, mt, nt, tm, tn, dtype, ctx: RunContext, rank_bound=0):
        """What the compressing constructor gives for an all-zero matrix (the drivers' C0, omp_main.cpp:243,355):
        rank-1 tiles whose product is zero, maxRank = min(tm,tn)/3."""
        return cls(mt, nt, tm, tn, dtype, ctx, compressed=True, rank_bound=rank_bound)

    def reset_to_zero(self):
        """Back to the drivers' C0: rank-1 tiles whose product is zero (only the live column/row is cleared)."""
        assert self.compressed
        v = self.buf.view(self.mt * self.nt, self.tile_elems)
        v[:, : self.tm].zero_()
        v[:, self.tm * self.max_rank: self.tm * self.max_rank + self.tn].zero_()
        self.ranks.fill_(1)
        self.state.zero_()

    def load_factors(self, U: torch.Tensor, V: torch.Tensor, rank: int):
        """Fill every tile from compact factor stacks: U (ntiles, rank, tm) [= column-major tm x rank], V (ntiles, tn,
        rank) [= column-major rank x tn], tile order lin = row + col*mt. Device-to-device or pinned-host-to-device."""
        v = self.buf.view(self.mt * self.nt, self.tile_elems)
        v[:, : self.tm * rank].copy_(U.reshape(self.mt * self.nt, -1), non_blocking=True)
        v[:, self.tm * self.max_rank: self.tm * self.max_rank + rank * self.tn].copy_(
            V.reshape(self.mt * self.nt, -1), non_blocking=True)
        self.ranks.fill_(rank)
        self.state.zero_()

    def rank_table(self) -> np.ndarray:
        return self.ranks.cpu().numpy().reshape(self.nt, self.mt).T.copy()

    def ToRawMatrix(self) -> np.ndarray:
        """RawMatrix ToRawMatrix(ctx) (TileMatrix.cpp:187-252): U*V per tile, assembled on the host."""
        out = np.zeros((self.mt * self.tm, self.nt * self.tn), dtype=np.float64 if self.dtype == torch.float64 else np.float32)
        for i in range(self.nt):
            for j in range(self.mt):
                out[j * self.tm:(j + 1) * self.tm, i * self.tn:(i + 1) * self.tn] = self.GetTile(j, i).to_dense()
        return out

    # -- wire / on-disk format: per tile the reference's TileMetadata (Tile.hpp:30-52) + the packed
    #    buffer UnPackTile hands out (Compressed.cpp:769-779), frozen here as a JSON manifest + one little-endian binary
    def save(self, path: str):
        """Persist the matrix for parity runs: `path`.json (manifest: grid, dtype, one TileMetadata record per tile with its
        byte offset) + `path`.bin (per compressed tile the LIVE factors [U (m x rank) | V (rank x n, ld = rank)] column-major,
        per dense tile the m x n block)."""
        import json
        ranks = self.ranks.cpu().numpy()
        esz = self.buf.element_size()
        tiles, off = [], 0
        with open(path + ".bin", "wb") as f:
            for lin in range(self.mt * self.nt):
                o = lin * self.tile_elems
                if self.compressed:
                    rk = int(ranks[lin])
                    u = self.buf[o: o + self.tm * rk]
                    v = self.buf[o + self.tm * self.max_rank: o + self.tm * self.max_rank + rk * self.tn]
                    blob = torch.cat([u, v]).cpu().numpy().tobytes()
                    meta = dict(mNumOfRows=self.tm, mNumOfCols=self.tn, mMatrixRank=rk, mMaxRank=self.max_rank,
                                mLeadingDimension=self.tm, mLayout="C", mType="COMPRESSED")
                else:
                    blob = self.buf[o: o + self.tile_elems].cpu().numpy().tobytes()
                    meta = dict(mNumOfRows=self.tm, mNumOfCols=self.tn, mMatrixRank=0, mMaxRank=0, mLeadingDimension=self.tm,
                                mLayout="C", mType="DENSE")
                f.write(blob)
                tiles.append(dict(row=lin % self.mt, col=lin // self.mt, offset=off, nbytes=len(blob), metadata=meta))
                off += len(blob)
        manifest = dict(format="hcorepp-b200 tile manifest v1", dtype={8: "f64", 4: "f32"}[esz], byteorder="little",
                        mt=self.mt, nt=self.nt, tile_rows=self.tm, tile_cols=self.tn, compressed=bool(self.compressed),
                        max_rank=self.max_rank, tiles=tiles)
        with open(path + ".json", "w") as f:
            json.dump(manifest, f)

    @classmethod
    def load(cls, path: str, ctx: RunContext):
        """Inverse of save(): every tile is re-packed from its (metadata, buffer) record (PackTile, Compressed.cpp:780-805)."""
        import json
        man = json.load(open(path + ".json"))
        assert man["format"].startswith("hcorepp-b200 tile manifest")
        dtype = torch.float64 if man["dtype"] == "f64" else torch.float32
        npdt = np.float64 if man["dtype"] == "f64" else np.float32
        tm = cls(man["mt"], man["nt"], man["tile_rows"], man["tile_cols"], dtype, ctx, compressed=man["compressed"],
                 max_rank=man["max_rank"] or None)
        raw = np.fromfile(path + ".bin", dtype=np.uint8)
        for t in man["tiles"]:
            lin = t["row"] + t["col"] * tm.mt
            a = torch.from_numpy(raw[t["offset"]: t["offset"] + t["nbytes"]].view(npdt).copy()).to(ctx.device)
            o = lin * tm.tile_elems
            md = t["metadata"]
            if md["mType"] == "COMPRESSED":
                rk, m, n = md["mMatrixRank"], md["mNumOfRows"], md["mNumOfCols"]
                tm.buf[o: o + m * rk] = a[: m * rk]
                tm.buf[o + m * tm.max_rank: o + m * tm.max_rank + rk * n] = a[m * rk:]
                tm.ranks[lin] = rk
            else:
                tm.buf[o: o + tm.tile_elems] = a
        tm.state.zero_()
        return tm

    def GetMemoryFootprint(self) -> int:
        """bytes actually holding data (TileMatrix.cpp:176-184)."""
        esz = self.buf.element_size()
        if not self.compressed:
            return self.mt * self.nt * self.tm * self.tn * esz
        return int(self.ranks.sum().item()) * (self.tm + self.tn) * esz


def tile_matrix_multiplication(A: TileMatrix, B: TileMatrix, Cm: TileMatrix, alpha, beta, ctx: RunContext,
                               params: CompressionParameters, owned=None, info: torch.Tensor | None = None,
                               k_range=None):
    """for i, j: for k: HCore::Gemm(A(j,k), B(k,i), C(j,i)) (omp_main.cpp:112-126) -- all owned (j,i) of one k in one
    batched call, k sequential.  `owned`: optional list of linear C indices (j + i*mt) this rank owns."""
    assert A.nt == B.mt and A.mt == Cm.mt and B.nt == Cm.nt
    ct = _CT[Cm.dtype]
    prm = params.c()
    if owned is None:
        o_ptr, n_owned = None, 0
    else:
        n_owned = len(owned)
        o_ptr = (C.c_int64 * n_owned)(*owned)
    k0, k1 = k_range if k_range is not None else (0, A.nt)
    check(_fn("tlr_matmul", Cm.dtype)(ctx.h, Cm.mt, Cm.nt, A.nt, A.descs, B.descs, Cm.descs, o_ptr, n_owned, k0, k1,
                                      ct(alpha), ct(beta), C.byref(prm), None if info is None else info.data_ptr()))


HCore.Potrf = staticmethod(_hcore_potrf)
HCore.Trsm = staticmethod(_hcore_trsm)
HCore.Syrk = staticmethod(_hcore_syrk)


class SymTileMatrix:
    """A symmetric positive definite matrix in tile-low-rank form for the Cholesky driver (BASELINE.json configs[4]):
    nt dense nb x nb DIAGONAL tiles (one pooled buffer, column-major each) + compressed tiles strictly BELOW the diagonal
    (a pooled nt x nt TileMatrix of which the entries i > j are used).  The reference has the tile routines but neither
    this container nor a driver."""

    def __init__(self, nt, nb, dtype, ctx: RunContext, max_rank=None, rank_bound=0):
        self.nt, self.nb, self.dtype, self.ctx = nt, nb, dtype, ctx
        self.diag = torch.zeros(nt, nb * nb, dtype=dtype, device=ctx.device)
        self.low = TileMatrix(nt, nt, nb, nb, dtype, ctx, compressed=True, max_rank=max_rank, rank_bound=rank_bound)

    @classmethod
    def from_tiles(cls, tile_fn, nt, nb, dtype, ctx: RunContext, params: CompressionParameters, chunk=64, max_rank=None):
        """tile_fn(i, j) -> dense nb x nb block (torch tensor on the device or numpy, A[i-block, j-block]); diagonal blocks
        are stored dense, blocks below the diagonal are compressed on the device in batches (hcb_?compress_batched)."""
        S = cls(nt, nb, dtype, ctx, max_rank=max_rank)
        dev = ctx.device
        as_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev).to(dtype)
        for k in range(nt):
            S.diag[k] = as_t(tile_fn(k, k)).t().contiguous().reshape(-1)
        pairs = [(i, j) for j in range(nt) for i in range(j + 1, nt)]
        prm = params.c()
        for c0 in range(0, len(pairs), chunk):
            part = pairs[c0:c0 + chunk]
            dense = torch.stack([as_t(tile_fn(i, j)).t().contiguous() for (i, j) in part])  # column-major tiles
            n = len(part)
            ptrs = (C.c_void_p * n)(*[dense.data_ptr() + dense.element_size() * nb * nb * q for q in range(n)])
            descs = (hcb_tile * n)(*[S.low.descs[S.low.lin(i, j)] for (i, j) in part])
            check(_fn("compress_batched", dtype)(ctx.h, n, ptrs, nb, descs, C.byref(prm), None))
            ctx.Sync()
        return S

    def lower_factor_dense(self) -> np.ndarray:
        """the lower-triangular factor L (after tlr_cholesky) / the lower triangle of A (before) as one dense matrix"""
        nt, nb = self.nt, self.nb
        out = np.zeros((nt * nb, nt * nb))
        d = self.diag.cpu().numpy()
        for k in range(nt):
            out[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = np.tril(d[k].reshape(nb, nb).T)
        for j in range(nt):
            for i in range(j + 1, nt):
                out[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = self.low.GetTile(i, j).to_dense()
        return out


def tlr_cholesky(S: SymTileMatrix, ctx: RunContext, params: CompressionParameters, info: torch.Tensor | None = None,
                 potrf_info: torch.Tensor | None = None):
    """Right-looking tile Cholesky A = L L^T in place (hcb_?tlr_potrf): per step k -- potrf of the dense diagonal tile, the
    panel solve A(i,k) := A(i,k) L_kk^-T on the V factors of the block column, the symmetric updates of the diagonal
    tiles, and ONE batched recompressing update A(i,j) -= A(i,k) A(j,k)^T for all i > j > k (HCore::Gemm with
    opB = Trans).  Asynchronous; potrf_info (int32[nt], optional) receives LAPACK's info per diagonal tile."""
    nt, nb = S.nt, S.nb
    esz = S.diag.element_size()
    dptr = (C.c_void_p * nt)(*[S.diag.data_ptr() + esz * nb * nb * k for k in range(nt)])
    prm = params.c()
    check(_fn("tlr_potrf", S.dtype)(ctx.h, nt, nb, dptr, nb, S.low.descs, C.byref(prm),
                                    None if info is None else info.data_ptr(),
                                    None if potrf_info is None else potrf_info.data_ptr()))
