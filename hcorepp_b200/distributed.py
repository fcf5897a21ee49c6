"""Multi-GPU TLR matrix product: one process per GPU, C tiles 2D block-cyclic over a P x Q grid, A row-panels and B
column-panels moved by NCCL broadcast over NVLink (BASELINE.json configs[3]).

The reference has no distributed layer (its parallel axis is OpenMP over C tiles, omp_main.cpp:112-113); this module is
the multi-GPU form of that driver loop.  Every tile (r, c) of every matrix lives on grid position (r mod P, c mod Q):
    C(j, i) on (j mod P, i mod Q)      A(j, k) on (j mod P, k mod Q)      B(k, i) on (k mod P, i mod Q)
so for step k the ranks of grid column k mod Q hold the row-panel A(:, k) -- each broadcasts ITS rows along its grid
row -- and the ranks of grid row k mod P hold B(k, :) and broadcast along their grid column.  A C tile's k-sum is
sequential (every step recompresses), so there is no split-k and no reduction: the panel broadcast is the only
data-path collective.  Transfers run on a side stream, double buffered against the batched recompression of the
previous step; the local step itself is ONE C-ABI call (hcb_?tlr_matmul_panel_step).

torch.distributed is plumbing here (process group, NCCL broadcast); all arithmetic is in libhcore_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import partition as part
from ._capi import check, lib
from .api import CompressionParameters, RunContext, TileMatrix, _CT, _PFX


class Grid2D:
    """P x Q process grid over the default torch.distributed group (or a single process when it is not initialised)."""

    def __init__(self, P: int | None = None, Q: int | None = None):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        if P is None or Q is None:
            P, Q = part.grid_shape(self.world)
        assert P * Q == self.world, f"grid {P}x{Q} does not match world size {self.world}"
        self.P, self.Q = P, Q
        self.pr, self.pc = part.grid_pos(self.rank, P, Q)
        self.row_group = self.col_group = None
        if self.dist and self.world > 1:
            # every rank creates every group (torch.distributed requires it), and keeps its own
            rows = [self.dist.new_group([r * Q + c for c in range(Q)]) for r in range(P)]
            cols = [self.dist.new_group([r * Q + c for r in range(P)]) for c in range(Q)]
            self.row_group, self.col_group = rows[self.pr], cols[self.pc]

    def owner(self, r: int, c: int) -> int:
        return (r % self.P) * self.Q + (c % self.Q)


class DistTileMatrix:
    """An mt x nt grid of (compressed) tiles distributed 2D block-cyclically; the local tiles live in one pooled
    TileMatrix.  `panel_rows=False` (A, C): local storage is column-major over (local row, local col), so a COLUMN
    panel (all local rows of one column) is one contiguous slab.  `panel_rows=True` (B): the local grid is stored
    transposed, so a ROW panel B(k, :) is one contiguous slab."""

    def __init__(self, mt, nt, tm, tn, dtype, ctx: RunContext, grid: Grid2D, panel_rows=False, max_rank=None,
                 rank_bound=0):
        self.mt, self.nt, self.tm, self.tn, self.dtype, self.ctx, self.grid = mt, nt, tm, tn, dtype, ctx, grid
        self.rows = part.owned_indices(mt, grid.P, grid.pr)
        self.cols = part.owned_indices(nt, grid.Q, grid.pc)
        self.panel_rows = panel_rows
        a, b = (len(self.cols), len(self.rows)) if panel_rows else (len(self.rows), len(self.cols))
        self.local = TileMatrix(max(a, 1), max(b, 1), tm, tn, dtype, ctx, compressed=True, max_rank=max_rank,
                                rank_bound=rank_bound)
        self.max_rank = self.local.max_rank

    def local_index(self, r: int, c: int):
        """(local row, local col) inside self.local of global tile (r, c); the caller must own it."""
        lr, lc = r // self.grid.P, c // self.grid.Q
        return (lc, lr) if self.panel_rows else (lr, lc)

    def owns(self, r: int, c: int) -> bool:
        return r % self.grid.P == self.grid.pr and c % self.grid.Q == self.grid.pc

    def GetTile(self, r: int, c: int):
        assert self.owns(r, c), f"tile ({r}, {c}) lives on rank {self.grid.owner(r, c)}"
        return self.local.GetTile(*self.local_index(r, c))

    def global_coords(self):
        """[(r, c)] of the local tiles in the order of self.local's linear index."""
        out = []
        if self.panel_rows:
            for lr, r in enumerate(self.rows):
                for lc, c in enumerate(self.cols):
                    out.append((r, c))      # linear = lc + lr * len(cols)
        else:
            for lc, c in enumerate(self.cols):
                for lr, r in enumerate(self.rows):
                    out.append((r, c))      # linear = lr + lc * len(rows)
        return out

    def panel_slab(self, p_local: int):
        """(buffer slab, rank slab) of local panel p_local: local column p_local of self.local."""
        n = self.local.mt
        e = self.local.tile_elems
        return self.local.buf[p_local * n * e:(p_local + 1) * n * e], self.local.ranks[p_local * n:(p_local + 1) * n]


def exchange_panels(g: Grid2D, k: int, a_slab, b_slab, pan_a, pan_b, have_rows=True, have_cols=True):
    """The data-path collective of step k (device-agnostic: NCCL on the GPUs, gloo in the CPU tests).  a_slab(kl) /
    b_slab(kl) return this rank's (factor slab, rank slab) of its kl-th local A column panel / B row panel; pan_a /
    pan_b = (buffer, ranks) receive the panels of step k: A(rows of this grid row, k) from grid column k mod Q along the
    grid row, B(k, columns of this grid column) from grid row k mod P along the grid column."""
    dist = g.dist
    src_c, src_r = k % g.Q, k % g.P
    if have_rows:
        if src_c == g.pc:
            buf, rk = a_slab(k // g.Q)
            pan_a[0].copy_(buf, non_blocking=True)
            pan_a[1].copy_(rk, non_blocking=True)
        if g.Q > 1:
            dist.broadcast(pan_a[0], src=g.pr * g.Q + src_c, group=g.row_group)
            dist.broadcast(pan_a[1], src=g.pr * g.Q + src_c, group=g.row_group)
    if have_cols:
        if src_r == g.pr:
            buf, rk = b_slab(k // g.P)
            pan_b[0].copy_(buf, non_blocking=True)
            pan_b[1].copy_(rk, non_blocking=True)
        if g.P > 1:
            dist.broadcast(pan_b[0], src=src_r * g.Q + g.pc, group=g.col_group)
            dist.broadcast(pan_b[1], src=src_r * g.Q + g.pc, group=g.col_group)


class MatmulPlan:
    """Buffers, streams and descriptor arrays of C = alpha * A * B + beta * C on a grid; reusable across calls."""

    def __init__(self, A: DistTileMatrix, B: DistTileMatrix, Cm: DistTileMatrix, ctx: RunContext):
        g = Cm.grid
        assert A.grid is g and B.grid is g
        assert A.nt == B.mt and A.mt == Cm.mt and B.nt == Cm.nt and not A.panel_rows and B.panel_rows and not Cm.panel_rows
        self.A, self.B, self.Cm, self.ctx, self.grid = A, B, Cm, ctx, g
        self.kt = A.nt
        dev = ctx.device
        self.mt_l, self.nt_l = len(Cm.rows), len(Cm.cols)
        mk = lambda n, src: TileMatrix(max(n, 1), 1, src.tm, src.tn, src.dtype, ctx, compressed=True, max_rank=src.max_rank,
                                       rank_bound=src.local.rank_bound)
        # two panel buffers per operand (double buffering); with one process they alias the local tiles (no copies)
        self.single = g.world == 1
        if not self.single:
            self.panA = [mk(self.mt_l, A), mk(self.mt_l, A)]
            self.panB = [mk(self.nt_l, B), mk(self.nt_l, B)]
            self.comm = torch.cuda.Stream(device=dev)
            self.ready = [torch.cuda.Event(), torch.cuda.Event()]
            self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.fn = getattr(lib, f"hcb_{_PFX[Cm.dtype]}tlr_matmul_panel_step")

    def _fetch(self, k: int, b: int):
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.done[b])
            exchange_panels(self.grid, k, self.A.panel_slab, self.B.panel_slab, (self.panA[b].buf, self.panA[b].ranks),
                            (self.panB[b].buf, self.panB[b].ranks), self.mt_l > 0, self.nt_l > 0)
            self.ready[b].record(self.comm)

    def run(self, alpha, beta, params: CompressionParameters, info: torch.Tensor | None = None, k_range=None):
        """for k: C(j, i) += A(j, k) * B(k, i) for the local (j, i); info (int32[local C tiles], optional) is sticky over k."""
        g, ctx = self.grid, self.ctx
        k0, k1 = k_range if k_range is not None else (0, self.kt)
        ct = _CT[self.Cm.dtype]
        prm = params.c()
        iptr = None if info is None else info.data_ptr()
        n = self.mt_l * self.nt_l
        if n == 0 and self.single:
            return
        main = torch.cuda.current_stream(ctx.device)
        step = lambda pa, ao, pb, bo, first: check(self.fn(
            ctx.h, self.mt_l, self.nt_l, C.cast(pa, C.c_void_p).value + ao, C.cast(pb, C.c_void_p).value + bo,
            self.Cm.local.descs, ct(alpha), ct(beta), C.byref(prm), iptr, int(first)))
        tsz = C.sizeof(self.Cm.local.descs._type_)
        if self.single:
            for k in range(k0, k1):
                step(self.A.local.descs, tsz * k * self.mt_l, self.B.local.descs, tsz * k * self.nt_l, k == k0)
            return
        self.done[0].record(main)
        self.done[1].record(main)
        if k0 < k1:
            self._fetch(k0, 0)
        for k in range(k0, k1):
            b = (k - k0) & 1
            if k + 1 < k1:
                self._fetch(k + 1, b ^ 1)
            main.wait_event(self.ready[b])
            if n:
                step(self.panA[b].descs, 0, self.panB[b].descs, 0, k == k0)
            self.done[b].record(main)


def tlr_matmul_distributed(A: DistTileMatrix, B: DistTileMatrix, Cm: DistTileMatrix, alpha, beta, ctx: RunContext,
                           params: CompressionParameters, info: torch.Tensor | None = None, k_range=None):
    """C = alpha * A * B + beta * C on the process grid (the multi-GPU form of tile_matrix_multiplication,
    examples/matrix_multiplication/omp_main.cpp:112-126).  The plan (panel buffers, streams) is cached on C."""
    plan = getattr(Cm, "_plan", None)
    if plan is None or plan.A is not A or plan.B is not B:
        plan = Cm._plan = MatmulPlan(A, B, Cm, ctx)
    plan.run(alpha, beta, params, info=info, k_range=k_range)
    return plan
