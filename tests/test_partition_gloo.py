"""CPU tests of the multi-GPU host logic: 2D block-cyclic ownership + the per-k panel broadcast schedule,
exercised with a world_size-2 (and 4) gloo process group on CPU tensors -- no CUDA involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hcorepp_b200 import partition as part


def test_ownership_maps_cover_everything_once():
    for n in (1, 2, 4, 8):
        P, Q = part.grid_shape(n)
        assert P * Q == n and P <= Q
        mt, nt, kt = 2 * P + 1, 3 * Q, 5
        owned = [part.owned_c_tiles(r, mt, nt, P, Q) for r in range(n)]
        flat = sorted(x for o in owned for x in o)
        assert flat == list(range(mt * nt))  # every C tile has exactly one owner
        for r in range(n):
            for lin in owned[r]:
                j, i = lin % mt, lin // mt
                assert part.owner_of_c(j, i, P, Q) == r
        for k in range(kt):
            for r in range(n):
                pr, pc = part.grid_pos(r, P, Q)
                ra, rb = part.panel_schedule(kt, r, P, Q)[k]
                # the A panel root sits in my grid row and owns A(j, k) for my rows; same for B in my grid column
                assert part.grid_pos(ra, P, Q)[0] == pr and part.owner_of_a(pr, k, P, Q) == ra
                assert part.grid_pos(rb, P, Q)[1] == pc and part.owner_of_b(k, pc, P, Q) == rb


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mt_l, nt_l, kt, out):
    """Each rank owns synthetic 'tiles' (a single float = global tile id); per k the owners broadcast their panels along
    grid rows / columns exactly as bench.py does; every rank then checks it holds the tiles its C tiles need."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, Q = part.grid_shape(world)
    pr, pc = part.grid_pos(rank, P, Q)
    row_groups = [dist.new_group([r * Q + c for c in range(Q)]) for r in range(P)]
    col_groups = [dist.new_group([r * Q + c for r in range(P)]) for c in range(Q)]
    kA, kB = part.owned_indices(kt, Q, pc), part.owned_indices(kt, P, pr)
    a_id = lambda j, k: 1000.0 * j + k          # global A(j, k) id
    b_id = lambda k, i: -(1000.0 * k + i) - 1   # global B(k, i) id
    # local storage: A_loc[kl][jl] with global j = jl*P + pr, k = kA[kl]; B_loc[kl][il] with global i = il*Q + pc, k = kB[kl]
    A_loc = torch.tensor([[a_id(jl * P + pr, k) for jl in range(mt_l)] for k in kA], dtype=torch.float64).reshape(len(kA), mt_l)
    B_loc = torch.tensor([[b_id(k, il * Q + pc) for il in range(nt_l)] for k in kB], dtype=torch.float64).reshape(len(kB), nt_l)
    ok = True
    sched = part.panel_schedule(kt, rank, P, Q)
    for k in range(kt):
        ra, rb = sched[k]
        pa = A_loc[k // Q].clone() if ra == rank else torch.zeros(mt_l, dtype=torch.float64)
        pb = B_loc[k // P].clone() if rb == rank else torch.zeros(nt_l, dtype=torch.float64)
        if Q > 1:
            dist.broadcast(pa, src=ra, group=row_groups[pr])
        if P > 1:
            dist.broadcast(pb, src=rb, group=col_groups[pc])
        for jl in range(mt_l):
            ok &= pa[jl].item() == a_id(jl * P + pr, k)
        for il in range(nt_l):
            ok &= pb[il].item() == b_id(k, il * Q + pc)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_panel_broadcast_schedule_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 3, 2, 5, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world


def _worker_lib(rank, world, port, mt, nt, kt, out):
    """The LIBRARY's exchange step (hcorepp_b200.distributed.Grid2D + exchange_panels, the code bench.py and
    tlr_matmul_distributed run with NCCL) under gloo on CPU tensors: a 'tile' is two floats (payload id, rank id); after
    step k every rank must hold A(j, k) for its C rows and B(k, i) for its C columns, in local order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hcorepp_b200 import distributed as D
    g = D.Grid2D()
    assert (g.P, g.Q) == part.grid_shape(world) and g.rank == rank
    rows, cols = part.owned_indices(mt, g.P, g.pr), part.owned_indices(nt, g.Q, g.pc)
    kA, kB = part.owned_indices(kt, g.Q, g.pc), part.owned_indices(kt, g.P, g.pr)
    a_id = lambda j, k: 1000.0 * j + k
    b_id = lambda k, i: -(1000.0 * k + i) - 1
    A_buf = [torch.tensor([a_id(j, k) for j in rows], dtype=torch.float64) for k in kA]
    A_rk = [torch.tensor([int(a_id(j, k)) % 97 for j in rows], dtype=torch.int32) for k in kA]
    B_buf = [torch.tensor([b_id(k, i) for i in cols], dtype=torch.float64) for k in kB]
    B_rk = [torch.tensor([int(-b_id(k, i)) % 89 for i in cols], dtype=torch.int32) for k in kB]
    pan_a = (torch.zeros(len(rows), dtype=torch.float64), torch.zeros(len(rows), dtype=torch.int32))
    pan_b = (torch.zeros(len(cols), dtype=torch.float64), torch.zeros(len(cols), dtype=torch.int32))
    ok = True
    for k in range(kt):
        D.exchange_panels(g, k, lambda kl: (A_buf[kl], A_rk[kl]), lambda kl: (B_buf[kl], B_rk[kl]), pan_a, pan_b,
                          len(rows) > 0, len(cols) > 0)
        ok &= pan_a[0].tolist() == [a_id(j, k) for j in rows] and pan_a[1].tolist() == [int(a_id(j, k)) % 97 for j in rows]
        ok &= pan_b[0].tolist() == [b_id(k, i) for i in cols] and pan_b[1].tolist() == [int(-b_id(k, i)) % 89 for i in cols]
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (5, 4, 7)), (4, (6, 7, 5)), (4, (3, 8, 9))])
def test_library_panel_exchange_gloo(world, shape):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_lib, args=(world, port, *shape, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world
