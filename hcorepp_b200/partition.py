"""2D block-cyclic ownership of tiles over a P x Q grid of GPUs -- pure host logic, no torch, no CUDA.

  C(j, i) lives on grid position (j mod P, i mod Q)        (the k-sum of one C tile is sequential, so no split-k)
  A(j, k) lives on               (j mod P, k mod Q)        -> row-panel A(:, k) is broadcast along grid rows
  B(k, i) lives on               (k mod P, i mod Q)        -> column-panel B(k, :) is broadcast along grid columns
The reference has no distributed layer at all (its parallel axis is OpenMP over C tiles, omp_main.cpp:112-113).
"""
from __future__ import annotations

import math


def grid_shape(n_ranks: int):
    """P x Q with P <= Q and P*Q == n_ranks: 1x1, 1x2, 2x2, 2x4, ..."""
    p = int(math.sqrt(n_ranks))
    while n_ranks % p:
        p -= 1
    return p, n_ranks // p


def grid_pos(rank: int, P: int, Q: int):
    return rank // Q, rank % Q


def owned_indices(n: int, parts: int, idx: int):
    """Indices in range(n) that the `idx`-th of `parts` cyclic owners holds."""
    return [k for k in range(n) if k % parts == idx]


def owner_of_c(j: int, i: int, P: int, Q: int) -> int:
    return (j % P) * Q + (i % Q)


def owner_of_a(j: int, k: int, P: int, Q: int) -> int:
    return (j % P) * Q + (k % Q)


def owner_of_b(k: int, i: int, P: int, Q: int) -> int:
    return (k % P) * Q + (i % Q)


def owned_c_tiles(rank: int, mt: int, nt: int, P: int, Q: int):
    """Linear (column-major, j + i*mt) indices of the C tiles `rank` owns."""
    pr, pc = grid_pos(rank, P, Q)
    return [j + i * mt for i in range(nt) for j in range(mt) if j % P == pr and i % Q == pc]


def panel_schedule(kt: int, rank: int, P: int, Q: int):
    """For every k: (root of the A row-panel broadcast inside my grid row, root of the B column-panel broadcast inside
    my grid column), as global ranks."""
    pr, pc = grid_pos(rank, P, Q)
    return [(pr * Q + (k % Q), (k % P) * Q + pc) for k in range(kt)]
