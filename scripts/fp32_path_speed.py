"""FP32 tiles through the fused path: promoted (FP64 shadows, round 2 default) vs the native FP32 kernels (HCB_FP32_NATIVE=1).
8 x 8 tiles of 1024, rank-44 inputs with the reference spectrum law, accuracy 1e-4, full k-sum (512 tile-GEMMs)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hcorepp_b200 as hc  # noqa: E402


def run(native):
    if native:
        os.environ["HCB_FP32_NATIVE"] = "1"
    else:
        os.environ.pop("HCB_FP32_NATIVE", None)
    ctx = hc.RunContext(0)
    T, nb, k, acc = int(os.environ.get("FPT", 8)), 1024, int(os.environ.get("FPK", 44)), 1e-4
    dt = torch.float32
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    sig = torch.logspace(0, -6, k, device="cuda", dtype=torch.float64)

    def synth(n):
        qu, _ = torch.linalg.qr(torch.randn(n, nb, k, generator=g, dtype=torch.float64, device="cuda"))
        qv, _ = torch.linalg.qr(torch.randn(n, nb, k, generator=g, dtype=torch.float64, device="cuda"))
        return qu.transpose(1, 2).contiguous().to(dt), (qv * sig[None, None, :]).contiguous().to(dt)
    A = hc.TileMatrix(T, T, nb, nb, dt, ctx, compressed=True, max_rank=k, rank_bound=k)
    B = hc.TileMatrix(T, T, nb, nb, dt, ctx, compressed=True, max_rank=k, rank_bound=k)
    A.load_factors(*synth(T * T), k)
    B.load_factors(*synth(T * T), k)
    C = hc.TileMatrix.zeros_compressed(T, T, nb, nb, dt, ctx)
    prm = hc.CompressionParameters(acc)
    hc.tile_matrix_multiplication(A, B, C, 1.0, 1.0, ctx, prm)       # calibration pass: how far do the ranks grow
    ctx.Sync()
    C.set_rank_bound(int(min(C.max_rank, (int(C.ranks.max().item()) + 8 + 7) // 8 * 8)))
    for it in range(3):
        C.reset_to_zero()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hc.tile_matrix_multiplication(A, B, C, 1.0, 1.0, ctx, prm)
        ctx.Sync()
        dtm = time.perf_counter() - t0
    rk = C.rank_table()
    d = C.GetTile(0, 0).to_dense().astype(np.float64)
    return dtm, float(rk.mean()), d


if __name__ == "__main__":
    tp, rp, dp = run(False)
    tn, rn, dn = run(True)
    print("promoted: %.1f ms per pass (%.0f tile-GEMM/s), mean rank %.1f" % (tp * 1e3, (int(os.environ.get("FPT", 8)) ** 3) / tp, rp))
    print("native  : %.1f ms per pass (%.0f tile-GEMM/s), mean rank %.1f" % (tn * 1e3, (int(os.environ.get("FPT", 8)) ** 3) / tn, rn))
    print("tile (0,0): ||promoted - native|| / ||native|| = %.2e" % (np.linalg.norm(dp - dn) / np.linalg.norm(dn)))
