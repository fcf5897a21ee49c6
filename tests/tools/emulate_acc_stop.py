"""CPU emulation (numpy, test infrastructure): what does an ACCURACY-AWARE stop of the one-sided Jacobi core SVD cost in
parity over a whole k-sum?  The oracle's recompression is run twice on the same seeded tiles -- once with LAPACK's SVD
(the reference's algorithm), once with a round-robin one-sided Jacobi that stops after the first sweep whose largest
|cos| satisfies cos^2 < accuracy, V Sigma derived as M^T U like the CUDA path -- and the final C tiles are compared.
usage: python tests/tools/emulate_acc_stop.py [nb] [rank] [ksteps] [acc]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import tlr_oracle as O  # noqa: E402


def rr_pairs(n2, rnd):
    mod = n2 - 1
    out = []
    for slot in range(n2 // 2):
        x, y = (mod, rnd % mod) if slot == 0 else ((rnd + slot) % mod, (rnd - slot + mod) % mod)
        out.append((min(x, y), max(x, y)))
    return out


def make_svd(stop2, log):
    def svd(A, svd="gesdd"):
        M = np.array(A, dtype=np.float64, order="F")
        a, b = M.shape
        W = M.copy()
        tol = np.sqrt(a) * 2.22e-16
        n2 = (b + 1) & ~1
        for sweep in range(60):
            maxc = 0.0
            for rnd in range(n2 - 1):
                ps = [(x, y) for x, y in rr_pairs(n2, rnd) if y < b]
                X, Y = np.array([p[0] for p in ps]), np.array([p[1] for p in ps])
                U, V = W[:, X], W[:, Y]
                al, be, ga = (U * U).sum(0), (V * V).sum(0), (U * V).sum(0)
                cosv = np.abs(ga) / np.sqrt(np.maximum(al * be, 1e-300))
                rot = cosv > tol
                maxc = max(maxc, float(cosv.max()))
                g = np.where(rot, ga, 1.0)
                zeta = (be - al) / (2 * g)
                t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = np.where(rot, c * t, 0.0)
                c = np.where(rot, c, 1.0)
                W[:, X], W[:, Y] = c * U - s * V, s * U + c * V
            if maxc * maxc < max(stop2, 16 * tol):
                break
        log.append(sweep + 1)
        sig = np.linalg.norm(W, axis=0)
        order = np.argsort(-sig, kind="stable")
        sig = sig[order]
        Us = W[:, order] / np.where(sig > 0, sig, 1)
        VS = M.T @ Us                                  # V Sigma = M^T U (what the CUDA path stores)
        VT = (VS / np.where(sig > 0, sig, 1)).T
        return np.asfortranarray(Us), sig, np.asfortranarray(VT)
    return svd


def run(nb, rank, ksteps, acc, svd_fn):
    saved = O.k_svd
    if svd_fn is not None:
        O.k_svd = svd_fn
    try:
        p = O.CompressionParameters(acc)
        C = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
        ranks = []
        for k in range(ksteps):
            A, B = O.synth_compressed_tile(nb, rank, 100 + k), O.synth_compressed_tile(nb, rank, 200 + k)
            O.hcore_gemm(1.0, A, False, B, False, 1.0, C, p)
            ranks.append(C.rank)
        return C.to_dense(), ranks
    finally:
        O.k_svd = saved


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    ksteps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    acc = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
    ref, r0 = run(nb, rank, ksteps, acc, None)
    for name, stop2 in (("machine precision", 0.0), ("cos^2 < acc", acc)):
        log = []
        got, r1 = run(nb, rank, ksteps, acc, make_svd(stop2, log))
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"{name:18s}: rel Frobenius vs LAPACK oracle {err:.2e} (gate {10 * acc:.0e}), ranks {r1} vs {r0}, "
              f"max rank diff {max(abs(a - b) for a, b in zip(r0, r1))}, Jacobi sweeps per step {log}")
