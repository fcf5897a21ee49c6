"""CPU emulation (numpy, test infrastructure) of the INCREMENTAL recompression used by the CUDA path in round 2.

After any recompression the C tile is in "SVD form": CU has orthonormal columns and CV = diag(sigma) W^T with W
orthonormal (Compressed.cpp:558-560, 598-622).  The next update C := beta*C + P Y^T then only needs the NEW columns
orthogonalised against the old bases (block classical Gram-Schmidt, twice):
    P = CU Gu + Q2u R2u ,   Y = W Gv + Q2v R2v
    [CU | P] = [CU | Q2u] RU,  RU = [[I, Gu], [0, R2u]] ;  [beta W S | Y] = [W | Q2v] RV,  RV = [[beta S, Gv], [0, R2v]]
    core K = RU RV^T = diag(beta S, 0) + [Gu; R2u] [Gv; R2v]^T          (a rank-kp update of a diagonal matrix)
    K = Us S' Vs^T (one-sided Jacobi: left vectors, V S' = K^T Us), truncate, CU' = [CU | Q2u] Us, CV'^T = [W | Q2v] (Vs S')
Both decompositions are exact identities whatever the orthogonality of CU / W, so the represented matrix only differs
from the full-QR path by the truncation decision (multiplicative perturbation of sigma by ||Q^T Q - I||).

The script runs a k-sum three ways -- LAPACK oracle (the reference's algorithm), full-QR + Jacobi, incremental + Jacobi
-- and prints parity (rel. Frobenius, ranks), the loss of orthogonality of CU / W over the steps and the Jacobi sweep
counts for different column orders of K.
usage: python tests/tools/emulate_incremental.py [nb] [rank] [ksteps] [acc]"""
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


def jacobi_left(M, stop2=0.0):
    """one-sided Jacobi on the columns of M: returns (Us normalised, sigma descending, sweeps)."""
    W = np.array(M, dtype=np.float64, order="F")
    a, b = W.shape
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
        if maxc < 4 * np.sqrt(tol) or maxc * maxc < stop2:   # predictive stop of the CUDA kernel / accuracy-aware stop
            break
    sig = np.linalg.norm(W, axis=0)
    order = np.argsort(-sig, kind="stable")
    sig = sig[order]
    Us = W[:, order] / np.where(sig > 0, sig, 1)
    return Us, sig, sweep + 1


def product_term(A, B, alpha):
    """P (m x ka), Y (n x ka) with alpha*A*B = P Y^T and mutually orthogonal Y columns (the CUDA preconditioner)."""
    T1 = A.V @ B.U                               # ka x kb
    T2 = T1 @ B.V                                # ka x n
    J, _, _ = np.linalg.svd(T2, full_matrices=False)   # J^T T2 has orthogonal rows
    return alpha * (A.U @ J), (J.T @ T2).T


def new_rank(sig, acc):
    for i in range(1, len(sig)):
        if sig[i] < acc:
            return i
    return len(sig)


def step_full(CU, CV, P, Y, beta, acc, maxrank, order="vnorm", stop2=0.0):
    SU, SV = np.hstack([CU, P]), np.hstack([beta * CV.T, Y])
    if order == "vnorm":
        perm = np.argsort(-np.linalg.norm(SV, axis=0), kind="stable")
        SU, SV = SU[:, perm], SV[:, perm]
    QU, RU = np.linalg.qr(SU)
    QV, RV = np.linalg.qr(SV)
    K = RU @ RV.T
    Us, sig, sw = jacobi_left(K, stop2)
    VS = K.T @ Us
    rk = min(new_rank(sig, acc), maxrank)
    return QU @ Us[:, :rk], (QV @ VS[:, :rk]).T, sw


def step_incremental(CU, CV, P, Y, beta, acc, maxrank, order="knorm", stop2=0.0):
    kc = CU.shape[1]
    sig_old = np.linalg.norm(CV, axis=1)
    Gu = CU.T @ P
    P1 = P - CU @ Gu
    G2 = CU.T @ P1
    P2 = P1 - CU @ G2
    Gu += G2
    Q2u, R2u = np.linalg.qr(P2)
    H = CV @ Y
    Z = H / sig_old[:, None] ** 2
    Y1 = Y - CV.T @ Z
    H2 = CV @ Y1
    Z2 = H2 / sig_old[:, None] ** 2
    Y2 = Y1 - CV.T @ Z2
    Z += Z2
    Gv = Z * sig_old[:, None]
    Q2v, R2v = np.linalg.qr(Y2)
    Xu, Xv = np.vstack([Gu, R2u]), np.vstack([Gv, R2v])
    K = Xu @ Xv.T
    K[np.arange(kc), np.arange(kc)] += beta * sig_old
    if order == "knorm":
        perm = np.argsort(-np.linalg.norm(K, axis=0), kind="stable")
    else:
        perm = np.arange(K.shape[1])
    Us, sig, sw = jacobi_left(K[:, perm], stop2)
    VS = K.T @ Us                                 # V S' = K^T Us (column order of K is irrelevant to Us / sigma)
    rk = min(new_rank(sig, acc), maxrank)
    CUn = np.hstack([CU, Q2u]) @ Us[:, :rk]
    VSt = VS[:, :rk].copy()
    VSt[:kc] /= sig_old[:, None]
    CVn = (np.hstack([CV.T, Q2v]) @ VSt).T
    return CUn, CVn, sw


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    ksteps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    acc = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
    p = O.CompressionParameters(acc)
    Cref = O.CompressedTile(np.zeros((nb, 1), order="F"), np.zeros((1, nb), order="F"), nb // 3)
    tiles = [(O.synth_compressed_tile(nb, rank, 100 + k), O.synth_compressed_tile(nb, rank, 200 + k)) for k in range(ksteps)]
    ref_ranks = []
    for A, B in tiles:
        O.hcore_gemm(1.0, A, False, B, False, 1.0, Cref, p)
        ref_ranks.append(Cref.rank)
    ref = Cref.to_dense()
    print("oracle ranks", ref_ranks)
    for name, fn, order, stop2 in (("full / vnorm order", step_full, "vnorm", 0.0),
                                   ("incremental / knorm order", step_incremental, "knorm", 0.0),
                                   ("incremental / natural order", step_incremental, "nat", 0.0),
                                   ("incremental / knorm + acc stop", step_incremental, "knorm", acc)):
        CU, CV = np.zeros((nb, 1)), np.zeros((1, nb))
        ranks, sweeps, orthU, orthW = [], [], [], []
        for k, (A, B) in enumerate(tiles):
            P, Y = product_term(A, B, 1.0)
            if k == 0 or fn is step_full:
                CU, CV, sw = step_full(CU, CV, P, Y, 1.0, acc, nb // 3, stop2=stop2)
            else:
                CU, CV, sw = fn(CU, CV, P, Y, 1.0, acc, nb // 3, order=order, stop2=stop2)
            ranks.append(CU.shape[1])
            sweeps.append(sw)
            s = np.linalg.norm(CV, axis=1)
            W = CV.T / s
            orthU.append(np.abs(CU.T @ CU - np.eye(CU.shape[1])).max())
            orthW.append(np.abs(W.T @ W - np.eye(W.shape[1])).max())
        got = CU @ CV
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"{name:32s}: rel Fro vs oracle {err:.2e} (gate {10 * acc:.0e}), ranks {ranks}, "
              f"max rank diff {max(abs(a - b) for a, b in zip(ref_ranks, ranks))}, sweeps {sweeps}, "
              f"|CU^T CU - I| {max(orthU):.1e}, |W^T W - I| {max(orthW):.1e}")
