"""CPU emulation (numpy, test infrastructure) of the two Cholesky-based fast paths of the incremental recompression, as the
CUDA kernels compute them (hcorepp_b200/csrc/kernels_tlr.cuh):

  * `cholqr2(X)`            -- k_cholqr_pass: CholeskyQR2 of the kp new columns with column scaling, the pivot test on the
                               scaled Gram matrix (>= 1e-6 in pass 0, within [1/4, 4] in pass 1) and the deflation of columns
                               below 1e-13 of the largest one; returns (Q, R, info) or (None, None, reason) = "fall back".
  * `graded_factor(M, pos, kc, kp)` -- k_vcore_chol: the triangular factor R' of M = RV * Pi from a blocked left-looking
                               Cholesky of the ASSEMBLED, column-scaled Gram matrix (spike-spike: identity, spike-dense: one
                               product, dense-dense: Gram of the dense columns), upper storage only.

`incremental_inputs()` builds the (X, M, pos, kc, kp) a k-sum of reference-law tiles produces at each step, the way
tests/tools/emulate_incremental.py does.  Used by tests/test_oracle.py; nothing in the product imports this."""
import sys

import numpy as np

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))
sys.path.insert(0, _HERE)
from oracle import tlr_oracle as O  # noqa: E402
from emulate_incremental import new_rank, product_term  # noqa: E402


def _chol_upper_with_test(G, lo, hi):
    """right-looking Cholesky of the upper triangle; returns (R, None) or (None, (j, pivot))."""
    R = np.triu(G).astype(np.float64).copy()
    n = R.shape[0]
    for j in range(n):
        piv = R[j, j]
        if not (lo <= piv <= hi):
            return None, (j, float(piv))
        R[j + 1:, j + 1:] -= np.triu(np.outer(R[j, j + 1:], R[j, j + 1:])) / piv
    d = np.sqrt(np.diag(R)).copy()
    R = R / d[:, None]
    R[np.arange(n), np.arange(n)] = d
    return np.triu(R), None


def cholqr2(X):
    kp = X.shape[1]
    Q, Rtot = X, np.eye(kp)
    deflated = np.zeros(kp, bool)
    for ps in range(2):
        G = Q.T @ Q
        g = np.diag(G).copy()
        if ps == 0:
            deflated = ~(g > 1e-26 * g.max())
            d = np.where(deflated, 0.0, 1.0 / np.sqrt(np.where(g > 0, g, 1.0)))
        else:
            deflated = g == 0.0
            d = np.where(deflated, 0.0, 1.0)
        Gs = G * d[:, None] * d[None, :]
        Gs[deflated, :] = 0.0
        Gs[:, deflated] = 0.0
        Gs[deflated, deflated] = 1.0
        Rc, bad = _chol_upper_with_test(Gs, 1e-6 if ps == 0 else 0.25, 4.0)
        if Rc is None:
            return None, None, ("pass %d" % ps,) + bad
        S = np.linalg.inv(Rc) * d[:, None]                     # D Rc^-1 (zero rows / columns where deflated)
        S[:, deflated] = 0.0
        Q = Q @ S
        with np.errstate(divide="ignore", invalid="ignore"):
            R1 = np.where((d[None, :] != 0) & (d[:, None] != 0), Rc / np.where(d != 0, d, 1.0)[None, :], 0.0)
        Rtot = R1 @ Rtot if ps == 1 else R1
    return Q, np.triu(Rtot), {"deflated": int(deflated.sum())}


def graded_factor(M, pos, kc, kp, nb=32):
    r = kc + kp
    ipos = np.zeros(r, int)
    ipos[pos] = np.arange(r)
    dscale = np.zeros(r)
    for cs in range(r):
        o = ipos[cs]
        if o < kc:
            v = M[o, cs]
            dscale[cs] = 1.0 / abs(v) if v != 0 else 0.0
    dense = [pos[kc + l] for l in range(kp)]
    Nm = M[:, dense].T @ M[:, dense]
    for l in range(kp):
        dscale[dense[l]] = 1.0 / np.sqrt(Nm[l, l]) if Nm[l, l] > 0 else 0.0
    dmin = dscale[dscale > 0].min()
    dscale[dscale > 1e13 * dmin] = 0.0                          # negligible columns: identity row / column
    S = np.full((r, r), np.nan)                                 # (the lower triangle is never read)
    for j in range(r):
        for i in range(j + 1):
            oi, oj = ipos[i], ipos[j]
            if i == j:
                v = 1.0
            elif dscale[i] == 0 or dscale[j] == 0 or (oi < kc and oj < kc):
                v = 0.0
            elif oi >= kc and oj >= kc:
                v = Nm[oi - kc, oj - kc] * dscale[i] * dscale[j]
            else:
                cs, cd, os_ = (i, j, oi) if oi < kc else (j, i, oj)
                v = M[os_, cs] * M[os_, cd] * dscale[i] * dscale[j]
            S[i, j] = v
    Gs = np.triu(S).copy()
    for n0 in range(0, r, nb):
        jw = min(nb, r - n0)
        if n0 > 0:
            upd = S[:n0, n0:n0 + jw].T @ S[:n0, n0:]
            for a in range(jw):
                S[n0 + a, n0 + a:] -= upd[a, a:]
        D, bad = _chol_upper_with_test(S[n0:n0 + jw, n0:n0 + jw], 1e-6, 4.0)
        if D is None:
            return None, None, (n0,) + bad
        S[n0:n0 + jw, n0:n0 + jw] = D + np.tril(S[n0:n0 + jw, n0:n0 + jw], -1)
        if n0 + jw < r:
            S[n0:n0 + jw, n0 + jw:] = np.linalg.solve(D.T, S[n0:n0 + jw, n0 + jw:])
    Rc = np.triu(S)
    with np.errstate(divide="ignore", invalid="ignore"):
        Rp = np.where(dscale[None, :] != 0, Rc / np.where(dscale != 0, dscale, 1.0)[None, :], 0.0)
    return np.triu(Rp), (Gs, Rc, dscale), None


def incremental_inputs(nb=256, rank=20, ksteps=5, acc=1e-8):
    """yields (k, X_u, X_v, M, pos, kc, kp) for the incremental steps of a k-sum of reference-law tiles"""
    CU = CV = None
    for k in range(ksteps):
        A, B = O.synth_compressed_tile(nb, rank, 100 + k), O.synth_compressed_tile(nb, rank, 200 + k)
        P, Y = product_term(A, B, 1.0)
        if k > 0:
            kc, kp = CU.shape[1], Y.shape[1]
            sig = np.linalg.norm(CV, axis=1)
            W = CV.T / sig
            Gv = W.T @ Y
            Y2 = Y - W @ Gv
            G2 = W.T @ Y2
            Y2 -= W @ G2
            Gv += G2
            P2 = P - CU @ (CU.T @ P)
            P2 -= CU @ (CU.T @ P2)
            _, R2v = np.linalg.qr(Y2)
            RV = np.block([[np.diag(sig), Gv], [np.zeros((kp, kc)), R2v]])
            SV = np.hstack([CV.T, Y])
            order = np.argsort(-np.linalg.norm(SV, axis=0), kind="stable")
            pos = np.zeros(kc + kp, int)
            pos[order] = np.arange(kc + kp)
            yield k, P2, Y2, RV[:, order], pos, kc, kp
        SU, SV = (P, Y) if k == 0 else (np.hstack([CU, P]), np.hstack([CV.T, Y]))
        QU, RU = np.linalg.qr(SU)
        QV, RVf = np.linalg.qr(SV)
        U_, s_, Vt_ = np.linalg.svd(RU @ RVf.T)
        rk = min(new_rank(s_, acc), nb // 3)
        CU, CV = QU @ U_[:, :rk], (QV @ (Vt_[:rk].T * s_[:rk])).T


if __name__ == "__main__":
    for k, Xu, Xv, M, pos, kc, kp in incremental_inputs():
        for nm, X in (("U", Xu), ("V", Xv)):
            Q, R, info = cholqr2(X)
            print(k, nm, "fallback" if Q is None else "orth %.1e  |X - QR| %.1e  deflated %d" % (
                np.abs(Q.T @ Q - np.diag((np.abs(np.diag(Q.T @ Q)) > 0.5).astype(float))).max(), np.abs(X - Q @ R).max(), info["deflated"]))
        Rp, aux, bad = graded_factor(M, pos, kc, kp)
        ref = np.linalg.qr(M)[1]
        print(k, "graded factor:", "fallback %s" % (bad,) if Rp is None else "| |R'| - |R_qr| | / |R| %.1e" % (
            np.abs(np.abs(Rp) - np.abs(ref)).max() / np.abs(ref).max()))
