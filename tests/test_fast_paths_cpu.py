"""CPU checks of the MATH behind the two Cholesky-based fast paths of the incremental recompression (round 2), on the numpy
transcriptions of the kernels in tests/tools/emulate_cholesky_paths.py (k_cholqr_pass, k_vcore_chol): they must reproduce a
Householder QR where their safety tests pass and must refuse (-> Householder fallback on the device) where CholeskyQR would
be inaccurate.  The GPU behaviour itself is tested in tests/test_gpu_parity.py (test_cholqr_fast_path_fallback_and_deflation)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import emulate_cholesky_paths as E  # noqa: E402

STEPS = list(E.incremental_inputs(nb=256, rank=20, ksteps=4))


def test_cholqr2_matches_householder_on_generic_updates():
    assert len(STEPS) == 3
    for k, Xu, Xv, M, pos, kc, kp in STEPS:
        for X in (Xu, Xv):
            Q, R, info = E.cholqr2(X)
            assert Q is not None, (k, info)
            live = np.abs(np.diag(Q.T @ Q)) > 0.5
            assert np.abs(Q[:, live].T @ Q[:, live] - np.eye(live.sum())).max() < 1e-13
            assert np.abs(X - Q @ R).max() <= 1e-13 * np.abs(X).max()
            assert np.allclose(np.tril(R, -1), 0)
            # same factor as Householder up to the signs of the rows (and zero rows where a column was deflated)
            Rh = np.linalg.qr(X)[1]
            assert np.abs(np.abs(R[live][:, live]) - np.abs(np.linalg.qr(X[:, live])[1])).max() <= 1e-10 * np.abs(Rh).max()


def test_cholqr2_refuses_nearly_dependent_columns_and_deflates_noise():
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((300, 12)))
    bad = q.copy()
    bad[:, 5] = bad[:, 4] + 1e-9 * bad[:, 5]                 # condition ~1e9: CholeskyQR would lose all accuracy
    Q, R, why = E.cholqr2(bad)
    assert Q is None and why[0] == "pass 0"
    noisy = q * (0.3 ** np.arange(12))
    noisy[:, 11] = 1e-17 * q[:, 3]                            # rounding-level copy of another direction
    Q, R, info = E.cholqr2(noisy)
    assert Q is not None and info["deflated"] == 1
    assert np.all(Q[:, 11] == 0) and np.all(R[11, :] == 0) and np.all(R[:, 11] == 0)
    assert np.abs(noisy[:, :11] - (Q @ R)[:, :11]).max() <= 1e-14


def test_graded_factor_by_cholesky_of_the_assembled_gram_matrix():
    for k, Xu, Xv, M, pos, kc, kp in STEPS:
        Rp, (Gs, Rc, dscale), bad = E.graded_factor(M, pos, kc, kp)
        assert Rp is not None, (k, bad)
        assert np.abs(Rc.T @ Rc - (Gs + np.triu(Gs, 1).T)).max() < 1e-13            # Cholesky of the scaled Gram matrix
        ref = np.linalg.qr(M)[1]
        assert np.abs(np.abs(Rp) - np.abs(ref)).max() <= 1e-12 * np.abs(ref).max()      # = the Householder R up to row signs
        # the scaled Gram matrix is well conditioned although M spans many decades
        full = Gs + np.triu(Gs, 1).T
        assert np.linalg.cond(full) < 50 and np.linalg.cond(M) > 1e6


def test_graded_factor_deflates_columns_inside_the_span_of_the_spikes_and_refuses_dependent_ones():
    k, Xu, Xv, M, pos, kc, kp = STEPS[-1]
    M2 = M.copy()
    c = pos[kc + kp - 1]                                       # a dense column whose part outside span(W) was deflated
    M2[kc:, c] = 0.0
    M2[:kc, c] *= 1e-15 / max(np.abs(M2[:kc, c]).max(), 1e-300)
    Rp, aux, bad = E.graded_factor(M2, pos, kc, kp)
    assert Rp is not None and np.all(Rp[:, c] == 0)
    keep = np.arange(kc + kp) != c
    ref = np.linalg.qr(M2[:, keep])[1]
    got = Rp[keep][:, keep]
    assert np.abs(np.abs(got) - np.abs(ref)).max() <= 1e-12 * np.abs(ref).max()
    M3 = M.copy()
    a, b = pos[kc], pos[kc + 1]
    M3[:, b] = M3[:, a] * 0.5 + 1e-10 * M3[:, b]              # two nearly parallel dense columns
    Rp, aux, bad = E.graded_factor(M3, pos, kc, kp)
    assert Rp is None
