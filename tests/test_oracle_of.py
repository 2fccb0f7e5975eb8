"""Oracle self-checks for the output-feedback controllers LQG / DRC / GRC / SFC / DSC (CPU only).

The reference ships no test or golden vector for these agents and cannot run here, so the common
filter-bank form of ``oracle/of.py`` is pinned to einsum-by-einsum restatements of the reference ``policy_loss``
functions, and the hand adjoints to finite differences of those literal restatements."""
import numpy as np
import pytest

from oracle import of as O


def _sys(rng, N, d, n, p, rho=0.9):
    A = np.zeros((N, d, d))
    for i in range(N):
        A[i] = np.diag(rng.uniform(0.5, rho, d)) + 0.05 * rng.standard_normal((d, d))
    B = rng.standard_normal((N, d, n))
    C = rng.standard_normal((N, p, d))
    return A, B, C


def test_spectral_filters_are_the_scaled_top_eigenvectors():
    m, h, gamma = 12, 4, 0.2
    filt = O.spectral_filters(m, h, gamma)
    assert filt.shape == (h, m)
    idx = np.arange(1, m + 1)
    Hm = ((1 - gamma) ** (idx[:, None] + idx[None, :] - 1)) / (idx[:, None] + idx[None, :] - 1)
    vals = np.linalg.eigvalsh(Hm)[-h:]
    unflipped = filt[:, ::-1]
    # rows are sigma^(1/4) v: H v = sigma v and |row| = sigma^(1/4)
    for j in range(h):
        v = unflipped[j] / vals[j] ** 0.25
        np.testing.assert_allclose(Hm @ v, vals[j] * v, atol=1e-12)
        np.testing.assert_allclose(np.linalg.norm(v), 1.0, atol=1e-12)


@pytest.mark.parametrize("kind,m,h", [("grc", 5, None), ("sfc", 6, 3), ("dsc", 4, 2), ("dsc", 5, 3)])
def test_filter_bank_form_equals_the_literal_policy_loss(kind, m, h):
    rng = np.random.default_rng(0)
    d, n, p = 6, 2, 3
    A, B, C = _sys(rng, 1, d, n, p)
    Phi, L = O.filter_bank(kind, m, h)
    ynats = rng.standard_normal((1, L + m, p))
    if kind == "grc":
        M = rng.standard_normal((m, n, p))
        Theta = M[None]
        lit = O.grc_policy_loss_literal(M, ynats[0][..., None], A[0], B[0], C[0], m)
    elif kind == "sfc":
        M0, M = rng.standard_normal((n, p)), rng.standard_normal((h, n, p))
        Theta = np.concatenate([M0[None], M])[None]
        lit = O.sfc_policy_loss_literal(M0, M, O.spectral_filters(m, h), ynats[0][..., None], A[0], B[0], C[0], m)
    else:
        M0, M1, M2 = rng.standard_normal((n, p)), rng.standard_normal((h, n, p)), rng.standard_normal((h, n, p))
        M3 = rng.standard_normal((h, h, n, p))
        Theta = O.dsc_pack_theta(M0, M1, M2, M3)[None]
        lit = O.dsc_policy_loss_literal(M0, M1, M2, M3, O.spectral_filters(m, h), ynats[0][..., None], A[0], B[0],
                                        C[0], m)
    got = O.fgrc_policy_loss(Theta, Phi, ynats, A, B, C, m)
    np.testing.assert_allclose(got[0], lit, rtol=1e-11)


@pytest.mark.parametrize("kind,m,h", [("grc", 5, None), ("sfc", 6, 3), ("dsc", 4, 2)])
def test_fgrc_hand_adjoint_matches_finite_differences(kind, m, h):
    rng = np.random.default_rng(1)
    d, n, p = 5, 2, 3
    A, B, C = _sys(rng, 2, d, n, p)
    Phi, L = O.filter_bank(kind, m, h)
    F = Phi.shape[0]
    Theta = rng.standard_normal((2, F, n, p))
    ynats = rng.standard_normal((2, L + m, p))
    G = O.fgrc_grad(Theta, Phi, ynats, A, B, C, m)
    eps = 1e-6
    for _ in range(25):
        f, a, q = rng.integers(F), rng.integers(n), rng.integers(p)
        Tp, Tm = Theta.copy(), Theta.copy()
        Tp[:, f, a, q] += eps
        Tm[:, f, a, q] -= eps
        fd = (O.fgrc_policy_loss(Tp, Phi, ynats, A, B, C, m) - O.fgrc_policy_loss(Tm, Phi, ynats, A, B, C, m)) / (2 * eps)
        np.testing.assert_allclose(G[:, f, a, q], fd, rtol=1e-6, atol=1e-7)


def test_markov_form_of_the_evolve_recursion():
    """sum_k C A^(m-1-k) B a(k) == C delta_m (the identity the CUDA kernel uses, lds_of.cu (i))."""
    rng = np.random.default_rng(2)
    d, n, p, m = 7, 3, 2, 6
    A, B, C = _sys(rng, 1, d, n, p)
    acts = rng.standard_normal((m, n))
    delta = np.zeros(d)
    for k in range(m):
        delta = A[0] @ delta + B[0] @ acts[k]
    G = O.drc_markov(A, B, C, m)[0]
    alt = sum(G[m - 1 - k] @ acts[k] for k in range(m))
    np.testing.assert_allclose(alt, C[0] @ delta, rtol=1e-12, atol=1e-12)


def test_drc_update_is_the_gradient_of_the_literal_policy_loss():
    rng = np.random.default_rng(3)
    d, n, p, m, h = 5, 2, 3, 4, 6
    A, B, C = _sys(rng, 1, d, n, p)
    G = O.drc_markov(A, B, C, h)
    K = rng.standard_normal((1, n, p))
    M0 = 0.1 * rng.standard_normal((1, m, n, p))
    st = O.drc_init(M0, h, n)
    st["y_nat"] = rng.standard_normal((1, m, p))
    st["us"] = rng.standard_normal((1, h, n))
    st["t"] = 2
    obs = rng.standard_normal((1, p))
    ref = dict(M=st["M"].copy(), y_nat=st["y_nat"].copy(), us=st["us"].copy())
    u = O.drc_call(st, obs, G, K, lr_scale=0.03, decay=True, RM=1000.0)
    # replay by hand with finite differences of the literal loss
    y_nat_new = obs[0] - np.einsum("hpa,ha->p", G[0], ref["us"][0])
    yn = np.concatenate([y_nat_new[None], ref["y_nat"][0][:-1]])
    np.testing.assert_allclose(st["y_nat"][0], yn)
    np.testing.assert_allclose(u[0], -K[0] @ obs[0] + np.einsum("map,mp->a", ref["M"][0], yn))
    loss = lambda M: O.drc_policy_loss_literal(M, G[0], yn[..., None], ref["us"][0][..., None], K[0])
    eps = 1e-6
    gfd = np.zeros_like(ref["M"][0])
    for idx in np.ndindex(*gfd.shape):
        Mp, Mm = ref["M"][0].copy(), ref["M"][0].copy()
        Mp[idx] += eps
        Mm[idx] -= eps
        gfd[idx] = (loss(Mp) - loss(Mm)) / (2 * eps)
    np.testing.assert_allclose(st["M"][0], ref["M"][0] - 0.03 / 3 * gfd, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(st["us"][0], np.concatenate([u, ref["us"][0][:-1]]))


def test_drc_without_decay_uses_unit_learning_rate_as_written():
    """``_drc.py:184``: ``lax.cond(decay, x/(t+1), 1.0)`` -- the non-decayed rate is 1.0, not lr_scale."""
    rng = np.random.default_rng(4)
    A, B, C = _sys(rng, 1, 4, 2, 2)
    G = O.drc_markov(A, B, C, 3)
    K = np.zeros((1, 2, 2))
    M0 = 0.01 * rng.standard_normal((1, 2, 2, 2))
    obs = rng.standard_normal((1, 2))
    a, b = O.drc_init(M0, 3, 2), O.drc_init(M0, 3, 2)
    O.drc_call(a, obs, G, K, lr_scale=0.03, decay=False)
    O.drc_call(b, obs, G, K, lr_scale=7.0, decay=False)
    np.testing.assert_allclose(a["M"], b["M"])


def test_lqg_gains_and_closed_loop():
    rng = np.random.default_rng(5)
    d, n, p = 6, 2, 3
    A, B, C = _sys(rng, 1, d, n, p)
    K, Lk = O.lqg_gains(A[0], B[0], C[0])
    assert K.shape == (n, d) and Lk.shape == (d, p)
    # separation principle: controller and (predictor-form) estimator error dynamics are both stable
    assert np.max(np.abs(np.linalg.eigvals(A[0] - B[0] @ K))) < 1
    T, N = 300, 4
    W = 0.1 * rng.standard_normal((T, N, d))
    rep = lambda M: np.repeat(M, N, axis=0)
    out = O.rollout_of(rep(A), rep(B), rep(C), np.zeros((N, d)), W, "lqg", dtype=np.float64,
                       K=rep(K[None]), Lk=rep(Lk[None]))
    assert np.isfinite(out["costs"]).all()


@pytest.mark.parametrize("kind", ["grc", "sfc", "dsc"])
def test_projection_keeps_theta_in_the_ball_and_episode_resumes(kind):
    rng = np.random.default_rng(6)
    N, d, n, p, m, h = 3, 6, 2, 2, 4, 2
    A, B, C = _sys(rng, N, d, n, p)
    Phi, L = O.filter_bank(kind, m, h)
    Theta0 = rng.standard_normal((N, Phi.shape[0], n, p))
    T = 40
    W = 0.5 * rng.standard_normal((T, N, d))
    x0 = np.zeros((N, d))
    full = O.rollout_of(A, B, C, x0, W, "fgrc", dtype=np.float64, Theta0=Theta0, Phi=Phi, m=m, R_M=1.5)
    assert (np.sqrt((full["state"]["Theta"] ** 2).sum(axis=(1, 2, 3))) <= 1.5 + 1e-9).all()
    a = O.rollout_of(A, B, C, x0, W[:15], "fgrc", dtype=np.float64, Theta0=Theta0, Phi=Phi, m=m, R_M=1.5)
    b = O.rollout_of(A, B, C, a["x_final"], W[15:], "fgrc", dtype=np.float64, Theta0=Theta0, Phi=Phi, m=m, R_M=1.5,
                     state=a["state"])
    np.testing.assert_allclose(np.concatenate([a["costs"], b["costs"]]), full["costs"], rtol=1e-12)
