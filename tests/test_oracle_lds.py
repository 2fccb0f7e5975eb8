"""Oracle self-checks for LDS + GPC/BPC/LQR (CPU only)."""
import numpy as np
import pytest
import torch

from oracle import lds as O
from oracle import philox as P


def _problem(N=3, d=10, m=3, H=3, HH=10, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    A, B, _ = O.lds_default_params(rng, N, d, m, dtype)
    K = O.lqr_gain(A, B, dtype=dtype)
    M = (0.1 * rng.standard_normal((N, H, m, d))).astype(dtype)
    w = rng.standard_normal((N, H + HH, d)).astype(dtype)
    return A, B, K, M, w


def _torch_policy_loss(M, w, A, B, K, H, HH):
    """Literal restatement of deluca/agents/_gpc.py:109-125 in torch (for autograd)."""
    d = A.shape[0]

    def action(state, h):
        return -K @ state + torch.tensordot(M, w[h:h + H], dims=([0, 2], [0, 1]))

    y = torch.zeros(d, dtype=M.dtype)
    for h in range(H - 1):
        y = A @ y + B @ action(y, h) + w[h + H]
    u = action(y, HH - 1)
    return y @ y + u @ u


@pytest.mark.parametrize("d,m,H,HH", [(10, 3, 3, 2), (10, 3, 3, 10), (10, 3, 10, 10),
                                      (10, 3, 10, 9), (4, 2, 5, 7), (3, 1, 1, 1), (5, 2, 2, 1)])
def test_hand_adjoint_matches_autograd(d, m, H, HH):
    A, B, K, M, w = _problem(2, d, m, H, HH)
    G = O.gpc_grad(M, w, A, B, K, H, HH)
    loss = O.gpc_policy_loss(M, w, A, B, K, H, HH)
    for n in range(2):
        Mt = torch.tensor(M[n], requires_grad=True)
        lt = _torch_policy_loss(Mt, torch.tensor(w[n]), torch.tensor(A[n]), torch.tensor(B[n]),
                                torch.tensor(K[n]), H, HH)
        lt.backward()
        np.testing.assert_allclose(loss[n], lt.item(), rtol=1e-12)
        np.testing.assert_allclose(G[n], Mt.grad.numpy(), rtol=1e-10, atol=1e-12)


def test_hand_adjoint_matches_finite_differences():
    A, B, K, M, w = _problem(1, 6, 2, 4, 5)
    G = O.gpc_grad(M, w, A, B, K, 4, 5)
    eps = 1e-6
    rng = np.random.default_rng(1)
    for _ in range(20):
        i, a, b = rng.integers(4), rng.integers(2), rng.integers(6)
        Mp, Mm = M.copy(), M.copy()
        Mp[0, i, a, b] += eps
        Mm[0, i, a, b] -= eps
        fd = (O.gpc_policy_loss(Mp, w, A, B, K, 4, 5) - O.gpc_policy_loss(Mm, w, A, B, K, 4, 5)) / (2 * eps)
        np.testing.assert_allclose(G[0, i, a, b], fd[0], rtol=1e-6, atol=1e-8)


def test_lqr_gain_is_dare_fixed_point():
    rng = np.random.default_rng(0)
    A, B, _ = O.lds_default_params(rng, 4, 10, 3, np.float64)
    K = O.lqr_gain(A, B, dtype=np.float64)
    # closed loop must be stable and K must satisfy the Riccati fixed point
    for n in range(4):
        assert np.max(np.abs(np.linalg.eigvals(A[n] - B[n] @ K[n]))) < 1.0
        X = np.eye(10)
        for _ in range(2000):
            Kx = np.linalg.solve(B[n].T @ X @ B[n] + np.eye(3), B[n].T @ X @ A[n])
            X = np.eye(10) + A[n].T @ X @ (A[n] - B[n] @ Kx)
        np.testing.assert_allclose(K[n], Kx, rtol=1e-8, atol=1e-10)


def test_gpc_with_zero_lr_is_lqr():
    rng = np.random.default_rng(3)
    N, d, m, T = 2, 10, 3, 50
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    a = O.rollout_lds(A, B, K, x0, W, "gpc", lr_scale=0.0)
    b = O.rollout_lds(A, B, K, x0, W, "lqr")
    np.testing.assert_array_equal(a["costs"], b["costs"])
    np.testing.assert_array_equal(a["X"], b["X"])


def test_gpc_f32_tracks_f64_and_learns():
    rng = np.random.default_rng(4)
    N, d, m, T = 4, 10, 3, 300
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    # lr_scale 5e-4: the as-written recursion is unstable at the default 5e-3 (next test)
    r32 = O.rollout_lds(A, B, K, x0, W, "gpc", lr_scale=5e-4, dtype=np.float32)
    r64 = O.rollout_lds(A, B, K, x0, W, "gpc", lr_scale=5e-4, dtype=np.float64)
    rel = np.abs(r32["costs"] - r64["costs"]) / np.maximum(np.abs(r64["costs"]), 1e-6)
    assert rel[r64["costs"] > 1e-3].max() < 1e-4
    assert np.abs(r64["state"]["M"]).max() > 0


def test_reference_noise_quirk_is_unstable_and_intended_variant_is_not():
    """SURVEY A-3.2: `noise = x - A x_prev - B u_t` uses the action just computed."""
    rng = np.random.default_rng(4)
    N, d, m, T = 16, 10, 3, 1500
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    with np.errstate(all="ignore"):
        ref = O.rollout_lds(A, B, K, x0, W, "gpc", dtype=np.float64, keep_x=False)
    fix = O.rollout_lds(A, B, K, x0, W, "gpc", dtype=np.float64, keep_x=False,
                        noise_uses_prev_action=True)
    assert (~np.isfinite(ref["costs"]).all(axis=0)).sum() > N // 2
    assert np.isfinite(fix["costs"]).all()


def test_rollout_resume_is_identical():
    rng = np.random.default_rng(5)
    N, d, m, T = 2, 10, 3, 40
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    full = O.rollout_lds(A, B, K, x0, W, "gpc")
    h1 = O.rollout_lds(A, B, K, x0, W[:15], "gpc")
    h2 = O.rollout_lds(A, B, K, h1["x_final"], W[15:], "gpc", state=h1["state"])
    np.testing.assert_array_equal(full["costs"][15:], h2["costs"])
    np.testing.assert_array_equal(full["x_final"], h2["x_final"])


def test_bpc_runs_and_perturbs():
    rng = np.random.default_rng(6)
    N, d, m, H, T = 2, 10, 3, 5, 30
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    M0 = O.unit_norm(rng.standard_normal((N, H, m, d)).astype(np.float32))
    eps0 = O.unit_norm(rng.standard_normal((N, H, H, m, d)).astype(np.float32))
    eps = O.unit_norm(rng.standard_normal((T * N, H, m, d)).astype(np.float32)).reshape(T, N, H, m, d)
    st = O.bpc_init(N, d, m, H, M0, eps0)
    r = O.rollout_lds(A, B, K, x0, W, "bpc", H=H, state=st, bpc_eps=eps, decay=False, lr_scale=1e-5)
    assert np.isfinite(r["costs"]).all()
    np.testing.assert_allclose(np.sqrt((eps[0, 0] ** 2).sum()), 1.0, rtol=1e-5)


# ------------------------------------------------------------------ philox
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kat:
        r = P.philox4x32(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))
        assert tuple(int(v) for v in r[0]) == out


def test_gaussian_disturbance_moments_and_sharding():
    W = P.gaussian_disturbance(7, np.arange(512), 0, 64, 10, 0.5)
    assert W.shape == (64, 512, 10) and W.dtype == np.float32
    assert abs(W.mean()) < 5e-3 and abs(W.std() - 0.5) < 5e-3
    # a shard of envs / a later start draws the same numbers
    W2 = P.gaussian_disturbance(7, np.arange(100, 164), 10, 20, 10, 0.5)
    np.testing.assert_array_equal(W[10:30, 100:164], W2)


def test_c_restatement_matches_numpy_oracle():
    """oracle/lds_c.c (the CPU-baseline port) against the line-by-line NumPy restatement."""
    from oracle.lds_c import gpc_rollout_c
    rng = np.random.default_rng(7)
    for (d, m, H, HH) in ((10, 3, 3, 10), (10, 3, 10, 10), (4, 2, 5, 7)):
        N, T = 12, 80
        A, B, _ = O.lds_default_params(rng, N, d, m)
        K = O.lqr_gain(A, B)
        x0 = (0.1 * rng.standard_normal((N, d))).astype(np.float32)
        W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
        for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
            r = O.rollout_lds(A, B, K, x0, W, "gpc", H=H, HH=HH, lr_scale=1e-4, dtype=dt)
            c, x = gpc_rollout_c(A, B, K, x0, W, H, HH, 1e-4, True, dt)
            np.testing.assert_allclose(c, r["costs"], rtol=tol, atol=tol)
            np.testing.assert_allclose(x, r["x_final"], rtol=tol, atol=tol)
