"""Oracle checks for the experience post-processing (CPU): the loop restatement of gae_advantages
against the closed form of the recurrence and against torch autograd-free float64 evaluation."""
import numpy as np

from oracle import igpc as OG
from oracle import ppo as OP


def _series(T, N, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(T, N)).astype(dtype), (rng.random((T, N)) > 0.1).astype(dtype),
            rng.normal(size=(T, N)).astype(dtype))


def test_gae_matches_closed_form_sum():
    T, N, g, l = 17, 5, 0.99, 0.95
    r, mk, v = _series(T, N)
    adv = OP.gae_advantages(r, mk, v, g, l)
    # A_t = sum_{k>=t} (prod_{j=t}^{k-1} g l mask_j) delta_k, delta_{T-1} = r - v (no bootstrap, ppo.py:64)
    delta = np.empty_like(r)
    delta[:-1] = r[:-1] + g * v[1:] * mk[:-1] - v[:-1]
    delta[-1] = r[-1] - v[-1]
    ref = np.zeros_like(r)
    for t in range(T):
        w = np.ones(N)
        for k in range(t, T):
            ref[t] += w * delta[k]
            w = w * g * l * mk[k]
    np.testing.assert_allclose(adv, ref, rtol=1e-12, atol=1e-12)


def test_gae_lambda_one_no_masks_is_discounted_return_minus_value():
    T, N, g = 12, 3, 0.9
    r, _, v = _series(T, N, seed=1)
    adv = OP.gae_advantages(r, np.ones_like(r), v, g, 1.0)
    G = np.zeros_like(r)
    acc = np.zeros(N)
    for t in reversed(range(T)):
        acc = r[t] + g * acc
        G[t] = acc
    np.testing.assert_allclose(adv, G - v, rtol=1e-12, atol=1e-12)


def test_ac_tail_returns_as_written_and_intended():
    T, N = 9, 4
    r, mk, v = _series(T, N, seed=2)
    losses, logp = (-r).T.copy(), _series(T, N, seed=3)[0].T.copy()
    adv, ret = OP.ac_tail(losses, logp, v.T.copy(), np.ones((N, T)), 0.99, 0.95)
    np.testing.assert_allclose(adv.T, OP.gae_advantages(r, np.ones_like(r), v, 0.99, 0.95), rtol=1e-14)
    np.testing.assert_array_equal(ret, adv + logp)          # ppo.py:133 adds t[4] = log_probs
    _, ret2 = OP.ac_tail(losses, logp, v.T.copy(), np.ones((N, T)), 0.99, 0.95, returns_use_values=True)
    np.testing.assert_array_equal(ret2, adv + v.T)


def test_float32_gae_close_to_float64():
    r, mk, v = _series(200, 8, seed=4)
    a64 = OP.gae_advantages(r, mk, v, 0.99, 0.95)
    a32 = OP.gae_advantages(r.astype(np.float32), mk.astype(np.float32), v.astype(np.float32), 0.99, 0.95)
    assert a32.dtype == np.float32
    np.testing.assert_allclose(a32, a64, rtol=0, atol=5e-5)


def test_generate_trajectories_chains_obs_and_next_obs():
    env = OG.Pendulum()
    N, T = 6, 20
    act = OP.random_actions(7, np.arange(N), T, env.m)
    assert act.shape == (N, T, 1) and abs(float(act.std()) - 1.0) < 0.25
    obs, a, nxt = OP.generate_trajectories(env, env.init(N, np.float32), act)
    np.testing.assert_array_equal(obs[:, 1:], nxt[:, :-1])
    np.testing.assert_array_equal(obs[:, 0], env.init(N, np.float32))
    np.testing.assert_allclose(nxt[:, 3], env.step(obs[:, 3], a[:, 3]), rtol=0, atol=0)
    # distinct global env ids draw distinct streams; the same id draws the same stream on any shard
    np.testing.assert_array_equal(OP.random_actions(7, [3, 4], T, 2)[0], OP.random_actions(7, [3], T, 2)[0])
