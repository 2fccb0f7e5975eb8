"""Parity of dlc_lds_open_rollout_f32 (LDS.__call__ x T, LDS.generate_random_trajectory; deluca/envs/_lds.py:102-140)
with the CPU oracle, and of the LDS mirror built on it (B200 only)."""
import numpy as np
import pytest
import torch

from oracle import lds as O
from oracle import philox

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def _sys(N, d, n, p, T, seed=0, shared=False):
    rng = np.random.default_rng(seed)
    M = 1 if shared else N
    A = np.zeros((M, d, d), np.float32)
    for i in range(M):
        Q = np.linalg.qr(rng.standard_normal((d, d)))[0]
        A[i] = Q @ np.diag(rng.uniform(0.5, 0.95, d)) @ Q.T
    B = (rng.standard_normal((M, d, n)) / np.sqrt(d)).astype(np.float32)
    C = (rng.standard_normal((M, p, d)) / np.sqrt(d)).astype(np.float32)
    if shared:
        A, B, C = A[0], B[0], C[0]
    x0 = (0.1 * rng.standard_normal((N, d))).astype(np.float32)
    U = rng.standard_normal((T, N, n)).astype(np.float32)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    return A, B, C, x0, U, W


def _close(got, want):
    scale = np.abs(want).mean()
    assert (np.abs(got - want) <= RTOL * np.maximum(np.abs(want), scale)).all(), np.abs(got - want).max()


@pytest.mark.parametrize("d,n,p,shared", [(10, 3, 2, False), (10, 3, 2, True), (10, 3, 10, False), (4, 1, 1, False),
                                          (25, 1, 1, True), (40, 5, 7, False), (64, 64, 64, False)])
def test_given_actions_match_oracle(d, n, p, shared):
    from deluca_b200.fused import lds_open_rollout
    N, T = 37, 150
    A, B, C, x0, U, W = _sys(N, d, n, p, T, seed=d + p, shared=shared)
    Y, xf = O.open_rollout(A, B, C, x0, U, W)
    x, y, u = lds_open_rollout(_dev(A), _dev(B), _dev(C), _dev(x0), T, u_in=_dev(U), W=_dev(W), observe0=True,
                               keep_actions=True)
    assert y.shape == (T + 1, N, p)
    y0 = np.einsum("npd,nd->np", C, x0) if not shared else x0 @ C.T
    _close(y[0].cpu().numpy(), y0)
    _close(y[1:].cpu().numpy(), Y)
    _close(x.cpu().numpy(), xf)
    assert torch.equal(u, _dev(U))


def test_row_rings_equal_register_buffers(monkeypatch):
    """Register-resident kernel (d, n, p) = (10, 3, 2) with given actions and disturbances: the cp.async rings
    (row_ring.cuh; taken when N % 4 == 0) against the register double buffers (DLC_ROW_RING=0), bit for bit, for
    horizons around the ring depth and a ragged last warp; and against the oracle."""
    from deluca_b200.fused import lds_open_rollout
    d, n, p = 10, 3, 2
    for N, T in ((132, 1), (132, 3), (132, 4), (132, 5), (132, 61), (4096 + 8, 33)):
        A, B, C, x0, U, W = _sys(N, d, n, p, T, seed=N + T, shared=N > 1000)
        args = (_dev(A), _dev(B), _dev(C), _dev(x0), T)
        kw = dict(u_in=_dev(U), W=_dev(W), observe0=True, keep_actions=True)
        monkeypatch.setenv("DLC_ROW_RING", "0")
        x0_, y0_, u0_ = lds_open_rollout(*args, **kw)
        monkeypatch.delenv("DLC_ROW_RING")
        x1_, y1_, u1_ = lds_open_rollout(*args, **kw)
        assert torch.equal(x0_, x1_) and torch.equal(y0_, y1_) and torch.equal(u0_, u1_), (N, T)
        if T == 61:
            Y, xf = O.open_rollout(A, B, C, x0, U, W)
            _close(y1_[1:].cpu().numpy(), Y)
            _close(x1_.cpu().numpy(), xf)


def test_random_policy_streams_and_resume():
    from deluca_b200.fused import gaussian_disturbance, lds_open_rollout
    N, T, d, n, p = 70, 64, 10, 3, 2
    A, B, C, x0, _, _ = _sys(N, d, n, p, T, seed=1)
    kw = dict(seed=99, env_offset=11, w_std=0.5, action_std=1.0)
    x, y, u = lds_open_rollout(_dev(A), _dev(B), _dev(C), _dev(x0), T, keep_actions=True, **kw)
    # actions: stream 3 of the library's generator; disturbances: the stream the controller kernels draw
    U = O.random_actions_stream3(99, np.arange(N) + 11, 0, T, n)
    np.testing.assert_allclose(u.cpu().numpy(), U, rtol=0, atol=5e-6)
    W = gaussian_disturbance(N, T, d, seed=99, env_offset=11, std=0.5)
    np.testing.assert_allclose(W.cpu().numpy(), philox.gaussian_disturbance(99, np.arange(N) + 11, 0, T, d), atol=5e-6)
    Y, xf = O.open_rollout(A, B, C, x0, u.cpu().numpy(), W.cpu().numpy())
    _close(y.cpu().numpy(), Y)
    # same numbers with the streams materialised by the caller
    x2, y2, _ = lds_open_rollout(_dev(A), _dev(B), _dev(C), _dev(x0), T, u_in=u, W=W)
    assert torch.equal(y, y2) and torch.equal(x, x2)
    # resume and sharding: bit for bit
    h = 24
    xa, ya, _ = lds_open_rollout(_dev(A), _dev(B), _dev(C), _dev(x0), h, **kw)
    xb, yb, _ = lds_open_rollout(_dev(A), _dev(B), _dev(C), xa, T - h, t0=h, **kw)
    assert torch.equal(torch.cat([ya, yb]), y) and torch.equal(xb, x)
    s = slice(32, 70)
    xs, ys, _ = lds_open_rollout(_dev(A[s]), _dev(B[s]), _dev(C[s]), _dev(x0[s]), T, seed=99, env_offset=11 + 32)
    assert torch.equal(ys, y[:, s])


def test_lds_mirror_generate_random_trajectory_and_step():
    """The stateful face (env(u), generate_random_trajectory) and the functional one (step / rollout) agree, and
    non-Gaussian disturbances go through the materialised-W path."""
    from deluca_b200.envs import LDS, SinusDisturbance, ZeroDisturbance
    from deluca_b200.envs._lds import LDSState
    env = LDS()
    env.init(3, 10, 2, key=5)                       # the colab's LDS: d_action = 3, d_hidden = 10, d_obs = 2
    sig = env.generate_random_trajectory(200, key=7)
    assert sig.shape == (200,) and env.t == 200 and torch.isfinite(sig).all()
    env2 = LDS()
    env2.init(3, 10, 2, key=5)
    st, _ = env2.reset(1)
    st, Y, U = env2.rollout(st, 200, key=7, keep_actions=True)
    assert torch.equal(Y[:, 0, 0], sig)
    # stepping with the same actions one call at a time gives the same observations
    env3 = LDS()
    env3.init(3, 10, 2, key=5)
    ys = torch.stack([env3(U[t, 0].reshape(3, 1), key=7)[0, 0] for t in range(20)])
    assert torch.equal(ys, sig[:20])
    # oracle parity of the whole signal
    W = philox.gaussian_disturbance(7, np.arange(1), 0, 200, 10)
    Yo, _ = O.open_rollout(env.A.cpu().numpy(), env.B.cpu().numpy(), env.C.cpu().numpy(), np.zeros((1, 10), np.float32),
                           U.cpu().numpy(), W)
    _close(sig.cpu().numpy(), Yo[:, 0, 0])
    for dist in (SinusDisturbance(), ZeroDisturbance()):
        e = LDS()
        e.init(3, 10, 2, key=5, disturbance=dist)
        st, y0 = e.reset(4)
        assert y0.shape == (4, 2)
        st, Yd, Ud = e.rollout(st, 30, key=3, keep_actions=True)
        Wd = np.stack([dist(t, 3, N=4).cpu().numpy() for t in range(30)])
        Yo, _ = O.open_rollout(e.A.cpu().numpy(), e.B.cpu().numpy(), e.C.cpu().numpy(), np.zeros((4, 10), np.float32),
                               Ud.cpu().numpy(), Wd)
        _close(Yd.cpu().numpy(), Yo)


def test_argument_errors():
    from deluca_b200 import _abi
    from deluca_b200.fused import lds_open_rollout
    z = lambda *s: torch.zeros(s, device="cuda")
    with pytest.raises(_abi.DlcError):
        lds_open_rollout(z(2, 70, 70), z(2, 70, 1), z(2, 1, 70), z(2, 70), 3)       # d > 64
    with pytest.raises(TypeError):
        lds_open_rollout(z(2, 4, 4).cpu(), z(2, 4, 1), z(2, 1, 4), z(2, 4), 3)
    x, y, _ = lds_open_rollout(z(0, 4, 4), z(0, 4, 1), z(0, 1, 4), z(0, 4), 3)
    assert y.shape == (3, 0, 1)


def test_sinus_disturbance_kernel():
    """SinusDisturbance (_lds.py:44-52): sin(t * phases) * amplitude in float32, one launch for a whole series."""
    from deluca_b200.envs import SinusDisturbance
    from deluca_b200.fused import sinus_disturbance
    rng = np.random.default_rng(0)
    ph = rng.standard_normal(10).astype(np.float32)
    T, N, t0 = 300, 5, 9000          # large t: the argument reduction of sinf matters
    W = sinus_disturbance(N, T, _dev(ph), 0.5, t0=t0).cpu().numpy()
    t = (np.arange(T) + t0).astype(np.float32)[:, None]
    want = (np.sin((t * ph[None, :]).astype(np.float32).astype(np.float64)) * 0.5).astype(np.float32)
    assert W.shape == (T, N, 10)
    np.testing.assert_allclose(W, np.broadcast_to(want[:, None, :], W.shape), rtol=0, atol=2e-7)
    dist = SinusDisturbance()
    dist.init(10)
    assert dist.phases.shape == (10,)
    one = dist(7, 0, N=3)
    ser = dist.series(5, 4, 3)
    assert one.shape == (3, 10) and torch.equal(ser[2], one)


def test_learner_generate_trajectories_on_lds():
    """Learner.generate_trajectories (learners/core.py:17-41) with the LDS env: (obs, action, next_obs)[N,T,.]."""
    from deluca_b200.envs import LDS
    from deluca_b200.learners import Learner
    env = LDS()
    env.init(3, 10, 2, key=1)
    N, T = 33, 40
    obs, act, nxt = Learner(env).generate_trajectories(N, T, key=4)
    assert obs.shape == (N, T, 2) and act.shape == (N, T, 3) and nxt.shape == (N, T, 2)
    assert torch.equal(obs[:, 1:], nxt[:, :-1]) and float(obs[:, 0].abs().max()) == 0.0   # x0 = 0
    W = philox.gaussian_disturbance(4, np.arange(N), 0, T, 10)
    Y, _ = O.open_rollout(env.A.cpu().numpy(), env.B.cpu().numpy(), env.C.cpu().numpy(), np.zeros((N, 10), np.float32),
                          act.transpose(0, 1).cpu().numpy(), W)
    _close(nxt.transpose(0, 1).cpu().numpy(), Y)
