"""The reference-shaped Python interface (Obj/Env/Agent mirrors, apg drivers) on the CUDA path (B200)."""
import numpy as np
import pytest
import torch

from oracle import lds as O

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def test_stateful_lds_gpc_loop_matches_oracle():
    """The colab-style loop of SURVEY S2 (config 1: single env, d=10, m=3, H=3, HH=10)."""
    from deluca_b200.agents import GPC
    from deluca_b200.envs import LDS
    from deluca_b200.fused import gaussian_disturbance
    rng = np.random.default_rng(0)
    d, m, T = 10, 3, 1000
    A, B, _ = O.lds_default_params(rng, 1, d, m)
    env = LDS()
    env.init(d_in=m, d_hidden=d, d_out=d, A=A[0], B=B[0], C=np.eye(d), key=5)
    agent = GPC(env.A, env.B, H=3, HH=10, lr_scale=1e-4)
    K = agent.K.cpu().numpy()[None]
    np.testing.assert_allclose(K, O.lqr_gain(A, B), rtol=1e-4, atol=1e-6)
    W = gaussian_disturbance(1, T, d, seed=5).cpu().numpy()
    ref = O.rollout_lds(A, B, K, np.zeros((1, d), np.float32), W, "gpc", H=3, HH=10, lr_scale=1e-4, dtype=np.float64)
    x = env.x
    costs = []
    for t in range(T):
        u = agent(x)
        costs.append(float((x * x).sum() + (u * u).sum()))
        x = env(u)
    rel = np.abs(np.array(costs) - ref["costs"][:, 0]) / np.maximum(ref["costs"][:, 0], 1e-2 * ref["costs"].mean())
    assert rel.max() < 1e-4
    assert agent.t == T and env.t == T


def test_apg_parallel_rollouts_equals_stepping_and_oracle():
    from deluca_b200.agents import GPC, LQR
    from deluca_b200.agents._gpc import quad_loss
    from deluca_b200.envs import LDS
    from deluca_b200.fused import gaussian_disturbance
    from deluca_b200.training.apg import apg_parallel_rollouts
    rng = np.random.default_rng(1)
    N, d, m, T = 48, 10, 3, 60
    A, B, _ = O.lds_default_params(rng, N, d, m)
    env = LDS()
    env.init(d_in=m, d_hidden=d, d_out=d, A=A, B=B, C=np.eye(d), key=9)
    agent = GPC(env.A, env.B, H=3, HH=10, lr_scale=1e-4)
    st, obs = env.reset(N)
    ro, env_final = apg_parallel_rollouts(env, st, obs, agent, agent.init(N), T, quad_loss)
    assert ro.losses.shape == (N, T) and ro.actions.shape == (N, T, m)
    W = gaussian_disturbance(N, T, d, seed=9).cpu().numpy()
    ref = O.rollout_lds(A, B, agent.K.cpu().numpy(), np.zeros((N, d), np.float32), W, "gpc", H=3, HH=10,
                        lr_scale=1e-4, dtype=np.float64)
    rel = np.abs(ro.losses.T.cpu().numpy() - ref["costs"]) / np.maximum(ref["costs"], 1e-2 * ref["costs"].mean())
    assert rel.max() < 1e-4
    # step-by-step through Env.step / Agent.step reproduces the fused rollout
    s, a = env.reset(N)[0], agent.init(N)
    for t in range(T):
        a, u = agent.step(a, s.x)
        s, _ = env.step(s, u)
    assert torch.allclose(s.x, env_final.x, rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.M, ro.agent_states.M, rtol=1e-4, atol=1e-7)
    lq = LQR(env.A, env.B)
    ro2, _ = apg_parallel_rollouts(env, st, obs, lq, None, T, quad_loss)
    assert float(ro2.losses.mean()) > 0


def test_obj_contract_and_errors():
    import dataclasses
    from deluca_b200.agents._gpc import GPCState
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.training.apg import apg_parallel_rollouts
    w = BreathWaveform.create()
    with pytest.raises(dataclasses.FrozenInstanceError):
        w.pip = 3
    s = GPCState(M=torch.zeros(1), noise_history=torch.zeros(2), state=torch.zeros(3), action=torch.zeros(4))
    assert [t.numel() for t in s.flatten()] == [1, 2, 3, 4]
    with pytest.raises(NotImplementedError):
        apg_parallel_rollouts(object(), None, None, None, None, 1, None)


def test_value_and_grad_apg_loss_matches_autograd():
    """Reverse mode through the LDS rollout (apg.py:103-122,152-167): dlc_lds_lqr_grad_f32 against torch.autograd
    on a float64 restatement of the same scan."""
    from deluca_b200.agents import LQR
    from deluca_b200.agents._gpc import quad_loss
    from deluca_b200.envs import LDS
    from deluca_b200.fused import gaussian_disturbance, lds_lqr_value_and_grad
    from deluca_b200.training.apg import value_and_grad_apg_loss
    rng = np.random.default_rng(2)
    N, d, m, T = 37, 10, 3, 80
    A, B, _ = O.lds_default_params(rng, N, d, m)
    K = O.lqr_gain(A, B) + 0.05 * rng.standard_normal((N, m, d)).astype(np.float32)
    x0 = rng.standard_normal((N, d)).astype(np.float32)
    W = gaussian_disturbance(N, T, d, seed=4).cpu().numpy()

    def torch_loss(Kt, x0t):
        At, Bt, Wt = (torch.tensor(a, dtype=torch.float64) for a in (A, B, W))
        x, tot = x0t, 0
        for t in range(T):
            u = -torch.einsum("nij,nj->ni", Kt, x)
            tot = tot + (x * x).sum(1) + (u * u).sum(1)
            x = torch.einsum("nij,nj->ni", At, x) + torch.einsum("nij,nj->ni", Bt, u) + Wt[t]
        return tot

    Kt = torch.tensor(K, dtype=torch.float64, requires_grad=True)
    xt = torch.tensor(x0, dtype=torch.float64, requires_grad=True)
    lt = torch_loss(Kt, xt)
    lt.sum().backward()
    loss, gK, gx, sums = lds_lqr_value_and_grad(_dev(A), _dev(B), _dev(K), _dev(x0), T, W=_dev(W), want_x0_grad=True)
    np.testing.assert_allclose(loss.cpu().numpy(), lt.detach().numpy(), rtol=1e-4)
    gref = Kt.grad.numpy()
    assert np.abs(gK.cpu().numpy() - gref).max() <= 1e-3 * np.abs(gref).max()
    assert np.abs(gx.cpu().numpy() - xt.grad.numpy()).max() <= 1e-3 * np.abs(xt.grad.numpy()).max()
    np.testing.assert_allclose(sums[1:].cpu().numpy().reshape(m, d), gref.sum(0), rtol=1e-3, atol=1e-3 * np.abs(gref.sum(0)).max())
    # in-kernel disturbances == materialised stream
    l2 = lds_lqr_value_and_grad(_dev(A), _dev(B), _dev(K), _dev(x0), T, W=None, seed=4)[0]
    assert torch.equal(l2, loss)
    # mirror: shared dynamics + shared K through the Env/Agent objects
    env = LDS()
    env.init(d_in=m, d_hidden=d, d_out=d, A=A[0], B=B[0], C=np.eye(d), key=4)
    agent = LQR(env.A, env.B)
    st, obs = env.reset(N, x0=x0)
    value, grad = value_and_grad_apg_loss(agent, env, st, obs, None, T, quad_loss)
    Ks = torch.tensor(agent.K.cpu().numpy(), dtype=torch.float64, requires_grad=True)
    A, B = np.repeat(A[:1], N, 0), np.repeat(B[:1], N, 0)
    ls = torch_loss(Ks.expand(N, m, d), torch.tensor(x0, dtype=torch.float64)).sum() / (N * T)
    ls.backward()
    assert abs(float(value) - float(ls)) <= 1e-4 * float(ls)
    assert np.abs(grad.cpu().numpy() - Ks.grad.numpy()).max() <= 2e-3 * np.abs(Ks.grad.numpy()).max() + 1e-5


def test_chunked_rollouts_continue_the_noise_stream():
    """the ENV's step counter keys the in-kernel disturbance stream.  Two T/2 chunks must equal one T
    rollout for LQR, and for GPC also when the agent is reset mid-way (fresh agent, continued env): the chunks see
    noise steps [0, T/2) and [T/2, T), never [0, T/2) twice."""
    from deluca_b200.agents import GPC, LQR
    from deluca_b200.agents._gpc import quad_loss
    from deluca_b200.envs import LDS
    from deluca_b200.fused import gaussian_disturbance
    from deluca_b200.training.apg import apg_parallel_rollouts
    rng = np.random.default_rng(3)
    N, d, m, T = 40, 10, 3, 80
    A, B, _ = O.lds_default_params(rng, N, d, m)
    env = LDS()
    env.init(d_in=m, d_hidden=d, d_out=d, A=A, B=B, C=np.eye(d), key=21)
    lqr = LQR(env.A, env.B)
    st, obs = env.reset(N)
    whole, _ = apg_parallel_rollouts(env, st, obs, lqr, None, T, quad_loss)
    a, mid = apg_parallel_rollouts(env, st, obs, lqr, None, T // 2, quad_loss)
    assert mid.t == T // 2
    b, end = apg_parallel_rollouts(env, mid, None, lqr, None, T // 2, quad_loss)
    assert end.t == T
    assert torch.equal(torch.cat([a.losses, b.losses], 1), whole.losses)
    assert not torch.equal(a.losses, b.losses)
    # GPC: agent reset mid-way on the continued env == the oracle fed the continuous stream W[0:T]
    gpc = GPC(env.A, env.B, H=3, HH=10, lr_scale=1e-4)
    g1, mid = apg_parallel_rollouts(env, st, obs, gpc, gpc.init(N), T // 2, quad_loss)
    g2, _ = apg_parallel_rollouts(env, mid, None, gpc, gpc.init(N), T // 2, quad_loss)
    W = gaussian_disturbance(N, T, d, seed=21).cpu().numpy()
    K = gpc.K.cpu().numpy()
    r1 = O.rollout_lds(A, B, K, np.zeros((N, d), np.float32), W[:T // 2], "gpc", H=3, HH=10, lr_scale=1e-4,
                       dtype=np.float64)
    r2 = O.rollout_lds(A, B, K, r1["x_final"], W[T // 2:], "gpc", H=3, HH=10, lr_scale=1e-4, dtype=np.float64)
    ref = np.concatenate([r1["costs"], r2["costs"]], 0)
    got = torch.cat([g1.losses, g2.losses], 1).T.cpu().numpy()
    rel = np.abs(got - ref) / np.maximum(ref, 1e-2 * ref.mean())
    assert rel.max() < 1e-4


def test_agents_built_from_numpy_inputs():
    """GPC / BPC / LQR take numpy (CPU) A, B, K the way the reference is called and keep them on the
    device; the first call must not fail with 'must be a CUDA tensor'."""
    from deluca_b200.agents import BPC, GPC, LQR
    rng = np.random.default_rng(4)
    d, m = 10, 3
    A, B, _ = O.lds_default_params(rng, 1, d, m)
    x = (0.1 * rng.standard_normal((d, 1))).astype(np.float32)
    Kref = O.lqr_gain(A, B)[0]
    for agent in (LQR(A[0], B[0]), GPC(A[0], B[0], H=3, HH=10), BPC(A[0], B[0], H=3)):
        assert agent.K.is_cuda and agent.A.is_cuda
        np.testing.assert_allclose(agent.K.cpu().numpy(), Kref, rtol=1e-4, atol=1e-6)
        u = agent(x, 0.0) if isinstance(agent, BPC) else agent(x)
        assert u.is_cuda and u.shape == (m, 1) and torch.isfinite(u).all()
    np.testing.assert_allclose(LQR(A[0], B[0])(x).cpu().numpy(), -Kref @ x, rtol=1e-5, atol=1e-6)
