"""Parity of the experience kernels (dlc_gae_f32, dlc_env_collect_f32) with the CPU oracle (B200 only).

GAE is float32 with every operation rounded separately in the reference's order, so the comparison
with the float32 oracle is BIT-EXACT; the float64 oracle bounds the rounding error (<= 1e-4)."""
import numpy as np
import pytest
import torch

from oracle import igpc as OG
from oracle import ppo as OP

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def _series(N, T, seed, masks=True):
    rng = np.random.default_rng(seed)
    f = lambda: rng.normal(size=(N, T)).astype(np.float32)
    mk = (rng.random((N, T)) > 0.1).astype(np.float32) if masks else None
    return f(), f(), f(), mk   # losses, values, log_probs, masks


@pytest.mark.parametrize("N,T", [(1, 1), (3, 29), (100, 29), (64, 7), (257, 32), (300, 33), (70, 130), (40, 200), (17, 64), (1000, 100), (513, 1000)])
@pytest.mark.parametrize("masks", [False, True])
def test_gae_batch_major_bit_exact(N, T, masks):
    from deluca_b200 import fused
    losses, values, logp, mk = _series(N, T, N + T, masks)
    adv, ret = fused.gae(_t(losses), _t(values), _t(logp), None if mk is None else _t(mk), 0.99, 0.95)
    o_adv, o_ret = OP.ac_tail(losses, logp, values, np.ones_like(losses) if mk is None else mk, 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), o_adv)
    np.testing.assert_array_equal(ret.cpu().numpy(), o_ret)
    a64, _ = OP.ac_tail(losses.astype(np.float64), logp.astype(np.float64), values.astype(np.float64),
                        np.ones(losses.shape) if mk is None else mk.astype(np.float64), 0.99, 0.95)
    np.testing.assert_allclose(adv.cpu().numpy(), a64, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("N,T", [(1, 5), (129, 29), (1000, 64)])
def test_gae_time_major_and_mirror(N, T):
    from deluca_b200 import fused
    from deluca_b200.training import ppo
    losses, values, logp, mk = _series(N, T, 5 * N + T)
    r_tm, v_tm, m_tm = -losses.T.copy(), values.T.copy(), mk.T.copy()
    adv = ppo.gae_advantages(_t(r_tm), _t(m_tm), _t(v_tm), 0.99, 0.95)     # reference signature, [T,N]
    np.testing.assert_array_equal(adv.cpu().numpy(), OP.gae_advantages(r_tm, m_tm, v_tm, 0.99, 0.95))
    adv2, ret2 = fused.gae(_t(losses.T.copy()), _t(v_tm), _t(logp.T.copy()), _t(m_tm), 0.99, 0.95, time_major=True,
                           returns_use_values=True)
    np.testing.assert_array_equal(adv2.cpu().numpy(), adv.cpu().numpy())
    np.testing.assert_array_equal(ret2.cpu().numpy(), adv.cpu().numpy() + v_tm)


def test_ac_get_experience_layout():
    from deluca_b200.training import ppo
    N, T = 28, 29          # ppo_run.py: batch of 28 envs, horizon 29
    losses, values, logp, _ = _series(N, T, 11, masks=False)
    actions = torch.randn(N, T, 1, device="cuda")
    ex = ppo.ac_get_experience(None, actions, _t(logp), _t(values), _t(losses), 0.99, 0.95)
    assert ex.actions.shape == (N * T, 1) and ex.advantages.shape == (N * T,) and ex.returns.shape == (N * T,)
    assert float(ex.done_masks.min()) == 1.0
    o_adv, o_ret = OP.ac_tail(losses, logp, values, np.ones_like(losses), 0.99, 0.95)
    np.testing.assert_array_equal(ex.advantages.cpu().numpy(), o_adv.reshape(-1))   # row n*T + t
    np.testing.assert_array_equal(ex.returns.cpu().numpy(), o_ret.reshape(-1))


def test_gae_full_size_property():
    """2^20 agents x T = 29: lambda = 1 without masks telescopes to discounted return - value."""
    from deluca_b200 import fused
    N, T, g = 1 << 20, 29, 0.9
    gen = torch.Generator(device="cuda").manual_seed(0)
    losses = torch.randn(N, T, device="cuda", generator=gen)
    values = torch.randn(N, T, device="cuda", generator=gen)
    adv, _ = fused.gae(losses, values, None, None, g, 1.0, want_returns=False)
    G = torch.zeros(N, T, device="cuda", dtype=torch.float64)
    acc = torch.zeros(N, device="cuda", dtype=torch.float64)
    for t in reversed(range(T)):
        acc = -losses[:, t].double() + g * acc
        G[:, t] = acc
    assert float((adv.double() - (G - values.double())).abs().max()) < 1e-4


def test_gae_rejects_bad_arguments():
    from deluca_b200 import fused
    x = torch.zeros(4, 5, device="cuda")
    with pytest.raises(ValueError):
        fused.gae(x, x)                       # returns need log_probs
    with pytest.raises(TypeError):
        fused.gae(x.cpu(), x.cpu(), x.cpu())  # no CPU path


def _pairs():
    from deluca_b200.envs.classic import Cartpole, Pendulum, PlanarQuadrotor, Reacher
    return [(OG.Pendulum(), Pendulum.create()), (OG.Cartpole(), Cartpole.create()),
            (OG.PlanarQuadrotor(wind=0.1), PlanarQuadrotor.create(wind=0.1)), (OG.Reacher(), Reacher.create())]


@pytest.mark.parametrize("which", [0, 1, 2, 3])
@pytest.mark.parametrize("N,T", [(1, 1), (37, 16), (200, 50)])
def test_generate_trajectories_matches_oracle(which, N, T):
    from deluca_b200.learners import Learner
    o_env, d_env = _pairs()[which]
    obs, act, nxt = Learner(d_env).generate_trajectories(N, T, key=123)
    assert obs.shape == (N, T, o_env.d) and act.shape == (N, T, o_env.m) and nxt.shape == obs.shape
    o_act = OP.random_actions(123, np.arange(N), T, o_env.m)
    np.testing.assert_allclose(act.cpu().numpy(), o_act, rtol=0, atol=5e-6)     # MUFU vs libm Box-Muller
    # step-by-step parity on the device's own (obs, action): one env step each, no chaotic drift
    x = obs.cpu().numpy().reshape(N * T, o_env.d).astype(np.float64)
    u = act.cpu().numpy().reshape(N * T, o_env.m).astype(np.float64)
    np.testing.assert_allclose(nxt.cpu().numpy().reshape(N * T, -1), o_env.step(x, u), rtol=1e-4, atol=1e-5)
    # the dataset chains: obs[t+1] == next_obs[t], obs[0] == init
    assert torch.equal(obs[:, 1:], nxt[:, :-1])
    np.testing.assert_array_equal(obs[:, 0].cpu().numpy(), o_env.init(N, np.float32))


def test_generate_trajectories_shard_invariance():
    from deluca_b200.envs.classic import Pendulum
    from deluca_b200.learners import Learner
    L = Learner(Pendulum.create())
    full = L.generate_trajectories(96, 40, key=5)
    part = L.generate_trajectories(32, 40, key=5, env_offset=64)
    for a, b in zip(full, part):
        assert torch.equal(a[64:], b)
