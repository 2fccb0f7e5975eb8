"""Parity of the fused ventilator episode kernel (PID + Expiratory + Balloon/DelayLung) and the lung
mirrors against the CPU oracle (B200 only).  Tolerance: 1e-4 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import lung as OL

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _ok(a, b, scale=None):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    scale = scale if scale is not None else max(np.abs(b).mean(), 1e-3)
    return np.abs(a - b) <= RTOL * np.maximum(np.abs(b), scale)


def _close(a, b, scale=None, env_axis=None, min_frac=1.0):
    """Every element within RTOL -- or, for long episodes, every element of at least ``min_frac`` of
    the envs: the simulators branch on `pressure > peep_valve` / `u > 0`, so a 1-ulp difference in
    tanhf/powf can flip a branch in a rare env and that env then follows another (valid) path."""
    ok = _ok(a, b, scale)
    if env_axis is None:
        assert ok.all(), np.abs((a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a) - b).max()
    else:
        per_env = ok.all(axis=tuple(i for i in range(ok.ndim) if i != env_axis))
        assert per_env.mean() >= min_frac, per_env.mean()


@pytest.mark.parametrize("sim,T", [("balloon", 29), ("balloon", 400), ("delay", 29), ("delay", 1000)])
def test_run_controller_scan_matches_oracle(sim, T):
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung, DelayLung
    from deluca_b200.lung.utils.scripts.run_controller import run_controller_scan
    rng = np.random.default_rng(0)
    N = 300
    K = rng.uniform(0, 5, (N, 3)).astype(np.float32)
    pip = rng.choice([10, 15, 20, 25, 30, 35], N).astype(np.float32)
    ref = OL.run_controller_scan(K, pip, 5.0, T=T, sim=sim, dtype=np.float32)
    env = (BalloonLung if sim == "balloon" else DelayLung).create()
    wave = BreathWaveform.create(pip=torch.tensor(pip), peep=5.0)
    res = run_controller_scan(PID.create(params=torch.tensor(K)), T=T, env=env, waveform=wave, init_controller=True)
    for k in ("timestamp", "pressure", "flow", "u_in", "u_out", "target"):
        got = res["timeseries"][k]
        assert got.shape == (N, T)
        # strict while no env has met a discontinuity: the whole default episode (T = 29) and the first
        # inhale (34 steps) of long ones.  From the first exhale on BalloonLung chatters around the
        # `pressure > peep_valve` switch (_balloon_lung.py:114-119), where 1-ulp differences between
        # CUDA and libm tanh/pow flip the branch; long episodes are then compared through the
        # reference's own episode score (test_controller.py:40-45).
        _close(got.T[:34], ref["timeseries"][k][:34])
    score = (res["timeseries"]["pressure"] - res["timeseries"]["target"]).abs().mean(1).cpu().numpy()
    score_ref = OL.test_controller_score(ref["timeseries"])
    assert (np.abs(score - score_ref) <= 2e-2 * np.maximum(score_ref, 1.0)).mean() >= 0.97
    b = res["final"]["buffers"]
    frac = 1.0 if T <= 29 else 0.0
    _close(b["volume_io"], ref["env_state"]["volume"], env_axis=0, min_frac=frac)
    _close(b["pressure_io"], ref["env_state"]["pressure"], env_axis=0, min_frac=frac)
    _close(b["time_io"], ref["env_state"]["time"])
    _close(b["pid_i_io"], ref["controller_state"]["i"], env_axis=0, min_frac=frac)
    assert (b["pid_steps_io"].cpu().numpy() == T).all()
    if sim == "delay":
        _close(b["in_history_io"], ref["env_state"]["in_history"], 1.0)
        _close(b["out_history_io"], ref["env_state"]["out_history"], 1.0)
    assert int(res["status"].sum()) == 0


def test_waveform_known_answers_on_device():
    """deluca/tests/lung/core_test.py:55-62 through the kernel's interpolator (ctrl-only step:
    p = target - pressure with pressure = 0)."""
    from deluca_b200.lung import _launch
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung
    ts = [0.0, 0.5, 1.0, 1.25, 1.5, 3.0, 2.0, 2.5]
    want = [35.0, 35.0, 35.0, 20.0, 5.0, 35.0, 15.0, 25.0]   # last two: the fp32 ramp (SURVEY A-9)
    b, out = _launch.launch(BalloonLung.create(), BreathWaveform.create(), len(ts), 1, "ctrl_only", "cuda",
                            obs=dict(predicted_pressure=0.0, time=torch.tensor(ts)))
    assert b["pid_p_io"].cpu().tolist() == want
    # is_in at 0.5 / 1.0 / 1.5 (core_test.py:43-47): u_out = 0 while inhaling
    assert out["u_out"][0, [1, 2, 4]].cpu().tolist() == [0.0, 0.0, 1.0]
    w = BreathWaveform.create()
    assert [w.at(t) for t in ts] == want and w.is_in(1.0) and not w.is_in(1.5)


def test_single_step_env_and_controller_calls_compose_to_the_fused_episode():
    """Env.__call__ / Controller.__call__ one step at a time == the fused scan."""
    from deluca_b200.lung.controllers import PID, Expiratory
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung
    from deluca_b200.lung.utils.scripts.run_controller import run_controller_scan
    N, T = 5, 12
    K = torch.tensor([[3.0, 4.0, 0.0]]).expand(N, 3).contiguous()
    wave = BreathWaveform.create(pip=torch.tensor([10.0, 15.0, 20.0, 25.0, 35.0]), peep=5.0)
    env = BalloonLung.create(waveform=wave)
    pid, exp = PID.create(params=K), Expiratory.create(waveform=wave)
    fusedr = run_controller_scan(pid, T=T, env=env, waveform=wave, init_controller=True)
    state, obs = env.reset(N)
    cs, es = pid.init(wave), exp.init()
    us, ps = [], []
    for _ in range(T):
        ps.append(obs.predicted_pressure)
        cs, u_in = pid(cs, obs)
        es, u_out = exp(es, obs)
        state, obs = env(state, (u_in, u_out))
        us.append(u_in)
    assert torch.equal(torch.stack(us, 1), fusedr["timeseries"]["u_in"])
    assert torch.equal(torch.stack(ps, 1), fusedr["timeseries"]["pressure"])


def test_generic_pid_agent_matches_oracle():
    from deluca_b200.agents import PID
    rng = np.random.default_rng(1)
    N, T = 64, 50
    kp, ki, kd = (rng.uniform(0, 3, N).astype(np.float32) for _ in range(3))
    err = rng.standard_normal((T, N)).astype(np.float32)
    agent = PID.create(K_P=torch.tensor(kp).cuda(), K_I=torch.tensor(ki).cuda(), K_D=torch.tensor(kd).cuda())
    st = agent.init()
    ost = dict(P=np.zeros(N, np.float32), I=np.zeros(N, np.float32), D=np.zeros(N, np.float32))
    for t in range(T):
        st, u = agent(st, torch.tensor(err[t]).cuda())
        uo = OL.generic_pid_call(ost, kp, ki, kd, err[t])
        _close(u, uo, 1.0)


def test_million_breaths_properties():
    """BASELINE config 4 size (2^20 breaths, T = 29): identical breaths give identical series, u_in
    stays inside the clamp, u_out follows is_in, and the emitted pressure is the previous step's state."""
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung
    from deluca_b200.lung.utils.scripts.run_controller import run_controller_scan
    N, T = 1 << 20, 29
    g = torch.Generator().manual_seed(0)
    K = (5 * torch.rand(N // 2, 3, generator=g)).repeat(2, 1)
    pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N // 2,), generator=g)].repeat(2)
    res = run_controller_scan(PID.create(params=K), T=T, env=BalloonLung.create(),
                              waveform=BreathWaveform.create(pip=pip, peep=5.0), init_controller=True)
    ts = res["timeseries"]
    assert ts["u_in"].shape == (N, T)
    assert torch.equal(ts["u_in"][: N // 2], ts["u_in"][N // 2:])
    assert float(ts["u_in"].min()) >= 0.0 and float(ts["u_in"].max()) <= 100.0
    assert float(ts["u_out"].abs().max()) == 0.0           # 29 steps of 0.03 s stay inside the inhale
    assert float(ts["pressure"][:, 0].abs().max()) == 0.0
    assert int(res["status"].sum()) == 0


@pytest.mark.parametrize("loss_name,use_noise", [("l1", False), ("l2", False), ("l1", True)])
def test_rollout_value_and_grad_matches_autograd(loss_name, use_noise):
    """Reverse mode through the episode (train_controller.py:33-92,151-153): the hand adjoint in
    dlc_lung_pid_grad_f32 against torch.autograd on the differentiable restatement (float64)."""
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.envs import BalloonLung
    from deluca_b200.lung.utils.scripts import train_controller as TC
    from oracle import lung_grad as OG
    rng = np.random.default_rng(3)
    N = 257
    K = rng.uniform(0.5, 4.0, (N, 3)).astype(np.float32)
    pips = rng.choice([10, 15, 20, 25, 30, 35], N).astype(np.float32)
    tt = np.linspace(0, 0.87, 29).astype(np.float32)
    noise = 1.5 if use_noise else 0.0
    lo, go = OG.value_and_grad(K, pips, 5.0, tt, loss=loss_name, use_noise=float(use_noise), noise=noise)
    loss_fn = TC.l1_loss if loss_name == "l1" else TC.l2_loss
    sim = BalloonLung.create()
    ctrl = PID.create(params=torch.tensor(K))
    l = TC.rollout(ctrl, sim, tt, use_noise, 5.0, torch.tensor(pips), loss_fn, noise_value=1.5)
    value, g = TC.value_and_grad_rollout_parallel(ctrl, sim, tt, use_noise, 5.0, torch.tensor(pips), loss_fn,
                                                  noise_value=1.5)
    np.testing.assert_allclose(l.cpu().numpy(), lo, rtol=1e-4, atol=1e-3)
    assert abs(float(value) - lo.mean()) <= 1e-4 * abs(lo.mean())
    got = g.cpu().numpy() * N
    # gradients flip sign where a clamp / the L1 kink is crossed within rounding; require the bulk
    scale = np.abs(go).mean(axis=0, keepdims=True)
    ok = (np.abs(got - go) <= 1e-3 * np.maximum(np.abs(go), scale)).all(axis=1)
    assert ok.mean() >= 0.97, ok.mean()
    # shared gains: the batch-mean gradient
    ctrl1 = PID.create(params=torch.tensor([3.0, 4.0, 0.0]))
    v1, g1 = TC.value_and_grad_rollout_parallel(ctrl1, sim, tt, False, 5.0, torch.tensor([10., 15, 20, 25, 30, 35]),
                                                TC.l1_loss)
    lo1, go1 = OG.value_and_grad(np.tile([3.0, 4.0, 0.0], (6, 1)), np.array([10., 15, 20, 25, 30, 35]), 5.0, tt)
    assert abs(float(v1) - lo1.mean()) <= 1e-4 * lo1.mean()
    np.testing.assert_allclose(g1.cpu().numpy(), go1.mean(0), rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("sim", ["balloon", "delay"])
def test_test_controller_score(sim):
    """``test_controller`` (test_controller.py:23-47): PIPs as the batch of one fused launch; shared gains give the
    reference's scalar, per-env gains one score per env."""
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.envs import BalloonLung, DelayLung
    from deluca_b200.lung.utils.scripts.test_controller import test_controller as score_fn
    pips = [10, 15, 20, 25, 30, 35]
    env = (BalloonLung if sim == "balloon" else DelayLung).create()
    K = np.array([[3.0, 4.0, 0.0], [1.0, 2.0, 0.5], [0.5, 0.1, 0.0]], np.float32)
    per_env = score_fn(PID.create(params=torch.tensor(K)), env, pips, 5.0)
    assert per_env.shape == (3,)
    for i in range(3):
        ref = OL.run_controller_scan(np.repeat(K[i:i + 1], len(pips), 0), np.array(pips, np.float32), 5.0, T=29, sim=sim,
                                     dtype=np.float32)
        want = OL.test_controller_score(ref["timeseries"]).mean()
        one = score_fn(PID.create(params=torch.tensor(K[i:i + 1])), env, pips, 5.0)
        assert one.dim() == 0
        assert abs(float(one) - want) <= 1e-4 * max(want, 1.0)
        assert abs(float(per_env[i]) - want) <= 1e-4 * max(want, 1.0)


@pytest.mark.parametrize("sim", ["balloon", "delay"])
@pytest.mark.parametrize("N,T", [(2050, 29), (4096, 137), (302, 1000)])
def test_two_breaths_per_thread_kernel_is_bit_identical(sim, N, T):
    """The hot instance runs two breaths per thread (float2 rows) from ~50-100 steps on; `kernel_variant` = 2 / 1
    forces the two- / one-per-thread kernel.  Same operations per env in the same order: every series and every carried leaf must be equal bit for
    bit, also across a chunked resume and with a partial last CTA.  An odd batch falls back by itself."""
    from deluca_b200 import fused
    from deluca_b200.lung import _launch
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung, DelayLung
    g = torch.Generator().manual_seed(N + T)
    K = 5 * torch.rand(N, 3, generator=g)
    pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
    wave = BreathWaveform.create(pip=pip, peep=5.0)
    env = (BalloonLung if sim == "balloon" else DelayLung).create()
    st, obs = env.reset()

    def run(variant, chunks):
        kw = dict(K=K, env_state=env.state_dict(st), obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time))
        outs, b = [], None
        for Tc in chunks:
            if b is not None:           # resume from the carried leaves
                kw = dict(K=K, env_state={k[:-3]: b[k] for k in ("steps_io", "time_io", "volume_io", "target_io")},
                          obs=dict(predicted_pressure=b["obs_pressure_io"], time=b["obs_time_io"]),
                          pid_state=dict(p=b["pid_p_io"], i=b["pid_i_io"], d=b["pid_d_io"], time=b["pid_time_io"],
                                         dt=b["pid_dt_io"], steps=b["pid_steps_io"]),
                          exp_state=dict(time=b["exp_time_io"], dt=b["exp_dt_io"], steps=b["exp_steps_io"]))
                kw["env_state"]["predicted_pressure"] = b["pressure_io"]
                if sim == "delay":
                    kw["env_state"].update(pipe_pressure=b["pipe_pressure_io"], in_history=b["in_history_io"],
                                           out_history=b["out_history_io"])
            p, b, series = _launch.prepare(env, wave, N, Tc, "closed_loop", "cuda", kernel_variant=variant, **kw)
            fused.lung_rollout(p, b)
            outs.append(series)
        cat = {k: torch.cat([o[k] for o in outs]) for k in outs[0]}
        return cat, b

    s0, b0 = run(2, [T])
    s1, b1 = run(1, [T])
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k
    for k, v in b0.items():
        if isinstance(v, torch.Tensor) and k.endswith(("_io", "_out")):
            assert torch.equal(v, b1[k]), k
    assert int((b0["status_out"] != 0).sum()) == 0
    s2, b2 = run(2, [T // 3, T - T // 3])            # chunked resume through the pair kernel (T // 3 < delay: old
                                                     # history entries survive the first chunk)
    for k in s0:
        assert torch.equal(s0[k], s2[k]), k
    assert torch.equal(b0["volume_io"], b2["volume_io"]) and torch.equal(b0["pid_i_io"], b2["pid_i_io"])


def test_odd_batch_takes_the_one_per_thread_kernel():
    from deluca_b200 import fused
    from deluca_b200.lung import _launch
    from deluca_b200.lung.envs import BalloonLung
    env = BalloonLung.create()
    st, obs = env.reset()
    N, T = 1001, 40
    K = torch.tensor([[1.0, 0.5, 0.1]]).expand(N, 3)
    kw = dict(K=K, env_state=env.state_dict(st), obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time))
    p, b, s = _launch.prepare(env, None, N, T, "closed_loop", "cuda", **kw)
    fused.lung_rollout(p, b)
    p2, b2, s2 = _launch.prepare(env, None, N - 1, T, "closed_loop", "cuda", **dict(kw, K=K[:-1]))
    fused.lung_rollout(p2, b2)
    assert torch.equal(s["pressure"][:, :-1], s2["pressure"]) and torch.equal(s["u_in"][:, :-1], s2["u_in"])


@pytest.mark.parametrize("sim", ["balloon", "delay"])
@pytest.mark.parametrize("T", [29, 130])
def test_lean_series_equal_materialised_series(sim, T):
    """For a batch `run_controller_scan` stores four series per breath and returns `timestamp` / `flow` as broadcast
    views of one column (they are the same for every breath that starts from `env.reset()`): every value must equal the
    all-six-series launch bit for bit."""
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung, DelayLung
    from deluca_b200.lung.utils.scripts.run_controller import run_controller_scan
    g = torch.Generator().manual_seed(T)
    N = 514
    K = 5 * torch.rand(N, 3, generator=g)
    pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
    wave = BreathWaveform.create(pip=pip, peep=5.0)
    env = (BalloonLung if sim == "balloon" else DelayLung).create()
    ctrl = PID.create(params=K)
    a = run_controller_scan(ctrl, T=T, env=env, waveform=wave, init_controller=True)
    b = run_controller_scan(ctrl, T=T, env=env, waveform=wave, init_controller=True, materialize=True)
    for k in ("timestamp", "pressure", "flow", "u_in", "u_out", "target"):
        assert a["timeseries"][k].shape == (N, T)
        assert torch.equal(a["timeseries"][k], b["timeseries"][k]), k
    assert a["timeseries"]["timestamp"].stride(0) == 0 and b["timeseries"]["timestamp"].stride(0) != 0
