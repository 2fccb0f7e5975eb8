"""Builds the ``DlcLungParams`` / ``DlcLungBuffers`` of one ``dlc_lung_pid_rollout_f32`` call from the
mirror objects.  Host-side packing only; every number comes out of the kernel."""
import math

import torch

from .. import _abi, fused
from .core import DEFAULT_DT, BreathWaveform, f32

MODE = {"closed_loop": 0, "ctrl_only": 1, "env_only": 2}


def _vec(v, N, dev, dtype=torch.float32):
    """Leaf -> contiguous [N] tensor of ``dtype`` on ``dev`` (scalars are broadcast)."""
    t = torch.as_tensor(v, dtype=dtype, device=dev) if not isinstance(v, torch.Tensor) else v.to(dev, dtype)
    return t.reshape(-1).expand(N).contiguous().clone() if t.numel() in (1, N) else _bad(v, N)


def _bad(v, N):
    raise ValueError(f"leaf with {getattr(v, 'shape', None)} does not match batch size {N}")


def sim_params(env, p):
    """Static simulator fields -> params (float32 rounding mirrors JAX weak typing)."""
    from .envs._balloon_lung import BalloonLung
    from .envs._delay_lung import DelayLung
    if isinstance(env, BalloonLung):
        p.sim = 0
        r0 = (3.0 * env.min_volume / (4.0 * math.pi)) ** (1.0 / 3.0)       # _balloon_lung.py:76
        p.leak, p.peep_valve, p.PC, p.P0, p.R = int(bool(env.leak)), env.peep_valve, env.PC, env.P0, env.R
        p.dt, p.min_volume, p.r0, p.r0_sq = env.dt, env.min_volume, r0, r0 ** 2.0
        p.leak_coef = env.dt / (5.0 + env.dt)
        p.delay = 1
    elif isinstance(env, DelayLung):
        p.sim = 1
        r0 = (3 * env.min_volume / (4 * math.pi)) ** (1 / 3)               # _delay_lung.py:59
        p.dt, p.min_volume, p.r0, p.r0_sq = env.dt, env.min_volume, r0, r0 ** 2
        p.delay, p.d_R, p.d_C, p.inertia, p.control_gain = env.delay, env.R, env.C, env.inertia, env.control_gain
    else:
        raise TypeError(f"no fused kernel for simulator {type(env).__name__}")
    p.env_flow = float(env.flow)


def launch(env, waveform, N, T, mode, dev, **kw):
    """prepare() + one kernel call.  Returns (dict of updated leaves, dict of [T,N] series)."""
    p, b, out = prepare(env, waveform, N, T, mode, dev, **kw)
    fused.lung_rollout(p, b)
    return b, out


def prepare(env, waveform, N, T, mode, dev, *, K=None, env_state=None, obs=None, pid_state=None, exp_state=None,
           actions=None, pid_decay=DEFAULT_DT / (DEFAULT_DT + 0.5), run_dt=DEFAULT_DT, series=True,
           kernel_variant=0, lean=False):
    """Pack one kernel call.  ``env_state`` / ``obs`` / ``pid_state`` / ``exp_state`` are dicts of
    leaves (scalars or [N] tensors); missing pieces get neutral values.  Returns (params, buffers,
    dict of [T,N] series tensors)."""
    p = _abi.DlcLungParams()
    p.N, p.T, p.mode, p.kernel_variant = N, T, MODE[mode], int(kernel_variant)
    waveform = waveform or BreathWaveform.create()
    kx, kf, period, xp_in = waveform.knots(N=N, device=dev)
    p.nk = len(kx)
    for i, v in enumerate(kx):
        p.kx[i] = v
    p.kf_per_env = int(kf.dim() == 2)
    p.period, p.xp_in, p.pid_decay, p.default_dt, p.run_dt = period, xp_in, pid_decay, DEFAULT_DT, run_dt
    sim_params(env, p)
    es, ob, ps, xs = (dict(d or {}) for d in (env_state, obs, pid_state, exp_state))
    inf = float("inf")
    b = dict.fromkeys(_abi.LUNG_BUFFER_FIELDS)
    Kt = torch.zeros(1, 3) if K is None else torch.as_tensor(K, dtype=torch.float32).reshape(-1, 3)
    if Kt.shape[0] not in (1, N):
        raise ValueError(f"PID gains describe {Kt.shape[0]} controllers, expected {N}")
    b["K"] = Kt.to(dev).expand(N, 3).contiguous()
    b["kf"] = kf
    b["steps_io"] = _vec(es.get("steps", 0.0), N, dev)
    b["time_io"] = _vec(es.get("time", 0.0), N, dev)
    b["volume_io"] = _vec(es.get("volume", env.min_volume), N, dev)
    b["pressure_io"] = _vec(es.get("predicted_pressure", 0.0), N, dev)
    b["target_io"] = _vec(es.get("target", 0.0), N, dev)
    if p.sim == 1:
        b["pipe_pressure_io"] = _vec(es.get("pipe_pressure", 0.0), N, dev)
        for k in ("in_history", "out_history"):
            h = es.get(k)
            h = torch.zeros(N, env.delay, device=dev) if h is None else torch.as_tensor(
                h, dtype=torch.float32, device=dev).reshape(-1, env.delay).expand(N, env.delay).contiguous().clone()
            b[k + "_io"] = h
    b["obs_pressure_io"] = _vec(ob.get("predicted_pressure", 0.0), N, dev)
    b["obs_time_io"] = _vec(ob.get("time", 0.0), N, dev)
    for k, dflt in (("p", 0.0), ("i", 0.0), ("d", 0.0), ("time", inf), ("dt", DEFAULT_DT)):
        b[f"pid_{k}_io"] = _vec(ps.get(k, dflt), N, dev)
    b["pid_steps_io"] = _vec(ps.get("steps", 0), N, dev, torch.int32)
    b["exp_time_io"] = _vec(xs.get("time", inf), N, dev)
    b["exp_dt_io"] = _vec(xs.get("dt", DEFAULT_DT), N, dev)
    b["exp_steps_io"] = _vec(xs.get("steps", 0), N, dev, torch.int32)
    if mode == "env_only":
        u_in, u_out = actions
        b["u_in_in"] = torch.as_tensor(u_in, dtype=torch.float32, device=dev).reshape(T, N).contiguous()
        b["u_out_in"] = torch.as_tensor(u_out, dtype=torch.float32, device=dev).reshape(T, N).contiguous()
    out = {}
    if series:
        # lean: the two series that are the same for every breath of a batch with a common start time (the timestamp
        # depends on the step only, flow is the env's static field) are not written per breath -- 16 B per env-step
        # instead of 24; the caller keeps one column of each (run_controller_scan does)
        names = ("u_in", "u_out", "pressure", "target") if lean else ("timestamp", "u_in", "u_out", "pressure", "flow",
                                                                      "target")
        for k in names:
            out[k] = torch.empty(T, N, device=dev)
            b[k + "_out"] = out[k]
    b["status_out"] = torch.empty(N, dtype=torch.int32, device=dev)
    return p, b, out
