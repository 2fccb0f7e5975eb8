"""``run_controller_scan`` (``deluca/lung/utils/scripts/run_controller.py:144-220``) for a batch of
breaths: the whole ``lax.scan(loop_over_tt)`` is one launch of ``dlc_lung_pid_rollout_f32``.

Batching follows ``jax.vmap`` over the leaves: ``controller.params`` [N,3] and/or a waveform whose
``pip`` / ``peep`` are tensors [N] give N independent episodes; series come back as [T] for a single
episode (the reference's shape) or [N,T] for a batch (vmap puts the batch axis first).
"""
import torch

from ...controllers._pid import PID
from ...core import BreathWaveform
from ...envs._balloon_lung import BalloonLung
from ... import _launch


def run_controller_scan(controller, T=1000, dt=0.03, abort=60, env=None, waveform=None, use_tqdm=False,
                        directory=None, init_controller=False, N=None, materialize=False):
    if not isinstance(controller, PID):
        raise TypeError("the fused episode kernel drives the lung PID controller; got "
                        f"{type(controller).__name__}")
    env = env or BalloonLung.create()
    waveform = waveform or BreathWaveform.create()
    ctrl_wave = waveform if init_controller else BreathWaveform.create()   # :162-165
    dev = env.device
    n_k = controller.params.shape[0]
    kf_n = ctrl_wave.knots(device="cpu")[1]
    n_w = kf_n.shape[0] if kf_n.dim() == 2 else 1
    N = N or max(n_k, n_w)
    env_state, obs = env.reset()                                          # :178
    # Every episode starts from env.reset(), so `timestamp` depends on the step only and `flow` is the env's static
    # field: for a batch they are computed once (a one-breath launch of the same kernel: bit-identical arithmetic) and
    # returned as broadcast VIEWS [N,T] instead of being written N times (16 B per env-step instead of 24).
    # `materialize=True` writes all six series per breath, as round 1 did.
    lean = N > 1 and not materialize
    kw = dict(env_state=env.state_dict(env_state), obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time),
              pid_decay=controller.dt / (controller.dt + controller.RC), run_dt=dt)
    b, out = _launch.launch(env, ctrl_wave, N, T, "closed_loop", dev, K=controller.params, lean=lean, **kw)
    if lean:
        one_wave = ctrl_wave if ctrl_wave.knots(device="cpu")[1].dim() == 1 else BreathWaveform.create()
        first = lambda v: v.reshape(-1)[:1] if isinstance(v, torch.Tensor) and v.numel() > 1 else v
        kw1 = dict(kw, env_state={k: first(v) for k, v in kw["env_state"].items()},
                   obs={k: first(v) for k, v in kw["obs"].items()})
        if "in_history" in kw1["env_state"]:      # DelayLung: [N, delay] histories -> the first breath's
            for k in ("in_history", "out_history"):
                h = kw["env_state"][k]
                kw1["env_state"][k] = h.reshape(-1, env.delay)[:1] if isinstance(h, torch.Tensor) else h
        _, col = _launch.launch(env, one_wave, 1, T, "closed_loop", dev, K=controller.params[:1], **kw1)
        out["timestamp"] = col["timestamp"].expand(T, N)
        out["flow"] = col["flow"].expand(T, N)
    single = N == 1
    shape = (lambda s: s[:, 0]) if single else (lambda s: s.T)
    ts = {k: shape(out[k]) for k in ("timestamp", "pressure", "flow", "u_in", "u_out")}
    if init_controller or waveform is ctrl_wave:
        ts["target"] = shape(out["target"])                               # waveform.at(timestamps), :207
    else:   # the reference evaluates the *caller's* waveform here even when the controller got the default
        ts["target"] = _target_series(env, waveform, out["timestamp"], N, dev, shape)
    result = {"timeseries": ts, "waveform": waveform, "status": b["status_out"],
              "final": {"env": env.state_from(b, single), "buffers": b}}
    if directory is not None:
        import datetime
        import os
        import pickle
        os.makedirs(directory, exist_ok=True)
        stamp = datetime.datetime.now().strftime("%Y-%m-%dT%H-%M-%S")
        with open(f"{directory}/{stamp}.pkl", "wb") as f:
            pickle.dump({"timeseries": {k: v.cpu() for k, v in ts.items()}}, f)
    return result


def _target_series(env, waveform, timestamps, N, dev, shape):
    """waveform.at(timestamps) through the kernel's interpolator: an env-only pass whose emitted
    `target` series is evaluated at the given timestamps is not available, so evaluate it with the
    controller-only mode: obs.time = timestamp gives target - pressure = err, and p = err with
    pressure = 0."""
    T = timestamps.shape[0]
    cols = []
    for t in range(T):
        b, _ = _launch.launch(env, waveform, N, 1, "ctrl_only", dev,
                              obs=dict(predicted_pressure=0.0, time=timestamps[t]), series=False)
        cols.append(b["pid_p_io"])
    return shape(torch.stack(cols))
