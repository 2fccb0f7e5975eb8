"""``train_controller.rollout`` / ``rollout_parallel`` (deluca/lung/utils/scripts/train_controller.py:33-92)
restated in differentiable torch (CPU, float64 or float32) -- TEST INFRASTRUCTURE (see
oracle/__init__.py).  ``torch.autograd`` then plays the role of ``jax.value_and_grad`` (``:151-153``)
and is the oracle for the hand-derived adjoint in ``dlc_lung_pid_grad_f32``.

Pinned to the reference's own files: ``tests/golden/ref_lung_train.npz`` holds ``rollout`` and
``jax.value_and_grad(rollout_parallel)`` executed literally (tests/golden/make_ref_fixtures.py, family ``lung_train``)
and ``tests/test_ref_fixtures_oracle.py`` holds this restatement to them.
PID (controllers/_pid.py:77-102) + Expiratory (_expiratory.py:35-45) + BalloonLung
(_balloon_lung.py:100-146) or DelayLung (_delay_lung.py:75-133); loss += loss_fn(waveform.at(tt[i]), pressure)
while u_out == 0.
"""
import math

import numpy as np
import torch

from . import lung as OL


def _interp(x, kx, kf):
    """jnp.interp on the period-padded table for a tensor of already-reduced abscissae x [N]."""
    i = torch.clamp(torch.searchsorted(kx, x.detach().contiguous(), right=True), 1, len(kx) - 1)
    n = torch.arange(kf.shape[0])
    f0, f1 = kf[n, i - 1], kf[n, i]
    dx = kx[i] - kx[i - 1]
    val = f0 + (x - kx[i - 1]) / torch.where(dx == 0, torch.ones_like(dx), dx) * (f1 - f0)
    # `.astype(self.dtype)` with dtype = float32 (lung/core.py:45,60): the target is rounded to float32 even under x64
    return torch.where(dx == 0, f0, val).float().to(x.dtype)


def rollout_loss(K, pip, peep, tt, loss="l1", use_noise=0.0, noise=0.0, dtype=torch.float64, sim_kwargs=None,
                 sim="balloon"):
    """Per-breath loss[N] of ``rollout`` for PID gains K[N,3] (a leaf; may require grad)."""
    if sim == "delay":
        return _rollout_loss_delay(K, pip, peep, tt, loss, use_noise, noise, dtype, sim_kwargs)
    kw = dict(leak=False, peep_valve=5.0, PC=40.0, P0=0.0, R=15.0, dt=0.03, min_volume=1.5)
    kw.update(sim_kwargs or {})
    npd = np.float64 if dtype == torch.float64 else np.float32
    kx, kf, period = OL.waveform_knots(pip, peep, dtype=npd)   # float32 (the reference's precision) folds 3 - 1e-8 to 3
    kx, kf = torch.tensor(kx.astype(npd)), torch.tensor(kf.astype(npd))
    N = K.shape[0]
    kf = kf.expand(N, kf.shape[1])
    period = float(period)
    c = lambda v: torch.tensor(v, dtype=dtype)
    r0 = (3.0 * kw["min_volume"] / (4.0 * math.pi)) ** (1.0 / 3.0)
    decay = c(0.03 / (0.03 + 0.5))
    V = torch.full((N,), kw["min_volume"], dtype=dtype)
    P = torch.zeros(N, dtype=dtype)
    time = torch.zeros(N, dtype=dtype)
    p = i = d = torch.zeros(N, dtype=dtype)
    total = torch.zeros(N, dtype=dtype)
    for step in range(len(tt)):
        P = P + use_noise * noise                                          # :51-53 (noise also enters the state)
        target = _interp(torch.remainder(time, period), kx, kf)           # PID: waveform.at(obs.time)
        err = target - P
        np_, ni, nd = err, i + decay * (err - i), d + decay * (err - p - d)
        p, i, d = np_, ni, nd
        u_in = torch.clamp(p * K[:, 0] + i * K[:, 1] + d * K[:, 2], 0.0, 100.0)
        u_out = (torch.remainder(time, period) > 1.0).to(dtype)            # Expiratory: 0 while inhaling
        pressure_obs = P
        # BalloonLung.__call__
        valve = torch.clamp(torch.tanh(0.03 * (3.0 * u_in - 130.0)) + 1.0, 0.0, 1.72)
        flow = torch.clamp(valve * kw["R"], 0.0, 2.0)
        flow = flow - torch.where(P > kw["peep_valve"], (u_out > 0).to(dtype) * 0.05 * P, torch.zeros_like(P))
        V = V + flow * kw["dt"]
        if kw["leak"]:
            V = V + kw["dt"] / (5.0 + kw["dt"]) * (kw["min_volume"] - V)
        r = (3.0 * V / (4.0 * math.pi)) ** (1.0 / 3.0)
        P = kw["P0"] + kw["PC"] * (1.0 - (r0 / r) ** 6) / (r0 ** 2 * r)
        time = time + kw["dt"]
        tgt_loss = _interp(torch.remainder(torch.full((N,), float(tt[step]), dtype=dtype), period), kx, kf)
        diff = tgt_loss - pressure_obs
        term = diff.abs() if loss == "l1" else diff * diff                 # :62-67, loss_fn = |x - y|
        total = total + torch.where(u_out == 0, term, torch.zeros_like(term))
    return total


def _rollout_loss_delay(K, pip, peep, tt, loss, use_noise, noise, dtype, sim_kwargs):
    """``rollout`` on DelayLung.  What differs from BalloonLung: the observation is the PRE-dynamics state
    (``_delay_lung.py:128-131``), so the controllers' clock lags one step; the action enters through
    ``in_history[0]``, which is the action just stored (``:120-123``), gated by ``state.time < self.delay`` -- seconds
    against a step count (``:88-96``); the clocks accumulate in ``dtype`` (float32 is the reference's arithmetic:
    where ``time`` crosses ``delay`` / the inhale end depends on it)."""
    kw = dict(min_volume=1.5, R=10.0, C=6.0, delay=25, inertia=0.995, control_gain=0.02, dt=0.03)
    kw.update(sim_kwargs or {})
    npd = np.float64 if dtype == torch.float64 else np.float32
    kx, kf, period = OL.waveform_knots(pip, peep, dtype=npd)
    kx, kf = torch.tensor(kx.astype(npd)), torch.tensor(kf.astype(npd))
    N = K.shape[0]
    kf = kf.expand(N, kf.shape[1])
    period = float(period)
    r0 = (3.0 * kw["min_volume"] / (4.0 * math.pi)) ** (1.0 / 3.0)
    decay = torch.tensor(0.03 / (0.03 + 0.5), dtype=dtype)
    V = torch.full((N,), kw["min_volume"], dtype=dtype)
    P = torch.zeros(N, dtype=dtype)
    pipe = torch.zeros(N, dtype=dtype)
    time = torch.zeros(N, dtype=dtype)
    obs_t = torch.zeros(N, dtype=dtype)
    p = i = d = torch.zeros(N, dtype=dtype)
    total = torch.zeros(N, dtype=dtype)
    for step in range(len(tt)):
        P = P + use_noise * noise
        phase = torch.remainder(obs_t, period)
        err = _interp(phase, kx, kf) - P
        np_, ni, nd = err, i + decay * (err - i), d + decay * (err - p - d)
        p, i, d = np_, ni, nd
        u_in = torch.clamp(p * K[:, 0] + i * K[:, 1] + d * K[:, 2], 0.0, 100.0)
        u_out = (phase > 1.0).to(dtype)
        pressure_obs = P
        # DelayLung.__call__ / dynamics
        u_pos = torch.where(u_in > 0.0, u_in, torch.zeros_like(u_in))
        obs_t = time
        V = torch.maximum(V + (P / kw["R"]) * kw["dt"], torch.full_like(V, kw["min_volume"]))
        r = (3.0 * V / (4.0 * math.pi)) ** (1.0 / 3.0)
        lung_p = kw["C"] * (1.0 - (r0 / r) ** 6) / (r0 ** 2 * r)
        early = time < kw["delay"]
        impulse = torch.where(early, torch.zeros_like(u_pos), kw["control_gain"] * u_pos)
        peep_on = (~early) & (u_out != 0)
        pipe = kw["inertia"] * pipe + impulse
        P = torch.clamp(pipe - lung_p, min=0.0)
        pipe = torch.where(peep_on, pipe * 0.995, pipe)
        time = time + kw["dt"]
        tgt_loss = _interp(torch.remainder(torch.full((N,), float(tt[step]), dtype=dtype), period), kx, kf)
        diff = tgt_loss - pressure_obs
        term = diff.abs() if loss == "l1" else diff * diff
        total = total + torch.where(u_out == 0, term, torch.zeros_like(term))
    return total


def value_and_grad(K, pip, peep, tt, **kw):
    """(loss[N], dloss/dK[N,3]) by reverse-mode autodiff."""
    Kt = torch.tensor(np.asarray(K), dtype=kw.get("dtype", torch.float64), requires_grad=True)
    loss = rollout_loss(Kt, pip, peep, tt, **kw)
    loss.sum().backward()
    return loss.detach().numpy(), Kt.grad.numpy()
