"""``deluca/lung/utils/scripts/train_controller.py``: the differentiable episode and its gradient.

``rollout`` / ``rollout_parallel`` (``:33-92``) return the inhale-only tracking loss of a lung PID
controller on BalloonLung or DelayLung; ``value_and_grad_rollout_parallel`` is what
``jax.value_and_grad(rollout_parallel)`` (``:151-153``) yields for the controller's ``params``
leaf.  Value and gradient come from one launch of ``dlc_lung_pid_grad_f32`` (forward mode: the three gains'
tangents ride along the episode).  The optimiser loop around it (optax, ``:116-186``) is out of scope.
"""
import torch

from ....core import Obj
from ... import _launch
from ...controllers._pid import PID
from ...core import DEFAULT_DT, BreathWaveform
from ...envs._balloon_lung import BalloonLung
from ...envs._delay_lung import DelayLung
from .... import _abi, fused


def l1_loss(x, y):
    """The reference's default ``loss_fn = lambda x, y: (jnp.abs(x - y)).mean()`` (``:106``)."""
    return (x - y).abs().mean()


def l2_loss(x, y):
    return ((x - y) ** 2).mean()


_LOSS = {l1_loss: 0, l2_loss: 1}


def _episode(controller, sim, tt, use_noise, peep, pips, loss_fn, noise_value):
    if not isinstance(controller, PID) or not isinstance(sim, (BalloonLung, DelayLung)):
        raise NotImplementedError("the fused gradient kernel covers the lung PID controller on BalloonLung / DelayLung")
    if loss_fn not in _LOSS:
        raise NotImplementedError("loss_fn must be train_controller.l1_loss (the reference default) or l2_loss")
    dev = sim.device
    pips = torch.as_tensor(pips, dtype=torch.float32).reshape(-1)
    N = max(pips.numel(), controller.params.shape[0])
    wave = BreathWaveform.create(peep=peep, pip=pips.expand(N) if pips.numel() == 1 and N > 1 else pips)
    tt = torch.as_tensor(tt, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    p = _abi.DlcLungParams()
    p.N, p.T, p.mode = N, tt.numel(), 0
    kx, kf, period, xp_in = wave.knots(N=N, device=dev)
    p.nk = len(kx)
    for i, v in enumerate(kx):
        p.kx[i] = v
    p.kf_per_env = int(kf.dim() == 2)
    p.period, p.xp_in, p.default_dt = period, xp_in, DEFAULT_DT
    p.pid_decay = controller.dt / (controller.dt + controller.RC)
    _launch.sim_params(sim, p)
    K = controller.params.to(dev).expand(N, 3).contiguous()
    noise = float(use_noise) * float(noise_value)
    reset_pressure = sim.reset_normalized_peep if isinstance(sim, BalloonLung) else 0.0     # _delay_lung.py:62-74
    return fused.lung_pid_value_and_grad(p, K, kf, tt, _LOSS[loss_fn], noise, reset_pressure), N


def rollout(controller, sim, tt, use_noise, peep, pip, loss_fn, loss=0.0, noise_value=1.5):
    """``rollout`` (``:33-67``) for one pip (or a tensor of pips -> per-breath losses).  ``noise_value``
    stands for ``mean + std * normal(PRNGKey(0))`` (``:47-50``), a constant the reference redraws
    identically every step; Threefry is not reproducible here, so it is an argument."""
    (l, _, _), N = _episode(controller, sim, tt, use_noise, peep, pip, loss_fn, noise_value)
    l = l + loss
    return l[0] if N == 1 else l


def rollout_parallel(controller, sim, tt, use_noise, peep, pips, loss_fn, noise_value=1.5):
    """``rollout_parallel`` (``:70-92``): sum over pips / len(pips)."""
    (_, _, sums), N = _episode(controller, sim, tt, use_noise, peep, pips, loss_fn, noise_value)
    return sums[0] / N


def value_and_grad_rollout_parallel(controller, sim, tt, use_noise, peep, pips, loss_fn, noise_value=1.5):
    """``jax.value_and_grad(rollout_parallel)`` w.r.t. ``controller.params`` (``:151-153``): returns
    (value, grad[3]) when the gains are shared by all pips, (value, grad[N,3]) when they are per breath."""
    (_, g, sums), N = _episode(controller, sim, tt, use_noise, peep, pips, loss_fn, noise_value)
    value = sums[0] / N
    if controller.params.shape[0] == 1:
        return value, (sums[1:] / N).float()
    return value, g / N
