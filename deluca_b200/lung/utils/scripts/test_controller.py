"""``test_controller`` (``deluca/lung/utils/scripts/test_controller.py:23-47``): the validation score of a lung
controller -- mean absolute tracking error over the default 29-step horizon, averaged over the PIP set.

The reference loops over ``pips`` in Python (one jitted scan each); here the PIPs are the batch axis of ONE fused
episode launch (``run_controller_scan`` with a waveform whose ``pip`` is a tensor), followed by the mean.  A
controller with per-env gains ``params[N,3]`` is scored per env: the result is then [N] (what ``jax.vmap`` of the
reference over the controller leaves would return); with shared gains it is the reference's scalar.
"""
import torch

from ...core import BreathWaveform
from .run_controller import run_controller_scan


def test_controller(controller, sim, pips, peep):
    horizon = 29                                                                 # :28
    pips = torch.as_tensor(pips, dtype=torch.float32).reshape(-1)
    n_p = pips.numel()
    n_k = controller.params.shape[0]
    if n_k == 1:
        ctrl, pip_b = controller, pips
    else:   # every (env, pip) pair is one episode: env-major batch of n_k * n_p breaths
        ctrl = controller.replace(params=controller.params.repeat_interleave(n_p, dim=0).contiguous())
        pip_b = pips.repeat(n_k)
    waveform = BreathWaveform.create(peep=peep, pip=pip_b)                       # :30
    result = run_controller_scan(ctrl, T=horizon, abort=horizon, env=sim, waveform=waveform,
                                 init_controller=True)                           # :31-38
    ts = result["timeseries"]
    err = (ts["pressure"] - ts["target"]).abs().reshape(-1, horizon).mean(dim=1)  # :40-44: mean over the 29 steps
    score = err.reshape(n_k, n_p).mean(dim=1)                                    # :45: average over the PIPs
    return score[0] if n_k == 1 else score


test_controller.__test__ = False   # not a pytest test
