"""Expiratory valve controller (``deluca/lung/controllers/_expiratory.py:25-45``)."""
import torch

from ...core import field
from ..core import BreathWaveform, Controller, ControllerState
from .. import _launch


class Expiratory(Controller):
    waveform: BreathWaveform = field(None, jaxed=False)
    device: str = field("cuda", jaxed=False)

    def setup(self):
        if self.waveform is None:
            self.waveform = BreathWaveform.create()

    def __call__(self, state, obs, *args, **kwargs):
        """u_out = 0 while inhaling else 1, plus the time bookkeeping (``:35-45``)."""
        from ..envs._balloon_lung import BalloonLung, _batch
        leaves = [state.time, state.steps, state.dt, obs.time]
        N = _batch(*leaves)
        scalar = not any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in leaves)
        env = BalloonLung.create(device=self.device)
        b, out = _launch.launch(env, self.waveform, N, 1, "ctrl_only", self.device,
                                obs=dict(predicted_pressure=0.0, time=obs.time),
                                exp_state=dict(time=state.time, steps=state.steps, dt=state.dt))
        g = (lambda t: float(t[0])) if scalar else (lambda t: t)
        gi = (lambda t: int(t[0])) if scalar else (lambda t: t)
        new = ControllerState(time=g(b["exp_time_io"]), steps=gi(b["exp_steps_io"]), dt=g(b["exp_dt_io"]))
        return new, out["u_out"][0]
