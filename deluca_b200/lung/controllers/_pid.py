"""Lung PID controller (``deluca/lung/controllers/_pid.py:29-102``).

``params`` is the pytree leaf: the 3 -> 1 bias-free Dense kernel ``[K_P, K_I, K_D]`` (``:39-44``),
a tensor [3] or [N,3] for a batch of controllers.  The reference's default (a flax lecun-normal
draw under ``PRNGKey(0)``, ``:57-67``) is not reproducible without JAX, so ``params`` is required.
"""
import torch

from ...core import Obj, field
from ..core import DEFAULT_DT, BreathWaveform, Controller
from .. import _launch


class PIDControllerState(Obj):
    waveform: Obj = field(None, jaxed=False)
    p: float = 0.0
    i: float = 0.0
    d: float = 0.0
    time: float = float("inf")
    steps: int = 0
    dt: float = DEFAULT_DT


class PID(Controller):
    params: torch.Tensor = field(None, jaxed=True)
    RC: float = field(0.5, jaxed=False)
    dt: float = field(0.03, jaxed=False)
    device: str = field("cuda", jaxed=False)

    def setup(self):
        if self.params is None:
            raise ValueError("PID.create(params=[K_P, K_I, K_D]) is required (the reference's flax default init "
                             "cannot be reproduced without JAX)")
        if isinstance(self.params, dict):   # flax layout {"K": {"kernel": [3,1]}}
            self.params = self.params["K"]["kernel"]
        self.params = torch.as_tensor(self.params, dtype=torch.float32).reshape(-1, 3)

    def init(self, waveform=None):
        if waveform is None:
            waveform = BreathWaveform.create()
        return PIDControllerState(waveform=waveform)

    def state_dict(self, st):
        return dict(p=st.p, i=st.i, d=st.d, time=st.time, steps=st.steps, dt=st.dt)

    def state_from(self, st, b, scalar):
        g = (lambda t: float(t[0])) if scalar else (lambda t: t)
        gi = (lambda t: int(t[0])) if scalar else (lambda t: t)
        return PIDControllerState(waveform=st.waveform, p=g(b["pid_p_io"]), i=g(b["pid_i_io"]), d=g(b["pid_d_io"]),
                                  time=g(b["pid_time_io"]), steps=gi(b["pid_steps_io"]), dt=g(b["pid_dt_io"]))

    def __call__(self, controller_state, obs):
        """``__call__`` (``:77-102``): returns (state, u_in); u_in is [N] (or [1] for scalar leaves)."""
        from ..envs._balloon_lung import BalloonLung, _batch
        leaves = list(self.state_dict(controller_state).values()) + [obs.predicted_pressure, obs.time]
        N = max(_batch(*leaves), self.params.shape[0])
        scalar = N == 1 and not any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in leaves)
        env = BalloonLung.create(device=self.device)   # simulator half is skipped in ctrl_only mode
        b, out = _launch.launch(env, controller_state.waveform, N, 1, "ctrl_only", self.device, K=self.params,
                                obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time),
                                pid_state=self.state_dict(controller_state),
                                pid_decay=self.dt / (self.dt + self.RC))
        return self.state_from(controller_state, b, scalar), out["u_in"][0]
