"""DelayLung (``deluca/lung/envs/_delay_lung.py:23-133``): state leaves may carry a batch axis [N]."""
import torch

from ...core import Obj, field
from ..core import BreathWaveform, LungEnv
from .. import _launch
from ._balloon_lung import _batch


class Observation(Obj):
    predicted_pressure: float = 0.0
    time: float = 0.0


class SimulatorState(Obj):
    in_history: torch.Tensor = None
    out_history: torch.Tensor = None
    steps: int = 0
    time: float = 0.0
    volume: float = 0.0
    predicted_pressure: float = 0.0
    pipe_pressure: float = 0.0
    target: float = 0.0


class DelayLung(LungEnv):
    """Delay lung."""

    min_volume: float = field(1.5, jaxed=False)
    R: int = field(10, jaxed=False)
    C: int = field(6, jaxed=False)
    delay: int = field(25, jaxed=False)
    inertia: float = field(0.995, jaxed=False)
    control_gain: float = field(0.02, jaxed=False)
    dt: float = field(0.03, jaxed=False)
    waveform: BreathWaveform = field(None, jaxed=False)
    r0: float = field(None, jaxed=False)
    flow: float = field(0.0, jaxed=False)
    device: str = field("cuda", jaxed=False)

    def setup(self):
        if self.waveform is None:
            self.waveform = BreathWaveform.create()
        self.r0 = (3 * self.min_volume / (4 * 3.141592653589793)) ** (1 / 3)

    def reset(self, N=None):
        """``reset`` (``:61-73``)."""
        tgt = self.waveform.at(0.0)
        n = 1 if N is None else N
        z = torch.zeros(n, self.delay, device=self.device)
        if N is None:
            st = SimulatorState(steps=0, time=0.0, volume=self.min_volume, predicted_pressure=0.0, pipe_pressure=0.0,
                                target=tgt, in_history=z[0], out_history=z[0].clone())
            return st, Observation(predicted_pressure=0.0, time=0.0)
        f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=self.device).reshape(-1).expand(N).contiguous()
        st = SimulatorState(steps=f(0.0), time=f(0.0), volume=f(self.min_volume), predicted_pressure=f(0.0),
                            pipe_pressure=f(0.0), target=f(tgt), in_history=z, out_history=z.clone())
        return st, Observation(predicted_pressure=f(0.0), time=f(0.0))

    def state_dict(self, state):
        return dict(steps=state.steps, time=state.time, volume=state.volume,
                    predicted_pressure=state.predicted_pressure, pipe_pressure=state.pipe_pressure,
                    target=state.target, in_history=state.in_history, out_history=state.out_history)

    def state_from(self, b, scalar):
        g = (lambda t: float(t[0])) if scalar else (lambda t: t)
        h = (lambda t: t[0]) if scalar else (lambda t: t)
        st = SimulatorState(steps=g(b["steps_io"]), time=g(b["time_io"]), volume=g(b["volume_io"]),
                            predicted_pressure=g(b["pressure_io"]), pipe_pressure=g(b["pipe_pressure_io"]),
                            target=g(b["target_io"]), in_history=h(b["in_history_io"]),
                            out_history=h(b["out_history_io"]))
        return st, Observation(predicted_pressure=g(b["obs_pressure_io"]), time=g(b["obs_time_io"]))

    def __call__(self, state, action):
        """``__call__`` + ``dynamics`` (``:75-133``)."""
        u_in, u_out = action
        sd = self.state_dict(state)
        scal = [v for k, v in sd.items() if "history" not in k]
        N = _batch(*scal, u_in, u_out)
        scalar = not any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in scal + [u_in, u_out])
        bc = lambda v: torch.as_tensor(v, dtype=torch.float32, device=self.device).reshape(-1).expand(N)
        b, _ = _launch.launch(self, self.waveform, N, 1, "env_only", self.device, env_state=sd,
                              actions=(bc(u_in), bc(u_out)), series=False)
        return self.state_from(b, scalar)
