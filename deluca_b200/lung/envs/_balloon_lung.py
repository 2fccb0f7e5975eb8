"""BalloonLung (``deluca/lung/envs/_balloon_lung.py:26-146``): state leaves may carry a batch axis [N]."""
import torch

from ...core import Obj, field
from ..core import BreathWaveform, LungEnv
from .. import _launch


class Observation(Obj):
    predicted_pressure: float = 0.0
    time: float = 0.0


class SimulatorState(Obj):
    steps: int = 0
    time: float = 0.0
    volume: float = 0.0
    predicted_pressure: float = 0.0
    target: float = 0.0


def _batch(*leaves):
    n = 1
    for v in leaves:
        if isinstance(v, torch.Tensor) and v.dim() > 0:
            n = max(n, v.numel())
    return n


class BalloonLung(LungEnv):
    """Lung simulator based on the two-balloon experiment."""

    leak: bool = field(False, jaxed=False)
    peep_valve: float = field(5.0, jaxed=False)
    PC: float = field(40.0, jaxed=False)
    P0: float = field(0.0, jaxed=False)
    R: float = field(15.0, jaxed=False)
    C: float = field(10.0, jaxed=False)
    dt: float = field(0.03, jaxed=False)
    min_volume: float = field(1.5, jaxed=False)
    r0: float = field(None, jaxed=False)
    waveform: BreathWaveform = field(None, jaxed=False)
    reset_normalized_peep: float = field(0.0, jaxed=False)
    flow: float = field(0, jaxed=False)
    device: str = field("cuda", jaxed=False)

    def setup(self):
        if self.waveform is None:
            self.waveform = BreathWaveform.create()
        self.r0 = (3.0 * self.min_volume / (4.0 * 3.141592653589793)) ** (1.0 / 3.0)

    def reset(self, N=None):
        """``reset`` (``:87-98``); with ``N`` the leaves are tensors [N]."""
        tgt = self.waveform.at(0.0)
        if N is None:
            st = SimulatorState(steps=0, time=0.0, volume=self.min_volume,
                                predicted_pressure=self.reset_normalized_peep, target=tgt)
            return st, Observation(predicted_pressure=self.reset_normalized_peep, time=0.0)
        f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=self.device).reshape(-1).expand(N).contiguous()
        st = SimulatorState(steps=f(0.0), time=f(0.0), volume=f(self.min_volume),
                            predicted_pressure=f(self.reset_normalized_peep), target=f(tgt))
        return st, Observation(predicted_pressure=f(self.reset_normalized_peep), time=f(0.0))

    def state_dict(self, state):
        return dict(steps=state.steps, time=state.time, volume=state.volume,
                    predicted_pressure=state.predicted_pressure, target=state.target)

    def state_from(self, b, scalar):
        g = (lambda t: float(t[0])) if scalar else (lambda t: t)
        st = SimulatorState(steps=g(b["steps_io"]), time=g(b["time_io"]), volume=g(b["volume_io"]),
                            predicted_pressure=g(b["pressure_io"]), target=g(b["target_io"]))
        return st, Observation(predicted_pressure=g(b["obs_pressure_io"]), time=g(b["obs_time_io"]))

    def __call__(self, state, action):
        """``__call__`` (``:100-146``): one simulator step for every breath of the batch."""
        u_in, u_out = action
        leaves = list(self.state_dict(state).values())
        N = _batch(*leaves, u_in, u_out)
        scalar = not any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in leaves + [u_in, u_out])
        bc = lambda v: torch.as_tensor(v, dtype=torch.float32, device=self.device).reshape(-1).expand(N)
        b, _ = _launch.launch(self, self.waveform, N, 1, "env_only", self.device, env_state=self.state_dict(state),
                              actions=(bc(u_in), bc(u_out)), series=False)
        return self.state_from(b, scalar)
