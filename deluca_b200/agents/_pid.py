"""Generic leaky PID agent (``deluca/agents/_pid.py:20-48``); leaves may carry a batch axis [N]."""
import torch

from ..core import Agent, Obj, field
from .. import fused


class PIDState(Obj):
    P: float = field(0.0, jaxed=True)
    I: float = field(0.0, jaxed=True)
    D: float = field(0.0, jaxed=True)


class PID(Agent):
    K_P: float = field(0.0, jaxed=True)
    K_I: float = field(0.0, jaxed=True)
    K_D: float = field(0.0, jaxed=True)
    RC: float = field(0.5, jaxed=False)
    dt: float = field(0.03, jaxed=False)
    decay: float = field(None, jaxed=False)
    device: str = field("cuda", jaxed=False)

    def init(self):
        return PIDState()

    def setup(self):
        self.decay = self.dt / (self.dt + self.RC)

    def __call__(self, state, obs):
        """Assume the observation is the error (``:40-48``)."""
        leaves = [state.P, state.I, state.D, obs, self.K_P, self.K_I, self.K_D]
        N = max([v.numel() for v in leaves if isinstance(v, torch.Tensor)] + [1])
        scalar = not any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in leaves)
        f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=self.device).reshape(-1).expand(N).contiguous()
        u, P, I, D = fused.pid_rollout(f(self.K_P), f(self.K_I), f(self.K_D), f(obs).reshape(1, N), f(state.P),
                                       f(state.I), f(state.D), self.decay)
        g = (lambda t: float(t[0])) if scalar else (lambda t: t)
        return state.replace(P=g(P), I=g(I), D=g(D)), g(u[0])
