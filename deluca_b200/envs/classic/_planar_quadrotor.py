"""PlanarQuadrotor (``deluca/envs/classic/_planar_quadrotor.py:36-104``): 6-state planar quadrotor with a
wind field, explicit Euler.  ``arr`` may be [6] or batched [N,6]; this is one of the two envs that
satisfy ``deluca/igpc``'s protocol in the reference (state_dim / action_dim / flatten / unflatten)."""
import torch

from ...core import Obj, field
from ._pendulum import _ClassicEnv


def dissipative(x, y, wind):
    """``dissipative`` (``:26-28``)."""
    return [wind * x, wind * y]


def constant(x, y, wind):
    """``constant`` (``:31-33``)."""
    c = 0.7071067811865476
    return [wind * c, wind * c]


class PlanarQuadrotorState(Obj):
    arr: torch.Tensor = field(None, jaxed=True)
    h: float = field(0.0, jaxed=True)
    last_action: torch.Tensor = field(None, jaxed=True)

    def flatten(self):
        return self.arr

    def unflatten(self, arr):
        return PlanarQuadrotorState(arr=arr, h=self.h, last_action=self.last_action)


class PlanarQuadrotor(_ClassicEnv):
    m: float = field(0.1, jaxed=False)
    l: float = field(0.2, jaxed=False)
    g: float = field(9.81, jaxed=False)
    dt: float = field(0.05, jaxed=False)
    H: int = field(100, jaxed=False)
    wind: float = field(0.0, jaxed=False)
    wind_func: object = field(dissipative, jaxed=False)
    goal_state: object = field(None, jaxed=False)
    goal_action: object = field(None, jaxed=False)
    device: str = field("cuda", jaxed=False)

    state_dim, action_dim = 6, 2

    def setup(self):
        self.goal_action = [self.m * self.g / 2.0, self.m * self.g / 2.0]     # :73
        if self.goal_state is None:
            self.goal_state = [0.0] * 6
        if self.wind_func not in (dissipative, constant):
            raise NotImplementedError("the fused kernels implement the reference's two wind fields "
                                      "(dissipative, constant)")

    def spec(self):
        return 2, [self.m, self.l, self.g, self.dt, self.wind, 0.0 if self.wind_func is dissipative else 1.0]

    def init(self, N=None):
        """``init`` (``:84-90``): returns (state, arr)."""
        arr = torch.tensor([1.0, 1.0, 0.0, 0.0, 0.0, 0.0], device=self.device)
        la = torch.zeros(2, device=self.device)
        if N is not None:
            arr, la = arr.expand(N, 6).contiguous(), la.expand(N, 2).contiguous()
        st = PlanarQuadrotorState(arr=arr, last_action=la)
        return st, st.arr

    def __call__(self, state, action):
        arr = self._step(state, action)
        return state.replace(arr=arr, last_action=action, h=state.h + 1), arr

    @property
    def action_size(self) -> int:
        return 2
