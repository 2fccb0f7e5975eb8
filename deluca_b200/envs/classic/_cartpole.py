"""Cartpole (``deluca/envs/classic/_cartpole.py:26-93``): the linearised 4-state model.

WAIVER (SURVEY A-7): as shipped the reference's step raises ``TypeError`` -- stray trailing commas
make ``B`` and ``arr`` 1-tuples (``:83-88``).  The evident intent ``arr' = A arr + B (u + offset)``
is what runs here."""
import torch

from ...core import Obj, field
from ._pendulum import _ClassicEnv


class CartpoleState(Obj):
    arr: torch.Tensor = field(None, jaxed=True)
    h: int = field(0, jaxed=True)
    offset: float = field(0.0, jaxed=False)

    def flatten(self):
        return self.arr

    def unflatten(self, arr):
        return CartpoleState(arr=arr, h=self.h, offset=self.offset)


class Cartpole(_ClassicEnv):
    m: float = field(0.1, jaxed=False)
    M: float = field(1.0, jaxed=False)
    l: float = field(1.0, jaxed=False)
    g: float = field(9.81, jaxed=False)
    dt: float = field(0.02, jaxed=False)
    H: int = field(10, jaxed=False)
    goal_state: object = field(None, jaxed=False)
    dynamics: bool = field(False, jaxed=False)
    offset: float = field(0.0, jaxed=False)   # CartpoleState.offset is static; carried by the env here
    device: str = field("cuda", jaxed=False)

    state_dim, action_dim = 4, 1

    def setup(self):
        if self.goal_state is None:
            self.goal_state = [0.0, 0.0, 0.0, 0.0]

    def spec(self):
        return 1, [self.m, self.M, self.l, self.g, self.dt, self.offset]

    def init(self, N=None):
        arr = torch.zeros(4, device=self.device) if N is None else torch.zeros(N, 4, device=self.device)
        state = CartpoleState(arr=arr, offset=self.offset)
        return state, state

    def __call__(self, state, action):
        arr = self._step(state, action)
        return state.replace(arr=arr, h=state.h + 1), arr

    @property
    def action_size(self) -> int:
        return 1
