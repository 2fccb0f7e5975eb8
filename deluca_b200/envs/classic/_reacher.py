"""Reacher (``deluca/envs/classic/_reacher.py:25-141``): two-link planar arm; the state carries the joint angles,
their rates and the tip's offset from ``goal_coord``.  ``arr`` may be [6] or batched [N,6].  With
PlanarQuadrotor, one of the two envs that satisfy ``deluca/igpc``'s protocol in the reference."""
import math

import torch

from ...core import Obj, field
from ._pendulum import _ClassicEnv


class ReacherState(Obj):
    arr: torch.Tensor = field(None, jaxed=True)
    h: float = field(0.0, jaxed=True)

    def flatten(self):
        return self.arr

    def unflatten(self, arr):
        return ReacherState(arr=arr, h=self.h)


class Reacher(_ClassicEnv):
    m1: float = field(1.0, jaxed=False)
    m2: float = field(1.0, jaxed=False)
    l1: float = field(1.0, jaxed=False)
    l2: float = field(1.0, jaxed=False)
    g: float = field(0.0, jaxed=False)
    max_torque: float = field(1.0, jaxed=False)     # declared but unused by the reference step (:112-141)
    dt: float = field(0.01, jaxed=False)
    H: int = field(200, jaxed=False)
    goal_coord: object = field(None, jaxed=False)
    device: str = field("cuda", jaxed=False)

    state_dim, action_dim = 6, 2

    def setup(self):
        if self.goal_coord is None:
            self.goal_coord = [0.0, 1.8]                                        # :95-97

    def spec(self):
        return 3, [self.m1, self.m2, self.l1, self.l2, self.g, self.dt, float(self.goal_coord[0]),
                   float(self.goal_coord[1])]

    def init(self, N=None):
        """``init`` (``:69-93``): returns (state, state)."""
        th = (math.pi / 4, math.pi / 2)
        arr = torch.tensor([th[0], th[1], 0.0, 0.0,
                            self.l1 * math.cos(th[0]) + self.l2 * math.cos(th[0] + th[1]) - float(self.goal_coord[0]),
                            self.l1 * math.sin(th[0]) + self.l2 * math.sin(th[0] + th[1]) - float(self.goal_coord[1])],
                           dtype=torch.float32, device=self.device)
        if N is not None:
            arr = arr.expand(N, 6).contiguous()
        st = ReacherState(arr=arr)
        return st, st

    def __call__(self, state, action):
        arr = self._step(state, action)
        return state.replace(arr=arr, h=state.h + 1), arr

    @property
    def action_size(self) -> int:
        return 2
