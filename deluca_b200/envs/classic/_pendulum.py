"""Pendulum (``deluca/envs/classic/_pendulum.py:23-77``), reproduced as written (SURVEY A-6: the
velocity update is ``thdot' = th + accel``).  ``arr`` may be [3] or batched [N,3]."""
import torch

from ...core import Env, Obj, field
from ... import fused


class PendulumState(Obj):
    arr: torch.Tensor = field(None, jaxed=True)
    h: int = field(0, jaxed=True)

    def flatten(self):          # igpc protocol (rollout.py:55-56), mirrors _planar_quadrotor.py:36-47
        return self.arr

    def unflatten(self, arr):
        return PendulumState(arr=arr, h=self.h)


class QuadCost:
    """c(x, u, env) = sum q (x - goal)^2 + sum r u^2: the cost the fused igpc kernels implement
    (the reference ships none for this path; q = 1, r = 0.1 is this repo's stated choice)."""

    def __init__(self, goal, q=1.0, r=0.1, goal_action=None):
        self.goal, self.q, self.r, self.goal_action = goal, q, r, goal_action

    def action_goal(self, m):
        return [0.0] * m if self.goal_action is None else [float(x) for x in self.goal_action]

    def vectors(self, d, m):
        ex = lambda v, n: [float(x) for x in (v if hasattr(v, "__len__") else [v] * n)]
        return ex(self.q, d), ex(self.r, m), [float(x) for x in self.goal]

    def __call__(self, state, u, env=None):
        x = state.arr if hasattr(state, "arr") else state
        q, r, goal = self.vectors(x.shape[-1], u.shape[-1])
        g = torch.tensor(goal, device=x.device)
        gu = torch.tensor(self.action_goal(u.shape[-1]), device=x.device)
        return ((x - g) ** 2 * torch.tensor(q, device=x.device)).sum(-1) + (
            (u - gu) ** 2 * torch.tensor(r, device=x.device)).sum(-1)


class _ClassicEnv(Env):
    """Shared single-step plumbing: ``env(state, action)`` is one H = 1 launch of dlc_env_rollout_f32."""

    def _step(self, state, action):
        arr = torch.as_tensor(state.arr, dtype=torch.float32, device=self.device)
        single = arr.dim() == 1
        x = arr.reshape(-1, self.state_dim).contiguous()
        u = torch.as_tensor(action, dtype=torch.float32, device=self.device).reshape(-1, 1, self.action_dim)
        if u.shape[0] not in (1, x.shape[0]):
            raise ValueError("action batch does not match the state batch")
        u = u.expand(x.shape[0], 1, self.action_dim).contiguous()
        p = fused.plan_params(self, self, QuadCost([0.0] * self.state_dim, 0.0, 0.0), x.shape[0], 1)
        X, _, _ = fused.env_rollout(p, u, x)
        new = X[:, 1]
        return new[0] if single else new


class Pendulum(_ClassicEnv):
    m: float = field(1.0, jaxed=False)
    l: float = field(1.0, jaxed=False)
    g: float = field(9.81, jaxed=False)
    max_torque: float = field(1.0, jaxed=False)
    dt: float = field(0.02, jaxed=False)
    H: int = field(300, jaxed=False)
    goal_state: object = field(None, jaxed=False)
    device: str = field("cuda", jaxed=False)

    state_dim, action_dim = 3, 1

    def setup(self):
        if self.goal_state is None:
            self.goal_state = [0.0, -1.0, 0.0]

    def spec(self):
        return 0, [self.m, self.l, self.g, self.max_torque, self.dt]

    def init(self, N=None):
        """``init`` (``:41-48``): returns (state, state); with ``N`` the batch [N,3]."""
        arr = torch.tensor([0.0, 1.0, 0.0], device=self.device)
        if N is not None:
            arr = arr.expand(N, 3).contiguous()
        state = PendulumState(arr=arr)
        return state, state

    def __call__(self, state, action):
        arr = self._step(state, action)
        return PendulumState(arr=arr, h=state.h + 1), arr

    @property
    def action_size(self) -> int:
        return 1
