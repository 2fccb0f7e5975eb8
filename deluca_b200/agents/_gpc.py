"""Gradient Perturbation Controller (``deluca/agents/_gpc.py:44-183``) on the fused CUDA path.

Stateful face (the reference's): ``agent = GPC(A, B, H=3, HH=2); u = agent(x)`` mutates
``agent.M / noise_history / state / t``.  Functional face (SURVEY A-1):
``st = agent.init(N)``; ``st, u = agent.step(st, x)`` with ``GPCState`` leaves carrying a leading
env axis.  Both run one agent-only step of ``dlc_lds_rollout_f32`` (hand-derived adjoint of
``policy_loss``, SURVEY A-3); whole episodes should go through ``deluca_b200.training.apg``.
"""
import torch

from ..core import Agent, Obj, field
from .. import fused
from ._lqr import LQR


def quad_loss(x, u):
    """``quad_loss`` (``_gpc.py:29-40``)."""
    return (x * x).sum() + (u * u).sum()


class GPCState(Obj):
    M: torch.Tensor = field(None, jaxed=True)              # [N,H,m,d]
    noise_history: torch.Tensor = field(None, jaxed=True)  # [N,H+HH,d] oldest -> newest
    state: torch.Tensor = field(None, jaxed=True)          # [N,d] previous observed state
    action: torch.Tensor = field(None, jaxed=True)         # [N,m] previous action
    t: int = field(0, jaxed=False)


class GPC(Agent):
    def __init__(self, A, B, Q=None, R=None, K=None, start_time=0, cost_fn=None, H=3, HH=2, lr_scale=0.005,
                 decay=True, noise_uses_prev_action=False, device=None):
        # cost_fn (_gpc.py:52,76): quad_loss, or a diagonal quadratic QuadCost(goal = 0, q, r) -- the family the fused
        # kernels seed their hand adjoint with (dlc_lds_rollout_cost_f32)
        self.cost_q = self.cost_r = None
        if cost_fn is not None and cost_fn is not quad_loss:
            from ..envs.classic._pendulum import QuadCost
            goal = getattr(cost_fn, "goal", None)
            if (not isinstance(cost_fn, QuadCost) or cost_fn.goal_action is not None
                    or (goal is not None and any(float(g) != 0.0 for g in goal))):
                raise NotImplementedError("the fused kernels take cost_fn = quad_loss or a diagonal quadratic "
                                          "QuadCost(goal=None or zeros, q, r) without an action goal")
        # numpy / CPU inputs (how the reference is normally called) are moved to the device once, here
        dev = torch.device(device) if device is not None else (
            A.device if isinstance(A, torch.Tensor) and A.is_cuda else torch.device("cuda"))
        as_t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()
        self.A, self.B = as_t(A), as_t(B)
        self.d_state, self.d_action = self.B.shape[-2], self.B.shape[-1]
        self.t = start_time * 0   # the reference ignores start_time (_gpc.py:82)
        self.H, self.HH = H, HH
        self.lr_scale, self.decay = lr_scale, decay
        self.noise_uses_prev_action = noise_uses_prev_action
        self.bias = 0
        self.K = as_t(K) if K is not None else LQR(self.A, self.B, as_t(Q), as_t(R)).K      # _gpc.py:93
        if cost_fn is not None and cost_fn is not quad_loss:
            ex = lambda v, n: torch.tensor([float(x) for x in (v if hasattr(v, "__len__") else [v] * n)], device=dev)
            self.cost_q, self.cost_r = ex(cost_fn.q, self.d_state), ex(cost_fn.r, self.d_action)
        st = self.init(1, device=dev)
        self.M = st.M[0]
        self.noise_history = st.noise_history[0].unsqueeze(-1)
        self.state = st.state[0].unsqueeze(-1)
        self.action = st.action[0].unsqueeze(-1)

    # ---- functional face
    def init(self, N=1, device=None):
        dev = device or self.A.device
        z = lambda *s: torch.zeros(*s, device=dev)
        st = GPCState(M=z(N, self.H, self.d_action, self.d_state),
                      noise_history=z(N, self.H + self.HH, self.d_state), state=z(N, self.d_state),
                      action=z(N, self.d_action), t=0)
        st.freeze()
        return st

    def step(self, st, obs):
        """``Agent.__call__(state, obs)`` for N envs: get_action then update (``_gpc.py:130-183``)."""
        x = torch.as_tensor(obs, dtype=torch.float32, device=st.M.device).reshape(st.M.shape[0], -1).contiguous()
        r = fused.lds_rollout(self.A.to(x.device), self.B.to(x.device), self.K.to(x.device), x, 1, policy="gpc",
                              H=self.H, HH=self.HH, M=st.M, hist=st.noise_history, xprev=st.state,
                              uprev=st.action, t0=st.t, lr_scale=self.lr_scale, decay=self.decay,
                              noise_uses_prev_action=self.noise_uses_prev_action, agent_only=True, traj_stride=1,
                              keep_costs=False, cost_q=self.cost_q, cost_r=self.cost_r)
        new = GPCState(M=r.M, noise_history=r.hist, state=r.xprev, action=r.uprev, t=st.t + 1)
        new.freeze()
        return new, r.U[0]

    # ---- stateful face (the reference's)
    def __call__(self, state):
        st = GPCState(M=self.M.unsqueeze(0).contiguous(), noise_history=self.noise_history.reshape(1, -1, self.d_state),
                      state=self.state.reshape(1, -1), action=self.action.reshape(1, -1), t=self.t)
        st, u = self.step(st, torch.as_tensor(state, dtype=torch.float32).reshape(1, -1))
        self.M, self.noise_history = st.M[0], st.noise_history[0].unsqueeze(-1)
        self.state, self.action = st.state[0].unsqueeze(-1), st.action[0].unsqueeze(-1)
        self.t += 1
        return u.reshape(-1, 1)
