"""LQG (``deluca/agents/_lqg.py:27-101``): Kalman predictor + LQR gain on the fused CUDA path.

Setup solves two DAREs (the reference calls scipy on the host, ``:61-79``; here the batched doubling solver of
``deluca_b200.riccati``): ``K`` from (A, B, Q, R) and the Kalman gain ``L = S C'(C S C' + V)^-1`` from the dual
problem (A', C', W, V).  ``u = -K x_hat; x_hat <- A x_hat + B u + L (y - C x_hat)`` runs in the kernel.
"""
import torch

from ..core import Agent, Obj, field
from ..riccati import dare_gain
from .. import fused


class LQGState(Obj):
    x_hat: torch.Tensor = field(None, jaxed=True)  # [N,d]
    t: int = field(0, jaxed=False)


class LQG(Agent):
    def __init__(self, A, B, C, Q=None, R=None, W=None, V=None):
        as_t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        A, B, C = as_t(A).contiguous(), as_t(B).contiguous(), as_t(C).contiguous()
        if not A.is_cuda:
            A, B, C = A.cuda(), B.cuda(), C.cuda()
        self.A, self.B, self.C = A, B, C
        self.d, self.n, self.p = A.shape[-1], B.shape[-1], C.shape[-2]
        dev = A.device
        mv = lambda a: None if a is None else as_t(a).to(dev)
        self.K, self.P = dare_gain(A, B, mv(Q), mv(R))
        self.K = self.K.contiguous()
        At, Ct = A.transpose(-1, -2).contiguous(), C.transpose(-1, -2).contiguous()
        _, Sigma = dare_gain(At, Ct, mv(W), mv(V))                                   # _lqg.py:72-77
        V64 = (torch.eye(self.p, dtype=torch.float64, device=dev) if V is None else mv(V).double())
        C64 = C.double()
        S = C64 @ Sigma @ C64.transpose(-1, -2) + V64
        self.L = (Sigma @ C64.transpose(-1, -2) @ torch.linalg.inv(S)).float().contiguous()   # _lqg.py:79
        self.x_hat = torch.zeros(self.d, 1, device=dev)
        self.t = 0

    def init(self, N=1):
        st = LQGState(x_hat=torch.zeros(N, self.d, device=self.A.device), t=0)
        st.freeze()
        return st

    def step(self, st, obs):
        N = st.x_hat.shape[0]
        y = torch.as_tensor(obs, dtype=torch.float32, device=st.x_hat.device).reshape(1, N, self.p).contiguous()
        r = fused.lds_of_rollout(self.A, self.B, self.C, None, 1, agent="lqg", y_in=y, K=self.K, Lk=self.L,
                                 z=st.x_hat, t0=st.t, agent_only=True, keep_costs=False)
        new = LQGState(x_hat=r.z, t=st.t + 1)
        new.freeze()
        return new, r.U[0]

    def rollout(self, x, T, W=None, st=None, **kw):
        st = st or self.init(x.shape[0])
        args = dict(agent="lqg", W=W, K=self.K, Lk=self.L, z=st.x_hat, t0=st.t)
        args.update(kw)   # the caller's keywords override the agent's own
        return fused.lds_of_rollout(self.A, self.B, self.C, x, T, **args)

    def __call__(self, y):
        st = LQGState(x_hat=self.x_hat.reshape(1, -1).contiguous(), t=self.t)
        st, u = self.step(st, torch.as_tensor(y, dtype=torch.float32).reshape(1, -1))
        self.x_hat = st.x_hat[0].unsqueeze(-1)
        self.t += 1
        return u.reshape(-1, 1)

    def update(self, y, u):
        """``_lqg.py:100-101``: no-op."""
        return None
