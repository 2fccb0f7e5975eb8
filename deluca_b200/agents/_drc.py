"""Disturbance Response Controller (``deluca/agents/_drc.py:43-203``) on the fused CUDA path.

``u = -K y + sum_j M[j] y_nat[j]`` with y_nat reconstructed through the truncated Markov operator
``G[i] = C A^i B`` (``:97-102``; formed inside the kernel) and an online gradient step on M.  As written, the
learning rate without decay is 1.0, not ``lr_scale`` (``:184``); reproduced.  The initial ``M`` is
``lr_scale * normal`` under ``PRNGKey(0)`` in the reference (``:105``); here the draw comes from
``torch.Generator().manual_seed(seed)``, or pass ``M`` to ``init``.
"""
import torch

from ..core import Agent, Obj, field
from .. import fused
from ._of import quad_loss


class DRCState(Obj):
    M: torch.Tensor = field(None, jaxed=True)      # [N,m,n,p]
    y_nat: torch.Tensor = field(None, jaxed=True)  # [N,m,p] newest first
    us: torch.Tensor = field(None, jaxed=True)     # [N,h,n] newest first
    t: int = field(0, jaxed=False)


class DRC(Agent):
    def __init__(self, A, B, C=None, K=None, cost_fn=None, m=10, h=50, lr_scale=0.03, decay=True, RM=1000, seed=0):
        if cost_fn is not None and cost_fn is not quad_loss:
            raise NotImplementedError("the fused kernel implements the reference's default quad_loss only")
        as_t = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous()
        self.A, self.B = as_t(A), as_t(B)
        if not self.A.is_cuda:
            self.A, self.B = self.A.cuda(), self.B.cuda()
        dev = self.A.device
        d_state, d_action = self.B.shape[-2], self.B.shape[-1]
        self.C = as_t(C).to(dev) if C is not None else torch.eye(d_state, device=dev).expand(
            *self.A.shape[:-2], d_state, d_state).contiguous()                       # _drc.py:87
        self.d, self.n, self.p = d_state, d_action, self.C.shape[-2]
        self.t = 0
        self.m, self.h = m, h
        self.lr_scale, self.decay, self.RM, self.seed = lr_scale, decay, RM, int(seed)
        self.K = as_t(K).to(dev) if K is not None else torch.zeros(*self.A.shape[:-2], self.n, self.p, device=dev)
        # the truncated response decays (|A^h B|_F <= |B|_F for every env): the kernel may carry sum_i G[i] us[i] as a
        # d-vector recursion (DLC_OF_FLAG_STABLE_A; it re-checks the promise per env)
        AhB = torch.linalg.matrix_power(self.A.double(), h) @ self.B.double()
        self.stable_a = bool((AhB.flatten(-2).norm(dim=-1) <= 0.999 * self.B.double().flatten(-2).norm(dim=-1)).all())
        st = self.init(1)
        self.M, self.y_nat, self.us = st.M[0], st.y_nat[0].unsqueeze(-1), st.us[0].unsqueeze(-1)

    @property
    def G(self):
        """Truncated Markov operator G[h,p,n] (``_drc.py:97-102``), for inspection; the kernel forms its own copy."""
        out, Ap = [], torch.eye(self.d, device=self.A.device).expand_as(self.A)
        for _ in range(self.h):
            out.append(self.C @ Ap @ self.B)
            Ap = Ap @ self.A
        return torch.stack(out, dim=-3)

    def init(self, N=1, M=None):
        dev = self.A.device
        if M is None:
            g = torch.Generator(device=dev).manual_seed(self.seed)
            M = self.lr_scale * torch.randn(N, self.m, self.n, self.p, device=dev, generator=g)
        st = DRCState(M=torch.as_tensor(M, dtype=torch.float32, device=dev).contiguous(),
                      y_nat=torch.zeros(N, self.m, self.p, device=dev), us=torch.zeros(N, self.h, self.n, device=dev),
                      t=0)
        st.freeze()
        return st

    def _kw(self, st):
        return dict(agent="drc", K=self.K, theta=st.M, ynat=st.y_nat, us=st.us, m=self.m, h=self.h, t0=st.t,
                    lr=self.lr_scale, decay=self.decay, R_M=float(self.RM), stable_a=self.stable_a)

    def step(self, st, obs):
        """update_noise, get_action, update_params for N envs (``_drc.py:131-149``)."""
        N = st.M.shape[0]
        y = torch.as_tensor(obs, dtype=torch.float32, device=st.M.device).reshape(1, N, self.p).contiguous()
        r = fused.lds_of_rollout(self.A, self.B, self.C, None, 1, y_in=y, agent_only=True, keep_costs=False,
                                 **self._kw(st))
        new = DRCState(M=r.theta, y_nat=r.ynat, us=r.us, t=st.t + 1)
        new.freeze()
        return new, r.U[0]

    def rollout(self, x, T, W=None, st=None, **kw):
        st = st or self.init(x.shape[0])
        args = dict(self._kw(st), W=W)
        args.update(kw)   # the caller's keywords (t0, lr, env_t0, ...) override the agent's own
        return fused.lds_of_rollout(self.A, self.B, self.C, x, T, **args)

    def __call__(self, obs):
        st = DRCState(M=self.M.unsqueeze(0).contiguous(), y_nat=self.y_nat.reshape(1, self.m, self.p).contiguous(),
                      us=self.us.reshape(1, self.h, self.n).contiguous(), t=self.t)
        st, u = self.step(st, torch.as_tensor(obs, dtype=torch.float32).reshape(1, -1))
        self.M, self.y_nat, self.us = st.M[0], st.y_nat[0].unsqueeze(-1), st.us[0].unsqueeze(-1)
        self.t += 1
        return u.reshape(-1, 1)
