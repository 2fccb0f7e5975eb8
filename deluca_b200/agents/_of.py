"""Output-feedback controllers on the fused CUDA path: shared plumbing for GRC / SFC / DSC / DRC / LQG
(``deluca/agents/_grc.py``, ``_sfc.py``, ``_dsc.py``, ``_drc.py``, ``_lqg.py``).

GRC, SFC and DSC are one family for the kernel: the action is linear in the parameters and in a fixed bank of linear
filters of the y_nat history,  a(k) = sum_f Theta[f] (sum_o Phi[f, o] ynat[k + o]).  ``filter_bank`` builds Phi for
each of them on the host (setup time, like the reference's ``np.linalg.eigh`` at ``_sfc.py:98-112``); every number of a
step is produced by ``dlc_lds_of_rollout_f32``.

Faces, as for GPC: stateful (``u = agent(y)`` mutates the agent, the reference's own) and functional
(``st = agent.init(N)``; ``st, u = agent.step(st, y)`` with leaves carrying a leading env axis N).  The reference
draws its initial parameters from ``jax.random.normal(PRNGKey(0))`` (e.g. ``_grc.py:77``), which cannot be reproduced
without JAX: the draw here comes from ``torch.Generator().manual_seed(key)``, or pass the parameters explicitly.
"""
import numpy as np
import torch

from ..core import Agent, Obj, field
from .. import fused


def quad_loss(y, u):
    """``quad_loss`` (``_grc.py:27-39``)."""
    return (y * y).sum() + (u * u).sum()


def spectral_filters(m, h, gamma=0.2):
    """``_sfc.py:98-112`` / ``_dsc.py:111-125``: sigma^(1/4)-scaled top-h eigenvectors of the Hankel matrix
    Z[i,j] = (1-gamma)^(i+j-1)/(i+j-1), flipped along the window axis -> filters[h, m] (float64, host)."""
    idx = np.arange(1, m + 1)
    I, J = np.meshgrid(idx, idx, indexing="ij")
    Z = ((1 - gamma) ** (I + J - 1)) / (I + J - 1)
    vals, vecs = np.linalg.eigh(Z)
    return np.flip((vals[-h:] ** 0.25)[:, None] * vecs[:, -h:].T, axis=1).copy()


def filter_bank(kind, m, h=None, gamma=0.2):
    """Phi[F, L] such that the reference's ``action(k)`` is sum_f Theta[f] @ (sum_o Phi[f, o] ynats[k + o])."""
    if kind == "grc":                                   # _grc.py:92-95
        return np.eye(m)
    filt = spectral_filters(m, h, gamma)
    if kind == "sfc":                                   # _sfc.py:131-140: [M_0 | M]
        Phi = np.zeros((h + 1, m + 1))
        Phi[0, m] = 1.0
        Phi[1:, :m] = filt
        return Phi
    if kind == "dsc":                                   # _dsc.py:151-196: [M_0 | M_1 | M_2 | M_3]
        Phi = np.zeros((1 + 2 * h + h * h, 2 * m + 1))
        Phi[0, 2 * m] = 1.0
        Phi[1:1 + h, m:2 * m] = filt
        Phi[1 + h:1 + 2 * h, m + 1:2 * m + 1] = filt
        for a in range(h):
            for b in range(h):                          # filters[a,i] filters[b,q] ynats[k+1+i+q]
                Phi[1 + 2 * h + a * h + b, 1:2 * m] = np.convolve(filt[a], filt[b])
        return Phi
    raise ValueError(kind)


class OFState(Obj):
    """Functional state of a batch of filter-family agents."""
    theta: torch.Tensor = field(None, jaxed=True)         # [N,F,n,p]
    z: torch.Tensor = field(None, jaxed=True)             # [N,d]
    ynat_history: torch.Tensor = field(None, jaxed=True)  # [N,L+m,p] oldest -> newest
    t: int = field(0, jaxed=False)


class FilterAgent(Agent):
    """Common body of GRC / SFC / DSC."""
    kind = None

    def _setup(self, A, B, C, cost_fn, R_M, key, init_scale, m, h, gamma, lr, decay):
        if cost_fn is not None and cost_fn is not quad_loss:
            raise NotImplementedError("the fused kernel implements the reference's default quad_loss only")
        as_t = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous()
        self.A, self.B, self.C = as_t(A), as_t(B), as_t(C)
        if not self.A.is_cuda:
            self.A, self.B, self.C = self.A.cuda(), self.B.cuda(), self.C.cuda()
        self.d, self.n, self.p = self.A.shape[-1], self.B.shape[-1], self.C.shape[-2]
        self.cost_fn = quad_loss
        self.m, self.h, self.R_M, self.lr, self.decay = m, h, R_M, lr, decay
        self.init_scale, self.key = init_scale, int(key)
        self.t = 0
        self.Phi = torch.as_tensor(filter_bank(self.kind, m, h, gamma), dtype=torch.float32, device=self.A.device)
        self.F, self.L = self.Phi.shape
        st = self.init(1)
        self._theta, self.z, self.ynat_history = st.theta[0], st.z[0].unsqueeze(-1), st.ynat_history[0].unsqueeze(-1)

    # ---- functional face
    def init(self, N=1, theta=None):
        dev = self.A.device
        if theta is None:
            g = torch.Generator(device=dev).manual_seed(self.key)
            theta = self.init_scale * torch.randn(N, self.F, self.n, self.p, device=dev, generator=g)
        st = OFState(theta=torch.as_tensor(theta, dtype=torch.float32, device=dev).contiguous(),
                     z=torch.zeros(N, self.d, device=dev),
                     ynat_history=torch.zeros(N, self.L + self.m, self.p, device=dev), t=0)
        st.freeze()
        return st

    def step(self, st, obs):
        """``Agent.__call__(state, obs)`` for N envs: get_action then update (``_grc.py:110-169``)."""
        N = st.theta.shape[0]
        y = torch.as_tensor(obs, dtype=torch.float32, device=st.theta.device).reshape(1, N, self.p).contiguous()
        r = fused.lds_of_rollout(self.A, self.B, self.C, None, 1, agent="fgrc", y_in=y, Phi=self.Phi, theta=st.theta,
                                 phi_is_identity=self.kind == "grc",
                                 z=st.z, ynat=st.ynat_history, m=self.m, t0=st.t, lr=self.lr, decay=self.decay,
                                 R_M=self.R_M, agent_only=True, keep_costs=False)
        new = OFState(theta=r.theta, z=r.z, ynat_history=r.ynat, t=st.t + 1)
        new.freeze()
        return new, r.U[0]

    def rollout(self, x, T, W=None, st=None, **kw):
        """T fused closed-loop steps on the LDS this agent was built for (the scan x vmap body)."""
        st = st or self.init(x.shape[0])
        args = dict(agent="fgrc", W=W, Phi=self.Phi, theta=st.theta, phi_is_identity=self.kind == "grc", z=st.z,
                    ynat=st.ynat_history, m=self.m, t0=st.t, lr=self.lr, decay=self.decay, R_M=self.R_M)
        args.update(kw)   # the caller's keywords (t0, lr, env_t0, lanes_per_env, ...) override the agent's own
        return fused.lds_of_rollout(self.A, self.B, self.C, x, T, **args)

    # ---- stateful face (the reference's)
    def __call__(self, y):
        st = OFState(theta=self._theta.unsqueeze(0).contiguous(), z=self.z.reshape(1, -1).contiguous(),
                     ynat_history=self.ynat_history.reshape(1, -1, self.p).contiguous(), t=self.t)
        st, u = self.step(st, torch.as_tensor(y, dtype=torch.float32).reshape(1, -1))
        self._theta, self.z, self.ynat_history = st.theta[0], st.z[0].unsqueeze(-1), st.ynat_history[0].unsqueeze(-1)
        self.t += 1
        return u.reshape(-1, 1)
