"""Bandit Perturbation Controller (``deluca/agents/_bpc.py:33-158``) on the fused CUDA path.

The reference draws its perturbations from the unseeded global numpy RNG (``:20,26-30``); here
they come from an explicit ``torch.Generator`` (initial ``M``/``eps``) and either an explicit
stream or the library's counter-based Philox stream (per-step ``eps``), so runs are reproducible
(SURVEY A-4).
"""
import torch

from ..core import Agent, Obj, field
from .. import fused
from ._lqr import LQR


def generate_uniform(shape, norm=1.00, generator=None, device="cuda"):
    """``generate_uniform`` (``_bpc.py:26-30``): a Gaussian tensor scaled to Frobenius norm ``norm``."""
    v = torch.randn(*shape, device=device, generator=generator)
    return norm * v / torch.linalg.vector_norm(v)


class BPCState(Obj):
    M: torch.Tensor = field(None, jaxed=True)              # [N,H,m,d]
    noise_history: torch.Tensor = field(None, jaxed=True)  # [N,H,d]
    state: torch.Tensor = field(None, jaxed=True)          # [N,d]
    eps: torch.Tensor = field(None, jaxed=True)            # [N,H,H,m,d]
    cost: torch.Tensor = field(None, jaxed=True)           # [N] cost handed to the next call
    t: int = field(0, jaxed=False)


class BPC(Agent):
    def __init__(self, A, B, Q=None, R=None, K=None, start_time=0, H=5, lr_scale=0.005, decay=False, delta=0.01,
                 seed=0, device=None):
        dev = torch.device(device) if device is not None else (
            A.device if isinstance(A, torch.Tensor) and A.is_cuda else torch.device("cuda"))
        as_t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()
        self.A, self.B = as_t(A), as_t(B)
        self.d_state, self.d_action = self.B.shape[-2], self.B.shape[-1]
        self.t, self.H = 0, H
        self.lr_scale, self.decay, self.delta = lr_scale, decay, delta
        self.K = as_t(K) if K is not None else LQR(self.A, self.B, as_t(Q), as_t(R)).K
        self.seed = seed
        self._st = None

    def init(self, N=1, device=None):
        dev = device or self.A.device
        g = torch.Generator(device=dev).manual_seed(self.seed)
        H, m, d = self.H, self.d_action, self.d_state
        M = torch.stack([self.delta * generate_uniform((H, m, d), generator=g, device=dev) for _ in range(N)])
        eps = torch.stack([generate_uniform((H, H, m, d), generator=g, device=dev) for _ in range(N)])
        z = lambda *s: torch.zeros(*s, device=dev)
        st = BPCState(M=M, noise_history=z(N, H, d), state=z(N, d), eps=eps, cost=z(N), t=0)
        st.freeze()
        return st

    def step(self, st, obs, cost=None, eps_new=None):
        """``BPC.__call__(state, cost)`` for N envs (``_bpc.py:97-158``)."""
        N = st.M.shape[0]
        x = torch.as_tensor(obs, dtype=torch.float32, device=st.M.device).reshape(N, -1).contiguous()
        c = st.cost if cost is None else torch.as_tensor(cost, dtype=torch.float32, device=x.device).reshape(N)
        r = fused.lds_rollout(self.A.to(x.device), self.B.to(x.device), self.K.to(x.device), x, 1, policy="bpc",
                              H=self.H, M=st.M, hist=st.noise_history, xprev=st.state, eps=st.eps,
                              eps_new=eps_new, bpc_cost=c.contiguous(), t0=st.t, lr_scale=self.lr_scale,
                              decay=self.decay, bpc_delta=self.delta, seed=self.seed, agent_only=True,
                              traj_stride=1, keep_costs=False)
        new = BPCState(M=r.M, noise_history=r.hist, state=r.xprev, eps=r.eps, cost=r.bpc_cost, t=st.t + 1)
        new.freeze()
        return new, r.U[0]

    def __call__(self, state, cost):
        if self._st is None:
            self._st = self.init(1)
        self._st, u = self.step(self._st, torch.as_tensor(state, dtype=torch.float32).reshape(1, -1),
                                cost=torch.as_tensor(float(cost)).reshape(1))
        self.M, self.t = self._st.M[0], self._st.t
        return u.reshape(-1, 1)
