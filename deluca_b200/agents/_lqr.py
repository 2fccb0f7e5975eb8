"""Infinite-horizon LQR (``deluca/agents/_lqr.py:22-72``).

The gain is a setup-time solve (the reference calls scipy's DARE on the host, ``:48-60``); here it
is the batched doubling solver in ``deluca_b200.riccati``.  ``u = -K x`` runs in the fused kernel.
"""
import torch

from ..core import Agent
from ..riccati import dare_gain
from .. import fused


class LQR(Agent):
    def __init__(self, A, B, Q=None, R=None, device=None):
        # numpy / CPU inputs (how the reference is normally called) are moved to the device once, here
        dev = torch.device(device) if device is not None else (
            A.device if isinstance(A, torch.Tensor) and A.is_cuda else torch.device("cuda"))
        as_t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()
        A, B = as_t(A), as_t(B)
        self.A, self.B = A, B
        self.K, self.X = dare_gain(A, B, as_t(Q), as_t(R))
        self.K = self.K.contiguous()

    def __call__(self, state):
        """u = -K x (``_lqr.py:62-72``); state [d,1] or batched [N,d]."""
        x = torch.as_tensor(state, dtype=torch.float32, device=self.K.device)
        col = x.dim() == 2 and x.shape[-1] == 1 and self.K.dim() == 2
        xb = x.reshape(1, -1) if col else x
        N, d = xb.shape
        r = fused.lds_rollout(self.A.to(xb.device), self.B.to(xb.device), self.K, xb.contiguous(), 1, policy="lqr",
                              agent_only=True, traj_stride=1, keep_costs=False)
        u = r.U[0]
        return u.reshape(-1, 1) if col else u
