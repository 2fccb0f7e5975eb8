"""Spectral Filtering Controller (``deluca/agents/_sfc.py:42-214``) on the fused CUDA path."""
from ._of import FilterAgent, quad_loss  # noqa: F401


class SFC(FilterAgent):
    kind = "sfc"

    def __init__(self, A, B, C, cost_fn=None, R_M=10.0, gamma=0.2, key=0, init_scale=1.0, h=10, m=30, lr=0.005,
                 decay=True):
        self._setup(A, B, C, cost_fn, R_M, key, init_scale, m, h, gamma, lr, decay)

    @property
    def M_0(self):
        """M_0[n,p] (``_sfc.py:94``)."""
        return self._theta[0]

    @property
    def M(self):
        """M[h,n,p] (``_sfc.py:93``)."""
        return self._theta[1:]

    @property
    def filters(self):
        """filters[h,m] (``_sfc.py:109-111``)."""
        return self.Phi[1:, :self.m]
