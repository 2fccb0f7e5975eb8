"""Double Spectral Controller (``deluca/agents/_dsc.py:42-292``) on the fused CUDA path."""
from ._of import FilterAgent, quad_loss  # noqa: F401


class DSC(FilterAgent):
    kind = "dsc"

    def __init__(self, A, B, C, cost_fn=None, R_M=10.0, gamma=0.2, key=0, init_scale=1.0, h=10, m=30, lr=0.005,
                 decay=True):
        self._setup(A, B, C, cost_fn, R_M, key, init_scale, m, h, gamma, lr, decay)

    @property
    def M_0(self):
        return self._theta[0]

    @property
    def M_1(self):
        return self._theta[1:1 + self.h]

    @property
    def M_2(self):
        return self._theta[1 + self.h:1 + 2 * self.h]

    @property
    def M_3(self):
        """M_3[h,h,n,p] (``_dsc.py:101-103``)."""
        return self._theta[1 + 2 * self.h:].reshape(self.h, self.h, self.n, self.p)

    @property
    def filters(self):
        return self.Phi[1:1 + self.h, self.m:2 * self.m]
