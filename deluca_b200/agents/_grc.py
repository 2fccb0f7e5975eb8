"""Gradient Response Controller (``deluca/agents/_grc.py:42-169``) on the fused CUDA path."""
from ._of import FilterAgent, quad_loss  # noqa: F401


class GRC(FilterAgent):
    kind = "grc"

    def __init__(self, A, B, C, cost_fn=None, R_M=10.0, key=0, init_scale=1.0, m=10, lr=0.005, decay=True):
        self._setup(A, B, C, cost_fn, R_M, key, init_scale, m, None, 0.2, lr, decay)

    @property
    def M(self):
        """M[m,n,p] (``_grc.py:77``)."""
        return self._theta
