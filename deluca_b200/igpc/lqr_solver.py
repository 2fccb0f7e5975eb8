"""``deluca/igpc/lqr_solver.py``: the Riccati backward pass, on the device instead of host NumPy."""
from .. import fused


def LQR(F, C):
    """``LQR`` (``lqr_solver.py:40-86``): F = (F_x[N,H,d,d], F_u[N,H,d,m]),
    C = (c_x, c_u, c_xx, c_uu) -> (k[N,H,m], K[N,H,m,d])."""
    squeeze = F[0].dim() == 3
    if squeeze:
        F = tuple(f.unsqueeze(0).contiguous() for f in F)
        C = tuple(c.unsqueeze(0).contiguous() for c in C)
    k, K, _ = fused.riccati_backward(F, C)
    return (k[0], K[0]) if squeeze else (k, K)
