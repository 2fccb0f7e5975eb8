"""``deluca/igpc/lqr_solver.py``: the Riccati backward passes, on the device instead of host NumPy."""
import torch

from .. import fused


def lambda_adjust(reg_lambda, multiplier, consts, increase):
    """``lambda_adjust`` (``lqr_solver.py:21-37``), host scalars (the caller's outer loop uses the decreasing
    branch after a successful forward pass; the increasing branch also runs inside ``LQG`` on the device)."""
    factor, lambda_min = consts
    if increase:
        multiplier = max(multiplier * factor, factor)
        reg_lambda = max(reg_lambda * multiplier, lambda_min)
    else:
        multiplier = min(1 / factor, multiplier / factor)
        reg_lambda = reg_lambda * multiplier if reg_lambda * multiplier >= lambda_min else 0
    reg_lambda = max(reg_lambda, lambda_min)
    if reg_lambda > 1e10:
        reg_lambda = 1e10
    return reg_lambda, multiplier


def LQR(F, C):
    """``LQR`` (``lqr_solver.py:40-86``): F = (F_x[N,H,d,d], F_u[N,H,d,m]),
    C = (c_x, c_u, c_xx, c_uu) -> (k[N,H,m], K[N,H,m,d])."""
    squeeze = F[0].dim() == 3
    if squeeze:
        F = tuple(f.unsqueeze(0).contiguous() for f in F)
        C = tuple(c.unsqueeze(0).contiguous() for c in C)
    k, K, _ = fused.riccati_backward(F, C)
    return (k[0], K[0]) if squeeze else (k, K)


def LQG(F, C, reg_lambda=0.0, multiplier=1.0, damping_consts=(None, None), reg_type="V"):
    """``LQG`` (``lqr_solver.py:89-154``): regularised backward pass; C = (c_x, c_u, c_xx, c_uu, c_ux) with
    c_ux[N,H,m,d] is required, like the reference (``:92``).  Returns ``k, K, [reg_lambda, multiplier, gains_lin,
    gains_quad]``.  Batched inputs give [N] tensors for the four scalars, NaN gains and zero k/K rows standing for
    the reference's ``None`` when no lambda makes every Q_uu positive definite within nine attempts (``:100-103``);
    unbatched inputs return Python scalars and ``None`` exactly as the reference does."""
    if len(C) != 5:
        raise ValueError("LQG needs C = (c_x, c_u, c_xx, c_uu, c_ux)")
    if reg_type not in ("Q", "V"):
        raise ValueError("reg_type must be 'Q' or 'V'")
    factor, lambda_min = damping_consts
    if factor is None or lambda_min is None:
        raise TypeError("damping_consts = [factor, lambda_min] is required (the reference's default [None, None] "
                        "fails inside lambda_adjust, lqr_solver.py:24)")
    squeeze = F[0].dim() == 3
    if squeeze:
        F = tuple(f.unsqueeze(0).contiguous() for f in F)
        C = tuple(c.unsqueeze(0).contiguous() for c in C)
    N, dev = F[0].shape[0], F[0].device
    vec = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1).expand(N).contiguous().clone()
    lam, mult = vec(reg_lambda), vec(multiplier)
    k, K, gains, status = fused.lqg_backward(F, C, lam, mult, factor, lambda_min, 0 if reg_type == "Q" else 1)
    if squeeze:
        if int(status[0]) != 0:
            return None, None, [float(lam[0]), float(mult[0]), None, None]
        return k[0], K[0], [float(lam[0]), float(mult[0]), float(gains[0, 0]), float(gains[0, 1])]
    return k, K, [lam, mult, gains[:, 0], gains[:, 1]]
