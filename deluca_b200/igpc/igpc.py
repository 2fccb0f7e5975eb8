"""``deluca/igpc/igpc.py``: GPC_rollout and the iGPC_closed loop on the fused CUDA path."""
import torch

from .. import fused
from ._common import W_IS, as_cost, batch_alpha, batch_U, batch_x0


def GPC_rollout(env_true, env_sim, cost_func, U_old, k, K, X_old, D, F, alpha, w_is, ref_lr=0.1, start_state=None):
    """``GPC_rollout`` (``igpc.py:63-126``).  ``D`` and ``F`` are accepted for signature parity; the
    kernel re-derives them from (X_old, U_old) under env_sim, which is how ``iGPC_closed`` obtains
    them (``:150``)."""
    cost = as_cost(cost_func, env_sim)
    U_old, squeeze = batch_U(U_old, env_true)
    N, Hh, m = U_old.shape
    dev, d = env_true.device, env_true.state_dim
    x0 = batch_x0(start_state, env_true, N)
    p = fused.plan_params(env_true, env_sim, cost, N, Hh, w_is=W_IS[w_is], lr=ref_lr)
    t = lambda a, *s: torch.as_tensor(a, dtype=torch.float32, device=dev).reshape(*s).contiguous()
    X, U, c = fused.igpc_rollout(p, U_old, t(k, N, Hh, m), t(K, N, Hh, m, d), t(X_old, N, Hh + 1, d),
                                 batch_alpha(alpha, N, dev), x0)
    return (X[0], U[0], c[0]) if squeeze else (X, U, c)


def iGPC_closed(env_true, env_sim, cost_func, U, T, k=None, K=None, X=None, w_is="de", lr=None, ref_alpha=1.0,
                backtracking=True, verbose=False, start_state=None):
    """``iGPC_closed`` (``igpc.py:129-179``) with ``verbose=False`` semantics (the LR/100 annealing
    sits inside ``if verbose:`` in the reference, ``:153-156``).  Returns (X, U, k, K, c)."""
    assert w_is in W_IS
    if lr is None:
        raise ValueError("lr is required (the reference's default None fails at `lr / h`, igpc.py:124)")
    cost = as_cost(cost_func, env_sim)
    U, squeeze = batch_U(U, env_true)
    N, Hh, _ = U.shape
    x0 = batch_x0(start_state, env_true, N)
    p = fused.plan_params(env_true, env_sim, cost, N, Hh, T=T, algo=1, backtracking=backtracking, n_alpha=6,
                          w_is=W_IS[w_is], alpha0=ref_alpha, lr=lr)
    X, U, k, K, c, _, _ = fused.plan(p, U, x0)
    return (X[0], U[0], k[0], K[0], c[0]) if squeeze else (X, U, k, K, c)
