"""``deluca/igpc/ilc.py``: iLC_closed on the fused CUDA path."""
from .. import fused
from ._common import as_cost, batch_U, batch_x0


def iLC_closed(env_true, env_sim, cost_func, U, T, k=None, K=None, X=None, ref_alpha=1.0, verbose=False,
               backtracking=True, start_state=None):
    """``iLC_closed`` (``ilc.py:23-69``) with backtracking and ``verbose=False``."""
    if not backtracking:
        raise NotImplementedError("without backtracking the reference discards its single rollout and breaks "
                                  "(ilc.py:60-68): nothing to compute")
    cost = as_cost(cost_func, env_sim)
    U, squeeze = batch_U(U, env_true)
    N, Hh, _ = U.shape
    x0 = batch_x0(start_state, env_true, N)
    p = fused.plan_params(env_true, env_sim, cost, N, Hh, T=T, algo=2, backtracking=True, n_alpha=10,
                          alpha0=ref_alpha)
    X, U, k, K, c, _, _ = fused.plan(p, U, x0)
    return (X[0], U[0], k[0], K[0], c[0]) if squeeze else (X, U, k, K, c)
