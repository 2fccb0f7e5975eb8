"""``deluca/igpc/ilqr.py``: the iLQR loop, fused into one launch per call (no host round trips)."""
from .. import fused
from ._common import as_cost, batch_U, batch_x0


def iLQR(env, cost, U, T, k=None, K=None, X=None, info="true", start_state=None, H=None, verbose=False,
         alpha=0.5, backtracking=True, fix_backtracking=False):
    """``iLQR`` (``ilqr.py:27-93``).  Returns (X, U, k, K, c).  ``fix_backtracking=False`` keeps the
    reference's slip of passing ``alpha`` instead of ``alphaC`` to the candidate rollouts
    (``:58-67``, SURVEY A-10)."""
    if k is not None:
        raise NotImplementedError("warm-started gains are not supported by the fused loop")
    cost = as_cost(cost, env)
    U, squeeze = batch_U(U, env)
    N, Hh, _ = U.shape
    x0 = batch_x0(start_state, env, N)
    p = fused.plan_params(env, env, cost, N, Hh, T=T, algo=0, backtracking=backtracking,
                          fix_backtracking=fix_backtracking, n_alpha=10, alpha0=alpha)
    X, U, k, K, c, _, _ = fused.plan(p, U, x0)
    return (X[0], U[0], k[0], K[0], c[0]) if squeeze else (X, U, k, K, c)
