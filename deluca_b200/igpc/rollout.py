"""``deluca/igpc/rollout.py``: trajectory rollout and linearisation on the fused CUDA path."""
import torch

from .. import fused
from ._common import as_cost, batch_alpha, batch_U, batch_x0


def rollout(env, cost_func, U_old, k=None, K=None, X_old=None, alpha=1.0, H=None, start_state=None):
    """``rollout`` (``rollout.py:22-65``): U = U_old + alpha k + K (x - X_old); returns (X, U, cost)
    with X[N,H+1,d], U[N,H,m], cost[N] (or the unbatched shapes when U_old is [H,m])."""
    cost = as_cost(cost_func, env)
    U_old, squeeze = batch_U(U_old, env)
    N, Hh, m = U_old.shape
    x0 = batch_x0(start_state, env, N)
    p = fused.plan_params(env, env, cost, N, Hh)
    if k is not None:
        k = torch.as_tensor(k, dtype=torch.float32, device=env.device).reshape(N, Hh, m).contiguous()
        K = torch.as_tensor(K, dtype=torch.float32, device=env.device).reshape(N, Hh, m, env.state_dim).contiguous()
        X_old = torch.as_tensor(X_old, dtype=torch.float32, device=env.device).reshape(
            N, Hh + 1, env.state_dim).contiguous()
    X, U, c = fused.env_rollout(p, U_old, x0, k, K, X_old, batch_alpha(alpha, N, env.device))
    return (X[0], U[0], c[0]) if squeeze else (X, U, c)


def compute_ders(env, cost, X, U, H=None):
    """``compute_ders`` (``rollout.py:68-99``): D, F = (F_x, F_u), C = (c_x, c_u, c_xx, c_uu), each
    stacked over [N,H,...] (closed-form Jacobians of the env step; no c_ux, like the reference)."""
    cost = as_cost(cost, env)
    U, squeeze = batch_U(U, env)
    N, Hh, m = U.shape
    X = torch.as_tensor(X, dtype=torch.float32, device=env.device).reshape(N, Hh + 1, env.state_dim).contiguous()
    p = fused.plan_params(env, env, cost, N, Hh)
    D, F, C = fused.linearize(p, X, U, use_sim=False)
    if squeeze:
        return D[0], tuple(f[0] for f in F), tuple(c[0] for c in C)
    return D, F, C
