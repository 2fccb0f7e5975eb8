"""Shared argument handling of the igpc mirrors: everything is batched over a leading trajectory
axis N (the vmap axis); unbatched reference-shaped inputs ([H,m], state with arr [d]) are promoted
to N = 1 and the outputs squeezed back."""
import torch

from ..envs.classic import QuadCost

W_IS = {"de": 0, "dede": 1, "delin": 2}


def as_cost(cost_func, env):
    if isinstance(cost_func, QuadCost):
        return cost_func
    raise TypeError("the fused igpc kernels implement deluca_b200.envs.classic.QuadCost "
                    "(sum q (x-goal)^2 + sum r u^2); arbitrary Python cost functions have no CUDA path")


def batch_U(U, env):
    U = torch.stack(list(U)) if isinstance(U, (list, tuple)) else U
    U = torch.as_tensor(U, dtype=torch.float32, device=env.device)
    squeeze = U.dim() == 2
    if squeeze:
        U = U.unsqueeze(0)
    return U.contiguous(), squeeze


def batch_x0(start_state, env, N):
    if start_state is None:
        start_state = env.init()[0]
    arr = start_state.arr if hasattr(start_state, "arr") else start_state
    arr = torch.as_tensor(arr, dtype=torch.float32, device=env.device).reshape(-1, env.state_dim)
    return arr.expand(N, env.state_dim).contiguous()


def batch_alpha(alpha, N, dev):
    return torch.as_tensor(alpha, dtype=torch.float32, device=dev).reshape(-1).expand(N).contiguous()
