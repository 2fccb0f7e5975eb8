"""The scan x vmap rollout drivers of ``deluca/training/apg.py:39-100`` dispatched to fused kernels.

``apg_rollout`` is ``lax.scan`` of (agent -> env -> loss) over ``unroll_length`` steps and
``apg_parallel_rollouts`` is its ``jax.vmap`` with ``in_axes=(None, 0, 0, None, 0, None, None)``:
env and agent objects are shared, start states carry the batch axis.  Here the batch axis is always
present (N >= 1) and the whole scan x vmap is ONE kernel launch for the (env, agent) families that
have a fused kernel:

    LDS  x  GPC | BPC | LQR   with loss_func = quad_loss      -> dlc_lds_rollout_f32

Anything else raises ``NotImplementedError`` -- there is no eager / CPU stepping fallback.

Deviation from the reference (stated): ``Rollout.agent_states`` holds the FINAL agent state, not the
T-stacked states (for GPC that stack would be T x H x m x d floats per env; the kernels emit only
costs and actions, BASELINE.json north_star).
"""
from typing import Any, NamedTuple

import torch

from .. import fused
from ..agents._bpc import BPC, BPCState
from ..agents._gpc import GPC, GPCState, quad_loss
from ..agents._lqr import LQR
from ..envs._lds import LDS, LDSState


class Rollout(NamedTuple):
    """``Rollout`` (``apg.py:30-36``)."""
    agent_states: Any
    actions: Any
    losses: Any


def apg_parallel_rollouts(env, env_start_states, env_init_obs, agent, agent_start_states, unroll_length, loss_func,
                          env_offset=0, disturbances=None):
    """``apg_parallel_rollouts`` (``apg.py:78-100``).  Returns (Rollout, final env state)."""
    if not isinstance(env, LDS):
        raise NotImplementedError(f"no fused rollout kernel for env {type(env).__name__}")
    if loss_func is not quad_loss:
        raise NotImplementedError("the fused LDS kernels realise the reference's quad_loss (x'x + u'u) only")
    T = int(unroll_length)
    x = env_start_states.x
    std = env.disturbance.std()
    W = disturbances
    if W is None and (std is None or std == 0.0):
        W = torch.stack([env.disturbance(env_start_states.t + t, env.key, N=x.shape[0], device=x.device)
                         for t in range(T)]).contiguous()
    # the ENV's step counter keys the in-kernel disturbance stream; the agent's own counter (s.t) only drives its
    # learning-rate decay, so chunked rollouts and fresh agents on a continued env neither replay nor skip noise
    common = dict(W=W, w_std=std or 0.0, seed=env.key, env_offset=env_offset, traj_stride=1,
                  env_t0=env_start_states.t)
    if isinstance(agent, GPC):
        s = agent_start_states
        r = fused.lds_rollout(env.A, env.B, agent.K.to(x.device), x, T, policy="gpc", H=agent.H, HH=agent.HH, M=s.M,
                              hist=s.noise_history, xprev=s.state, uprev=s.action, t0=s.t, lr_scale=agent.lr_scale,
                              decay=agent.decay, noise_uses_prev_action=agent.noise_uses_prev_action, **common)
        final = GPCState(M=r.M, noise_history=r.hist, state=r.xprev, action=r.uprev, t=s.t + T)
    elif isinstance(agent, BPC):
        s = agent_start_states
        r = fused.lds_rollout(env.A, env.B, agent.K.to(x.device), x, T, policy="bpc", H=agent.H, M=s.M,
                              hist=s.noise_history, xprev=s.state, eps=s.eps, bpc_cost=s.cost, t0=s.t,
                              lr_scale=agent.lr_scale, decay=agent.decay, bpc_delta=agent.delta, **common)
        final = BPCState(M=r.M, noise_history=r.hist, state=r.xprev, eps=r.eps, cost=r.bpc_cost, t=s.t + T)
    elif isinstance(agent, LQR):
        r = fused.lds_rollout(env.A, env.B, agent.K.to(x.device), x, T, policy="lqr", **common)
        final = agent_start_states
    else:
        raise NotImplementedError(f"no fused rollout kernel for agent {type(agent).__name__} on LDS")
    env_final = LDSState(x=r.x, t=env_start_states.t + T)
    # vmap puts the batch axis first, scan stacks time after it: [N, T, ...]
    return Rollout(agent_states=final, actions=r.U.transpose(0, 1), losses=r.costs.T), env_final


def apg_rollout(env, env_start_state, env_init_obs, agent, agent_start_state, unroll_length, loss_func, **kw):
    """``apg_rollout`` (``apg.py:39-68``): same call, leaves already carry N (N = 1 for one episode)."""
    return apg_parallel_rollouts(env, env_start_state, env_init_obs, agent, agent_start_state, unroll_length,
                                 loss_func, **kw)


def value_and_grad_apg_loss(agent, env, env_start_states, env_init_obs, agent_init_states, unroll_length, loss_func,
                            env_offset=0, disturbances=None):
    """``jax.value_and_grad(apg_loss)`` (``apg.py:103-122,152-167``) for a linear state-feedback agent (LQR: the
    leaf is ``agent.K``) on LDS with ``quad_loss``: returns (mean loss over batch and horizon, d mean-loss / dK)
    with the gradient shaped like ``agent.K`` ([m,d] for a shared gain = summed over envs, [N,m,d] per env).
    Forward and reverse pass run in ``dlc_lds_lqr_grad_f32`` (hand adjoint)."""
    if not isinstance(env, LDS) or not isinstance(agent, LQR) or loss_func is not quad_loss:
        raise NotImplementedError("the fused reverse-mode kernel covers LDS x linear state feedback (LQR.K) x quad_loss")
    T = int(unroll_length)
    x = env_start_states.x
    N = x.shape[0]
    std = env.disturbance.std()
    W = disturbances
    if W is None and (std is None or std == 0.0):
        W = torch.stack([env.disturbance(env_start_states.t + t, env.key, N=N, device=x.device)
                         for t in range(T)]).contiguous()
    K = agent.K.to(x.device)
    per_env_dyn = env.A.dim() == 3
    if per_env_dyn and K.dim() == 2:
        K = K.expand(N, *K.shape).contiguous()
    A, B = env.A, env.B
    if K.dim() == 3 and not per_env_dyn:
        A, B = A.expand(N, *A.shape).contiguous(), B.expand(N, *B.shape).contiguous()
    loss, gK, _, sums = fused.lds_lqr_value_and_grad(A, B, K, x, T, W=W, seed=env.key, env_offset=env_offset,
                                                     t0=env_start_states.t, w_std=std or 0.0)
    denom = float(N * T)
    value = sums[0] / denom
    if agent.K.dim() == 2:
        return value, (sums[1:] / denom).float().reshape(agent.K.shape)
    return value, gK / denom


def apg_loss(agent, env, env_start_states, env_init_obs, agent_init_states, unroll_length, loss_func,
             env_offset=0, disturbances=None):
    """Mean loss over the batch and horizon, the forward value of ``apg_loss`` (``apg.py:103-122``).  Argument order
    as in the reference: the agent comes first because ``jax.value_and_grad`` differentiates argument 0."""
    ro, _ = apg_parallel_rollouts(env, env_start_states, env_init_obs, agent, agent_init_states, unroll_length,
                                  loss_func, env_offset=env_offset, disturbances=disturbances)
    return ro.losses.mean()
