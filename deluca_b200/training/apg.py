"""The scan x vmap rollout drivers of ``deluca/training/apg.py:39-100`` dispatched to fused kernels.

``apg_rollout`` is ``lax.scan`` of (agent -> env -> loss) over ``unroll_length`` steps and
``apg_parallel_rollouts`` is its ``jax.vmap`` with ``in_axes=(None, 0, 0, None, 0, None, None)``:
env and agent objects are shared, start states carry the batch axis.  Here the batch axis is always
present (N >= 1) and the whole scan x vmap is ONE kernel launch for the (env, agent, loss) families that
have a fused kernel:

    LDS            x GPC | BPC | LQR          x quad_loss                  -> dlc_lds_rollout_f32
    LDS            x LQR                      x QuadLoss (Q, R diagonal)   -> dlc_lds_rollout_f32 costs re-weighted
                                                                              from the emitted states / actions
    Pendulum | Cartpole | PlanarQuadrotor | Reacher
                   x generic PID | LQR (u = -K x)   x QuadLoss             -> dlc_env_agent_rollout_f32
    BalloonLung | DelayLung x lung PID (+ Expiratory, as run_controller_scan composes them)
                                              x LungLoss (l1 | l2 of target - pressure)
                                                                           -> dlc_lung_pid_rollout_f32

A Python ``loss_func`` outside these declarative families, or an (env, agent) pair outside the table, raises
``NotImplementedError`` -- there is no eager / CPU stepping fallback.

``Rollout.agent_states``: the reference stacks the agent state after EVERY step.  For the generic PID that stack is
3 d floats per step and is emitted in full by default.  For GPC / BPC it would be H m d floats per env per step, so
the default is the FINAL state only (a stated deviation); ``state_stride=k`` opts in to the stack after every k-th
step (the rollout is then launched in chunks of k steps, resumed from the carried state -- bit-identical to one
launch, tests/test_mirror_gpu.py).
"""
from typing import Any, NamedTuple

import torch

from .. import fused
from ..agents._bpc import BPC, BPCState
from ..agents._gpc import GPC, GPCState, quad_loss
from ..agents._lqr import LQR
from ..agents._pid import PID, PIDState
from ..envs._lds import LDS, LDSState
from ..envs.classic._pendulum import QuadCost, _ClassicEnv


class QuadLoss(QuadCost):
    """A declarative ``loss_func(env_state, env_obs, action, env, counter)`` (``apg.py:54``):
    sum q (x - goal)^2 + sum r (action[:m] - goal_action)^2, the family the fused env x agent kernel evaluates.
    ``goal=None`` takes ``env.goal_state``."""

    def __init__(self, goal=None, q=1.0, r=0.1, goal_action=None):
        super().__init__(goal, q, r, goal_action)

    def bound(self, env):
        goal = self.goal if self.goal is not None else getattr(env, "goal_state", None)
        if goal is None:
            goal = [0.0] * env.state_dim
        goal = goal.tolist() if hasattr(goal, "tolist") else list(goal)
        return QuadCost(goal, self.q, self.r, self.goal_action)

    def __call__(self, env_state, env_obs, action, env=None, counter=None):
        m = env.action_dim if env is not None else action.shape[-1]
        return self.bound(env)(env_state, action[..., :m], env)


class LungLoss:
    """``loss_fn(target, pressure)`` of the lung training scripts (``train_controller.py:34-52``: ``l1`` = |t - p|,
    ``l2`` = (t - p)^2) as an apg ``loss_func``."""

    def __init__(self, kind="l1"):
        if kind not in ("l1", "l2"):
            raise ValueError("kind must be 'l1' or 'l2'")
        self.kind = kind

    def series(self, target, pressure):
        e = target - pressure
        return e.abs() if self.kind == "l1" else e * e


class Rollout(NamedTuple):
    """``Rollout`` (``apg.py:30-36``)."""
    agent_states: Any
    actions: Any
    losses: Any


def _classic_rollout(env, env_start_states, agent, agent_start_states, T, loss_func, state_stride, want_grad=False):
    """Classic env x (generic PID | LQR-type linear feedback) x QuadLoss -> dlc_env_agent_rollout_f32."""
    if not isinstance(loss_func, QuadLoss):
        raise NotImplementedError("the fused env x agent kernel evaluates the declarative QuadLoss family; an arbitrary "
                                  "Python loss_func has no CUDA path")
    dev = torch.device(env.device)
    arr = torch.as_tensor(env_start_states.arr, dtype=torch.float32, device=dev)
    x = arr.reshape(-1, env.state_dim).contiguous()
    N, d = x.shape
    cost = loss_func.bound(env)
    if isinstance(agent, PID):
        tens = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
        kp, ki, kd = tens(agent.K_P), tens(agent.K_I), tens(agent.K_D)
        per_env = max(kp.dim(), ki.dim(), kd.dim()) > 0
        gains = torch.stack([g.reshape(-1).expand(N) if per_env else g for g in (kp, ki, kd)], -1).contiguous()
        s = agent_start_states
        def leaf(v):    # scalar (the reference's PIDState default 0.0), [d] (one episode) or [N,d]
            t = tens(v)
            if t.numel() == 1:
                return t.reshape(1, 1).expand(N, d)
            if t.numel() == d:
                return t.reshape(1, d).expand(N, d)
            return t.reshape(N, d)
        ps = [leaf(v) for v in (s.P, s.I, s.D)]
        stride = 1 if state_stride is None else int(state_stride)
        r = fused.env_agent_rollout(env, cost, "pid", gains, x, T, pid_state=ps, pid_decay=agent.decay,
                                    state_stride=stride, want_grad=want_grad)
        if stride > 0:       # [Ts,N,3,d] -> vmap-first stacks [N,Ts,d]
            stk = r.agent_stack.permute(2, 1, 0, 3)
            states = PIDState(P=stk[0], I=stk[1], D=stk[2])
        else:
            states = PIDState(P=r.agent_state[0], I=r.agent_state[1], D=r.agent_state[2])
    elif isinstance(agent, LQR):
        K = agent.K.to(dev).float().contiguous()
        stride = 0 if state_stride is None else int(state_stride)
        r = fused.env_agent_rollout(env, cost, "linear", K, x, T, state_stride=stride, want_grad=want_grad)
        states = agent_start_states
    else:
        raise NotImplementedError(f"no fused rollout kernel for agent {type(agent).__name__} on {type(env).__name__}")
    return r, states


def _lung_rollout(env, agent, T, loss_func, N):
    """BalloonLung / DelayLung x lung PID (+ Expiratory): the closed loop ``run_controller_scan`` composes
    (``run_controller.py:111-141``), with per-step loss = loss_fn(target, pressure)."""
    from ..lung.utils.scripts.run_controller import run_controller_scan
    if not isinstance(loss_func, LungLoss):
        raise NotImplementedError("lung rollouts take a LungLoss (l1 | l2 of target - pressure)")
    res = run_controller_scan(agent, T=T, dt=env.dt, env=env, N=N)
    ts = res["timeseries"]
    two = lambda v: v.reshape(1, -1) if v.dim() == 1 else v
    losses = loss_func.series(two(ts["target"]), two(ts["pressure"]))
    actions = torch.stack([two(ts["u_in"]), two(ts["u_out"])], -1)
    return Rollout(agent_states=res["final"]["buffers"], actions=actions, losses=losses), res["final"]["env"]


def _weighted_lds_losses(r, loss_func, x0):
    """Q, R-weighted quadratic loss on LDS from the emitted trajectory (x_out[t] is the state the agent saw at step t,
    include/deluca_b200.h): loss_t = sum q (x_t - goal)^2 + sum r (u_t - goal_u)^2   (apg.py:54)."""
    X = r.X
    d, m = X.shape[-1], r.U.shape[-1]
    q, rr, goal = loss_func.vectors(d, m) if loss_func.goal is not None else QuadCost([0.0] * d, loss_func.q,
                                                                                     loss_func.r).vectors(d, m)
    gu = loss_func.action_goal(m)
    tq, tr = torch.tensor(q, device=X.device), torch.tensor(rr, device=X.device)
    tg, tgu = torch.tensor(goal, device=X.device), torch.tensor(gu, device=X.device)
    return (((X - tg) ** 2) * tq).sum(-1) + (((r.U - tgu) ** 2) * tr).sum(-1)


def apg_parallel_rollouts(env, env_start_states, env_init_obs, agent, agent_start_states, unroll_length, loss_func,
                          env_offset=0, disturbances=None, state_stride=None):
    """``apg_parallel_rollouts`` (``apg.py:78-100``).  Returns (Rollout, final env state)."""
    T = int(unroll_length)
    if isinstance(env, _ClassicEnv):
        r, states = _classic_rollout(env, env_start_states, agent, agent_start_states, T, loss_func, state_stride)
        upd = dict(arr=r.x, h=env_start_states.h + T)
        if hasattr(env_start_states, "last_action"):                      # _planar_quadrotor.py:117
            upd["last_action"] = r.actions[-1, :, :env.action_dim]
        final = env_start_states.replace(**upd)
        return Rollout(agent_states=states, actions=r.actions.transpose(0, 1), losses=r.losses.T), final
    from ..lung.core import LungEnv
    if isinstance(env, LungEnv):
        N = None if env_start_states is None else getattr(env_start_states, "N", None)
        return _lung_rollout(env, agent, T, loss_func, N)
    if not isinstance(env, LDS):
        raise NotImplementedError(f"no fused rollout kernel for env {type(env).__name__}")
    weighted = isinstance(loss_func, QuadLoss)
    if loss_func is not quad_loss and not weighted:
        raise NotImplementedError("the fused LDS kernels realise quad_loss (x'x + u'u) and its diagonal Q, R-weighted "
                                  "form QuadLoss")
    # (loss_func only scores the rollout, apg.py:54; GPC / BPC keep learning from their own cost_fn = quad_loss inside
    # the kernel, deluca/agents/_gpc.py:55-56, so a QuadLoss here re-weights the reported losses and nothing else)
    if state_stride:
        return _lds_strided(env, env_start_states, env_init_obs, agent, agent_start_states, T, loss_func, env_offset,
                            disturbances, int(state_stride))
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
                              decay=agent.decay, noise_uses_prev_action=agent.noise_uses_prev_action,
                              cost_q=agent.cost_q, cost_r=agent.cost_r, **common)
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
    losses = _weighted_lds_losses(r, loss_func, x) if weighted else r.costs
    # vmap puts the batch axis first, scan stacks time after it: [N, T, ...]
    return Rollout(agent_states=final, actions=r.U.transpose(0, 1), losses=losses.T), env_final


def _lds_strided(env, es, obs, agent, ast, T, loss_func, env_offset, disturbances, k):
    """Opt-in ``Rollout.agent_states`` stack after every k-th step: the rollout runs in chunks of k steps, each resumed
    from the carried (env, agent) state -- the same arithmetic as one launch."""
    if T % k:
        raise ValueError("unroll_length must be a multiple of state_stride")
    acts, losses, stack = [], [], []
    for c in range(T // k):
        W = None if disturbances is None else disturbances[c * k:(c + 1) * k].contiguous()
        ro, es = apg_parallel_rollouts(env, es, obs, agent, ast, k, loss_func, env_offset=env_offset, disturbances=W)
        ast = ro.agent_states
        acts.append(ro.actions); losses.append(ro.losses); stack.append(ast)
    leaves = {}
    if hasattr(ast, "replace"):
        for name in ("M", "noise_history", "state", "action", "eps", "cost"):
            v = getattr(ast, name, None)
            if isinstance(v, torch.Tensor):
                leaves[name] = torch.stack([getattr(s, name) for s in stack], 1)     # [N, T/k, ...]
        stacked = ast.replace(**leaves) if leaves else ast
    else:
        stacked = ast
    return Rollout(agent_states=stacked, actions=torch.cat(acts, 1), losses=torch.cat(losses, 1)), es


def apg_rollout(env, env_start_state, env_init_obs, agent, agent_start_state, unroll_length, loss_func, **kw):
    """``apg_rollout`` (``apg.py:39-68``): same call, leaves already carry N (N = 1 for one episode)."""
    return apg_parallel_rollouts(env, env_start_state, env_init_obs, agent, agent_start_state, unroll_length,
                                 loss_func, **kw)


def value_and_grad_apg_loss(agent, env, env_start_states, env_init_obs, agent_init_states, unroll_length, loss_func,
                            env_offset=0, disturbances=None):
    """``jax.value_and_grad(apg_loss)`` (``apg.py:103-122,152-167``) for a linear state-feedback agent (LQR: the
    leaf is ``agent.K``) on LDS with ``quad_loss``: returns (mean loss over batch and horizon, d mean-loss / dK)
    with the gradient shaped like ``agent.K`` ([m,d] for a shared gain = summed over envs, [N,m,d] per env).
    Forward and reverse pass run in ``dlc_lds_lqr_grad_f32`` (hand adjoint).

    On a classic env with ``QuadLoss`` the leaves are the generic PID's gains (the gradient comes back as a ``PID`` whose
    ``K_P, K_I, K_D`` hold d value / d gain, the pytree ``jax.value_and_grad`` returns) or an LQR-type agent's ``K``;
    the tangents ride along the forward rollout in ``dlc_env_agent_rollout_f32`` (3 or m d leaves: forward mode needs
    no tape)."""
    if isinstance(env, _ClassicEnv):
        T = int(unroll_length)
        r, _ = _classic_rollout(env, env_start_states, agent, agent_init_states, T, loss_func, 0, want_grad=True)
        N = r.loss_sum.shape[0]
        denom = float(N * T)
        value = r.loss_sum.sum() / denom
        if isinstance(agent, PID):
            per_env = any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in (agent.K_P, agent.K_I, agent.K_D))
            g = r.grad / denom if per_env else r.grad.double().sum(0).float() / denom
            return value, agent.replace(K_P=g[..., 0], K_I=g[..., 1], K_D=g[..., 2])
        g = r.grad.reshape(N, *agent.K.shape[-2:])
        return value, (g / denom if agent.K.dim() == 3 else g.double().sum(0).float() / denom)
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
