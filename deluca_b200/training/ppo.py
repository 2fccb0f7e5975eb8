"""Experience post-processing of ``deluca/training/ppo.py`` on the fused CUDA path.

What is mirrored: the ``Rollout`` record (``ppo.py:27-39``), ``gae_advantages``
(``:44-84``), the tail of ``ac_rollout`` that turns stacked per-step series into advantages and
returns (``:131-134``) and the flattening of ``ac_get_experience`` (``:173-199``).  The scan itself
needs an actor-critic agent that emits ``(action, log_probs, values)``; the only one the reference
ships is the neural ``DeepAC`` (``lung/controllers/_deep_actor_critic.py``), which is out of scope
(DESIGN.md section 7), so ``ac_parallel_rollouts`` takes the stacked series of whichever fused
rollout produced them.

Reproduced as written: ``returns = advantages + log_probs`` (``t[4]`` at ``ppo.py:133`` is the
log-prob series, not the values); ``returns_use_values=True`` gives the intended sum.  The last
step's advantage is ``rewards[-1] - values[-1]`` with no bootstrap (``:64-65``).
"""
from typing import Any, NamedTuple

import torch

from .. import fused


class Rollout(NamedTuple):
    """``Rollout`` (``ppo.py:27-39``)."""
    agent_states: Any
    actions: Any
    log_probs: Any
    values: Any
    losses: Any
    done_masks: Any
    advantages: Any
    returns: Any


def gae_advantages(rewards, terminal_masks, values, discount, gae_param):
    """``gae_advantages`` (``ppo.py:44-84``): arrays shaped (actor_steps, num_agents)."""
    neg = torch.neg(rewards).contiguous()  # the kernel takes losses = -rewards (exact)
    adv, _ = fused.gae(neg, values.contiguous(), None, terminal_masks.contiguous(), discount, gae_param,
                       time_major=True, want_returns=False)
    return adv


def ac_parallel_rollouts(agent_states, actions, log_probs, values, losses, gamma, lambda_, done_masks=None,
                         returns_use_values=False):
    """The part of ``ac_parallel_rollouts`` / ``ac_rollout`` after the scan (``ppo.py:129-134``):
    [N,T] series in, ``Rollout`` with advantages and returns out.  ``done_masks`` defaults to the
    constant 1.0 the reference emits (``:113``) and is then materialised only in the record."""
    adv, ret = fused.gae(losses, values, log_probs, done_masks, gamma, lambda_, time_major=False,
                         returns_use_values=returns_use_values)
    masks = done_masks if done_masks is not None else torch.ones_like(losses)
    return Rollout(agent_states, actions, log_probs, values, losses, masks, adv, ret)


def ac_get_experience(*args, **kwargs):
    """``ac_get_experience`` (``ppo.py:173-199``): the [N,T,...] leaves flattened to [N*T,...] (views)."""
    r = ac_parallel_rollouts(*args, **kwargs)

    def flat(x):
        return x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:])) if isinstance(x, torch.Tensor) else x

    return Rollout(*[flat(x) for x in r])
