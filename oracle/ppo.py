"""CPU restatement of the experience post-processing (test infrastructure; see oracle/__init__.py).

Parity pinned to the reference's own file run here: ``gae_advantages`` of ``deluca/training/ppo.py:44-84`` executed
literally on tests/refshim (``tests/golden/ref_drivers.npz``, tests/test_ref_fixtures_oracle.py); this follows it
statement by statement in the dtype it is given (float32 arrays reproduce the reference's arithmetic; float64 is the
yardstick).
"""
import numpy as np

from . import philox


def gae_advantages(rewards, terminal_masks, values, discount, gae_param):
    """rewards, terminal_masks, values: [T, num_agents] (ppo.py:55-58)."""
    dt = rewards.dtype.type
    discount, gae_param = dt(discount), dt(gae_param)
    advantages = [rewards[-1] - values[-1]]                                          # ppo.py:64
    gae = rewards[-1] - values[-1]                                                   # ppo.py:65
    for t in reversed(range(len(rewards) - 1)):                                      # ppo.py:66
        delta = rewards[t] + discount * values[t + 1] * terminal_masks[t] - values[t]   # ppo.py:67
        gae = delta + discount * gae_param * terminal_masks[t] * gae                 # ppo.py:68
        advantages.append(gae)
    return np.stack(advantages[::-1])                                                # ppo.py:70,84


def ac_tail(losses, log_probs, values, done_masks, gamma, lambda_, returns_use_values=False):
    """The lines after the scan in ac_rollout (ppo.py:131-134), for ONE agent's [T] series or,
    vmapped, [N,T] series (batch-major)."""
    rewards = -1.0 * losses                                                          # ppo.py:132
    T_first = lambda a: np.moveaxis(np.atleast_2d(a), -1, 0)
    adv = gae_advantages(T_first(rewards).astype(losses.dtype), T_first(done_masks), T_first(values), gamma, lambda_)
    adv = np.moveaxis(adv, 0, -1).reshape(losses.shape)
    returns = adv + (values if returns_use_values else log_probs)                    # ppo.py:133 (t[4] = log_probs)
    return adv, returns


def random_actions(seed, env_ids, T, m, std=1.0):
    """Actions of the collect kernel -> [N, T, m] float32: component c of step t is normal number f = t*m + c of
    the trajectory's stream, element f % 4 of the Philox block with counter (env, f // 4, 0, 2)."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    N = env_ids.shape[0]
    nblk = (T * m + 3) // 4
    ctr = np.zeros((N, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = env_ids[:, None]
    ctr[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 3] = 2
    key = np.zeros((N, nblk, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox.philox4x32(ctr, key)
    z0, z1 = philox.box_muller_f32(r[..., 0], r[..., 1])
    z2, z3 = philox.box_muller_f32(r[..., 2], r[..., 3])
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(N, nblk * 4)[:, :T * m].reshape(N, T, m)
    return (np.float32(std) * z).astype(np.float32)


def generate_trajectories(env, x0, actions):
    """deluca/learners/core.py:17-41 with the action stream given: scan of
    (obs, action) -> env step -> (obs, action, next_obs); env is an oracle.igpc env (batched step)."""
    N, T, _ = actions.shape
    x = np.array(x0, dtype=actions.dtype)
    obs = np.empty((N, T, x.shape[1]), dtype=x.dtype)
    nxt = np.empty_like(obs)
    for t in range(T):
        obs[:, t] = x
        x = env.step(x, actions[:, t])
        nxt[:, t] = x
    return obs, actions, nxt
