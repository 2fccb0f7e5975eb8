"""``deluca/learners/core.py:12-41``: the dataset generator of the system-identification learners.

``Learner.generate_trajectories(N, T, key)`` is vmap(scan) of a ``SimpleRandom`` policy
(``deluca/agents/_random.py:39-53``: action ~ N(0, 1)) over an env, returning the triple
``(obs, action, next_obs)`` with shapes [N,T,.] that ``NNLearner.learn`` consumes
(``deluca/learners/nn_learner.py:127-131``).  Here it is one launch of ``dlc_env_collect_f32``.

Stated differences: the reference env there is the brax ``Pendulum2D`` (out of scope); this mirror
runs the classic envs the kernels implement.  The reference stores ``action[0]`` -- the first
component only, shape (1,) -- which coincides with the full action for its single-input envs; for
m > 1 all m components are drawn and stored.  Actions come from the library's Philox stream keyed
by ``key`` (an integer seed), not from jax.random.  ``learn`` / ``predict`` (flax training of an
MLP) are outside the rollout path.
"""
import torch

from .. import fused


class Learner:
    def __init__(self, env):
        self.env = env

    def generate_trajectories(self, N, T, key=0, start_state=None, env_offset=0):
        env = self.env
        if start_state is None:
            start_state = env.init()[0]
        arr = start_state.arr if hasattr(start_state, "arr") else start_state
        x0 = torch.as_tensor(arr, dtype=torch.float32, device=env.device).reshape(-1, env.state_dim)
        x0 = x0.expand(N, env.state_dim).contiguous()
        return fused.env_collect(env, x0, T, seed=int(key), env_offset=env_offset)
