"""``deluca/learners/core.py:12-41``: the dataset generator of the system-identification learners.

``Learner.generate_trajectories(N, T, key)`` is vmap(scan) of a ``SimpleRandom`` policy
(``deluca/agents/_random.py:39-53``: action ~ N(0, 1)) over an env, returning the triple
``(obs, action, next_obs)`` with shapes [N,T,.] that ``NNLearner.learn`` consumes
(``deluca/learners/nn_learner.py:127-131``).  Here it is one launch of ``dlc_env_collect_f32``.

Stated differences: the reference env there is the brax ``Pendulum2D`` (out of scope); this mirror
runs the classic envs the kernels implement and the LDS env (observations y = C x).  The reference stores ``action[0]`` -- the first
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
        from ..envs._lds import LDS, LDSState
        if isinstance(env, LDS):   # obs = y = C x: one launch of dlc_lds_open_rollout_f32 with the entry observation
            st = start_state if start_state is not None else env.reset(N)[0]
            x0 = st.x.expand(N, env.d).contiguous()
            _, Y, U = fused.lds_open_rollout(*env._abc(N), x0, T, W=env._w(st.t, T, int(key), N, x0.device),
                                             w_std=env.disturbance.std() or 0.0, seed=int(key), t0=st.t,
                                             env_offset=env_offset, observe0=True, keep_actions=True)
            return Y[:-1].transpose(0, 1), U.transpose(0, 1), Y[1:].transpose(0, 1)   # [N,T,.] views
        if start_state is None:
            start_state = env.init()[0]
        arr = start_state.arr if hasattr(start_state, "arr") else start_state
        x0 = torch.as_tensor(arr, dtype=torch.float32, device=env.device).reshape(-1, env.state_dim)
        x0 = x0.expand(N, env.state_dim).contiguous()
        return fused.env_collect(env, x0, T, seed=int(key), env_offset=env_offset)
