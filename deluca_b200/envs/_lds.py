"""LDS environment (``deluca/envs/_lds.py``) on the fused CUDA path.

Two faces, as in the reference:
  * stateful, the reference's own:  ``env = LDS(); env.init(d_in, d_hidden, d_out, ...)``,
    ``y = env(u)`` mutates ``env.x`` / ``env.t`` (``_lds.py:58-111``);
  * functional pytree, what ``apg_rollout`` x ``vmap`` needs: ``state = env.reset(N)``,
    ``state, obs = env.step(state, u)`` with ``LDSState(x[N,d], t)`` (SURVEY A-1).
Every state update and observation is computed by ``dlc_lds_open_rollout_f32`` (T = 1 for a single call).  Disturbances
come from the library's counter-based Gaussian stream (Threefry is not reproducible without JAX,
SURVEY a3); ``key`` is the integer stream seed.
"""
import math

import torch

from ..core import Disturbance, Env, Obj, field
from .. import fused


class ZeroDisturbance(Disturbance):
    """``_lds.py:24-32``."""

    def init(self, d):
        self.d = d

    def std(self):
        return 0.0

    def __call__(self, t, key, N=1, device="cuda"):
        return torch.zeros(N, self.d, device=device)


class GaussianDisturbance(Disturbance):
    """``_lds.py:35-41``: N(0, 1) * std, std = 0.5."""

    def init(self, d, std=0.5):
        self.d = d
        self._std = std

    def std(self):
        return self._std

    def __call__(self, t, key, N=1, device="cuda", env_offset=0):
        return fused.gaussian_disturbance(N, 1, self.d, seed=int(key), env_offset=env_offset, t0=int(t),
                                          std=self._std, device=device)[0]


class SinusDisturbance(Disturbance):
    """``_lds.py:44-52``: sin(t * phases) * amplitude; phases ~ N(0,1) from stream seed 0."""

    def init(self, d, amplitude=0.5, device="cuda"):
        self.d = d
        self.amplitude = amplitude
        self.phases = fused.gaussian_disturbance(1, 1, d, seed=0, std=1.0, device=device)[0, 0]

    def std(self):
        return None

    def __call__(self, t, key, N=1, device="cuda"):
        return self.series(t, 1, N)[0]

    def series(self, t0, T, N):
        """W[T,N,d] for steps t0 .. t0+T-1 in one launch of ``dlc_sinus_disturbance_f32``."""
        return fused.sinus_disturbance(N, T, self.phases.contiguous(), self.amplitude, t0=int(t0))


class LDSState(Obj):
    """Functional state of an LDS batch: x[N,d] and the step counter (SURVEY A-1)."""
    x: torch.Tensor = field(None, jaxed=True)
    t: int = field(0, jaxed=False)


class LDS(Env):
    """``deluca.envs.LDS`` (``_lds.py:55-111``).  A[d,d] / B[d,n] / C[p,d] may also be given with a
    leading env axis [N,...] (per-env systems, the vmapped case)."""

    def init(self, d_in=1, d_hidden=25, d_out=1, A=None, B=None, C=None, key=0, x0=None, disturbance=None,
             device="cuda"):
        g = torch.Generator(device=device).manual_seed(int(key))
        dev = torch.device(device)
        as_t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
        self.A = as_t(A) if A is not None else torch.diag(
            0.9 + 0.05 * torch.rand(d_hidden, device=dev, generator=g))            # _lds.py:71-77
        self.B = as_t(B) if B is not None else torch.randn(d_hidden, d_in, device=dev, generator=g)
        self.C = as_t(C) if C is not None else torch.randn(d_out, d_hidden, device=dev, generator=g)
        self.d, self.n, self.p = self.A.shape[-1], self.B.shape[-1], self.C.shape[-2]
        self.x = torch.zeros(self.d, 1, device=dev) if x0 is None else as_t(x0).reshape(self.d, 1)
        self.t = 0
        self.disturbance = disturbance if disturbance is not None else GaussianDisturbance()
        self.disturbance.init(self.d)
        self.key = int(key)
        return self.x

    # ---- functional face
    def reset(self, N=1, x0=None):
        dev = self.A.device
        x = torch.zeros(N, self.d, device=dev) if x0 is None else torch.as_tensor(
            x0, dtype=torch.float32, device=dev).reshape(N, self.d).contiguous()
        st = LDSState(x=x, t=0)
        st.freeze()
        return st, self.observe(st)

    def _abc(self, N):
        """A, B, C with one batching convention for the kernel: all shared, or all per env."""
        mats = (self.A, self.B, self.C)
        if all(m.dim() == 2 for m in mats) or all(m.dim() == 3 for m in mats):
            return mats
        return tuple(m if m.dim() == 3 else m.unsqueeze(0).expand(N, *m.shape).contiguous() for m in mats)

    def _w(self, t0, T, key, N, dev):
        """None when the kernel draws the Gaussian stream itself, else the materialised W[T,N,d]."""
        std = self.disturbance.std()
        if std is None or std == 0.0:
            if T <= 0:
                return None
            if hasattr(self.disturbance, "series"):
                return self.disturbance.series(t0, T, N)
            return torch.stack([self.disturbance(t0 + t, key, N=N, device=dev).reshape(N, self.d)
                                for t in range(T)]).contiguous()
        return None

    def observe(self, state):
        """y = C x (``_lds.py:108``), through the kernel (T = 0, observe-only)."""
        _, y, _ = fused.lds_open_rollout(*self._abc(state.x.shape[0]), state.x, 0, observe0=True)
        return y[0]

    def step(self, state, u, key=None, env_offset=0):
        """``Env.__call__(state, action)``: x <- A x + B u + w_t, y = C x, t <- t + 1 (``_lds.py:102-111``)."""
        N = state.x.shape[0]
        u = torch.as_tensor(u, dtype=torch.float32, device=state.x.device).reshape(1, N, self.n).contiguous()
        key = self.key if key is None else int(key)
        x, y, _ = fused.lds_open_rollout(*self._abc(N), state.x, 1, u_in=u,
                                         W=self._w(state.t, 1, key, N, state.x.device),
                                         w_std=self.disturbance.std() or 0.0, seed=key, t0=state.t,
                                         env_offset=env_offset)
        st = LDSState(x=x, t=state.t + 1)
        st.freeze()
        return st, y[0]

    def rollout(self, state, T, actions=None, key=None, env_offset=0, action_std=1.0, keep_actions=False):
        """T fused steps of ``step``: the given ``actions[T,N,n]`` or, when None, the N(0, 1) random policy of
        ``generate_random_trajectory`` (``_lds.py:117-140``).  Returns (state, Y[T,N,p], U[T,N,n] or None)."""
        N = state.x.shape[0]
        key = self.key if key is None else int(key)
        if actions is not None:
            actions = torch.as_tensor(actions, dtype=torch.float32, device=state.x.device).reshape(T, N, self.n).contiguous()
        x, Y, U = fused.lds_open_rollout(*self._abc(N), state.x, T, u_in=actions,
                                         W=self._w(state.t, T, key, N, state.x.device),
                                         w_std=self.disturbance.std() or 0.0, seed=key, t0=state.t,
                                         env_offset=env_offset, action_std=action_std, keep_actions=keep_actions)
        st = LDSState(x=x, t=state.t + T)
        st.freeze()
        return st, Y, U

    # ---- stateful face (the reference's)
    def __call__(self, u, key=None):
        st = LDSState(x=self.x.reshape(1, self.d).contiguous(), t=self.t)
        st, y = self.step(st, torch.as_tensor(u, dtype=torch.float32).reshape(1, self.n), key=key)
        self.x = st.x.reshape(self.d, 1)
        self.t += 1
        return y.reshape(self.p, 1)

    @property
    def action_size(self) -> int:
        return self.n

    def generate_random_trajectory(self, trajectory_length=1000, key=None):
        """``generate_random_trajectory`` (``_lds.py:117-140``): ``trajectory_length`` steps under N(0, 1) actions,
        one fused launch; returns the first coordinate of every observation [T] and advances ``self.x`` / ``self.t``
        like the reference's loop of ``self(rand_action, subkey)`` calls."""
        st = LDSState(x=self.x.reshape(1, self.d).contiguous(), t=self.t)
        st, Y, _ = self.rollout(st, int(trajectory_length), key=0 if key is None else key)
        self.x = st.x.reshape(self.d, 1)
        self.t = st.t
        return Y[:, 0, 0]

    def show_me_the_signal(self, length=1000, key=None):
        """``show_me_the_signal`` (``_lds.py:142-145``) without the plot (matplotlib is not part of this image)."""
        return self.generate_random_trajectory(length, key)
