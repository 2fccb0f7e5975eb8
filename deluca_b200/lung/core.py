"""Lung core (``deluca/lung/core.py``): waveform, env base, controller base.

``BreathWaveform.at`` is evaluated inside the kernels from the knot table that
``jnp.interp(t, xp, fp, period=period)`` builds (``core.py:54-60``); ``knots()`` builds that table
on the host in float32 -- which, like default JAX, turns ``3 - 1e-8`` into ``3.0`` and so makes the
exhale a ramp (SURVEY A-9).  ``at()`` / ``is_in()`` on their own evaluate the same table on the host
in float32 (scalar bookkeeping, not a rollout path).
"""
import math
import struct
from typing import Tuple

import torch

from ..core import Agent, Env, Obj, field

DEFAULT_XP = [0.0, 1.0, 1.5, 3.0 - 1e-8, 3.0]   # core.py:26
DEFAULT_DT = 0.03                               # core.py:27


def f32(x):
    """Round a Python float to float32 (the reference's weakly typed constants)."""
    return struct.unpack("f", struct.pack("f", float(x)))[0]


class BreathWaveform(Obj):
    r"""Waveform generator with shape |‾\_ (``core.py:33-69``).  ``pip`` / ``peep`` may be tensors
    [N] for a batch of breaths with different targets."""

    peep: float = field(5.0, jaxed=False)
    pip: float = field(35.0, jaxed=False)
    bpm: int = field(20, jaxed=False)
    fp: object = field(None, jaxed=False)
    xp: object = field(None, jaxed=False)
    in_bounds: Tuple[int, int] = field((0, 1), jaxed=False)
    ex_bounds: Tuple[int, int] = field((2, 4), jaxed=False)
    period: float = field(None, jaxed=False)
    dt: float = field(DEFAULT_DT, jaxed=False)

    def setup(self):
        if self.xp is None:
            self.xp = list(DEFAULT_XP)
        self.period = 60 / self.bpm

    # ---- knot table (host, float32)
    def knots(self, N=None, device="cuda"):
        """(kx list[nk], kf tensor [N,nk] or [nk], period, xp_in): the table of jnp.interp with
        ``period``: xp mod period (in float32), stable sort, one wrapped knot each side."""
        period = f32(self.period)
        xs = [f32(math.fmod(f32(x), period)) for x in self.xp]
        order = sorted(range(len(xs)), key=lambda i: xs[i])        # stable
        batched = any(isinstance(v, torch.Tensor) and v.dim() > 0 for v in (self.pip, self.peep)) or (
            isinstance(self.fp, torch.Tensor) and self.fp.dim() == 2)
        as_col = lambda v: torch.as_tensor(v, dtype=torch.float32).reshape(-1)
        if self.fp is None:
            pip, peep = as_col(self.pip), as_col(self.peep)
            n = max(pip.numel(), peep.numel())
            pip, peep = pip.expand(n), peep.expand(n)
            fp = torch.stack([pip, pip, peep, peep, pip], dim=1)   # core.py:48-49
        else:
            fp = torch.as_tensor(self.fp, dtype=torch.float32).reshape(-1, len(self.xp))
        fs = fp[:, order]
        kx = [f32(xs[order[-1]] - period)] + [xs[i] for i in order] + [f32(xs[order[0]] + period)]
        kf = torch.cat([fs[:, -1:], fs, fs[:, :1]], dim=1)
        if N is not None and kf.shape[0] not in (1, N):
            raise ValueError(f"waveform describes {kf.shape[0]} breaths, expected {N}")
        kf = kf.contiguous() if (batched or kf.shape[0] > 1) else kf[0].contiguous()
        return kx, kf.to(device), period, f32(self.xp[self.in_bounds[1]])

    def at(self, t):
        kx, kf, period, _ = self.knots(device="cpu")
        kf = kf.reshape(-1, len(kx))
        x = f32(math.fmod(f32(t), period))
        i = min(max(sum(1 for k in kx if k <= x), 1), len(kx) - 1)
        dx = f32(kx[i] - kx[i - 1])
        f0, f1 = kf[:, i - 1], kf[:, i]
        val = f0 if dx == 0 else f0 + (f32(f32(x - kx[i - 1]) / dx)) * (f1 - f0)
        return val if val.numel() > 1 else float(val)

    def elapsed(self, t):
        return f32(math.fmod(f32(t), f32(self.period)))

    def is_in(self, t):
        return self.elapsed(t) <= f32(self.xp[self.in_bounds[1]])

    def is_ex(self, t):
        return not self.is_in(t)


class LungEnv(Env):
    """``LungEnv`` (``core.py:72-93``)."""

    def time(self, state):
        return self.dt * state.steps

    @property
    def dt(self):
        return DEFAULT_DT

    @property
    def physical(self):
        return False

    def should_abort(self):
        return False

    def wait(self, duration):
        pass

    def cleanup(self):
        pass


class LungEnvState(Obj):
    t_in: int = 0
    steps: int = 0
    predicted_pressure: float = 0.0


class ControllerState(Obj):
    """``ControllerState`` (``core.py:103-106``); leaves may be tensors [N]."""
    time: float = float("inf")
    steps: int = 0
    dt: float = DEFAULT_DT


def proper_time(t):
    """``proper_time`` (``core.py:109-110``)."""
    if isinstance(t, torch.Tensor):
        return torch.where(torch.isinf(t), torch.zeros_like(t), t)
    return 0.0 if t == float("inf") else t


class Controller(Agent):
    """``Controller`` (``core.py:113-151``)."""

    def init(self, waveform=None):
        return ControllerState()

    @property
    def max(self):
        return 100

    @property
    def min(self):
        return 0

    def decay(self, waveform, t):
        elapsed = waveform.elapsed(t)
        if elapsed < f32(waveform.xp[waveform.in_bounds[1]]):
            return float("inf")
        if elapsed < f32(waveform.xp[waveform.ex_bounds[0]]):
            return 0.0
        return f32(5 * (1 - math.exp(5 * (f32(waveform.xp[waveform.ex_bounds[0]]) - elapsed))))
