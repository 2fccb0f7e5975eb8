"""`jax.numpy` subset used by the reference's hot-path files (TEST INFRASTRUCTURE, see tests/refshim/README.md)."""
import math

import numpy as _np
import torch

from . import _core
from ._core import Array, asarray, bool_, float32, float64, int32, int64, _raw

ndarray = Array
pi = math.pi
inf = math.inf
newaxis = None
float_ = float64
dtype = lambda x: x   # jnp.dtype is only used as an annotation


def _w(t):
    return Array(t)


def _t(x):
    r = _raw(x)
    return r if isinstance(r, torch.Tensor) else asarray(r)._t


def array(x, dtype=None):
    return asarray(x, dtype)


def zeros(shape, dtype=None):
    return _w(torch.zeros(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=_core.canon(dtype) or _core.float_dtype()))


def ones(shape, dtype=None):
    return _w(torch.ones(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=_core.canon(dtype) or _core.float_dtype()))


def zeros_like(x):
    return _w(torch.zeros_like(_t(x)))


def identity(n, dtype=None):
    return _w(torch.eye(n, dtype=_core.canon(dtype) or _core.float_dtype()))


eye = identity


def arange(*a, dtype=None):
    if any(isinstance(v, float) for v in a):
        return _w(torch.arange(*a, dtype=_core.canon(dtype) or _core.float_dtype()))
    return _w(torch.arange(*a, dtype=_core.canon(dtype) or _core.int_dtype()))


def diag(x):
    return _w(torch.diag(_t(x)))


def _un(fn):
    def f(x):
        t = _t(x)
        if not t.is_floating_point():
            t = t.to(_core.float_dtype())
        return _w(fn(t))
    return f


sin, cos, tanh, exp, sqrt, arctan, log = (_un(f) for f in (torch.sin, torch.cos, torch.tanh, torch.exp, torch.sqrt,
                                                             torch.atan, torch.log))
abs = lambda x: _w(torch.abs(_t(x)))
square = lambda x: _w(_t(x) * _t(x))


def arctan2(a, b):
    a, b = _t(a), _t(b)
    return _w(torch.atan2(a.to(_core.float_dtype()) if not a.is_floating_point() else a,
                          b.to(a.dtype) if b.is_floating_point() else b.to(_core.float_dtype())))


def _red(fn_all, fn_axis):
    def f(x, axis=None, keepdims=False):
        t = _t(x)
        if axis is None:
            r = fn_all(t)
            return _w(r.reshape([1] * t.dim()) if keepdims else r)
        return _w(fn_axis(t, axis, keepdims))
    return f


sum = _red(torch.sum, lambda t, a, k: torch.sum(t, a, keepdim=k))
mean = _red(torch.mean, lambda t, a, k: torch.mean(t, a, keepdim=k))
max = _red(torch.max, lambda t, a, k: torch.amax(t, a, keepdim=k))
min = _red(torch.min, lambda t, a, k: torch.amin(t, a, keepdim=k))


def _pair(a, b):
    ra, rb = _raw(a), _raw(b)
    ta = ra if isinstance(ra, torch.Tensor) else None
    tb = rb if isinstance(rb, torch.Tensor) else None
    if ta is None and tb is None:
        return asarray(ra)._t, asarray(rb)._t
    if ta is None:
        return torch.as_tensor(ra, dtype=tb.dtype if (tb.is_floating_point() or isinstance(ra, int)) else _core.float_dtype()), tb
    if tb is None:
        return ta, torch.as_tensor(rb, dtype=ta.dtype if (ta.is_floating_point() or isinstance(rb, int)) else _core.float_dtype())
    if ta.dtype != tb.dtype:
        dt = torch.promote_types(ta.dtype, tb.dtype)
        ta, tb = ta.to(dt), tb.to(dt)
    return ta, tb


def maximum(a, b):
    return _w(torch.maximum(*_pair(a, b)))


def minimum(a, b):
    return _w(torch.minimum(*_pair(a, b)))


def clip(x, a_min=None, a_max=None):
    t = _t(x)
    if t.dtype == torch.bool:
        t = t.to(_core.float_dtype() if isinstance(a_min, float) or isinstance(a_max, float) else _core.int_dtype())
    if not t.is_floating_point() and (isinstance(a_min, float) or isinstance(a_max, float)):
        t = t.to(_core.float_dtype())
    lo = _raw(a_min) if a_min is not None else None
    hi = _raw(a_max) if a_max is not None else None
    return _w(torch.clamp(t, lo, hi))


def where(c, a, b):
    ta, tb = _pair(a, b)
    return _w(torch.where(_t(c).bool(), ta, tb))


def roll(x, shift, axis=None):
    t = _t(x)
    if axis is None:
        return _w(torch.roll(t.reshape(-1), int(shift)).reshape(t.shape))
    return _w(torch.roll(t, int(shift), dims=axis))


def flip(x, axis=None):
    t = _t(x)
    return _w(torch.flip(t, dims=list(range(t.dim())) if axis is None else [axis]))


def squeeze(x, axis=None):
    return asarray(x).squeeze(axis)


def reshape(x, shape):
    return asarray(x).reshape(shape)


def concatenate(xs, axis=0):
    return _w(torch.cat([_t(x) for x in xs], dim=axis))


def stack(xs, axis=0):
    return _w(torch.stack([_t(x) for x in xs], dim=axis))


def tensordot(a, b, axes=2):
    ta, tb = _pair(a, b)
    if isinstance(axes, int):
        return _w(torch.tensordot(ta, tb, dims=axes))
    return _w(torch.tensordot(ta, tb, dims=(list(axes[0]), list(axes[1]))))


def einsum(spec, *ops):
    # the reference writes size-1 axes as the subscript "1" ("mnp,mp1->mn1", _grc.py:97); torch accepts letters only,
    # so digits are renamed to letters the expression does not use
    free = [c for c in "ZYXWVUTSRQ" if c not in spec]
    for dgt in sorted({c for c in spec if c.isdigit()}):
        spec = spec.replace(dgt, free.pop())
    ts = [_t(o) for o in ops]
    dt = ts[0].dtype
    for t in ts[1:]:
        dt = torch.promote_types(dt, t.dtype)
    return _w(torch.einsum(spec, *[t.to(dt) for t in ts]))


def dot(a, b):
    return _w(torch.matmul(*_pair(a, b)))


matmul = dot


def outer(a, b):
    return _w(torch.outer(*_pair(a, b)))


def isnan(x):
    return _w(torch.isnan(_t(x)))


def argsort(x):
    return _w(torch.argsort(_t(x), stable=True).to(_core.int_dtype()))


def searchsorted(a, v, side="left"):
    return _w(torch.searchsorted(_t(a).contiguous(), _t(v).contiguous(), right=(side == "right")).to(_core.int_dtype()))


def interp(x, xp, fp, left=None, right=None, period=None):
    """`jnp.interp` (jax/_src/numpy/lax_numpy.py::_interp), restated: with `period`, x and xp are reduced modulo the
    period, (xp, fp) are sorted by a STABLE argsort of xp and extended by one wrapped knot on each side; the segment
    index is clip(searchsorted(xp, x, side='right'), 1, len(xp) - 1) and coincident knots (|dx| <= eps) return
    fp[i - 1]."""
    xt, xpt, fpt = _t(x), _t(xp), _t(fp)
    if not xt.is_floating_point():
        xt = xt.to(_core.float_dtype())
    dt = torch.promote_types(torch.promote_types(xt.dtype, xpt.dtype), fpt.dtype)
    xt, xpt, fpt = xt.to(dt), xpt.to(dt), fpt.to(dt)
    if period is not None:
        if period == 0:
            raise ValueError("period must be a non-zero value")
        period = builtins_abs(period)
        xt = torch.remainder(xt, period)
        xpt = torch.remainder(xpt, period)
        order = torch.argsort(xpt, stable=True)
        xpt, fpt = xpt[order], fpt[order]
        xpt = torch.cat([xpt[-1:] - period, xpt, xpt[:1] + period])
        fpt = torch.cat([fpt[-1:], fpt, fpt[:1]])
    shape = xt.shape
    xf = xt.reshape(-1)
    i = torch.clamp(torch.searchsorted(xpt.contiguous(), xf.contiguous(), right=True), 1, xpt.numel() - 1)
    df = fpt[i] - fpt[i - 1]
    dx = xpt[i] - xpt[i - 1]
    delta = xf - xpt[i - 1]
    eps = float(_np.spacing(_np.finfo(_np.float32 if dt == torch.float32 else _np.float64).eps))
    dx0 = dx.abs() <= eps
    f = torch.where(dx0, fpt[i - 1], fpt[i - 1] + (delta / torch.where(dx0, torch.ones_like(dx), dx)) * df)
    if period is None:
        f = torch.where(xf < xpt[0], fpt[0] if left is None else torch.as_tensor(left, dtype=dt), f)
        f = torch.where(xf > xpt[-1], fpt[-1] if right is None else torch.as_tensor(right, dtype=dt), f)
    return _w(f.reshape(shape))


import builtins as _b
builtins_abs = _b.abs


class _Linalg:
    @staticmethod
    def inv(x):
        return _w(torch.linalg.inv(_t(x)))

    @staticmethod
    def norm(x, ord=None, axis=None):
        t = _t(x)
        if ord is None and axis is None:
            return _w(torch.sqrt((t * t).sum()))
        return _w(torch.linalg.norm(t, ord=ord, dim=axis))

    @staticmethod
    def cholesky(x):
        return _w(torch.linalg.cholesky(_t(x)))

    @staticmethod
    def eigh(x):
        w, v = torch.linalg.eigh(_t(x))
        return _w(w), _w(v)


linalg = _Linalg()
