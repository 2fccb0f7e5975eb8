"""A JAX-semantics shim backed by torch CPU tensors -- TEST INFRASTRUCTURE ONLY (tests/refshim/README.md).

JAX is not installable in this image.  This package restates, eagerly, the small part of the JAX API that
the reference's hot-path source files use, so that THOSE FILES can be executed here, unmodified, to generate golden
fixtures (tests/golden/make_ref_fixtures.py).  It is never imported by the product (`deluca_b200/`).
"""
import functools as _ft

from . import _core
from ._core import Array, config, grad, value_and_grad, stop_gradient      # noqa: F401
from . import numpy, lax, random, tree_util, scipy                          # noqa: F401
from . import tree_util as tree                                             # jax.tree.map

__version__ = "0.0.0+refshim"


def jit(fun=None, static_argnums=None, static_argnames=None, **_):
    """Eager, but with the one effect of tracing that changes VALUES: the non-static arguments' Python-scalar leaves
    become arrays of the default dtype (`_core.arrayify`).  Works as `@jax.jit`, `jax.jit(f, static_argnums=...)` and
    `functools.partial(jax.jit, static_argnums=...)`."""
    if fun is None:
        return lambda f: jit(f, static_argnums=static_argnums, static_argnames=static_argnames)
    static = () if static_argnums is None else ((static_argnums,) if isinstance(static_argnums, int) else tuple(static_argnums))
    names = () if static_argnames is None else ((static_argnames,) if isinstance(static_argnames, str) else tuple(static_argnames))

    @_ft.wraps(fun)
    def wrapped(*args, **kwargs):
        args = [a if i in static else _core.arrayify(a) for i, a in enumerate(args)]
        kwargs = {k: (v if k in names else _core.arrayify(v)) for k, v in kwargs.items()}
        return fun(*args, **kwargs)
    return wrapped


def jacfwd(f, argnums=0, has_aux=False):
    return _core.jacobian(f, argnums)


def jacrev(f, argnums=0, has_aux=False):
    return _core.jacobian(f, argnums)


def hessian(f, argnums=0):
    return jacfwd(jacrev(f, argnums), argnums)


def vmap(f, in_axes=0, out_axes=0):
    """Loop over the mapped axis and stack the results leaf by leaf."""
    import torch
    from ._core import asarray, tree_flatten, tree_unflatten

    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        flat = []
        for a, ax in zip(args, axes):
            if ax is None:
                flat.append(None)
                continue
            leaves, td = tree_flatten(a)
            leaves = [asarray(l) for l in leaves]
            n = leaves[0].shape[ax] if n is None else n
            flat.append((leaves, td, ax))
        outs = []
        for i in range(n):
            call = []
            for a, fl in zip(args, flat):
                if fl is None:
                    call.append(a)
                else:
                    leaves, td, ax = fl
                    call.append(tree_unflatten(td, [Array(l._t.select(ax, i)) for l in leaves]))
            outs.append(f(*call))
        leaves0, td0 = tree_flatten(outs[0])
        cols = [[] for _ in leaves0]
        for o in outs:
            for c, l in zip(cols, tree_flatten(o)[0]):
                c.append(asarray(l)._t)
        return tree_unflatten(td0, [Array(torch.stack(c, dim=out_axes)) for c in cols])
    return wrapped


def process_index():
    return 0


def device_get(x):
    return x


class _Debug:
    @staticmethod
    def callback(fn, *a, **k):
        fn(*a, **k)

    @staticmethod
    def print(fmt, *a, **k):
        print(fmt.format(*a, **k))


debug = _Debug()
