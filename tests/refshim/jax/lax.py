"""`jax.lax` subset (TEST INFRASTRUCTURE): scan, cond, dynamic_slice(_in_dim), clamp -- eager Python."""
import torch

from . import _core
from ._core import Array, asarray, tree_flatten, tree_map, tree_unflatten, _raw


def _to_int(i):
    return int(i._t) if isinstance(i, Array) else int(i)


def _start(i, size, dim):
    """JAX `dynamic_slice` index rule: a negative start is wrapped once (start + dim), then the start is clamped
    into [0, dim - size] so that the slice always lies inside the operand."""
    i = _to_int(i)
    if i < 0:
        i += dim
    return max(0, min(i, dim - size))


def dynamic_slice(operand, start_indices, slice_sizes):
    t = asarray(operand)._t
    idx = tuple(slice(_start(s, n, d), _start(s, n, d) + n) for s, n, d in zip(start_indices, slice_sizes, t.shape))
    return Array(t[idx])


def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    t = asarray(operand)._t
    s = _start(start_index, slice_size, t.shape[axis])
    return Array(t.narrow(axis, s, slice_size))


def dynamic_update_slice(operand, update, start_indices):
    t = asarray(operand)._t.clone()
    u = asarray(update)._t
    idx = tuple(slice(_start(s, n, d), _start(s, n, d) + n) for s, n, d in zip(start_indices, u.shape, t.shape))
    t[idx] = u
    return Array(t)


def clamp(lo, x, hi):
    t = asarray(x)._t
    return Array(torch.clamp(t, _raw(lo), _raw(hi)))


def cond(pred, true_fun, false_fun, *operands):
    """Only the selected branch runs (same value as JAX for pure branches).  A numeric predicate means pred != 0."""
    p = bool(pred._t != 0) if isinstance(pred, Array) else bool(pred)
    return true_fun(*operands) if p else false_fun(*operands)


def _stack_leaves(items):
    leaves0, td = tree_flatten(items[0])
    cols = [[] for _ in leaves0]
    for it in items:
        ls, _ = tree_flatten(it)
        for c, l in zip(cols, ls):
            c.append(asarray(l)._t)
    return tree_unflatten(td, [Array(torch.stack(c)) for c in cols])


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
        get = lambda i: None
    else:
        leaves, td = tree_flatten(xs)
        leaves = [asarray(l) for l in leaves]
        n = leaves[0].shape[0] if length is None else length
        get = lambda i: tree_unflatten(td, [l[i] for l in leaves])
    carry = _core.arrayify(init)          # tracing turns the carry's Python scalars into arrays (float32 by default)
    ys = []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        carry, y = f(carry, get(i))
        carry = _core.arrayify(carry)
        ys.append(y)
    if reverse:
        ys.reverse()
    if not ys or all(tree_flatten(y)[0] == [] for y in ys):
        return carry, None
    return carry, _stack_leaves(ys)


def fori_loop(lo, hi, body, init):
    v = init
    for i in range(_to_int(lo), _to_int(hi)):
        v = body(i, v)
    return v


def stop_gradient(x):
    return _core.stop_gradient(x)
