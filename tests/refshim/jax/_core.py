"""Array type, pytrees and autodiff of the JAX-semantics shim (see tests/refshim/README.md).

TEST INFRASTRUCTURE.  A minimal, eager restatement of the parts of the JAX API the reference's hot-path files use,
backed by torch CPU tensors (so that `jax.grad` / `jacfwd` / `jacrev` work through `torch.autograd`).  It exists so that
the reference's OWN source files can be executed in this container -- where JAX cannot be installed -- to produce
golden fixtures.  Semantics restated from JAX's documented behaviour:

* default dtype float32 (x64 disabled); `REFSHIM_X64=1` / `config.update("jax_enable_x64", True)` switches to float64;
  float64 requests are silently truncated to float32 when x64 is disabled, as JAX does;
* Python scalars are weakly typed (they never promote an array);
* functional updates `x.at[i].set(v)` / `.add(v)`; arrays are immutable;
* a scalar index that is itself an array (= traced under jit/scan) clamps into bounds, as JAX's gather does; a static
  Python index out of bounds raises, as in JAX;
* `jit` is the identity; `vmap` loops over the mapped axis and stacks; `lax.scan` is a Python loop; `lax.cond` evaluates
  the selected branch only (same value as JAX for pure branches).
"""
import dataclasses
import os

import numpy as np
import torch

_X64 = [os.environ.get("REFSHIM_X64", "0") == "1"]


def x64_enabled():
    return _X64[0]


def float_dtype():
    return torch.float64 if _X64[0] else torch.float32


def int_dtype():
    return torch.int64 if _X64[0] else torch.int32


class _Config:
    def update(self, name, value):
        if name == "jax_enable_x64":
            _X64[0] = bool(value)
            torch.set_default_dtype(float_dtype())
        # every other option is irrelevant to an eager interpreter

    @property
    def jax_enable_x64(self):
        return _X64[0]


config = _Config()
torch.set_default_dtype(float_dtype())


# ------------------------------------------------------------------------------------------ dtypes
class _DType:
    def __init__(self, name, t):
        self.name, self.t = name, t

    def __call__(self, x):
        return asarray(x, dtype=self)

    def __repr__(self):
        return f"refshim.{self.name}"

    def __eq__(self, o):
        return isinstance(o, _DType) and canon(o) == canon(self)

    def __hash__(self):
        return hash(self.name)


float32 = _DType("float32", torch.float32)
float64 = _DType("float64", torch.float64)
int32 = _DType("int32", torch.int32)
int64 = _DType("int64", torch.int64)
bool_ = _DType("bool", torch.bool)


def canon(dt):
    """Requested dtype -> torch dtype under the x64 flag (float64 -> float32, int64 -> int32 when disabled)."""
    if dt is None:
        return None
    if isinstance(dt, _DType):
        t = dt.t
    elif isinstance(dt, torch.dtype):
        t = dt
    elif dt is float:
        t = torch.float64
    elif dt is int:
        t = torch.int64
    elif dt is bool:
        t = torch.bool
    else:
        t = {"float32": torch.float32, "float64": torch.float64, "int32": torch.int32, "int64": torch.int64,
             "bool": torch.bool}[np.dtype(dt).name]
    if not _X64[0]:
        t = {torch.float64: torch.float32, torch.int64: torch.int32}.get(t, t)
    return t


# ------------------------------------------------------------------------------------------ Array
def _raw(x):
    """Operand -> torch tensor or Python scalar (weak type)."""
    if isinstance(x, Array):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, int, float)):
        return x
    if isinstance(x, np.generic):
        return x.item()            # numpy scalars behave as weakly typed Python scalars here
    if isinstance(x, np.ndarray):
        return asarray(x)._t
    if isinstance(x, (list, tuple)):
        return asarray(x)._t
    raise TypeError(f"refshim: unsupported operand {type(x)}")


def _idx(i):
    if isinstance(i, Array):
        t = i._t
        return int(t) if t.dim() == 0 and t.dtype != torch.bool else (t.long() if t.dtype != torch.bool else t)
    if isinstance(i, np.generic):
        return i.item()
    if isinstance(i, np.ndarray):
        return torch.as_tensor(i)
    if isinstance(i, tuple):
        return tuple(_idx(j) for j in i)
    return i


class _AtIndexer:
    def __init__(self, arr, idx=None):
        self.arr, self.idx = arr, idx

    def __getitem__(self, idx):
        return _AtIndexer(self.arr, idx)

    def _val(self, v, like):
        v = _raw(v)
        if isinstance(v, torch.Tensor):
            return v.to(like.dtype)
        return v

    def set(self, v):
        t = self.arr._t.clone()
        tgt = t[_idx(self.idx)]
        v = self._val(v, t)
        if isinstance(v, torch.Tensor) and v.dim() > tgt.dim():       # jax broadcasts / squeezes trailing 1s
            v = v.reshape(tgt.shape)
        t[_idx(self.idx)] = v
        return Array(t)

    def add(self, v):
        t = self.arr._t.clone()
        t[_idx(self.idx)] = t[_idx(self.idx)] + self._val(v, t)
        return Array(t)


def _binop(fn, reflect=False):
    def op(self, other):
        if isinstance(other, (list, tuple)):
            # JAX operators reject Python sequences ("... requires ndarray or scalar arguments, got <class 'tuple'>");
            # this is what makes Cartpole.__call__ a TypeError as written (_cartpole.py:83-88, SURVEY A-7)
            raise TypeError(f"unsupported operand type(s): 'Array' and {type(other).__name__!r} "
                            "(JAX operators require ndarray or scalar arguments)")
        try:
            a, b = _raw(self), _raw(other)
        except TypeError:
            return NotImplemented
        if isinstance(b, torch.Tensor) and isinstance(a, torch.Tensor) and a.dtype != b.dtype:
            dt = torch.promote_types(a.dtype, b.dtype)
            if not _X64[0]:
                dt = {torch.float64: torch.float32, torch.int64: torch.int32}.get(dt, dt)
            a, b = a.to(dt), b.to(dt)
        r = fn(b, a) if reflect else fn(a, b)
        return Array(r) if isinstance(r, torch.Tensor) else r
    return op


def _true_div(a, b):
    if isinstance(a, torch.Tensor) and not a.is_floating_point() and not isinstance(b, torch.Tensor):
        a = a.to(float_dtype())
    r = torch.true_divide(torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a, b)
    if r.dtype == torch.float64 and not _X64[0]:
        r = r.float()
    return r


def _pow(a, b):
    if not isinstance(a, torch.Tensor):
        # python scalar ** array: weak scalar takes the array's float type (or the default float for int arrays)
        bt = b if b.is_floating_point() else b.to(float_dtype())
        return torch.pow(torch.tensor(float(a), dtype=bt.dtype), bt)
    if not isinstance(b, torch.Tensor) and not a.is_floating_point() and isinstance(b, float):
        a = a.to(float_dtype())
    return torch.pow(a, b)


class Array:
    """Immutable n-d array (a torch CPU tensor underneath).  `jnp.ndarray` and `jax.Array` are this class."""
    __array_priority__ = 1000
    __slots__ = ("_t",)

    def __init__(self, t):
        if isinstance(t, Array):
            t = t._t
        elif not isinstance(t, torch.Tensor):
            t = asarray(t)._t
        object.__setattr__(self, "_t", t)

    # -- metadata
    shape = property(lambda s: tuple(s._t.shape))
    ndim = property(lambda s: s._t.dim())
    size = property(lambda s: s._t.numel())
    T = property(lambda s: Array(s._t.T if s._t.dim() == 2 else s._t.permute(*reversed(range(s._t.dim())))))
    at = property(lambda s: _AtIndexer(s))

    @property
    def dtype(self):
        return {torch.float32: float32, torch.float64: float64, torch.int32: int32, torch.int64: int64,
                torch.bool: bool_}[self._t.dtype]

    def __len__(self):
        return self._t.shape[0]

    def __iter__(self):
        if self._t.dim() == 0:
            raise TypeError("iteration over a 0-d array")
        return (Array(self._t[i]) for i in range(self._t.shape[0]))

    def __getitem__(self, i):
        if isinstance(i, Array) and i._t.dim() == 0 and i._t.dtype != torch.bool:
            # a TRACED scalar index (here: any index that is itself an array): JAX lowers it to a gather /
            # dynamic_slice whose out-of-bounds index is clamped into the array ("the index is clamped to the
            # bounds of the array since something must be returned", JAX sharp bits); negative wraps once first
            n = self._t.shape[0]
            k = int(i._t)
            if k < 0:
                k += n
            return Array(self._t[min(max(k, 0), n - 1)])
        return Array(self._t[_idx(i)])

    def __setitem__(self, i, v):
        raise TypeError("refshim arrays are immutable, like JAX arrays: use x.at[i].set(v)")

    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self._t)

    def __int__(self):
        return int(self._t)

    def __bool__(self):
        return bool(self._t)

    def __index__(self):
        return int(self._t)

    def __repr__(self):
        return f"Array({self._t.detach().numpy()!r}, dtype={self.dtype.name})"

    def __hash__(self):
        raise TypeError("unhashable type: Array")

    # -- arithmetic
    __add__ = _binop(torch.add); __radd__ = _binop(torch.add, True)
    __sub__ = _binop(torch.sub); __rsub__ = _binop(torch.sub, True)
    __mul__ = _binop(torch.mul); __rmul__ = _binop(torch.mul, True)
    __truediv__ = _binop(_true_div); __rtruediv__ = _binop(_true_div, True)
    __pow__ = _binop(_pow); __rpow__ = _binop(_pow, True)
    __mod__ = _binop(torch.remainder); __rmod__ = _binop(torch.remainder, True)
    __matmul__ = _binop(torch.matmul); __rmatmul__ = _binop(torch.matmul, True)
    __lt__ = _binop(torch.lt); __le__ = _binop(torch.le); __gt__ = _binop(torch.gt); __ge__ = _binop(torch.ge)
    __eq__ = _binop(torch.eq); __ne__ = _binop(torch.ne)
    __and__ = _binop(torch.logical_and); __or__ = _binop(torch.logical_or)

    def __neg__(self):
        return Array(-self._t)

    def __pos__(self):
        return self

    def __abs__(self):
        return Array(self._t.abs())

    def __invert__(self):
        return Array(~self._t)

    # -- methods the reference uses
    def astype(self, dt):
        return Array(self._t.to(canon(dt)))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self._t.reshape(*shape))

    def flatten(self):
        return Array(self._t.reshape(-1))

    def squeeze(self, axis=None):
        return Array(self._t.squeeze() if axis is None else self._t.squeeze(axis))

    def sum(self, axis=None):
        return Array(self._t.sum() if axis is None else self._t.sum(axis))

    def mean(self, axis=None):
        return Array(self._t.mean() if axis is None else self._t.mean(axis))

    def max(self, axis=None):
        return Array(self._t.max() if axis is None else self._t.amax(axis))

    def min(self, axis=None):
        return Array(self._t.min() if axis is None else self._t.amin(axis))

    def dot(self, other):
        return Array(torch.matmul(self._t, _raw(other)))

    def transpose(self, *axes):
        return self.T if not axes else Array(self._t.permute(*axes))

    def item(self):
        return self._t.item()

    def tolist(self):
        return self._t.tolist()

    def copy(self):
        return Array(self._t.clone())

    def block_until_ready(self):
        return self


def asarray(x, dtype=None):
    """`jnp.array` / `jnp.asarray`: nested lists may contain Arrays (stacked); numpy float64 -> default float dtype."""
    want = canon(dtype)
    if isinstance(x, Array):
        t = x._t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x)).reshape(x.shape)   # (ascontiguousarray makes 0-d arrays 1-d)
        if t.dtype == torch.float64 and not _X64[0]:
            t = t.float()
        if t.dtype == torch.int64 and not _X64[0]:
            t = t.int()
    elif isinstance(x, (list, tuple)):
        if len(x) == 0:
            t = torch.zeros(0, dtype=float_dtype())
        else:
            parts = [asarray(e)._t for e in x]
            dt = parts[0].dtype
            for p in parts[1:]:
                dt = torch.promote_types(dt, p.dtype)
            t = torch.stack([p.to(dt) for p in parts])
    elif isinstance(x, bool):
        t = torch.tensor(x)
    elif isinstance(x, (int, np.integer)):
        t = torch.tensor(int(x), dtype=int_dtype())
    elif isinstance(x, (float, np.floating)):
        t = torch.tensor(float(x), dtype=float_dtype())
    else:
        raise TypeError(f"refshim: cannot convert {type(x)} to an array")
    if want is not None and t.dtype != want:
        t = t.to(want)
    return Array(t)


# ------------------------------------------------------------------------------------------ pytrees
class _Leaf:
    pass


LEAF = _Leaf()
_registry = {}   # class -> (flatten(obj) -> (children, aux), unflatten(aux, children))


def register_pytree_node(cls, flatten, unflatten):
    _registry[cls] = (flatten, unflatten)


def _is_leaf(x):
    return x is None or not (isinstance(x, (list, tuple, dict)) or type(x) in _registry)


def tree_flatten(tree):
    leaves = []

    def go(x):
        if x is None:
            return ("none",)
        if type(x) in _registry:
            children, aux = _registry[type(x)][0](x)
            return ("node", type(x), aux, [go(c) for c in children])
        if isinstance(x, tuple) and hasattr(x, "_fields"):
            return ("namedtuple", type(x), [go(c) for c in x])
        if isinstance(x, (list, tuple)):
            return ("list" if isinstance(x, list) else "tuple", [go(c) for c in x])
        if isinstance(x, dict):
            ks = sorted(x)
            return ("dict", ks, [go(x[k]) for k in ks])
        leaves.append(x)
        return ("leaf",)

    return leaves, go(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def go(d):
        k = d[0]
        if k == "none":
            return None
        if k == "leaf":
            return next(it)
        if k == "node":
            return _registry[d[1]][1](d[2], [go(c) for c in d[3]])
        if k == "namedtuple":
            return d[1](*[go(c) for c in d[2]])
        if k == "list":
            return [go(c) for c in d[1]]
        if k == "tuple":
            return tuple(go(c) for c in d[1])
        if k == "dict":
            return {kk: go(c) for kk, c in zip(d[1], d[2])}
        raise ValueError(k)

    return go(treedef)


def arrayify(tree):
    """What tracing does to the arguments of a jitted function and to a scan carry: Python / NumPy scalar leaves
    become (weakly typed) arrays of the default dtype, so that e.g. `state.time + self.dt` accumulates in float32
    inside `lax.scan` instead of in Python's float64.  Other leaves (arrays, functions, strings) pass through."""
    def conv(l):
        if isinstance(l, (bool, int, float, np.generic)):
            return asarray(l)
        if isinstance(l, np.ndarray):
            return asarray(l)
        return l
    leaves, td = tree_flatten(tree)
    return tree_unflatten(td, [conv(l) for l in leaves])


def tree_map(f, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


# ------------------------------------------------------------------------------------------ autodiff
def _as_diff_leaf(x):
    """Leaf -> (Array usable as a differentiation variable).  Integer leaves are rejected as in JAX."""
    a = x if isinstance(x, Array) else asarray(x)
    if not a._t.is_floating_point():
        raise TypeError("grad / jacobian requires real-valued inputs (input dtype that is a sub-dtype of np.inexact), "
                        f"but got {a.dtype.name}")
    t = a._t
    if not t.requires_grad:
        t = t.detach().clone().requires_grad_(True)
    return Array(t)


def _prepare(args, argnums):
    """Returns (args with differentiation variables in place, argnums, [(leaves, treedef)], nested) where `nested` says
    that some input already carries a graph (an outer grad / jacobian is in progress, e.g. jacfwd(jacrev(f)))."""
    nums = (argnums,) if isinstance(argnums, int) else tuple(argnums)
    args = list(args)
    nested = any(isinstance(l, Array) and l._t.requires_grad for a in args for l in tree_flatten(a)[0])
    flat = []
    for n in nums:
        leaves, td = tree_flatten(args[n])
        leaves = [_as_diff_leaf(l) for l in leaves]
        args[n] = tree_unflatten(td, leaves)
        flat.append((leaves, td))
    return args, nums, flat, nested


def _scalar(out):
    t = out._t if isinstance(out, Array) else torch.as_tensor(out)
    if t.numel() != 1:
        raise TypeError(f"Gradient only defined for scalar-output functions. Output had shape: {tuple(t.shape)}.")
    return t.reshape(())


def _grads_of(scalar_t, leaves, nested):
    if not scalar_t.requires_grad:
        return [torch.zeros_like(l._t) for l in leaves]
    gs = torch.autograd.grad(scalar_t, [l._t for l in leaves], create_graph=nested, retain_graph=True,
                             allow_unused=True)
    return [torch.zeros_like(l._t) if g is None else (g if nested else g.detach()) for g, l in zip(gs, leaves)]


def _detach_tree(tree):
    return tree_map(lambda l: Array(l._t.detach()) if isinstance(l, Array) else l, tree)


def value_and_grad(f, argnums=0, has_aux=False):
    def wrapped(*args, **kw):
        a, nums, flat, nested = _prepare(args, argnums)
        out = f(*a, **kw)
        val, aux = (out if has_aux else (out, None))
        s = _scalar(val)
        res = []
        for leaves, td in flat:
            res.append(tree_unflatten(td, [Array(g) for g in _grads_of(s, leaves, nested)]))
        g = res[0] if isinstance(argnums, int) else tuple(res)
        val = Array(s if nested else s.detach()) if isinstance(val, Array) else val
        if not nested:
            aux = _detach_tree(aux)
        return ((val, aux), g) if has_aux else (val, g)
    return wrapped


def grad(f, argnums=0, has_aux=False):
    vg = value_and_grad(f, argnums, has_aux)

    def wrapped(*args, **kw):
        v, g = vg(*args, **kw)
        return (g, v[1]) if has_aux else g
    return wrapped


def jacobian(f, argnums=0):
    """Both `jacfwd` and `jacrev` (the values agree; only JAX's evaluation strategy differs).  Output pytree:
    the structure of f's output, each leaf replaced by the structure of the differentiated argument(s), leaves of shape
    out_leaf.shape + in_leaf.shape."""
    def wrapped(*args, **kw):
        a, nums, flat, nested = _prepare(args, argnums)
        out = f(*a, **kw)
        out_leaves, out_td = tree_flatten(out)
        res_leaves = []
        for o in out_leaves:
            ot = (o._t if isinstance(o, Array) else asarray(o)._t)
            if not ot.is_floating_point():
                raise TypeError("jacobian requires real-valued outputs, got an integer leaf (e.g. a step counter `h: int`)")
            of = ot.reshape(-1)
            per_arg = []
            for leaves, td in flat:
                rows = [_grads_of(of[j], leaves, nested) for j in range(of.numel())]
                jl = []
                for li, l in enumerate(leaves):
                    J = torch.stack([rows[j][li] for j in range(of.numel())]).reshape(tuple(ot.shape) + tuple(l._t.shape))
                    jl.append(Array(J))
                per_arg.append(tree_unflatten(td, jl))
            res_leaves.append(per_arg[0] if isinstance(argnums, int) else tuple(per_arg))
        return tree_unflatten(out_td, res_leaves)
    return wrapped


def stop_gradient(x):
    return tree_map(lambda l: Array(l._t.detach()) if isinstance(l, Array) else l, x)
