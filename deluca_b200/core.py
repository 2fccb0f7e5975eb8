"""Object / pytree contract of the reference (``deluca/core.py:48-152``) without flax or JAX.

``Obj`` subclasses are dataclasses whose fields are split into pytree LEAVES
(``field(jaxed=True)``: arrays that ``jax.vmap`` would batch -- here torch CUDA tensors with a
leading env axis N) and STATIC fields (``jaxed=False``: compile-time constants that end up in the
kernels' parameter structs).  ``create()`` runs ``setup()`` then freezes the instance;
``replace()`` returns a frozen copy; mutation of a frozen object raises
``dataclasses.FrozenInstanceError`` exactly like the reference (``core.py:66-85``).
"""
import dataclasses
import os
import pickle
from abc import abstractmethod


def save(obj, path):
    """``deluca.save`` (``core.py:35-41``)."""
    dirname = os.path.abspath(os.path.dirname(path))
    os.makedirs(dirname, exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def load(path):
    """``deluca.load`` (``core.py:44-45``)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def field(default=None, jaxed=True, **kwargs):
    """``deluca.field`` (``core.py:48-52``): ``jaxed`` marks a pytree leaf."""
    meta = dict(kwargs.pop("metadata", {}) or {})
    meta["pytree_node"] = jaxed
    if "default_factory" not in kwargs:
        kwargs["default"] = default
    return dataclasses.field(metadata=meta, **kwargs)


class Obj:
    """``deluca.Obj`` (``core.py:55-110``)."""

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        dataclasses.dataclass(eq=False)(cls)

    def __new__(cls, *args, **kwargs):
        obj = object.__new__(cls)
        object.__setattr__(obj, "__frozen__", False)
        return obj

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise dataclasses.FrozenInstanceError(f"cannot assign to field {name!r}")
        object.__setattr__(self, name, value)

    def freeze(self):
        object.__setattr__(self, "__frozen__", True)

    def unfreeze(self):
        object.__setattr__(self, "__frozen__", False)

    def is_frozen(self):
        return getattr(self, "__frozen__", True)

    def replace(self, **updates):
        obj = dataclasses.replace(self, **updates)
        obj.freeze()
        return obj

    @classmethod
    def create(cls, *args, **kwargs):
        obj = cls(*args, **kwargs)
        obj.setup()
        obj.freeze()
        return obj

    def setup(self):
        """Used in place of ``__init__`` (``core.py:105-106``)."""

    # -- pytree view: leaves in field order, nested Obj flattened depth-first
    def flatten(self):
        leaves = []
        for f in dataclasses.fields(self):
            if not f.metadata.get("pytree_node", True):
                continue
            v = getattr(self, f.name)
            leaves.extend(v.flatten() if isinstance(v, Obj) else [v])
        return leaves

    def unflatten(self, leaves):
        """Rebuild an object like ``self`` from ``leaves`` (inverse of ``flatten``)."""
        leaves = list(leaves)

        def build(proto):
            upd = {}
            for f in dataclasses.fields(proto):
                if not f.metadata.get("pytree_node", True):
                    continue
                v = getattr(proto, f.name)
                upd[f.name] = build(v) if isinstance(v, Obj) else leaves.pop(0)
            return proto.replace(**upd)

        return build(self)


class Env(Obj):
    """``deluca.Env`` (``core.py:113-126``)."""

    @abstractmethod
    def init(self, *args, **kwargs):
        """Return an initial (state, observation)."""

    @abstractmethod
    def __call__(self, state, action, *args, **kwargs):
        """Return (next state, observation)."""

    @property
    def action_size(self) -> int:
        raise NotImplementedError


class AgentState(Obj):
    """``deluca.AgentState`` (``core.py:129-131``)."""
    time: float = float("inf")
    steps: int = 0


class Agent(Obj):
    """``deluca.Agent`` (``core.py:134-141``)."""

    @abstractmethod
    def __call__(self, state, obs, *args, **kwargs):
        """Return (next agent state, action)."""

    def init(self):
        return AgentState()


class Disturbance(Obj):
    """``deluca.Disturbance`` (``core.py:144-152``)."""

    @abstractmethod
    def init(self, *args, **kwargs):
        """Initialise the disturbance."""

    @abstractmethod
    def __call__(self, *args, **kwargs):
        """Return the next disturbance."""
