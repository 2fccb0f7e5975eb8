"""`flax.struct.dataclass` / `field`: a frozen dataclass registered as a pytree; fields with `pytree_node=False` are
static aux data (TEST INFRASTRUCTURE).  Restated from flax/struct.py."""
import dataclasses

from jax._core import register_pytree_node


def field(pytree_node=True, **kwargs):
    meta = dict(kwargs.pop("metadata", None) or {})
    meta["pytree_node"] = pytree_node
    return dataclasses.field(metadata=meta, **kwargs)


def dataclass(clz=None, **kwargs):
    if clz is None:
        return lambda c: dataclass(c, **kwargs)
    if "_flax_dataclass" in clz.__dict__:
        return clz
    data_clz = dataclasses.dataclass(frozen=True, **kwargs)(clz)
    data_fields, meta_fields = [], []
    for f in dataclasses.fields(data_clz):
        (data_fields if f.metadata.get("pytree_node", True) else meta_fields).append(f.name)

    def replace(self, **updates):
        return dataclasses.replace(self, **updates)

    data_clz.replace = replace

    def flatten(x):
        return [getattr(x, n) for n in data_fields], tuple(getattr(x, n) for n in meta_fields)

    def unflatten(aux, children):
        kw = dict(zip(meta_fields, aux))
        kw.update(zip(data_fields, children))
        # build without running __init__ / setup side effects, exactly the stored fields
        obj = object.__new__(data_clz)
        for k, v in kw.items():
            object.__setattr__(obj, k, v)
        return obj

    register_pytree_node(data_clz, flatten, unflatten)
    data_clz._flax_dataclass = True
    return data_clz
