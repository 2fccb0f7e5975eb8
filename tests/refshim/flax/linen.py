"""The sliver of `flax.linen` that deluca/lung/controllers/_pid.py uses: a compact Module holding one bias-free Dense
(TEST INFRASTRUCTURE).  `init` draws lecun-normal kernels from the shim's RNG (not flax's numbers: fixtures always pass
explicit gains), `apply({"params": ...}, x)` evaluates with the given parameters."""
import math

import jax
import jax.numpy as jnp

_ctx = []


def compact(f):
    return f


class Module:
    name = None

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def init(self, key, *args):
        _ctx.append({"mode": "init", "params": {}, "key": key})
        try:
            self(*args)
            return {"params": _ctx[-1]["params"]}
        finally:
            _ctx.pop()

    def apply(self, variables, *args):
        _ctx.append({"mode": "apply", "params": variables["params"], "key": None})
        try:
            return self(*args)
        finally:
            _ctx.pop()


module = Module


class Dense(Module):
    def __init__(self, features, use_bias=True, name=None):
        self.features, self.use_bias, self.name = features, use_bias, name or "Dense_0"

    def __call__(self, x):
        c = _ctx[-1]
        fan_in = x.shape[-1]
        if c["mode"] == "init":
            c["key"], sub = jax.random.split(c["key"])
            p = {"kernel": jax.random.normal(sub, (fan_in, self.features)) * math.sqrt(1.0 / fan_in)}
            if self.use_bias:
                p["bias"] = jnp.zeros((self.features,))
            c["params"][self.name] = p
        p = c["params"][self.name]
        y = x @ jnp.array(p["kernel"])
        return y + p["bias"] if self.use_bias else y
