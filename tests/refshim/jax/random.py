"""`jax.random` subset (TEST INFRASTRUCTURE).  Keys are NumPy seed sequences, NOT Threefry: the numbers differ from
JAX's (SURVEY a3 -- Threefry streams cannot be reproduced without JAX).  The fixture generator therefore records every
random input (systems, disturbances, perturbations) next to the outputs; parity is defined on those inputs."""
import numpy as np

from . import _core
from ._core import Array, asarray


class Key:
    def __init__(self, entropy):
        self.entropy = tuple(int(e) for e in entropy)

    def _rng(self):
        return np.random.default_rng(np.random.SeedSequence(self.entropy))

    def __repr__(self):
        return f"Key{self.entropy}"


def PRNGKey(seed):
    return Key((int(seed),))


key = PRNGKey


def split(k, num=2):
    return [Key(k.entropy + (i,)) for i in range(num)]


def _np_dtype():
    return np.float64 if _core.x64_enabled() else np.float32


def normal(k, shape=(), dtype=None):
    return asarray(k._rng().standard_normal(shape).astype(_np_dtype()))


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return asarray(k._rng().uniform(minval, maxval, shape).astype(_np_dtype()))


def permutation(k, x):
    if isinstance(x, int):
        return asarray(k._rng().permutation(x))
    return asarray(k._rng().permutation(np.asarray(x)))
