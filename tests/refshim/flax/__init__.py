"""flax shim: `flax.struct` (dataclass pytrees) and the sliver of `flax.linen` the lung PID uses (TEST INFRASTRUCTURE)."""
from . import struct, linen  # noqa: F401
