"""Import stub (TEST INFRASTRUCTURE): deluca/training/apg.py imports clu.metric_writers for its train loop (out of scope)."""
from . import metric_writers  # noqa: F401
