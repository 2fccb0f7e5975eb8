"""Import stub (TEST INFRASTRUCTURE): deluca/training/apg.py and lung/utils/scripts/train_controller.py import optax
for their train loops (out of scope).  Attribute access succeeds -- train_controller.py:104 names ``optax.adamw`` as a
default argument at import -- but calling anything raises."""


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    def _missing(*a, **k):
        raise RuntimeError(f"optax.{name}: optax is not installed; the refshim stub only satisfies the import")

    _missing.__name__ = name
    return _missing
