"""Import stub (TEST INFRASTRUCTURE): deluca/training/apg.py imports optax for its train loop (out of scope)."""


def __getattr__(name):
    raise RuntimeError("optax is not installed; the refshim stub only satisfies the import")
