def __getattr__(name):
    raise RuntimeError("matplotlib is not installed; the refshim stub only satisfies the import")
