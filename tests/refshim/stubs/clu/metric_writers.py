def __getattr__(name):
    raise RuntimeError("clu is not installed; the refshim stub only satisfies the import")
