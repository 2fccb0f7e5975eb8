"""Makes the reference's hot-path source files importable here (TEST INFRASTRUCTURE, see README.md in this directory).

    from tests.refshim import loader; loader.install()          # then: import deluca.agents._gpc, ...

* puts the JAX/flax shim and the import stubs (matplotlib, clu, optax) first on sys.path;
* registers the reference's packages as NAMESPACE-style modules whose `__path__` points into /root/reference, so that
  `import deluca.agents._gpc` executes the reference's own file but NOT the packages' `__init__.py` (several of those
  fail at import: `deluca/agents/__init__` pulls agents that need the missing `deluca.agents.core`,
  `deluca/lung/envs/__init__` reaches `lung/utils/nn/lstm.py:75`, a SyntaxError -- SURVEY F3);
* supplies the one module the reference itself lacks: `deluca.agents.core` with a plain `Agent` base class
  (`_lqr.py:19`, `_bpc.py:23` import it; nothing in the tree defines it).
Nothing under /root/reference is modified or copied.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DELUCA_REFERENCE", "/root/reference")

_PACKAGES = [
    "deluca", "deluca.agents", "deluca.envs", "deluca.envs.classic", "deluca.igpc", "deluca.lung",
    "deluca.lung.envs", "deluca.lung.controllers", "deluca.lung.utils", "deluca.lung.utils.scripts",
    "deluca.lung.utils.data", "deluca.training", "deluca.learners", "deluca.utils",
]
_installed = [False]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "deluca"))


def install(x64=False):
    if _installed[0]:
        import jax
        jax.config.update("jax_enable_x64", bool(x64))
        return
    if not available():
        raise RuntimeError(f"the reference tree is not present at {REF_ROOT} (it does not travel to the GPU box)")
    for p in (os.path.join(HERE, "stubs"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in list(sys.modules):
        if name == "jax" or name.startswith(("jax.", "flax", "deluca")):
            raise RuntimeError(f"{name} is already imported; install the shim first")
    import jax
    jax.config.update("jax_enable_x64", bool(x64))
    for name in _PACKAGES:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
        m.__package__ = name
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    core = importlib.import_module("deluca.core")
    top = sys.modules["deluca"]
    for n in ("Agent", "AgentState", "Disturbance", "Env", "Obj", "field", "load", "save"):   # deluca/__init__.py:17-39
        setattr(top, n, getattr(core, n))
    agents_core = types.ModuleType("deluca.agents.core")
    agents_core.__doc__ = "supplied by tests/refshim: the module is missing from the reference tree (SURVEY F3)"

    class Agent:
        """Plain base class for the stateful agents (`LQR`, `BPC`), which define `__init__` / `__call__` themselves."""

    agents_core.Agent = Agent
    sys.modules["deluca.agents.core"] = agents_core
    sys.modules["deluca.agents"].core = agents_core
    _installed[0] = True
