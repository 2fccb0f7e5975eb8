from .core import Learner  # noqa: F401
