from ._expiratory import Expiratory  # noqa: F401
from ._pid import PID  # noqa: F401
