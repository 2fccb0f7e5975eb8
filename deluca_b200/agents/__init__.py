from ._bpc import BPC  # noqa: F401
from ._gpc import GPC  # noqa: F401
from ._lqr import LQR  # noqa: F401
from ._pid import PID  # noqa: F401
