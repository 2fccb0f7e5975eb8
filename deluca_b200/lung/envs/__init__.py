from ._balloon_lung import BalloonLung  # noqa: F401
from ._delay_lung import DelayLung  # noqa: F401
