from ._lds import LDS, LDSState, GaussianDisturbance, SinusDisturbance, ZeroDisturbance  # noqa: F401
from .classic import Cartpole, CartpoleState, Pendulum, PendulumState, PlanarQuadrotor, PlanarQuadrotorState  # noqa: F401
