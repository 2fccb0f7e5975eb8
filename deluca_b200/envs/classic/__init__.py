from ._cartpole import Cartpole, CartpoleState  # noqa: F401
from ._pendulum import Pendulum, PendulumState, QuadCost  # noqa: F401
from ._planar_quadrotor import PlanarQuadrotor, PlanarQuadrotorState, constant, dissipative  # noqa: F401
from ._reacher import Reacher, ReacherState  # noqa: F401
