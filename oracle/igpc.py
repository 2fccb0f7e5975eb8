"""deluca/igpc (rollout, compute_ders, LQR / LQG backward passes, iLQR, GPC_rollout, iGPC_closed, iLC_closed)
and the classic envs it plans on, restated on the CPU (test infrastructure; see oracle/__init__.py).

Parity pinned to the reference's own files run here: ``tests/golden/ref_igpc.npz`` / ``ref_classic.npz`` hold
``rollout``, ``compute_ders``, ``LQR``, ``iLQR``, ``GPC_rollout``, ``iGPC_closed``, ``iLC_closed`` and the env steps
executed literally on tests/refshim (the reference has no test of its own for any of this), reproduced here to rounding
in float64 (tests/test_ref_fixtures_oracle.py).  Analytic Jacobians replace ``jax.jacfwd`` / ``jax.grad`` and are
additionally cross-checked against ``torch.autograd`` on a literal restatement of the env step
(tests/test_oracle_igpc.py).

Batched over a leading trajectory axis N; all trajectories advance in lockstep and the
reference's data-dependent Python control flow (``if cC <= c: ... break``) becomes per-trajectory
masks with the same accept order.

Rulings:
  A-6  Pendulum is reproduced as written: thdot' = th + accel (no dt, old thdot ignored).
  A-7  Cartpole: the stray commas at _cartpole.py:83-88 make the literal code a TypeError; the
       evident intent arr' = A arr + B (u + offset) is implemented (a waiver).
  A-10 iLQR passes ``alpha`` instead of ``alphaC`` to the candidate rollout (ilqr.py:58-67);
       ``fix_backtracking=False`` reproduces that, ``True`` uses alphaC.
  Cost: the reference ships no cost function for this path; c(x,u) = sum q (x-goal)^2 + sum r u^2
       with q = 1, r = 0.1 is this repo's stated choice.
"""
import numpy as np


# ----------------------------------------------------------------------------- envs
class Pendulum:
    """``deluca/envs/classic/_pendulum.py:35-77``."""
    d, m = 3, 1

    def __init__(self, m=1.0, l=1.0, g=9.81, max_torque=1.0, dt=0.02, H=300, goal_state=(0.0, -1.0, 0.0)):
        self.mass, self.l, self.g, self.max_torque, self.dt, self.H = m, l, g, max_torque, dt, H
        self.goal_state = np.asarray(goal_state)

    def init(self, N, dtype):
        x = np.zeros((N, 3), dtype=dtype)
        x[:, 1] = 1.0                                                     # :47
        return x

    def step(self, x, u):
        """x[N,3], u[N,1] -> x'[N,3]   (:54-73, verbatim incl. thdot' = th + accel)."""
        c = x.dtype.type
        th = np.arctan2(x[:, 0], x[:, 1])
        a = c(self.max_torque) * np.tanh(u[:, 0])
        newthdot = th + (c(-3.0 * self.g / (2.0 * self.l)) * np.sin(th + c(np.pi))
                         + c(3.0 / (self.mass * self.l ** 2)) * a)
        newth = th + newthdot * c(self.dt)
        return np.stack([np.sin(newth), np.cos(newth), newthdot], axis=1).astype(x.dtype)

    def jac(self, x, u):
        """Analytic d step / d x [N,3,3] and d step / d u [N,3,1]."""
        c = x.dtype.type
        s, co = x[:, 0], x[:, 1]
        r2 = s * s + co * co
        th = np.arctan2(s, co)
        dth = np.stack([co / r2, -s / r2, np.zeros_like(s)], axis=1)      # d th / d arr
        tu = np.tanh(u[:, 0])
        a = c(self.max_torque) * tu
        k1, k2 = c(-3.0 * self.g / (2.0 * self.l)), c(3.0 / (self.mass * self.l ** 2))
        newthdot = th + (k1 * np.sin(th + c(np.pi)) + k2 * a)
        newth = th + newthdot * c(self.dt)
        dthd_dth = c(1) + k1 * np.cos(th + c(np.pi))
        dnth_dth = c(1) + c(self.dt) * dthd_dth
        sn, cn = np.sin(newth), np.cos(newth)
        Fx = np.stack([(cn * dnth_dth)[:, None] * dth, (-sn * dnth_dth)[:, None] * dth,
                       dthd_dth[:, None] * dth], axis=1)
        dthd_du = k2 * c(self.max_torque) * (c(1) - tu * tu)
        Fu = np.stack([cn * c(self.dt) * dthd_du, -sn * c(self.dt) * dthd_du, dthd_du], axis=1)[:, :, None]
        return Fx.astype(x.dtype), Fu.astype(x.dtype)


class Cartpole:
    """``deluca/envs/classic/_cartpole.py:36-93`` with the A-7 waiver (intent-fixed)."""
    d, m = 4, 1

    def __init__(self, m=0.1, M=1.0, l=1.0, g=9.81, dt=0.02, H=10, goal_state=(0.0, 0.0, 0.0, 0.0), offset=0.0):
        self.mass, self.M, self.l, self.g, self.dt, self.H, self.offset = m, M, l, g, dt, H, offset
        self.goal_state = np.asarray(goal_state)

    def init(self, N, dtype):
        return np.zeros((N, 4), dtype=dtype)

    def AB(self, dtype):
        dt, m, M, l, g = self.dt, self.mass, self.M, self.l, self.g
        A = np.array([[1.0, 0.0, dt, 0.0], [0.0, 1.0, 0.0, dt], [0.0, dt * m * g / M, 1.0, 0.0],
                      [0.0, dt * (m + M) * g / (M * l), 0.0, 1.0]], dtype=dtype)       # :70-82
        B = np.array([[0.0], [0.0], [dt / M], [dt / (M * l)]], dtype=dtype)           # :83-87
        return A, B

    def step(self, x, u):
        A, B = self.AB(x.dtype)
        return (x @ A.T + (u + x.dtype.type(self.offset)) @ B.T).astype(x.dtype)

    def jac(self, x, u):
        A, B = self.AB(x.dtype)
        N = x.shape[0]
        return np.broadcast_to(A, (N, 4, 4)).copy(), np.broadcast_to(B, (N, 4, 1)).copy()


class PlanarQuadrotor:
    """``deluca/envs/classic/_planar_quadrotor.py:51-104``: explicit Euler, wind field
    ``dissipative`` (wind * x, wind * y) or ``constant`` (wind * cos(pi/4), wind * sin(pi/4))."""
    d, m = 6, 2

    def __init__(self, m=0.1, l=0.2, g=9.81, dt=0.05, H=100, wind=0.0, wind_kind="dissipative",
                 goal_state=(0.0,) * 6):
        self.mass, self.l, self.g, self.dt, self.H, self.wind, self.wind_kind = m, l, g, dt, H, wind, wind_kind
        self.goal_state = np.asarray(goal_state)
        self.goal_action = np.array([m * g / 2.0, m * g / 2.0])            # :73

    def init(self, N, dtype):
        x = np.zeros((N, 6), dtype=dtype)
        x[:, 0] = x[:, 1] = 1.0                                            # :87
        return x

    def step(self, x, u):
        c = x.dtype.type
        m, g, l, dt = c(self.mass), c(self.g), c(self.l), c(self.dt)
        if self.wind_kind == "dissipative":
            wx, wy = c(self.wind) * x[:, 0], c(self.wind) * x[:, 1]
        else:
            wx = wy = np.full(x.shape[0], c(self.wind * np.cos(np.pi / 4.0)))
        th = x[:, 2]
        tu = u[:, 0] + u[:, 1]
        xdd = -tu * np.sin(th) / m + wx / m                                # :97
        ydd = tu * np.cos(th) / m - g + wy / m                             # :98
        tdd = l * (u[:, 1] - u[:, 0]) / (m * l ** 2)                       # :99
        dot = np.stack([x[:, 3], x[:, 4], x[:, 5], xdd, ydd, tdd], axis=1)
        return (x + dot * dt).astype(x.dtype)                              # :101

    def jac(self, x, u):
        c = x.dtype.type
        N = x.shape[0]
        m, l, dt = c(self.mass), c(self.l), c(self.dt)
        th, tu = x[:, 2], u[:, 0] + u[:, 1]
        Fx = np.broadcast_to(np.eye(6, dtype=x.dtype), (N, 6, 6)).copy()
        Fx[:, 0, 3] = Fx[:, 1, 4] = Fx[:, 2, 5] = dt
        if self.wind_kind == "dissipative":
            Fx[:, 3, 0] = Fx[:, 4, 1] = c(self.wind) / m * dt
        Fx[:, 3, 2] = -tu * np.cos(th) / m * dt
        Fx[:, 4, 2] = -tu * np.sin(th) / m * dt
        Fu = np.zeros((N, 6, 2), dtype=x.dtype)
        Fu[:, 3, 0] = Fu[:, 3, 1] = -np.sin(th) / m * dt
        Fu[:, 4, 0] = Fu[:, 4, 1] = np.cos(th) / m * dt
        Fu[:, 5, 0], Fu[:, 5, 1] = -l / (m * l ** 2) * dt, l / (m * l ** 2) * dt
        return Fx, Fu


class _Fwd:
    """Forward-mode value + tangent pair over a batch (what ``jax.jacfwd`` propagates): v[N], t[N,P]."""

    def __init__(self, v, t):
        self.v, self.t = v, t

    @staticmethod
    def lift(o, like):
        return o if isinstance(o, _Fwd) else _Fwd(np.full_like(like.v, o), np.zeros_like(like.t))

    def __add__(self, o):
        o = _Fwd.lift(o, self)
        return _Fwd(self.v + o.v, self.t + o.t)

    __radd__ = __add__

    def __neg__(self):
        return _Fwd(-self.v, -self.t)

    def __sub__(self, o):
        return self + (-_Fwd.lift(o, self))

    def __rsub__(self, o):
        return _Fwd.lift(o, self) - self

    def __mul__(self, o):
        o = _Fwd.lift(o, self)
        return _Fwd(self.v * o.v, self.v[:, None] * o.t + o.v[:, None] * self.t)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _Fwd.lift(o, self)
        q = self.v / o.v
        return _Fwd(q, (self.t - q[:, None] * o.t) / o.v[:, None])

    def sin(self):
        return _Fwd(np.sin(self.v), np.cos(self.v)[:, None] * self.t)

    def cos(self):
        return _Fwd(np.cos(self.v), -np.sin(self.v)[:, None] * self.t)


class Reacher:
    """``deluca/envs/classic/_reacher.py:56-141``: two-link arm, explicit Euler on the joint
    accelerations ``inv([[a11,a12],[a12,a22]]) @ b``, tip offset from ``goal_coord`` carried in the state.
    ``max_torque`` is a declared field the step never uses."""
    d, m = 6, 2

    def __init__(self, m1=1.0, m2=1.0, l1=1.0, l2=1.0, g=0.0, dt=0.01, H=200, goal_coord=(0.0, 1.8)):
        self.m1, self.m2, self.l1, self.l2, self.g, self.dt, self.H = m1, m2, l1, l2, g, dt, H
        self.goal_coord = np.asarray(goal_coord, dtype=np.float64)
        self.goal_state = np.zeros(6)

    def init(self, N, dtype):
        th1, th2 = np.pi / 4, np.pi / 2                                     # :75
        arr = np.array([th1, th2, 0.0, 0.0,
                        self.l1 * np.cos(th1) + self.l2 * np.cos(th1 + th2) - self.goal_coord[0],
                        self.l1 * np.sin(th1) + self.l2 * np.sin(th1 + th2) - self.goal_coord[1]])
        return np.broadcast_to(arr.astype(dtype), (N, 6)).copy()

    def _f(self, th1, th2, dth1, dth2, t1, t2, sin, cos, c):
        m1, m2, l1, l2, g, dt = (c(v) for v in (self.m1, self.m2, self.l1, self.l2, self.g, self.dt))
        a11 = ((m1 + m2) * l1 ** 2 + m2 * l2 ** 2) + (c(2) * m2 * l1 * l2) * cos(th2)            # :113
        a12 = m2 * l2 ** 2 + (m2 * l1 * l2) * cos(th2)                                            # :114
        a22 = m2 * l2 ** 2                                                                        # :115
        b1 = ((t1 + (m2 * l1 * l2) * ((c(2) * dth1 + dth2) * dth2 * sin(th2)))
              - (m2 * l2 * g) * sin(th1 + th2)) - ((m1 + m2) * l1 * g) * sin(th1)                 # :116-121
        b2 = (t2 - (m2 * l1 * l2) * (dth1 * dth1 * sin(th2))) - (m2 * l2 * g) * sin(th1 + th2)    # :122-126
        det = a22 * a11 - a12 * a12                                                               # :127-128
        dd1 = (a22 * b1 - a12 * b2) / det
        dd2 = (a11 * b2 - a12 * b1) / det
        nth1, nth2 = th1 + dth1 * dt, th2 + dth2 * dt                                             # :130
        ndth1, ndth2 = dth1 + dd1 * dt, dth2 + dd2 * dt                                           # :131
        Dx = (l1 * cos(nth1) + l2 * cos(nth1 + nth2)) - c(self.goal_coord[0])                     # :132-135
        Dy = (l1 * sin(nth1) + l2 * sin(nth1 + nth2)) - c(self.goal_coord[1])
        return [nth1, nth2, ndth1, ndth2, Dx, Dy]

    def step(self, x, u):
        c = x.dtype.type
        out = self._f(x[:, 0], x[:, 1], x[:, 2], x[:, 3], u[:, 0], u[:, 1], np.sin, np.cos, c)
        return np.stack(out, axis=1).astype(x.dtype)

    def jac(self, x, u):
        """Forward-mode Jacobians (the old tip offset x[4:6] has no influence: zero columns)."""
        c = x.dtype.type
        N = x.shape[0]
        seeds = np.eye(8, dtype=x.dtype)
        var = lambda v, i: _Fwd(v, np.broadcast_to(seeds[i], (N, 8)).copy())
        out = self._f(var(x[:, 0], 0), var(x[:, 1], 1), var(x[:, 2], 2), var(x[:, 3], 3), var(u[:, 0], 6),
                      var(u[:, 1], 7), lambda a: a.sin(), lambda a: a.cos(), c)
        J = np.stack([o.t for o in out], axis=1)                           # [N,6,8]
        return J[:, :, :6].astype(x.dtype), J[:, :, 6:].astype(x.dtype)


class QuadCost:
    """c(x,u) = sum_i q_i (x_i - goal_i)^2 + sum_j r_j (u_j - goal_u_j)^2  (this repo's stated choice)."""

    def __init__(self, goal, q=1.0, r=0.1, goal_action=None):
        self.goal, self.q, self.r = np.asarray(goal, dtype=np.float64), q, r
        self.goal_action = None if goal_action is None else np.asarray(goal_action, dtype=np.float64)

    def _du(self, u):
        return u if self.goal_action is None else u - self.goal_action.astype(u.dtype)

    def __call__(self, x, u):
        dx = x - self.goal.astype(x.dtype)
        du = self._du(u)
        q, r = np.asarray(self.q, dtype=x.dtype), np.asarray(self.r, dtype=x.dtype)     # scalars or per-coordinate
        return (q * (dx * dx)).sum(1) + (r * (du * du)).sum(1)

    def ders(self, x, u):
        N, d = x.shape
        m = u.shape[1]
        q = np.broadcast_to(np.asarray(self.q, dtype=x.dtype), (d,))
        r = np.broadcast_to(np.asarray(self.r, dtype=x.dtype), (m,))
        c_x = x.dtype.type(2) * q * (x - self.goal.astype(x.dtype))
        c_u = x.dtype.type(2) * r * self._du(u)
        c_xx = np.broadcast_to(np.diag(x.dtype.type(2) * q), (N, d, d))
        c_uu = np.broadcast_to(np.diag(x.dtype.type(2) * r), (N, m, m))
        return c_x, c_u, c_xx, c_uu


# ----------------------------------------------------------------------------- rollout / ders / LQR
def rollout(env, cost, U_old, k=None, K=None, X_old=None, alpha=1.0, x0=None):
    """``igpc.rollout`` (``rollout.py:22-65``).  U_old[N,H,m], k[N,H,m], K[N,H,m,d], X_old[N,H+1,d],
    alpha scalar or [N].  Returns X[N,H+1,d], U[N,H,m], cost[N] (terminal state not costed)."""
    N, H, m = U_old.shape
    dtype = U_old.dtype
    alpha = np.broadcast_to(np.asarray(alpha, dtype=dtype), (N,))
    X = np.zeros((N, H + 1, env.d), dtype=dtype)
    U = np.zeros((N, H, m), dtype=dtype)
    X[:, 0] = env.init(N, dtype) if x0 is None else x0
    c = np.zeros(N, dtype=dtype)
    for h in range(H):
        if k is None:
            U[:, h] = U_old[:, h]                                           # :52-53
        else:
            U[:, h] = (U_old[:, h] + alpha[:, None] * k[:, h]) + np.einsum(
                "nij,nj->ni", K[:, h], X[:, h] - X_old[:, h])               # :55-57
        X[:, h + 1] = env.step(X[:, h], U[:, h])                            # :59
        c = c + cost(X[:, h], U[:, h])                                      # :60
    return X, U, c


def compute_ders(env, cost, X, U):
    """``compute_ders`` (``rollout.py:68-99``): D[N,H,d], (F_x[N,H,d,d], F_u[N,H,d,m]),
    (c_x, c_u, c_xx, c_uu) per step."""
    N, H, m = U.shape
    d = env.d
    x0 = X[:, :-1].reshape(N * H, d)
    u = U.reshape(N * H, m)
    D = (X[:, 1:].reshape(N * H, d) - env.step(x0, u)).reshape(N, H, d)
    Fx, Fu = env.jac(x0, u)
    c_x, c_u, c_xx, c_uu = cost.ders(x0, u)
    r = lambda a: a.reshape((N, H) + a.shape[1:])
    return D, (r(Fx), r(Fu)), (r(c_x), r(c_u), r(c_xx), r(c_uu))


def lqr_backward(F, C):
    """``LQR`` (``lqr_solver.py:40-86``): Riccati backward pass.  Returns k[N,H,m], K[N,H,m,d]."""
    Fx, Fu = F
    c_x, c_u, c_xx, c_uu = C
    N, H, d, m = Fu.shape
    dtype = Fx.dtype
    k = np.zeros((N, H, m), dtype=dtype)
    K = np.zeros((N, H, m, d), dtype=dtype)
    V_x = np.zeros((N, d), dtype=dtype)
    V_xx = np.zeros((N, d, d), dtype=dtype)
    T = lambda a: np.swapaxes(a, 1, 2)
    for h in range(H - 1, -1, -1):
        fx, fu = Fx[:, h], Fu[:, h]
        Q_x = c_x[:, h] + np.einsum("nji,nj->ni", fx, V_x)
        Q_u = c_u[:, h] + np.einsum("nji,nj->ni", fu, V_x)
        Q_xx = c_xx[:, h] + T(fx) @ V_xx @ fx
        Q_ux = T(fu) @ V_xx @ fx
        Q_uu = c_uu[:, h] + T(fu) @ V_xx @ fu
        Qi = np.linalg.inv(Q_uu)
        K[:, h] = -Qi @ Q_ux
        k[:, h] = -np.einsum("nij,nj->ni", Qi, Q_u)
        V_x = Q_x - np.einsum("nji,nj->ni", K[:, h], np.einsum("nij,nj->ni", Q_uu, k[:, h]))
        V_xx = Q_xx - T(K[:, h]) @ Q_uu @ K[:, h]
        V_xx = (V_xx + T(V_xx)) / dtype.type(2)
    return k, K


def lambda_adjust(reg_lambda, multiplier, consts, increase):
    """``lambda_adjust`` (``lqr_solver.py:21-37``) on scalars."""
    factor, lambda_min = consts
    if increase:
        multiplier = max(multiplier * factor, factor)
        reg_lambda = max(reg_lambda * multiplier, lambda_min)
    else:
        multiplier = min(1 / factor, multiplier / factor)
        reg_lambda = reg_lambda * multiplier if reg_lambda * multiplier >= lambda_min else 0
    reg_lambda = max(reg_lambda, lambda_min)
    if reg_lambda > 1e10:
        reg_lambda = 1e10
    return reg_lambda, multiplier


def _is_pd_and_invert(Mx):
    """``isPD_and_invert`` (``lqr_solver.py:157-163``): Cholesky test, inverse as L^-T L^-1.  NumPy raises where
    jnp would return NaN; both mean "not positive definite"."""
    try:
        L = np.linalg.cholesky(Mx)
    except np.linalg.LinAlgError:
        return False, None
    if np.isnan(np.sum(L)):
        return False, None
    Li = np.linalg.inv(L)
    return True, Li.T @ Li


def lqg_backward(F, C, reg_lambda, multiplier, damping_consts, reg_type="V"):
    """``LQG`` (``lqr_solver.py:89-154``) per trajectory.  C = (c_x, c_u, c_xx, c_uu, c_ux[N,H,m,d]).
    reg_lambda / multiplier: scalars or [N].  Returns k, K (zero rows where no lambda was found), lam[N], mult[N],
    gains_lin[N], gains_quad[N] (NaN there), ok[N]."""
    Fx, Fu = F
    c_x, c_u, c_xx, c_uu, c_ux = C
    N, H, d, m = Fu.shape
    dtype = Fx.dtype
    lam = np.broadcast_to(np.asarray(reg_lambda, dtype=np.float64), (N,)).copy()
    mult = np.broadcast_to(np.asarray(multiplier, dtype=np.float64), (N,)).copy()
    k = np.zeros((N, H, m), dtype=dtype)
    K = np.zeros((N, H, m, d), dtype=dtype)
    gl = np.full(N, np.nan)
    gq = np.full(N, np.nan)
    ok = np.zeros(N, dtype=bool)
    for n in range(N):
        counter = 0
        while True:
            counter += 1
            if counter >= 10:                                               # :100-103
                k[n], K[n] = 0, 0
                break
            V_x, V_xx = np.zeros(d, dtype=dtype), np.zeros((d, d), dtype=dtype)
            g_lin, g_quad = 0.0, 0.0
            success = True
            rl = dtype.type(lam[n])
            for h in range(H - 1, -1, -1):
                fx, fu = Fx[n, h], Fu[n, h]
                Q_x, Q_u = c_x[n, h] + fx.T @ V_x, c_u[n, h] + fu.T @ V_x
                Q_xx = c_xx[n, h] + fx.T @ V_xx @ fx
                Q_ux = c_ux[n, h] + fu.T @ V_xx @ fx
                Q_uu = c_uu[n, h] + fu.T @ V_xx @ fu
                if reg_type == "Q":                                         # :121-125
                    Q_uu_reg, Q_ux_reg = Q_uu + rl * np.eye(m, dtype=dtype), Q_ux
                else:                                                       # :126-129
                    Q_uu_reg, Q_ux_reg = Q_uu + rl * fu.T @ fu, Q_ux + rl * fu.T @ fx
                pd, Qi = _is_pd_and_invert(Q_uu_reg)
                if not pd:
                    success = False
                    break
                K[n, h], k[n, h] = -Qi @ Q_ux_reg, -Qi @ Q_u
                V_x = Q_x + K[n, h].T @ Q_uu @ k[n, h] + K[n, h].T @ Q_u + Q_ux.T @ k[n, h]      # :134
                V_xx = Q_xx + K[n, h].T @ Q_uu @ K[n, h] + K[n, h].T @ Q_ux + Q_ux.T @ K[n, h]   # :135
                V_xx = (V_xx + V_xx.T) / dtype.type(2)
                g_lin += k[n, h] @ Q_u                                      # :143
                g_quad += k[n, h] @ Q_uu @ k[n, h]                          # :144
            if success:
                ok[n], gl[n], gq[n] = True, g_lin, g_quad
                break
            lam[n], mult[n] = lambda_adjust(lam[n], mult[n], damping_consts, True)
    return k, K, lam, mult, gl, gq, ok


# ----------------------------------------------------------------------------- iLQR
def ilqr(env, cost, U, T, x0=None, alpha=0.5, backtracking=True, fix_backtracking=False, n_alpha=10, trace=None):
    """``iLQR`` (``ilqr.py:27-93``).  Returns X, U, k, K, c, alpha[N].  ``trace``: a list that receives, per iteration,
    dict(decision[N] = index of the accepted candidate or -1, c_before[N], cC[tried,N], c_after[N])."""
    N = U.shape[0]
    dtype = U.dtype
    X, U, c = rollout(env, cost, U, x0=x0)
    alpha = np.full(N, alpha, dtype=dtype)
    k = K = None
    for _ in range(T):
        _, F, C = compute_ders(env, cost, X, U)
        k, K = lqr_backward(F, C)
        if backtracking:
            done = np.zeros(N, dtype=bool)
            dec, cCs, c_before = np.full(N, -1), [], c.copy()
            Xn, Un, cn, an = X.copy(), U.copy(), c.copy(), alpha.copy()
            # as written every candidate rollout uses `alpha` (A-10), so all ten are identical and only
            # j = 0 can ever be accepted; with the fix each candidate uses its own alphaC
            for j in range(n_alpha if fix_backtracking else 1):
                alphaC = (alpha * dtype.type(1.1)) * dtype.type(1.1) ** dtype.type(-(j ** 2))     # :55
                XC, UC, cC = rollout(env, cost, U, k, K, X, alphaC if fix_backtracking else alpha, x0=x0)
                acc = (~done) & (cC <= c)                                   # :73
                Xn[acc], Un[acc], cn[acc], an[acc] = XC[acc], UC[acc], cC[acc], alphaC[acc]
                dec[acc] = j
                cCs.append(cC.copy())
                done |= acc
                if done.all():
                    break
            X, U, c, alpha = Xn, Un, cn, an
            if trace is not None:
                trace.append(dict(decision=dec, c_before=c_before, cC=np.stack(cCs), c_after=c.copy()))
        else:
            X, U, c = rollout(env, cost, U, k, K, X, alpha, x0=x0)          # :78-90
    return X, U, k, K, c, alpha


# ----------------------------------------------------------------------------- iGPC
def igpc_counterfact_grad(env, cost, E, off, W, x, U_old, k, K, X_old, alpha, Cc=1.0, H=3, M=3):
    """grad of ``counterfact_loss`` wrt (E, off) (``igpc.py:22-60``) by a hand adjoint.

    Only the LAST step's cost is returned by the reference (``cost =``, ``:41``), so the loss is
    c(y_{H-1}, u_{H-1}) with y advanced H-1 times by env_sim + W[h+M].
    E[N,H,m,d], off[N,m], W[N,H+M,d], x[N,d]; U_old/k/K/X_old are the H-step windows."""
    N, d = x.shape
    m = off.shape[1]
    dtype = x.dtype
    ys, us, Fxs, Fus = [], [], [], []
    y = x
    for h in range(H):
        ud = np.einsum("nhij,nhj->ni", E, W[:, h:h + M]) + off              # :27-34
        u = ((U_old[:, h] + alpha[:, None] * k[:, h]) + np.einsum("nij,nj->ni", K[:, h], y - X_old[:, h])) \
            + dtype.type(Cc) * ud                                           # :35-40
        ys.append(y)
        us.append(u)
        if h < H - 1:
            Fx, Fu = env.jac(y, u)
            Fxs.append(Fx)
            Fus.append(Fu)
            y = env.step(y, u) + W[:, h + M]                                # :43-44
    c_x, c_u, _, _ = cost.ders(ys[-1], us[-1])
    gE = np.zeros_like(E)
    goff = np.zeros_like(off)
    lam_y, lam_u = c_x, c_u
    for h in range(H - 1, -1, -1):
        # u_h depends on y_h through K[h] and on (E, off) through C * u_delta
        gE += dtype.type(Cc) * np.einsum("ni,nhj->nhij", lam_u, W[:, h:h + M])
        goff += dtype.type(Cc) * lam_u
        lam_y = lam_y + np.einsum("nij,ni->nj", K[:, h], lam_u)
        if h > 0:
            Fx, Fu = Fxs[h - 1], Fus[h - 1]
            lam_u = np.einsum("nij,ni->nj", Fu, lam_y)
            lam_y = np.einsum("nij,ni->nj", Fx, lam_y)
    return gE, goff


def gpc_rollout(env_true, env_sim, cost, U_old, k, K, X_old, D, F, alpha, w_is="de", ref_lr=0.1, x0=None):
    """``GPC_rollout`` (``igpc.py:63-126``).  Returns X, U, cost[N]."""
    H_, M_, lr, Cc = 3, 3, ref_lr, 1.0                                      # :67
    N, Hh, m = U_old.shape
    d = env_sim.d
    dtype = U_old.dtype
    alpha = np.broadcast_to(np.asarray(alpha, dtype=dtype), (N,))
    X = np.zeros((N, Hh + 1, d), dtype=dtype)
    U = np.zeros((N, Hh, m), dtype=dtype)
    W = np.zeros((N, H_ + M_, d), dtype=dtype)
    E = np.zeros((N, H_, m, d), dtype=dtype)
    off = np.zeros((N, m), dtype=dtype)
    X[:, 0] = env_true.init(N, dtype) if x0 is None else x0
    c = np.zeros(N, dtype=dtype)
    for h in range(Hh):
        ud = np.einsum("nhij,nhj->ni", E, W[:, -M_:]) + off                 # :80
        U[:, h] = ((U_old[:, h] + alpha[:, None] * k[:, h]) + np.einsum(
            "nij,nj->ni", K[:, h], X[:, h] - X_old[:, h])) + dtype.type(Cc) * ud          # :81-86
        X[:, h + 1] = env_true.step(X[:, h], U[:, h])                       # :87
        c = c + cost(X[:, h], U[:, h])                                      # :88
        new_state = env_sim.step(X[:, h], U[:, h])                          # :91
        if w_is == "de":
            w = X[:, h + 1] - new_state
        elif w_is == "dede":
            w = (X[:, h + 1] - new_state) - D[:, h]
        else:
            w = X[:, h + 1] - (X_old[:, h + 1] + np.einsum("nij,nj->ni", F[0][:, h], X[:, h] - X_old[:, h])
                               + np.einsum("nij,nj->ni", F[1][:, h], U[:, h] - U_old[:, h]))
        W = np.concatenate([W[:, 1:], w[:, None]], axis=1)                  # :103-104
        if h >= H_:
            s = slice(h - H_, h)                                            # the first H entries of [h-H:]
            gE, goff = igpc_counterfact_grad(env_sim, cost, E, off, W, X[:, h - H_], U_old[:, s], k[:, s],
                                             K[:, s], X_old[:, s], alpha, Cc, H_, M_)
            E = E - dtype.type(lr / h) * gE                                 # :124-125
            off = off - dtype.type(lr / h) * goff
    return X, U, c


def igpc_closed(env_true, env_sim, cost, U, T, w_is="de", lr=0.1, ref_alpha=1.0, backtracking=True,
                x0=None, n_alpha=6, trace=None):
    """``iGPC_closed`` (``igpc.py:129-179``) with ``verbose=False`` (no LR annealing, ``:153-156``)."""
    N = U.shape[0]
    dtype = U.dtype
    alpha = np.full(N, ref_alpha, dtype=dtype)
    X, U, c = rollout(env_true, cost, U, x0=x0)
    k = K = None
    for _ in range(T):
        D, F, C = compute_ders(env_sim, cost, X, U)
        k, K = lqr_backward(F, C)
        if backtracking:
            done = np.zeros(N, dtype=bool)
            dec, cCs, c_before = np.full(N, -1), [], c.copy()
            Xn, Un, cn, an = X.copy(), U.copy(), c.copy(), alpha.copy()
            for j in range(n_alpha):
                alphaC = (alpha * dtype.type(1.1)) * dtype.type(1.1) ** dtype.type(-(j ** 2))     # :159
                XC, UC, cC = gpc_rollout(env_true, env_sim, cost, U, k, K, X, D, F, alphaC, w_is, lr, x0=x0)
                acc = (~done) & (cC < c)                                    # strict, :164
                Xn[acc], Un[acc], cn[acc], an[acc] = XC[acc], UC[acc], cC[acc], alphaC[acc]
                dec[acc] = j
                cCs.append(cC.copy())
                done |= acc
                if done.all():
                    break
            X, U, c, alpha = Xn, Un, cn, an
            if trace is not None:
                trace.append(dict(decision=dec, c_before=c_before, cC=np.stack(cCs), c_after=c.copy()))
        else:
            X, U, c = gpc_rollout(env_true, env_sim, cost, U, k, K, X, D, F, alpha, w_is, lr, x0=x0)
    return X, U, k, K, c, alpha


def ilc_closed(env_true, env_sim, cost, U, T, ref_alpha=1.0, x0=None, n_alpha=10, trace=None):
    """``iLC_closed`` (``ilc.py:23-69``) with backtracking and ``verbose=False``: derivatives from
    env_sim, candidate rollouts on env_true with alphaC, accept the first cC <= c.  (With
    ``verbose=False`` ``prev_loop_fail`` is never cleared but also never acted on, ``:44-47,58``.)"""
    N = U.shape[0]
    dtype = U.dtype
    alpha = np.full(N, ref_alpha, dtype=dtype)
    X, U, c = rollout(env_true, cost, U, x0=x0)
    k = K = None
    for _ in range(T):
        _, F, C = compute_ders(env_sim, cost, X, U)
        k, K = lqr_backward(F, C)
        done = np.zeros(N, dtype=bool)
        dec, cCs, c_before = np.full(N, -1), [], c.copy()
        Xn, Un, cn, an = X.copy(), U.copy(), c.copy(), alpha.copy()
        for j in range(n_alpha):
            alphaC = (alpha * dtype.type(1.1)) * dtype.type(1.1) ** dtype.type(-(j ** 2))
            XC, UC, cC = rollout(env_true, cost, U, k, K, X, alphaC, x0=x0)
            acc = (~done) & (cC <= c)
            Xn[acc], Un[acc], cn[acc], an[acc] = XC[acc], UC[acc], cC[acc], alphaC[acc]
            dec[acc] = j
            cCs.append(cC.copy())
            done |= acc
            if done.all():
                break
        X, U, c, alpha = Xn, Un, cn, an
        if trace is not None:
            trace.append(dict(decision=dec, c_before=c_before, cC=np.stack(cCs), c_after=c.copy()))
    return X, U, k, K, c, alpha
