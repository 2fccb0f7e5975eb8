"""Oracle self-checks for the igpc path (CPU only): analytic Jacobians and hand adjoints against
torch.autograd on literal restatements of the reference expressions; Riccati pass against the
infinite-horizon fixed point; iLQR monotonicity."""
import numpy as np
import pytest
import torch

from oracle import igpc as G


def _pend_step_torch(x, u, env):
    """Literal deluca/envs/classic/_pendulum.py:64-73 in torch."""
    sin, cos = x[0], x[1]
    a = env.max_torque * torch.tanh(u[0])
    th = torch.atan2(sin, cos)
    newthdot = th + (-3.0 * env.g / (2.0 * env.l) * torch.sin(th + np.pi) + 3.0 / (env.mass * env.l ** 2) * a)
    newth = th + newthdot * env.dt
    return torch.stack([torch.sin(newth), torch.cos(newth), newthdot])


def test_pendulum_jacobians_match_autograd():
    env = G.Pendulum()
    rng = np.random.default_rng(0)
    th = rng.uniform(-3, 3, 16)
    x = np.stack([np.sin(th), np.cos(th), rng.standard_normal(16)], 1)
    u = rng.standard_normal((16, 1))
    Fx, Fu = env.jac(x, u)
    for n in range(16):
        J = torch.autograd.functional.jacobian(lambda a, b: _pend_step_torch(a, b, env),
                                               (torch.tensor(x[n]), torch.tensor(u[n])))
        np.testing.assert_allclose(Fx[n], J[0].numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Fu[n], J[1].numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(env.step(x[n:n + 1], u[n:n + 1])[0],
                                   _pend_step_torch(torch.tensor(x[n]), torch.tensor(u[n]), env).numpy(), rtol=1e-12)


def test_cartpole_intent_fixed():
    env = G.Cartpole()
    x = np.array([[0.1, 0.2, 0.3, 0.4]])
    u = np.array([[0.5]])
    A, B = env.AB(np.float64)
    np.testing.assert_allclose(env.step(x, u)[0], A @ x[0] + B[:, 0] * 0.5)
    assert A[2, 1] == pytest.approx(0.02 * 0.1 * 9.81) and B[3, 0] == pytest.approx(0.02)


def test_riccati_backward_converges_to_dare_gain():
    import scipy.linalg
    env = G.Cartpole()
    A, B = env.AB(np.float64)
    H, N = 400, 2
    Fx = np.broadcast_to(A, (N, H, 4, 4)).copy()
    Fu = np.broadcast_to(B, (N, H, 4, 1)).copy()
    Q, R = 2 * np.eye(4), 0.2 * np.eye(1)
    C = (np.zeros((N, H, 4)), np.zeros((N, H, 1)), np.broadcast_to(Q, (N, H, 4, 4)).copy(),
         np.broadcast_to(R, (N, H, 1, 1)).copy())
    k, K = G.lqr_backward((Fx, Fu), C)
    X = scipy.linalg.solve_discrete_are(A, B, Q, R)
    Kinf = np.linalg.solve(R + B.T @ X @ B, B.T @ X @ A)
    np.testing.assert_allclose(-K[0, 0], Kinf, rtol=1e-6)
    assert np.abs(k).max() == 0


def _counterfact_loss_torch(E, off, W, x, env, cost, U_old, k, K, X_old, alpha, Cc=1.0, H=3, M=3):
    """Literal deluca/igpc/igpc.py:22-57 for one trajectory (pendulum env_sim)."""
    y = x
    c = 0
    for h in range(H):
        ud = torch.tensordot(E, W[h:h + M], dims=([0, 2], [0, 1])) + off
        u = U_old[h] + alpha * k[h] + K[h] @ (y - X_old[h]) + Cc * ud
        c = ((y - torch.tensor(cost.goal)) ** 2).sum() * cost.q + cost.r * (u ** 2).sum()
        y = _pend_step_torch(y, u, env) + W[h + M]
    return c


def test_counterfactual_adjoint_matches_autograd():
    env = G.Pendulum()
    cost = G.QuadCost(env.goal_state, 1.0, 0.1)
    rng = np.random.default_rng(1)
    N, H, M, d, m = 3, 3, 3, 3, 1
    E = 0.3 * rng.standard_normal((N, H, m, d))
    off = 0.1 * rng.standard_normal((N, m))
    W = 0.1 * rng.standard_normal((N, H + M, d))
    th = rng.uniform(-3, 3, N)
    x = np.stack([np.sin(th), np.cos(th), rng.standard_normal(N)], 1)
    U_old, k = rng.standard_normal((N, H, m)), rng.standard_normal((N, H, m))
    K, X_old = 0.2 * rng.standard_normal((N, H, m, d)), rng.standard_normal((N, H, d))
    alpha = np.full(N, 0.7)
    gE, goff = G.igpc_counterfact_grad(env, cost, E, off, W, x, U_old, k, K, X_old, alpha)
    for n in range(N):
        Et, ot = torch.tensor(E[n], requires_grad=True), torch.tensor(off[n], requires_grad=True)
        L = _counterfact_loss_torch(Et, ot, torch.tensor(W[n]), torch.tensor(x[n]), env, cost,
                                    torch.tensor(U_old[n]), torch.tensor(k[n]), torch.tensor(K[n]),
                                    torch.tensor(X_old[n]), 0.7)
        L.backward()
        np.testing.assert_allclose(gE[n], Et.grad.numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(goff[n], ot.grad.numpy(), rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("fix", [False, True])
def test_ilqr_cost_is_monotone_and_improves(fix):
    env = G.Pendulum(H=40)
    cost = G.QuadCost(env.goal_state, 1.0, 0.1)
    rng = np.random.default_rng(2)
    N = 6
    x0 = env.init(N, np.float64) + 0.05 * rng.standard_normal((N, 3))
    U0 = np.zeros((N, env.H, 1))
    _, _, c0 = G.rollout(env, cost, U0, x0=x0)
    cs = [c0]
    for T in (1, 3, 6):
        cs.append(G.ilqr(env, cost, U0, T, x0=x0, fix_backtracking=fix)[4])
    for a, b in zip(cs, cs[1:]):
        assert (b <= a + 1e-12).all()
    if fix:
        assert (cs[-1] < c0).any()


def test_igpc_and_ilc_run_and_never_increase_cost():
    env_true, env_sim = G.Pendulum(H=30, m=1.1), G.Pendulum(H=30)
    cost = G.QuadCost(env_sim.goal_state, 1.0, 0.1)
    rng = np.random.default_rng(3)
    N = 4
    x0 = env_true.init(N, np.float64) + 0.05 * rng.standard_normal((N, 3))
    U0 = np.zeros((N, 30, 1))
    _, _, c0 = G.rollout(env_true, cost, U0, x0=x0)
    for w_is in ("de", "dede", "delin"):
        c = G.igpc_closed(env_true, env_sim, cost, U0, 3, w_is=w_is, lr=0.01, x0=x0)[4]
        assert np.isfinite(c).all() and (c <= c0 + 1e-12).all()
    c = G.ilc_closed(env_true, env_sim, cost, U0, 3, x0=x0)[4]
    assert (c <= c0 + 1e-12).all()


def test_planar_quadrotor_jacobians_match_autograd():
    """deluca/envs/classic/_planar_quadrotor.py:92-104 restated literally in torch."""
    for kind in ("dissipative", "constant"):
        env = G.PlanarQuadrotor(wind=0.3, wind_kind=kind)
        rng = np.random.default_rng(5)
        x = rng.standard_normal((6, 6))
        u = 0.5 + 0.2 * rng.standard_normal((6, 2))

        def step_t(xt, ut):
            xx, yy, th, xd, yd, thd = xt
            wind = [env.wind * xx, env.wind * yy] if kind == "dissipative" else [
                env.wind * np.cos(np.pi / 4.0) + 0 * xx, env.wind * np.sin(np.pi / 4.0) + 0 * xx]
            xdd = -(ut[0] + ut[1]) * torch.sin(th) / env.mass + wind[0] / env.mass
            ydd = (ut[0] + ut[1]) * torch.cos(th) / env.mass - env.g + wind[1] / env.mass
            tdd = env.l * (ut[1] - ut[0]) / (env.mass * env.l ** 2)
            return xt + torch.stack([xd, yd, thd, xdd, ydd, tdd]) * env.dt

        Fx, Fu = env.jac(x, u)
        for n in range(6):
            J = torch.autograd.functional.jacobian(step_t, (torch.tensor(x[n]), torch.tensor(u[n])))
            np.testing.assert_allclose(Fx[n], J[0].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(Fu[n], J[1].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(env.step(x[n:n + 1], u[n:n + 1])[0],
                                       step_t(torch.tensor(x[n]), torch.tensor(u[n])).numpy(), rtol=1e-12)


def test_ilqr_stabilises_the_quadrotor():
    env = G.PlanarQuadrotor(H=60)
    cost = G.QuadCost(env.goal_state, 1.0, 0.1, goal_action=env.goal_action)
    N = 4
    x0 = env.init(N, np.float64) + 0.05 * np.random.default_rng(0).standard_normal((N, 6))
    U0 = np.tile(env.goal_action, (N, env.H, 1))
    _, _, c0 = G.rollout(env, cost, U0, x0=x0)
    c = G.ilqr(env, cost, U0, 8, x0=x0, fix_backtracking=True)[4]
    assert (c < 0.5 * c0).all()


def _reacher_step_torch(xt, ut, env):
    """Literal deluca/envs/classic/_reacher.py:109-141 in torch."""
    m1, m2, l1, l2, g = env.m1, env.m2, env.l1, env.l2, env.g
    th1, th2, dth1, dth2, Dx, Dy = xt
    t1, t2 = ut
    a11 = (m1 + m2) * l1 ** 2 + m2 * l2 ** 2 + 2 * m2 * l1 * l2 * torch.cos(th2)
    a12 = m2 * l2 ** 2 + m2 * l1 * l2 * torch.cos(th2)
    a22 = torch.as_tensor(m2 * l2 ** 2, dtype=xt.dtype)
    b1 = (t1 + m2 * l1 * l2 * (2 * dth1 + dth2) * dth2 * torch.sin(th2) - m2 * l2 * g * torch.sin(th1 + th2)
          - (m1 + m2) * l1 * g * torch.sin(th1))
    b2 = t2 - m2 * l1 * l2 * dth1 ** 2 * torch.sin(th2) - m2 * l2 * g * torch.sin(th1 + th2)
    A, b = torch.stack([torch.stack([a11, a12]), torch.stack([a12, a22])]), torch.stack([b1, b2])
    ddth1, ddth2 = torch.linalg.inv(A) @ b
    th1, th2 = th1 + dth1 * env.dt, th2 + dth2 * env.dt
    dth1, dth2 = dth1 + ddth1 * env.dt, dth2 + ddth2 * env.dt
    Dx = l1 * torch.cos(th1) + l2 * torch.cos(th1 + th2) - env.goal_coord[0]
    Dy = l1 * torch.sin(th1) + l2 * torch.sin(th1 + th2) - env.goal_coord[1]
    return torch.stack([th1, th2, dth1, dth2, Dx, Dy])


def test_reacher_step_and_jacobians_match_autograd():
    for env in (G.Reacher(), G.Reacher(m1=0.8, m2=0.7, l1=1.2, l2=0.8, g=9.81, dt=0.02, goal_coord=(0.3, 1.1))):
        rng = np.random.default_rng(7)
        x, u = rng.standard_normal((8, 6)), rng.standard_normal((8, 2))
        Fx, Fu = env.jac(x, u)
        assert np.abs(Fx[:, :, 4:]).max() == 0           # the carried tip offset never feeds back
        for n in range(8):
            f = lambda a, b: _reacher_step_torch(a, b, env)
            J = torch.autograd.functional.jacobian(f, (torch.tensor(x[n]), torch.tensor(u[n])))
            np.testing.assert_allclose(Fx[n], J[0].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(Fu[n], J[1].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(env.step(x[n:n + 1], u[n:n + 1])[0], f(torch.tensor(x[n]), torch.tensor(u[n])).numpy(),
                                       rtol=1e-12, atol=1e-14)
    # init() places the tip at l1 (cos, sin)(pi/4) + l2 (cos, sin)(3 pi/4) minus the goal (:75-91)
    x0 = G.Reacher().init(1, np.float64)[0]
    np.testing.assert_allclose(x0, [np.pi / 4, np.pi / 2, 0, 0, 0.0, np.sqrt(2.0) - 1.8], atol=1e-12)


def test_ilqr_moves_the_reacher_tip_to_the_goal():
    env = G.Reacher(H=60, dt=0.05)
    cost = G.QuadCost(env.goal_state, q=[0, 0, 0.1, 0.1, 10.0, 10.0], r=0.01)
    N = 3
    x0 = env.init(N, np.float64)
    x0[:, :2] += 0.05 * np.random.default_rng(0).standard_normal((N, 2))
    x0 = env.step(np.concatenate([x0[:, :2] , x0[:, 2:]], 1), np.zeros((N, 2)))     # make the tip offset consistent
    U0 = np.zeros((N, env.H, 2))
    _, _, c0 = G.rollout(env, cost, U0, x0=x0)
    c = G.ilqr(env, cost, U0, 10, x0=x0, fix_backtracking=True)[4]
    assert (c < 0.7 * c0).all()


def _lq_problem(rng, N, H, d, m, cuu_scale=1.0):
    Fx = np.eye(d) + 0.1 * rng.standard_normal((N, H, d, d))
    Fu = 0.3 * rng.standard_normal((N, H, d, m))
    S = rng.standard_normal((N, H, d, d))
    c_xx = np.einsum("nhij,nhkj->nhik", S, S) / d + 0.1 * np.eye(d)
    R = rng.standard_normal((N, H, m, m))
    c_uu = cuu_scale * (np.einsum("nhij,nhkj->nhik", R, R) / m + 0.2 * np.eye(m))
    return (Fx, Fu), (rng.standard_normal((N, H, d)), rng.standard_normal((N, H, m)), c_xx, c_uu)


def test_lqg_reduces_to_lqr_without_regularisation():
    rng = np.random.default_rng(11)
    F, C = _lq_problem(rng, 3, 12, 4, 2)
    k0, K0 = G.lqr_backward(F, C)
    for reg_type in ("Q", "V"):
        k, K, lam, mult, gl, gq, ok = G.lqg_backward(F, C + (np.zeros((3, 12, 2, 4)),), 0.0, 1.0, (2.0, 1e-6), reg_type)
        assert ok.all() and (lam == 0).all() and (mult == 1).all()
        np.testing.assert_allclose(k, k0, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(K, K0, rtol=1e-9, atol=1e-11)
        # with Q_uu k = -Q_u:  gains_lin = -gains_quad < 0   (lqr_solver.py:143-144)
        np.testing.assert_allclose(gl, -gq, rtol=1e-9)
        assert (gq > 0).all()


def test_lqg_lambda_schedule_on_indefinite_costs():
    rng = np.random.default_rng(12)
    N, H, d, m = 4, 6, 3, 2
    F, C = _lq_problem(rng, N, H, d, m)
    c_uu = C[3].copy()
    c_uu[1] = -0.5 * np.eye(m)          # trajectory 1: Q_uu indefinite until lambda is large enough
    c_uu[2] = -1e12 * np.eye(m)         # trajectory 2: no lambda <= 1e10 can fix it -> None, None
    c_ux = 0.05 * rng.standard_normal((N, H, m, d))
    Cc = (C[0], C[1], C[2], c_uu, c_ux)
    k, K, lam, mult, gl, gq, ok = G.lqg_backward(F, Cc, 0.0, 1.0, (10.0, 1e-3), "Q")
    assert ok.tolist() == [True, True, False, True]
    assert lam[0] == 0 and mult[0] == 1 and lam[3] == 0
    # the schedule: mult = 10, 100, ...; lam = max(lam * mult, 1e-3): 1e-3, 0.1, 100 -- the first value above 0.5
    assert lam[1] == pytest.approx(100.0) and mult[1] == pytest.approx(1000.0)
    # nine failed attempts: lam_j = 1e-3 * 10^(j(j+1)/2 - 1), capped at 1e10 (:35-36)
    assert lam[2] == 1e10 and mult[2] == pytest.approx(1e9) and np.isnan(gl[2]) and (K[2] == 0).all()
    # replay trajectory 1 with the lambda it found: one attempt, same gains
    k1, K1, lam1, *_ = G.lqg_backward(tuple(a[1:2] for a in F), tuple(a[1:2] for a in Cc), 100.0, 1000.0, (10.0, 1e-3), "Q")
    np.testing.assert_allclose(K1[0], K[1])
    assert lam1[0] == 100.0
    # scalar lambda_adjust, both directions (lqr_solver.py:21-37)
    assert G.lambda_adjust(0.0, 1.0, (10.0, 1e-3), True) == (1e-3, 10.0)
    assert G.lambda_adjust(1.0, 10.0, (10.0, 1e-3), False) == (pytest.approx(0.1), pytest.approx(0.1))
    assert G.lambda_adjust(1e-3, 0.1, (10.0, 1e-3), False) == (1e-3, pytest.approx(0.01))
    assert G.lambda_adjust(1e9, 100.0, (10.0, 1e-3), True)[0] == 1e10
