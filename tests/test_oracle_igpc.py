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
