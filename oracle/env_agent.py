"""The functional Env / Agent rollout (``deluca/training/apg.py:39-100``) on a classic env with the generic PID agent
(``deluca/agents/_pid.py:34-48``) or linear state feedback, restated on the CPU in numpy (test infrastructure; see
oracle/__init__.py).  Pinned: tests/test_ref_fixtures_oracle.py compares it with the reference's own ``apg_rollout`` /
``apg_parallel_rollouts`` / ``jax.value_and_grad(apg_loss)`` run on Pendulum x PID (tests/golden/ref_drivers.npz).

Per step (apg.py:47-63):  agent_state', action = agent(agent_state, obs);  env_state', obs' = env(env_state, action);
loss = loss_func(env_state, obs, action, env, t), with
    loss = sum q (x - goal)^2 + sum r (action[:m] - goal_u)^2
and d(sum_t loss)/d(gains) carried forward with the analytic env Jacobians of oracle/igpc.py.
"""
import numpy as np


def rollout(env, agent, gains, x0, T, q, goal, r, goal_u, pid_decay=None, pid_state=None, want_grad=False,
            dtype=np.float64):
    """env: an oracle/igpc.py env (step / jac batched over N).  agent "pid": gains [3] or [N,3]; "linear": gains
    [m,d] or [N,m,d] (u = -K x).  Returns a dict: losses [T,N], actions [T,N,da], X [T+1,N,d], and for PID the stacked
    agent states P/I/D [T,N,d]; with want_grad also grad [N,ng] = d(sum_t loss)/d(gains)."""
    c = dtype
    x = np.array(x0, dtype=c)
    N, D = x.shape
    M = env.m
    q, goal = np.asarray(q, c), np.asarray(goal, c)
    r, goal_u = np.asarray(r, c), np.asarray(goal_u, c)
    pid = agent == "pid"
    g = np.asarray(gains, c)
    if pid:
        g = np.broadcast_to(g.reshape(-1, 3), (N, 3))
        ng, da = 3, D
        P, I, Dv = [np.zeros((N, D), c) for _ in range(3)] if pid_state is None else [np.array(s, c) for s in pid_state]
        decay = c(pid_decay)
    else:
        g = np.broadcast_to(g.reshape(-1, M, D), (N, M, D))
        ng, da = M * D, M
    dx = np.zeros((N, D, ng), c)
    dP, dI, dD = np.zeros((N, D, ng), c), np.zeros((N, D, ng), c), np.zeros((N, D, ng), c)
    dloss = np.zeros((N, ng), c)
    out = dict(losses=np.zeros((T, N), c), actions=np.zeros((T, N, da), c), X=np.zeros((T + 1, N, D), c),
               P=np.zeros((T, N, D), c), I=np.zeros((T, N, D), c), D=np.zeros((T, N, D), c))
    out["X"][0] = x
    for t in range(T):
        if pid:
            In = I + decay * (x - I)                                   # _pid.py:42
            Dn = Dv + decay * ((x - P) - Dv)                           # _pid.py:43
            dI = dI + decay * (dx - dI)
            dD = dD + decay * ((dx - dP) - dD)
            dP = dx.copy()
            P, I, Dv = x.copy(), In, Dn
            act = g[:, 0:1] * P + g[:, 1:2] * I + g[:, 2:3] * Dv       # _pid.py:45
            dact = (g[:, 0, None, None] * dP + g[:, 1, None, None] * dI + g[:, 2, None, None] * dD)[:, :M].copy()
            dact[:, :, 0] += P[:, :M]; dact[:, :, 1] += I[:, :M]; dact[:, :, 2] += Dv[:, :M]
            out["P"][t], out["I"][t], out["D"][t] = P, I, Dv
        else:
            act = -np.einsum("nmd,nd->nm", g, x)
            dact = -np.einsum("nmd,ndk->nmk", g, dx)
            for cc in range(M):
                for j in range(D):
                    dact[:, cc, cc * D + j] -= x[:, j]
        u = act[:, :M]
        ex, eu = x - goal, u - goal_u
        out["losses"][t] = (q * ex * ex).sum(1) + (r * eu * eu).sum(1)
        out["actions"][t] = act
        if want_grad:
            dloss += np.einsum("nd,ndk->nk", 2 * q * ex, dx) + np.einsum("nm,nmk->nk", 2 * r * eu, dact)
            Fx, Fu = env.jac(x, u)
            dx = np.einsum("nij,njk->nik", Fx, dx) + np.einsum("nic,nck->nik", Fu, dact)
        x = env.step(x, u)
        out["X"][t + 1] = x
    if want_grad:
        out["grad"] = dloss
    return out
