#!/usr/bin/env python
"""Golden fixtures produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES (google/deluca, /root/reference).

    python tests/golden/make_ref_fixtures.py [family ...]        # writes tests/golden/ref_<family>.npz

JAX cannot be installed in this image, so the files run on `tests/refshim` -- an eager JAX-semantics shim
backed by torch CPU tensors (see its README) -- imported by path so that the reference's broken package `__init__`s are
bypassed.  Nothing of the reference is copied: the generator imports `deluca.agents._gpc.GPC`,
`deluca.envs._lds.LDS`, `deluca.lung.envs._balloon_lung.BalloonLung`, `deluca.igpc.ilqr.iLQR`, ... and drives them the
way the reference's own callers do (the loops are cited below).  Every random input (systems, disturbances,
perturbations, initial parameters) is recorded next to the outputs: the shim's `jax.random` is not Threefry, so parity
is defined on the recorded inputs.

Each family is run twice: float64 (`jax_enable_x64`) = the values the oracle and the kernels are compared with, and
float32 (JAX's default) = evidence of how far the reference's own default precision is from that (`*_f32` arrays).

This script needs /root/reference and therefore runs only in the build container; the `.npz` files travel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from refshim import loader  # noqa: E402


def npf(x):
    return np.asarray(x, dtype=np.float64)


def both_precisions(fn):
    """Run `fn()` under x64 and under default float32; merge the results (`name` and `name_f32`)."""
    import jax
    jax.config.update("jax_enable_x64", True)
    out = fn()
    jax.config.update("jax_enable_x64", False)
    out32 = fn()
    jax.config.update("jax_enable_x64", True)
    for k, v in out32.items():
        if k.startswith("out_"):
            out[k + "_f32"] = v
    return out


# ------------------------------------------------------------------------------------------ LDS + GPC / BPC / LQR
def _lds_systems(rng, N, d, m):
    """Per-env systems with the distributions of `LDS.init` (deluca/envs/_lds.py:71-82)."""
    A = np.zeros((N, d, d))
    A[:, np.arange(d), np.arange(d)] = rng.uniform(0.9, 0.95, (N, d))
    B = rng.standard_normal((N, d, m))
    return A, B


def family_lds():
    """`LDS.__call__` + `GPC.__call__` / `BPC.__call__` / `LQR.__call__` stepped as the stateful API is meant to be
    (the colab loop `u = agent(x); x = env(u)`; the notebook itself is missing from the snapshot):

        x = env.init(...);  for t: u = agent(x) [BPC: agent(x, cost_{t-1})];  cost_t = quad_loss(x, u);  x = env(u)

    with `C = I` so that the observation is the state, and a `Disturbance` subclass that replays a recorded W."""
    import jax.numpy as jnp
    from deluca.core import Disturbance
    from deluca.envs._lds import LDS
    from deluca.agents._gpc import GPC, quad_loss
    from deluca.agents._lqr import LQR
    import deluca.agents._bpc as bpc_mod

    class Replay(Disturbance):             # a Disturbance whose __call__(t, key) returns the recorded w_t
        def init(self, d):
            self.d = d

        def __call__(self, t, key):
            return jnp.array(self.W[t].reshape(self.d, 1))

    def make_env(A, B, W):
        dist = Replay()
        dist.W = W
        env = LDS()
        d, m = B.shape
        x = env.init(d_in=m, d_hidden=d, d_out=d, A=A, B=B, C=np.eye(d), disturbance=dist)
        return env, x

    cases = [  # name, d, m, H, HH, lr_scale, decay, T
        ("h3", 10, 3, 3, 10, 1e-4, True, 80),
        ("h10", 10, 3, 10, 10, 1e-4, True, 60),
        ("h3hh2", 10, 3, 3, 2, 1e-4, True, 60),
        ("d4", 4, 2, 5, 7, 2e-5, False, 60),
        ("default_lr", 10, 3, 3, 10, 0.005, True, 40),          # the reference's default learning rate
        ("clamp", 6, 2, 4, 1, 1e-4, True, 50),                  # HH < H-1: dynamic_slice clamping is exercised
        ("ref_defaults", 25, 1, 3, 2, 1e-4, True, 40),          # LDS(d_hidden=25, d_in=1) + GPC(H=3, HH=2) defaults
    ]

    def run():
        out = {}
        rng = np.random.default_rng(20241)
        N = 3
        for name, d, m, H, HH, lr, decay, T in cases:
            A, B = _lds_systems(rng, N, d, m)
            W = 0.5 * rng.standard_normal((T, N, d))
            costs = np.zeros((T, N)); X = np.zeros((T + 1, N, d)); U = np.zeros((T, N, m))
            Ks = np.zeros((N, m, d)); Mf = np.zeros((N, H, m, d)); hist = np.zeros((N, H + HH, d))
            for n in range(N):
                env, x = make_env(A[n], B[n], W[:, n])
                agent = GPC(jnp.array(A[n]), jnp.array(B[n]), H=H, HH=HH, lr_scale=lr, decay=decay)
                Ks[n] = npf(agent.K)                          # K = LQR(A, B).K, scipy DARE (_gpc.py:93, _lqr.py:57-60)
                for t in range(T):
                    X[t, n] = npf(x).ravel()
                    u = agent(x)
                    costs[t, n] = float(quad_loss(x, u))
                    U[t, n] = npf(u).ravel()
                    env(u)
                    x = env.x                                 # C = I: the observation is the state
                X[T, n] = npf(x).ravel()
                Mf[n], hist[n] = npf(agent.M), npf(agent.noise_history)[:, :, 0]
            out.update({f"in_gpc_{name}_A": A, f"in_gpc_{name}_B": B, f"in_gpc_{name}_W": W,
                        f"in_gpc_{name}_cfg": np.array([d, m, H, HH, lr, float(decay), T]),
                        f"out_gpc_{name}_K": Ks, f"out_gpc_{name}_costs": costs, f"out_gpc_{name}_X": X,
                        f"out_gpc_{name}_U": U, f"out_gpc_{name}_M": Mf, f"out_gpc_{name}_hist": hist})
        # ---- LQR on its own (u = -K x, with Q, R given)
        d, m, T = 10, 3, 50
        A, B = _lds_systems(rng, N, d, m)
        W = 0.5 * rng.standard_normal((T, N, d))
        Q = np.diag(rng.uniform(0.5, 2.0, d)); R = np.diag(rng.uniform(0.5, 2.0, m))
        costs = np.zeros((T, N)); X = np.zeros((T + 1, N, d)); Ks = np.zeros((N, m, d))
        for n in range(N):
            env, x = make_env(A[n], B[n], W[:, n])
            agent = LQR(jnp.array(A[n]), jnp.array(B[n]), jnp.array(Q), jnp.array(R))
            Ks[n] = npf(agent.K)
            for t in range(T):
                X[t, n] = npf(x).ravel()
                u = agent(x)
                costs[t, n] = float(quad_loss(x, u))
                env(u)
                x = env.x
            X[T, n] = npf(x).ravel()
        out.update(in_lqr_A=A, in_lqr_B=B, in_lqr_W=W, in_lqr_Q=Q, in_lqr_R=R, out_lqr_K=Ks, out_lqr_costs=costs,
                   out_lqr_X=X)
        # ---- BPC: the perturbations come from the global numpy RNG (_bpc.py:20,26-30); they are recorded as drawn.
        # WAIVER (found by running the file): BPC as written raises AttributeError on its first call -- `self.eps` is a
        # NumPy array (`generate_uniform` returns np.array, _bpc.py:29, and `np.roll` at :137 returns one again), and
        # NumPy arrays have no `.at` (:134).  The evident intent (a jnp array updated functionally) is restored by
        # re-wrapping `agent.eps = jnp.array(agent.eps)` before every call; no arithmetic changes.
        d, m, H, T = 10, 3, 5, 40
        A, B = _lds_systems(rng, N, d, m)
        W = 0.5 * rng.standard_normal((T, N, d))
        costs = np.zeros((T, N)); X = np.zeros((T + 1, N, d)); Ks = np.zeros((N, m, d))
        M0 = np.zeros((N, H, m, d)); eps0 = np.zeros((N, H, H, m, d)); eps_new = np.zeros((T, N, H, m, d))
        Mf = np.zeros((N, H, m, d))
        orig = bpc_mod.generate_uniform
        for n in range(N):
            drawn = []

            def recording(shape, norm=1.00):
                v = orig(shape, norm)
                drawn.append(np.array(v))
                return v

            bpc_mod.generate_uniform = recording
            try:
                np.random.seed(100 + n)
                env, x = make_env(A[n], B[n], W[:, n])
                # (the default lr_scale = 0.005 makes the bandit update blow up within 40 steps on the default LDS; 2e-5 stays finite)
                agent = bpc_mod.BPC(jnp.array(A[n]), jnp.array(B[n]), H=H, lr_scale=2e-5, decay=(n % 2 == 1), delta=0.01)
                Ks[n] = npf(agent.K)
                M0[n], eps0[n] = npf(agent.M), npf(agent.eps)      # drawn[0] * delta, drawn[1]; drawn[2] = eps_bias (unused)
                n_init = len(drawn)
                cost = 0.0
                for t in range(T):
                    X[t, n] = npf(x).ravel()
                    agent.eps = jnp.array(agent.eps)            # see WAIVER above
                    u = agent(x, cost)                         # the cost handed over is the one realised a step earlier
                    cost = float(quad_loss(x, u))
                    costs[t, n] = cost
                    env(u)
                    x = env.x
                X[T, n] = npf(x).ravel()
                eps_new[:, n] = np.stack(drawn[n_init:])
                Mf[n] = npf(agent.M)
            finally:
                bpc_mod.generate_uniform = orig
        out.update(in_bpc_A=A, in_bpc_B=B, in_bpc_W=W, in_bpc_M0=M0, in_bpc_eps0=eps0, in_bpc_eps_new=eps_new,
                   in_bpc_cfg=np.array([d, m, H, 2e-5, 0.01, T]), out_bpc_K=Ks, out_bpc_costs=costs, out_bpc_X=X,
                   out_bpc_M=Mf)
        # ---- LDS on its own: y = C x with C != I, actions given (LDS.__call__, _lds.py:102-111)
        d, n_in, p, T = 10, 3, 2, 30
        A, B = _lds_systems(rng, N, d, n_in)
        C = rng.standard_normal((N, p, d)); Uin = rng.standard_normal((T, N, n_in)); W = 0.5 * rng.standard_normal((T, N, d))
        Y = np.zeros((T, N, p)); xf = np.zeros((N, d))
        for n in range(N):
            dist = Replay(); dist.W = W[:, n]
            env = LDS()
            env.init(d_in=n_in, d_hidden=d, d_out=p, A=A[n], B=B[n], C=C[n], disturbance=dist)
            for t in range(T):
                Y[t, n] = npf(env(jnp.array(Uin[t, n].reshape(n_in, 1)))).ravel()
            xf[n] = npf(env.x).ravel()
            assert env.t == T
        out.update(in_open_A=A, in_open_B=B, in_open_C=C, in_open_U=Uin, in_open_W=W, out_open_Y=Y, out_open_x=xf)
        return out

    return both_precisions(run)


# ------------------------------------------------------------------------------------------ lung
def family_lung():
    """`run_controller_scan` (deluca/lung/utils/scripts/run_controller.py:144-220) and `test_controller`
    (test_controller.py:23-46) executed literally: lung `PID` (explicit gains) + `Expiratory` driving `BalloonLung` /
    `DelayLung`; plus the `BreathWaveform` / `Controller.decay` / `proper_time` helpers and the generic `PID` agent
    (deluca/agents/_pid.py)."""
    import jax.numpy as jnp
    from deluca.lung.core import BreathWaveform, Controller, proper_time
    from deluca.lung.controllers._pid import PID
    from deluca.lung.envs._balloon_lung import BalloonLung
    from deluca.lung.envs._delay_lung import DelayLung
    from deluca.lung.utils.scripts.run_controller import run_controller_scan
    from deluca.lung.utils.scripts.test_controller import test_controller
    from deluca.agents._pid import PID as GenericPID

    gains = np.array([[3.0, 4.0, 0.0], [1.0, 2.5, 0.3], [0.4, 0.05, 0.0], [4.5, 0.8, 1.5]])
    pips = [10.0, 20.0, 35.0]

    def run():
        out = dict(in_gains=gains, in_pips=np.array(pips), in_peep=np.array(5.0))
        for sim_name, make in (("balloon", lambda wf: BalloonLung.create(waveform=wf)),
                               ("delay", lambda wf: DelayLung.create(waveform=wf)),
                               ("balloon_leak", lambda wf: BalloonLung.create(waveform=wf, leak=True))):
            for T in (29, 150):
                series = {k: np.zeros((len(gains), len(pips), T)) for k in
                          ("timestamp", "pressure", "flow", "target", "u_in", "u_out")}
                for gi, g in enumerate(gains):
                    ctrl = PID.create(params={"K": {"kernel": jnp.array(g.reshape(3, 1))}})
                    for pi_, pip in enumerate(pips):
                        wf = BreathWaveform.create(peep=5.0, pip=pip)
                        res = run_controller_scan(ctrl, T=T, abort=T, env=make(wf), waveform=wf, init_controller=True)
                        for k in series:
                            series[k][gi, pi_] = npf(res["timeseries"][k]).reshape(T)
                for k, v in series.items():
                    out[f"out_{sim_name}_T{T}_{k}"] = v
            # test_controller: mean over pips of mean |pressure - target| at horizon 29 (test_controller.py:26-45)
            scores = np.zeros(len(gains))
            for gi, g in enumerate(gains):
                ctrl = PID.create(params={"K": {"kernel": jnp.array(g.reshape(3, 1))}})
                scores[gi] = float(test_controller(ctrl, make(BreathWaveform.create()), pips, 5.0))
            out[f"out_{sim_name}_score"] = scores
        # ---- helpers with reference-held known answers (deluca/tests/lung/core_test.py) plus a denser sweep
        wf = BreathWaveform.create()
        tt = np.linspace(0.0, 6.5, 131)
        out["in_wave_t"] = tt
        out["out_wave_at"] = np.array([float(wf.at(float(t))) for t in tt])
        out["out_wave_is_in"] = np.array([float(bool(wf.is_in(float(t)))) for t in tt])
        ctl = Controller()
        out["out_decay"] = np.array([float(ctl.decay(wf, float(t))) for t in tt])
        out["out_proper_time"] = np.array([float(proper_time(float("inf"))), float(proper_time(1.25))])
        # ---- generic PID agent (deluca/agents/_pid.py:20-48): obs is the error signal
        rng = np.random.default_rng(7)
        err = rng.standard_normal(40)
        agent = GenericPID.create(K_P=1.5, K_I=0.7, K_D=0.2)
        st = agent.init()
        u = []
        for e in err:
            st, a = agent(st, jnp.array(e))
            u.append(float(a))
        out.update(in_pid_err=err, in_pid_gains=np.array([1.5, 0.7, 0.2]), out_pid_u=np.array(u),
                   out_pid_state=np.array([float(st.P), float(st.I), float(st.D)]))
        return out

    return both_precisions(run)


# ------------------------------------------------------------------------------------------ classic envs + deluca/igpc
def _igpc_envs():
    """The reference envs behind the protocol `deluca/igpc` expects (`env.init()` -> bare state with
    `flatten()/unflatten()`, `env.state_dim/action_dim`; SURVEY A-6).  ADAPTERS ONLY -- every number still comes from the
    reference's `__call__`:
      * PlanarQuadrotor / Reacher: `init()` returns `(state, obs)`; igpc needs the state (`rollout.py:47-48`).
      * Pendulum: its state lacks `flatten/unflatten`, `h` is an int (jacfwd rejects integer leaves) and the env lacks
        `state_dim/action_dim`; the adapter re-wraps the state the reference's step returns."""
    import jax.numpy as jnp
    from deluca.envs.classic._pendulum import Pendulum, PendulumState
    from deluca.envs.classic._planar_quadrotor import PlanarQuadrotor, constant, dissipative
    from deluca.envs.classic._reacher import Reacher

    class Quad(PlanarQuadrotor):
        def init(self):
            return PlanarQuadrotor.init(self)[0]

    class Reach(Reacher):
        def init(self):
            return Reacher.init(self)[0]

    class PendState(PendulumState):
        def flatten(self):
            return self.arr

        def unflatten(self, arr):
            return PendState(arr=arr, h=self.h)

    class Pend(Pendulum):
        state_dim = 3
        action_dim = 1

        def init(self):
            return PendState(arr=Pendulum.init(self)[0].arr, h=0.0)

        def __call__(self, state, action):
            s, obs = Pendulum.__call__(self, state, action)
            return PendState(arr=s.arr, h=s.h), obs

    return dict(Quad=Quad, Reach=Reach, Pend=Pend, constant=constant, dissipative=dissipative)


def family_classic():
    """Single steps and short open-loop rollouts of the classic envs, `Env.__call__(state, action)` executed literally
    (deluca/envs/classic/_pendulum.py:54-73, _planar_quadrotor.py:92-104, _reacher.py:91-141); Cartpole's literal
    `__call__` is recorded as raising (SURVEY A-7)."""
    import jax.numpy as jnp
    from deluca.envs.classic._pendulum import Pendulum, PendulumState
    from deluca.envs.classic._cartpole import Cartpole
    from deluca.envs.classic._planar_quadrotor import PlanarQuadrotor, PlanarQuadrotorState, constant
    from deluca.envs.classic._reacher import Reacher, ReacherState

    def run():
        rng = np.random.default_rng(11)
        out = {}
        N, T = 4, 25
        specs = [
            ("pendulum", Pendulum.create(), lambda a: PendulumState(arr=jnp.array(a)), 3, 1),
            ("quad", PlanarQuadrotor.create(), lambda a: PlanarQuadrotorState(arr=jnp.array(a), last_action=jnp.zeros(2)), 6, 2),
            ("quad_wind", PlanarQuadrotor.create(wind=0.3), lambda a: PlanarQuadrotorState(arr=jnp.array(a), last_action=jnp.zeros(2)), 6, 2),
            ("quad_const", PlanarQuadrotor.create(wind=0.2, wind_func=constant), lambda a: PlanarQuadrotorState(arr=jnp.array(a), last_action=jnp.zeros(2)), 6, 2),
            ("reacher", Reacher.create(), lambda a: ReacherState(arr=jnp.array(a)), 6, 2),
        ]
        for name, env, mk, d, m in specs:
            x0 = np.stack([npf(env.init()[0].arr) for _ in range(N)]) + 0.05 * rng.standard_normal((N, d))
            if name == "pendulum":            # keep (sin, cos) on the unit circle, as every reachable state is
                th = rng.uniform(-3, 3, N)
                x0 = np.stack([np.sin(th), np.cos(th), rng.standard_normal(N)], 1)
            U = (0.5 if name != "reacher" else 0.2) * rng.standard_normal((T, N, m))
            if name.startswith("quad"):
                U = U * 0.1 + 0.4905
            X = np.zeros((T + 1, N, d))
            for n in range(N):
                st = mk(x0[n])
                X[0, n] = x0[n]
                for t in range(T):
                    st, obs = env(st, jnp.array(U[t, n]))
                    X[t + 1, n] = npf(st.arr)
                    assert np.array_equal(npf(obs), npf(st.arr))
            out.update({f"in_{name}_x0": x0, f"in_{name}_U": U, f"out_{name}_X": X})
        env = Cartpole.create()
        st = env.init()[0]
        try:
            env(st, jnp.array([0.1]))
            raised = 0.0
        except TypeError:
            raised = 1.0
        out["out_cartpole_literal_raises_typeerror"] = np.array(raised)
        return out

    return both_precisions(run)


def family_igpc():
    """`deluca/igpc`: `rollout` (rollout.py:22-65), `compute_ders` (:68-99, jacfwd / grad / hessian), `LQR`
    (lqr_solver.py:40-86), `iLQR` (ilqr.py:27-93, as written: `alpha` for `alphaC`), `GPC_rollout` (igpc.py:63-126),
    `iGPC_closed` (:129-179), `iLC_closed` (ilc.py:23-69), executed literally on PlanarQuadrotor (true env with wind,
    simulator without) and on Pendulum.  Cost (the reference ships none):
    c(x,u) = sum (x - goal_state)^2 + 0.1 sum (u - goal_action)^2."""
    import contextlib
    import io
    import jax.numpy as jnp
    from deluca.igpc.rollout import rollout, compute_ders
    from deluca.igpc.lqr_solver import LQR
    from deluca.igpc.ilqr import iLQR
    from deluca.igpc.igpc import GPC_rollout, iGPC_closed
    from deluca.igpc.ilc import iLC_closed
    E = _igpc_envs()

    def cost_quad(state, u, env):
        dx = state.arr - env.goal_state
        du = u - env.goal_action
        return jnp.sum(dx * dx) + 0.1 * jnp.sum(du * du)

    def cost_pend(state, u, env):
        dx = state.arr - env.goal_state
        return jnp.sum(dx * dx) + 0.1 * jnp.sum(u * u)

    def lists(X, U):
        return np.stack([npf(x.arr) for x in X]), np.stack([npf(u) for u in U])

    def run():
        out = {}
        H, T = 24, 4
        quiet = contextlib.redirect_stdout(io.StringIO())
        for name, env_true, env_sim, cost, m in (
                ("quad", E["Quad"].create(H=H, wind=0.1), E["Quad"].create(H=H), cost_quad, 2),
                ("quadc", E["Quad"].create(H=H, wind=0.1, wind_func=E["constant"]), E["Quad"].create(H=H), cost_quad, 2),
                ("pend", E["Pend"].create(H=H, g=9.0), E["Pend"].create(H=H), cost_pend, 1)):
            goal_u = npf(env_sim.goal_action) if hasattr(env_sim, "goal_action") and env_sim.goal_action is not None else np.zeros(m)
            U0 = np.tile(goal_u, (H, 1))
            if name == "pend":           # the default start [0, 1, 0] is a fixed point of the pendulum map under u = 0
                U0 = U0 + 0.4 * np.sin(0.7 * np.arange(H))[:, None] + 0.2
            U0 = jnp.array(U0)
            with quiet:
                # open-loop rollout, derivatives, Riccati pass, closed-loop rollout
                X, U, c = rollout(env_sim, cost, U0)
                D, F, C = compute_ders(env_sim, cost, X, U)
                k, K = LQR(F, C)
                X1, U1, c1 = rollout(env_sim, cost, U, k, K, X, 0.5)
                Xa, Ua = lists(X, U)
                X1a, U1a = lists(X1, U1)
                out.update({f"in_{name}_U0": npf(U0), f"out_{name}_X": Xa, f"out_{name}_c": npf(c),
                            f"out_{name}_D": npf(D),
                            f"out_{name}_Fx": np.stack([npf(f[0]) for f in F]), f"out_{name}_Fu": np.stack([npf(f[1]) for f in F]),
                            f"out_{name}_cx": np.stack([npf(cc[0]) for cc in C]), f"out_{name}_cu": np.stack([npf(cc[1]) for cc in C]),
                            f"out_{name}_cxx": np.stack([npf(cc[2]) for cc in C]), f"out_{name}_cuu": np.stack([npf(cc[3]) for cc in C]),
                            f"out_{name}_k": np.stack([npf(v) for v in k]), f"out_{name}_K": np.stack([npf(v) for v in K]),
                            f"out_{name}_X1": X1a, f"out_{name}_U1": U1a, f"out_{name}_c1": npf(c1)})
                # iLQR as written (on the simulator) and with the alpha/alphaC slip left in place
                Xi, Ui, ki, Ki, ci = iLQR(env_sim, cost, U0, T, verbose=False)
                Xia, Uia = lists(Xi, Ui)
                out.update({f"out_{name}_ilqr_X": Xia, f"out_{name}_ilqr_U": Uia, f"out_{name}_ilqr_c": npf(ci)})
                # one GPC_rollout on the true env around the iLQR solution (all three disturbance modes)
                Dg, Fg, Cg = compute_ders(env_sim, cost, Xi, Ui)
                kg, Kg = LQR(Fg, Cg)
                for w_is in ("de", "dede", "delin"):
                    Xg, Ug, cg = GPC_rollout(env_true, env_sim, cost, Ui, kg, Kg, Xi, Dg, Fg, 0.6, w_is, 0.05)
                    Xga, Uga = lists(Xg, Ug)
                    out.update({f"out_{name}_gpc_{w_is}_X": Xga, f"out_{name}_gpc_{w_is}_U": Uga,
                                f"out_{name}_gpc_{w_is}_c": npf(cg)})
                # the outer loops
                Xc, Uc, kc, Kc, cc_ = iGPC_closed(env_true, env_sim, cost, U0, T, lr=0.05, verbose=False)
                Xca, Uca = lists(Xc, Uc)
                out.update({f"out_{name}_igpc_X": Xca, f"out_{name}_igpc_U": Uca, f"out_{name}_igpc_c": npf(cc_)})
                Xl, Ul, kl, Kl, cl = iLC_closed(env_true, env_sim, cost, U0, T, verbose=False)
                Xla, Ula = lists(Xl, Ul)
                out.update({f"out_{name}_ilc_X": Xla, f"out_{name}_ilc_U": Ula, f"out_{name}_ilc_c": npf(cl)})
        return out

    return both_precisions(run)


# ------------------------------------------------------------------------------------------ output-feedback agents
def family_of():
    """`GRC` / `SFC` / `DSC` / `DRC` / `LQG` (deluca/agents/_grc.py, _sfc.py, _dsc.py, _drc.py, _lqg.py) executed
    literally on the reference's partially observed `LDS` (C != I, recorded disturbances):

        y = C x0;  for t: u = agent(y);  cost_t = quad_loss(y, u);  y = env(u)

    The agents draw their initial parameters from `jax.random` (not Threefry here): they are recorded after
    construction and handed to the oracle / kernels as inputs."""
    import jax
    import jax.numpy as jnp
    from deluca.core import Disturbance
    from deluca.envs._lds import LDS
    from deluca.agents._grc import GRC, quad_loss
    from deluca.agents._sfc import SFC
    from deluca.agents._dsc import DSC
    from deluca.agents._drc import DRC
    from deluca.agents._lqg import LQG

    class Replay(Disturbance):
        def init(self, d):
            self.d = d

        def __call__(self, t, key):
            return jnp.array(self.W[t].reshape(self.d, 1))

    def run():
        rng = np.random.default_rng(31)
        out = {}
        N, d, n, p, T = 2, 10, 3, 2, 60
        A = np.zeros((N, d, d))
        A[:, np.arange(d), np.arange(d)] = rng.uniform(0.9, 0.95, (N, d))
        B = rng.standard_normal((N, d, n)) / np.sqrt(d)
        C = rng.standard_normal((N, p, d)) / np.sqrt(d)
        W = 0.1 * rng.standard_normal((T, N, d))
        x0 = 0.3 * rng.standard_normal((N, d))
        out.update(in_A=A, in_B=B, in_C=C, in_W=W, in_x0=x0)

        def episode(make, grab_init, grab_final, name):
            costs = np.zeros((T, N)); U = np.zeros((T, N, n)); init, final = [], []
            xf = np.zeros((N, d))
            for e in range(N):
                dist = Replay(); dist.W = W[:, e]
                env = LDS()
                env.init(d_in=n, d_hidden=d, d_out=p, A=A[e], B=B[e], C=C[e], x0=x0[e].reshape(d, 1), disturbance=dist)
                agent = make(jnp.array(A[e]), jnp.array(B[e]), jnp.array(C[e]), e)
                init.append(grab_init(agent))
                y = jnp.array(C[e]) @ env.x
                for t in range(T):
                    u = agent(y)
                    costs[t, e] = float(quad_loss(y, u))
                    U[t, e] = npf(u).ravel()
                    y = env(u)
                final.append(grab_final(agent))
                xf[e] = npf(env.x).ravel()
            out.update({f"out_{name}_costs": costs, f"out_{name}_U": U, f"out_{name}_x": xf})
            for k in init[0]:
                out[f"in_{name}_{k}"] = np.stack([i[k] for i in init])
            for k in final[0]:
                out[f"out_{name}_{k}"] = np.stack([f[k] for f in final])

        episode(lambda a, b, c, e: GRC(a, b, c, key=jax.random.PRNGKey(e), init_scale=0.2, m=5, lr=0.002),
                lambda ag: dict(M=npf(ag.M)), lambda ag: dict(Mf=npf(ag.M), ynat=npf(ag.ynat_history)[:, :, 0]), "grc")
        episode(lambda a, b, c, e: GRC(a, b, c, key=jax.random.PRNGKey(e), init_scale=3.0, m=5, lr=0.05, R_M=2.0, decay=False),
                lambda ag: dict(M=npf(ag.M)), lambda ag: dict(Mf=npf(ag.M)), "grc_proj")      # the norm projection is active
        episode(lambda a, b, c, e: SFC(a, b, c, key=jax.random.PRNGKey(e), init_scale=0.2, h=4, m=8, lr=0.002),
                lambda ag: dict(M=npf(ag.M), M_0=npf(ag.M_0), filters=npf(ag.filters)),
                lambda ag: dict(Mf=npf(ag.M), M_0f=npf(ag.M_0)), "sfc")
        episode(lambda a, b, c, e: DSC(a, b, c, key=jax.random.PRNGKey(e), init_scale=0.2, h=3, m=6, lr=0.002),
                lambda ag: dict(M_0=npf(ag.M_0), M_1=npf(ag.M_1), M_2=npf(ag.M_2), M_3=npf(ag.M_3), filters=npf(ag.filters)),
                lambda ag: dict(M_0f=npf(ag.M_0), M_1f=npf(ag.M_1), M_2f=npf(ag.M_2), M_3f=npf(ag.M_3)), "dsc")
        Kdrc = 0.1 * rng.standard_normal((N, n, p))
        out["in_drc_K"] = Kdrc
        episode(lambda a, b, c, e: DRC(a, b, c, K=jnp.array(Kdrc[e]), m=4, h=6, lr_scale=0.03, decay=True),
                lambda ag: dict(M=npf(ag.M), G=npf(ag.G)), lambda ag: dict(Mf=npf(ag.M)), "drc")
        episode(lambda a, b, c, e: DRC(a, b, c, K=jnp.array(Kdrc[e]), m=4, h=6, lr_scale=0.03, decay=False, RM=0.5),
                lambda ag: dict(M=npf(ag.M)), lambda ag: dict(Mf=npf(ag.M)), "drc_nodecay")    # lr = 1.0 as written; projection active
        episode(lambda a, b, c, e: LQG(a, b, c), lambda ag: dict(K=npf(ag.K), L=npf(ag.L)),
                lambda ag: dict(x_hat=npf(ag.x_hat)[:, 0]), "lqg")
        return out

    return both_precisions(run)


# ------------------------------------------------------------------------------------------ apg / ppo drivers
def family_drivers():
    """`apg_rollout` / `apg_parallel_rollouts` (deluca/training/apg.py:39-100: lax.scan of agent -> env -> loss, vmapped
    with in_axes (None, 0, 0, None, 0, None, None)) executed literally for the one functional (env, agent) pair of the
    reference that composes as written -- `Pendulum` x generic `PID` (the PID acts on the observation vector; the
    pendulum consumes `action[0]`) -- with a user `loss_func`; and `gae_advantages` (deluca/training/ppo.py:44-84)."""
    import jax
    import jax.numpy as jnp
    from deluca.envs.classic._pendulum import Pendulum, PendulumState
    from deluca.agents._pid import PID, PIDState
    from deluca.training.apg import apg_parallel_rollouts, apg_rollout, apg_loss
    from deluca.training.ppo import gae_advantages

    def loss_func(env_state, env_obs, action, env, counter):
        dx = env_state.arr - env.goal_state
        return jnp.sum(dx * dx) + 0.1 * action[0] * action[0]

    def run():
        rng = np.random.default_rng(5)
        out = {}
        N, T = 6, 40
        th = rng.uniform(-3, 3, N)
        arr0 = np.stack([np.sin(th), np.cos(th), 0.5 * rng.standard_normal(N)], 1)
        gains = np.array([0.8, 0.3, 0.1])
        env = Pendulum.create()
        agent = PID.create(K_P=gains[0], K_I=gains[1], K_D=gains[2])
        st = PendulumState(arr=jnp.array(arr0), h=jnp.zeros(N, dtype=jnp.int32))
        ast = PIDState(P=jnp.zeros((N, 3)), I=jnp.zeros((N, 3)), D=jnp.zeros((N, 3)))
        ro = apg_parallel_rollouts(env, st, jnp.array(arr0), agent, ast, T, loss_func)
        out.update(in_apg_arr0=arr0, in_apg_gains=gains, out_apg_losses=npf(ro.losses), out_apg_actions=npf(ro.actions),
                   out_apg_P=npf(ro.agent_states.P), out_apg_I=npf(ro.agent_states.I), out_apg_D=npf(ro.agent_states.D),
                   out_apg_mean_loss=npf(apg_loss(agent, env, st, jnp.array(arr0), ast, T, loss_func)))
        one = apg_rollout(env, PendulumState(arr=jnp.array(arr0[0]), h=0), jnp.array(arr0[0]), agent,
                          PIDState(P=jnp.zeros(3), I=jnp.zeros(3), D=jnp.zeros(3)), T, loss_func)
        out["out_apg_single_losses"] = npf(one.losses)
        # value_and_grad of apg_loss w.r.t. the agent's leaves (K_P, K_I, K_D), as train_chunk takes it (apg.py:152-167)
        val, g = jax.value_and_grad(apg_loss)(agent, env, st, jnp.array(arr0), ast, T, loss_func)
        out.update(out_apg_grad=np.array([float(g.K_P), float(g.K_I), float(g.K_D)]), out_apg_value=npf(val))
        # ---- GAE
        Tg, Ng = 29, 5
        rewards = rng.standard_normal((Tg, Ng)); values = rng.standard_normal((Tg, Ng))
        masks = (rng.uniform(size=(Tg, Ng)) > 0.1).astype(np.float64)
        adv = gae_advantages(jnp.array(rewards), jnp.array(masks), jnp.array(values), 0.99, 0.95)
        out.update(in_gae_rewards=rewards, in_gae_values=values, in_gae_masks=masks, out_gae_adv=npf(adv))
        return out

    return both_precisions(run)


def family_lung_steps():
    """Every state the reference visits: `loop_over_tt` (run_controller.py:111-141, the body `run_controller_scan` scans)
    called step by step, recording the full carry (simulator state, observation, PID state, Expiratory state) BEFORE each
    step and the step's outputs.  Lets a test restart the kernel from any visited state and compare ONE step, which
    separates per-step arithmetic from the closed loop's amplification of rounding (BalloonLung + PID is locally
    expansive on the float32 waveform's 5 -> 35 ramp)."""
    import jax.numpy as jnp
    from deluca.lung.core import BreathWaveform
    from deluca.lung.controllers._expiratory import Expiratory
    from deluca.lung.controllers._pid import PID
    from deluca.lung.envs._balloon_lung import BalloonLung
    from deluca.lung.envs._delay_lung import DelayLung
    from deluca.lung.utils.scripts.run_controller import loop_over_tt

    gains = np.array([[3.0, 4.0, 0.0], [1.0, 2.5, 0.3], [0.4, 0.05, 0.0], [4.5, 0.8, 1.5]])
    pips = [10.0, 20.0, 35.0]
    T = 150
    f = lambda v: float(v)

    def run():
        out = dict(in_gains=gains, in_pips=np.array(pips))
        for sim_name, make in (("balloon", lambda wf: BalloonLung.create(waveform=wf)),
                               ("delay", lambda wf: DelayLung.create(waveform=wf, delay=3)),
                               ("balloon_leak", lambda wf: BalloonLung.create(waveform=wf, leak=True))):
            names = ["env_steps", "env_time", "env_volume", "env_pressure", "env_target", "obs_pressure", "obs_time",
                     "pid_p", "pid_i", "pid_d", "pid_time", "pid_dt", "pid_steps", "exp_time", "exp_dt", "exp_steps",
                     "u_in", "u_out", "pressure", "timestamp"]
            if sim_name == "delay":
                names += ["env_pipe"]
            rec = {k: np.zeros((len(gains) * len(pips), T + 1)) for k in names}
            hist = np.zeros((len(gains) * len(pips), T + 1, 2, 3)) if sim_name == "delay" else None
            b = 0
            for g in gains:
                ctrl = PID.create(params={"K": {"kernel": jnp.array(g.reshape(3, 1))}})
                for pip in pips:
                    wf = BreathWaveform.create(peep=5.0, pip=pip)
                    env = make(wf)
                    exp = Expiratory.create(waveform=wf)
                    cs, es = ctrl.init(wf), exp.init()
                    state, obs = env.reset()
                    # `lax.scan` traces its carry: Python-scalar leaves become arrays of the default dtype, which is what
                    # makes `state.time + self.dt` accumulate in float32 under run_controller_scan
                    from jax._core import arrayify
                    carry = arrayify((state, obs, cs, es, 0))
                    for t in range(T + 1):
                        state, obs, cs, es, _ = carry
                        r = rec
                        r["env_steps"][b, t], r["env_time"][b, t] = f(state.steps), f(state.time)
                        r["env_volume"][b, t], r["env_pressure"][b, t] = f(state.volume), f(state.predicted_pressure)
                        r["env_target"][b, t] = f(state.target)
                        r["obs_pressure"][b, t], r["obs_time"][b, t] = f(obs.predicted_pressure), f(obs.time)
                        r["pid_p"][b, t], r["pid_i"][b, t], r["pid_d"][b, t] = f(cs.p), f(cs.i), f(cs.d)
                        r["pid_time"][b, t], r["pid_dt"][b, t], r["pid_steps"][b, t] = f(cs.time), f(cs.dt), f(cs.steps)
                        r["exp_time"][b, t], r["exp_dt"][b, t], r["exp_steps"][b, t] = f(es.time), f(es.dt), f(es.steps)
                        if sim_name == "delay":
                            r["env_pipe"][b, t] = f(state.pipe_pressure)
                            hist[b, t, 0], hist[b, t, 1] = npf(state.in_history), npf(state.out_history)
                        if t == T:
                            break
                        carry, (ts, u_in, u_out, p, fl) = loop_over_tt(carry, None, ctrl, exp, env, 0.03)
                        carry = arrayify(carry)
                        r["timestamp"][b, t], r["u_in"][b, t], r["u_out"][b, t], r["pressure"][b, t] = (
                            f(ts), f(jnp.squeeze(u_in)), f(u_out), f(p))
                    b += 1
            for k, v in rec.items():
                out[f"out_{sim_name}_{k}"] = v
            if hist is not None:
                out[f"out_{sim_name}_hist"] = hist
        return out

    return both_precisions(run)


def family_lung_train():
    """`train_controller.rollout` / `rollout_parallel` (deluca/lung/utils/scripts/train_controller.py:33-92) and
    `jax.value_and_grad(rollout_parallel)` as `train_controller` takes it w.r.t. the controller's leaves (`:151-153`),
    executed literally for the lung `PID` (explicit gains) on `BalloonLung` (with and without leak) and `DelayLung`.
    DelayLung as written compares seconds with its step-count `delay` (`_delay_lung.py:88-96`): with the default
    delay = 25 nothing of the action reaches the pressure for 833 steps and the gradient is exactly zero at the training
    horizon.  The loss only counts inhale steps, so the action first shows in it during the breath after the impulses
    start: `delay=1` (impulses from step 34 on) over 130 steps, and the default delay over 940 steps (l1 only).  The noise of
    `:47-50` is `1.5 + normal(PRNGKey(0))`, the same constant every step; the shim's generator is not Threefry, so the
    value is recorded (`in_noise_value`)."""
    import jax
    import jax.numpy as jnp
    from deluca.lung.controllers._pid import PID
    from deluca.lung.envs._balloon_lung import BalloonLung
    from deluca.lung.envs._delay_lung import DelayLung
    from deluca.lung.utils.scripts import train_controller as TC

    gains = np.array([[3.0, 4.0, 0.0], [1.0, 2.5, 0.3], [0.4, 0.05, 0.0], [4.5, 0.8, 1.5]])
    pips = [10.0, 20.0, 35.0]
    losses = {"l1": lambda x, y: (jnp.abs(x - y)).mean(),           # the default loss_fn (:106)
              "l2": lambda x, y: ((x - y) ** 2).mean()}
    cases = [("balloon", lambda: BalloonLung.create(), 29), ("balloon_leak", lambda: BalloonLung.create(leak=True), 29),
             ("balloon", lambda: BalloonLung.create(), 60),
             ("delay", lambda: DelayLung.create(), 29), ("delay1", lambda: DelayLung.create(delay=1), 130),
             ("delay", lambda: DelayLung.create(), 940)]

    def run():
        out = dict(in_gains=gains, in_pips=np.array(pips), in_peep=np.array(5.0),
                   in_noise_value=np.array(float(1.5 + 1.0 * jax.random.normal(jax.random.PRNGKey(0), shape=()))))
        for sim_name, make, T in cases:
            tt = jnp.array(np.linspace(0.0, 0.03 * T, T))            # jnp.linspace(0, duration, int(duration / dt)), :148
            out[f"in_tt_T{T}"] = npf(tt)
            for loss_name, loss_fn in losses.items():
                for use_noise in (0.0, 1.0):
                    if use_noise and (loss_name == "l2" or sim_name == "balloon_leak"):
                        continue
                    if T > 500 and (use_noise or loss_name == "l2"):
                        continue
                    tag = f"{sim_name}_T{T}_{loss_name}_n{int(use_noise)}"
                    per = np.zeros((len(gains), len(pips)))
                    gper = np.zeros((len(gains), len(pips), 3))
                    val = np.zeros(len(gains))
                    grad = np.zeros((len(gains), 3))
                    for gi, g in enumerate(gains):
                        ctrl = PID.create(params={"K": {"kernel": jnp.array(g.reshape(3, 1))}})
                        sim = make()
                        for pi_, pip in enumerate(pips):
                            per[gi, pi_] = float(TC.rollout(ctrl, sim, tt, use_noise, 5.0, pip, loss_fn, 0.0))
                            v1, g1 = jax.value_and_grad(TC.rollout_parallel)(ctrl, sim, tt, use_noise, 5.0,
                                                                             jnp.array([pip]), loss_fn)
                            assert abs(float(v1) - per[gi, pi_]) <= 1e-9 * max(1.0, abs(per[gi, pi_]))
                            gper[gi, pi_] = npf(g1.params["K"]["kernel"]).reshape(3)
                        v, gr = jax.value_and_grad(TC.rollout_parallel)(ctrl, sim, tt, use_noise, 5.0, jnp.array(pips),
                                                                        loss_fn)
                        val[gi], grad[gi] = float(v), npf(gr.params["K"]["kernel"]).reshape(3)
                    out[f"out_{tag}_loss"], out[f"out_{tag}_loss_grad"] = per, gper
                    out[f"out_{tag}_value"], out[f"out_{tag}_grad"] = val, grad
        return out

    return both_precisions(run)


def family_lds_cost():
    """`GPC(cost_fn=...)` (deluca/agents/_gpc.py:52,76,125): the surrogate loss the controller learns from is a caller's
    function of the counterfactual final state and action.  Run with a diagonal quadratic
    `cost_fn(x, u) = sum(q x^2) + sum(r u^2)` -- the family the fused kernels take as `cost_q` / `cost_r` -- on recorded
    disturbances, stepped like `family_lds`."""
    import jax.numpy as jnp
    from deluca.core import Disturbance
    from deluca.envs._lds import LDS
    from deluca.agents._gpc import GPC, quad_loss

    class Replay(Disturbance):
        def init(self, d):
            self.d = d

        def __call__(self, t, key):
            return jnp.array(self.W[t].reshape(self.d, 1))

    cases = [("h3", 10, 3, 3, 10, 1e-4, True, 60), ("h10", 10, 3, 10, 10, 1e-4, True, 50), ("d4", 4, 2, 5, 7, 2e-5, False, 50)]

    def run():
        out = {}
        rng = np.random.default_rng(77)
        N = 3
        for name, d, m, H, HH, lr, decay, T in cases:
            A, B = _lds_systems(rng, N, d, m)
            W = 0.5 * rng.standard_normal((T, N, d))
            q, r = rng.uniform(0.3, 3.0, d), rng.uniform(0.3, 3.0, m)
            qc, rc = jnp.array(q.reshape(d, 1)), jnp.array(r.reshape(m, 1))
            cost_fn = lambda x, u: jnp.sum(qc * x * x) + jnp.sum(rc * u * u)
            costs = np.zeros((T, N)); X = np.zeros((T + 1, N, d)); U = np.zeros((T, N, m))
            Ks = np.zeros((N, m, d)); Mf = np.zeros((N, H, m, d)); hist = np.zeros((N, H + HH, d))
            for n in range(N):
                dist = Replay()
                dist.W = W[:, n]
                env = LDS()
                x = env.init(d_in=m, d_hidden=d, d_out=d, A=A[n], B=B[n], C=np.eye(d), disturbance=dist)
                agent = GPC(jnp.array(A[n]), jnp.array(B[n]), cost_fn=cost_fn, H=H, HH=HH, lr_scale=lr, decay=decay)
                Ks[n] = npf(agent.K)
                for t in range(T):
                    X[t, n] = npf(x).ravel()
                    u = agent(x)
                    costs[t, n] = float(quad_loss(x, u))             # the realised (unweighted) cost, as in family_lds
                    U[t, n] = npf(u).ravel()
                    env(u)
                    x = env.x
                X[T, n] = npf(x).ravel()
                Mf[n], hist[n] = npf(agent.M), npf(agent.noise_history)[:, :, 0]
            out.update({f"in_{name}_A": A, f"in_{name}_B": B, f"in_{name}_W": W, f"in_{name}_q": q, f"in_{name}_r": r,
                        f"in_{name}_cfg": np.array([d, m, H, HH, lr, float(decay), T]),
                        f"out_{name}_K": Ks, f"out_{name}_costs": costs, f"out_{name}_X": X, f"out_{name}_U": U,
                        f"out_{name}_M": Mf, f"out_{name}_hist": hist})
        return out

    return both_precisions(run)


FAMILIES = {"lds_cost": family_lds_cost, "lung_train": family_lung_train, "lung_steps": family_lung_steps, "drivers": family_drivers, "of": family_of, "lds": family_lds, "lung": family_lung, "classic": family_classic, "igpc": family_igpc}


def main():
    loader.install(x64=True)
    names = sys.argv[1:] or sorted(FAMILIES)
    for name in names:
        out = FAMILIES[name]()
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
