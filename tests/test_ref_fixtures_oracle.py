"""The CPU oracle against fixtures produced by EXECUTING THE REFERENCE'S OWN FILES (tests/golden/ref_*.npz, made by
tests/golden/make_ref_fixtures.py on the JAX-semantics shim in tests/refshim).  This is what pins `oracle/` to the
reference: same recorded inputs -> same numbers (float64: to rounding; the reference's float32 run is compared too).
"""
import os

import numpy as np
import pytest

from oracle import lds as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, f"ref_{name}.npz"))


GPC_CASES = ["h3", "h10", "h3hh2", "d4", "default_lr", "clamp", "ref_defaults"]


@pytest.mark.parametrize("case", GPC_CASES)
def test_gpc_oracle_equals_reference_file(case):
    """`deluca/agents/_gpc.py` + `deluca/envs/_lds.py` executed literally vs oracle/lds.py (float64: <= 1e-9 relative,
    i.e. rounding-order differences only -- including the `dynamic_slice` clamping case HH < H-1)."""
    g = _load("lds")
    d, m, H, HH, lr, decay, T = g[f"in_gpc_{case}_cfg"]
    d, m, H, HH, T = int(d), int(m), int(H), int(HH), int(T)
    A, B, W = g[f"in_gpc_{case}_A"], g[f"in_gpc_{case}_B"], g[f"in_gpc_{case}_W"]
    K = g[f"out_gpc_{case}_K"]
    np.testing.assert_allclose(O.lqr_gain(A, B, dtype=np.float64), K, rtol=1e-9, atol=1e-12)      # scipy DARE both sides
    ref = O.rollout_lds(A, B, K, np.zeros((A.shape[0], d)), W, "gpc", H=H, HH=HH, lr_scale=lr, decay=bool(decay),
                        dtype=np.float64)
    scale = np.abs(g[f"out_gpc_{case}_costs"]).max()
    np.testing.assert_allclose(ref["costs"], g[f"out_gpc_{case}_costs"], rtol=1e-9, atol=1e-12 * scale)
    np.testing.assert_allclose(ref["X"], g[f"out_gpc_{case}_X"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ref["state"]["M"], g[f"out_gpc_{case}_M"], rtol=1e-8, atol=1e-15)
    np.testing.assert_allclose(ref["state"]["hist"], g[f"out_gpc_{case}_hist"], rtol=1e-9, atol=1e-11)
    # the reference's own float32 run deviates from its float64 run by about as much as the oracle's float32 does
    ref32 = O.rollout_lds(A, B, K, np.zeros((A.shape[0], d)), W, "gpc", H=H, HH=HH, lr_scale=lr, decay=bool(decay),
                          dtype=np.float32)
    e_ref = np.abs(g[f"out_gpc_{case}_costs_f32"] - g[f"out_gpc_{case}_costs"]).max()
    e_ora = np.abs(ref32["costs"] - g[f"out_gpc_{case}_costs"]).max()
    assert e_ora < 20 * e_ref + 1e-6 * scale and e_ref < 1e-4 * scale


def test_c_oracle_equals_reference_file():
    """oracle/lds_c.c (the fp64 checker of the benchmarked horizons and the CPU baseline) on the same fixtures."""
    from oracle.lds_c import gpc_rollout_c
    g = _load("lds")
    for case in ("h3", "h10", "h3hh2", "d4", "ref_defaults"):        # (the C restatement requires HH >= H-1)
        d, m, H, HH, lr, decay, T = g[f"in_gpc_{case}_cfg"]
        A, B, W, K = (g[f"in_gpc_{case}_A"], g[f"in_gpc_{case}_B"], g[f"in_gpc_{case}_W"], g[f"out_gpc_{case}_K"])
        costs, xf = gpc_rollout_c(A, B, K, np.zeros((A.shape[0], int(d))), W, int(H), int(HH), lr, bool(decay),
                                  np.float64, 1)
        scale = np.abs(g[f"out_gpc_{case}_costs"]).max()
        np.testing.assert_allclose(costs, g[f"out_gpc_{case}_costs"], rtol=1e-9, atol=1e-12 * scale)
        np.testing.assert_allclose(xf, g[f"out_gpc_{case}_X"][-1], rtol=1e-9, atol=1e-11)


def test_lqr_oracle_equals_reference_file():
    g = _load("lds")
    A, B, W, Q, R = (g[k] for k in ("in_lqr_A", "in_lqr_B", "in_lqr_W", "in_lqr_Q", "in_lqr_R"))
    K = O.lqr_gain(A, B, Q, R, dtype=np.float64)
    np.testing.assert_allclose(K, g["out_lqr_K"], rtol=1e-9, atol=1e-12)
    ref = O.rollout_lds(A, B, K, np.zeros(A.shape[:2]), W, "lqr", dtype=np.float64)
    np.testing.assert_allclose(ref["costs"], g["out_lqr_costs"], rtol=1e-10)
    np.testing.assert_allclose(ref["X"], g["out_lqr_X"], rtol=1e-10, atol=1e-12)


def test_bpc_oracle_equals_reference_file():
    """`deluca/agents/_bpc.py` (with the `.at`-on-NumPy waiver recorded in make_ref_fixtures.py) vs oracle/lds.py, fed
    the perturbations the reference drew.  Odd envs ran with `decay=True`."""
    g = _load("lds")
    d, m, H, lr, delta, T = g["in_bpc_cfg"]
    d, m, H, T = int(d), int(m), int(H), int(T)
    A, B, W, K = g["in_bpc_A"], g["in_bpc_B"], g["in_bpc_W"], g["out_bpc_K"]
    for n in range(A.shape[0]):
        s = slice(n, n + 1)
        st = O.bpc_init(1, d, m, H, g["in_bpc_M0"][s] / delta, g["in_bpc_eps0"][s], delta=delta, dtype=np.float64)
        np.testing.assert_allclose(st["M"], g["in_bpc_M0"][s], rtol=1e-12)
        ref = O.rollout_lds(A[s], B[s], K[s], np.zeros((1, d)), W[:, s], "bpc", H=H, lr_scale=lr, decay=(n % 2 == 1),
                            state=st, bpc_eps=g["in_bpc_eps_new"][:, s], bpc_delta=delta, dtype=np.float64)
        np.testing.assert_allclose(ref["costs"], g["out_bpc_costs"][:, s], rtol=1e-9)
        np.testing.assert_allclose(ref["X"], g["out_bpc_X"][:, s], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(ref["state"]["M"], g["out_bpc_M"][s], rtol=1e-8, atol=1e-14)


def test_open_lds_oracle_equals_reference_file():
    g = _load("lds")
    Y, x = O.open_rollout(g["in_open_A"], g["in_open_B"], g["in_open_C"], np.zeros(g["in_open_A"].shape[:2]),
                          g["in_open_U"], g["in_open_W"])
    np.testing.assert_allclose(Y, g["out_open_Y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(x, g["out_open_x"], rtol=1e-10, atol=1e-12)


# ------------------------------------------------------------------------------------------ lung
@pytest.mark.parametrize("sim,kw", [("balloon", {}), ("delay", {}), ("balloon_leak", dict(leak=True))])
@pytest.mark.parametrize("T", [29, 150])
def test_lung_oracle_equals_reference_files(sim, kw, T):
    """`run_controller_scan` + lung `PID` + `Expiratory` + `BalloonLung` / `DelayLung` executed literally vs
    oracle/lung.py.  float64: rounding only at the reference's horizon T = 29 (1e-11) and <= 1e-4 at T = 150, where
    BalloonLung chatters around `pressure > peep_valve` and amplifies rounding; float32 (the reference's default
    precision, incl. the `3 - 1e-8 == 3` waveform knot): 1e-4 relative at T = 29."""
    from oracle import lung as L
    g = _load("lung")
    gains, pips = g["in_gains"], g["in_pips"]
    K = np.repeat(gains, len(pips), axis=0)
    pip = np.tile(pips, len(gains))
    for dtype, suf in ((np.float64, ""), (np.float32, "_f32")):
        r = L.run_controller_scan(K, pip, 5.0, T=T, sim="delay" if sim == "delay" else "balloon", dtype=dtype,
                                  sim_kwargs=kw)
        for k in ("timestamp", "pressure", "u_in", "u_out", "target", "flow"):
            ref = g[f"out_{sim}_T{T}_{k}{suf}"].reshape(-1, T).T
            got = r["timeseries"][k]
            scale = max(np.abs(ref).max(), 1.0)
            if dtype == np.float64:
                tol = 1e-11 if T == 29 else 1e-4
            else:
                tol = 1e-4 if T == 29 else None       # long float32 episodes diverge at branch flips (see the GPU test)
            if tol is not None:
                assert np.abs(got - ref).max() <= tol * scale, (k, suf, np.abs(got - ref).max())
    # test_controller's score (mean |pressure - target| over the pips), float64
    if T == 29:
        r = L.run_controller_scan(K, pip, 5.0, T=29, sim="delay" if sim == "delay" else "balloon", dtype=np.float64,
                                  sim_kwargs=kw)
        score = L.test_controller_score(r["timeseries"]).reshape(len(gains), len(pips)).mean(1)
        np.testing.assert_allclose(score, g[f"out_{sim}_score"], rtol=1e-10)


def test_lung_helpers_equal_reference_files():
    from oracle import lung as L
    g = _load("lung")
    tt = g["in_wave_t"]
    for dtype, suf in ((np.float64, ""), (np.float32, "_f32")):
        kx, kf, period = L.waveform_knots(35.0, 5.0, dtype=dtype)
        at = np.array([L.waveform_at(np.array([t], dtype), kx, kf, period)[0] for t in tt])
        np.testing.assert_allclose(at, g[f"out_wave_at{suf}"], rtol=1e-6 if suf else 1e-12, atol=1e-5 if suf else 1e-12)
        is_in = np.array([float(L.waveform_is_in(np.array(t, dtype), period)) for t in tt])
        np.testing.assert_array_equal(is_in, g[f"out_wave_is_in{suf}"])
    # the float32 waveform differs QUALITATIVELY from the float64 one on the exhale plateau (SURVEY A-9): the
    # reference's own files show it
    i = int(np.argmin(np.abs(tt - 2.0)))
    assert abs(g["out_wave_at"][i] - 5.0) < 1e-6 and abs(g["out_wave_at_f32"][i] - 15.0) < 1e-3
    dec = np.array([L.controller_decay(t) for t in tt])
    ref = g["out_decay_f32"]
    fin = np.isfinite(ref)
    assert (np.isfinite(dec) == fin).all()
    np.testing.assert_allclose(dec[fin], ref[fin], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(g["out_proper_time"], [0.0, 1.25])
    st = dict(P=np.zeros(1), I=np.zeros(1), D=np.zeros(1))
    kp, ki, kd = g["in_pid_gains"]
    u = [L.generic_pid_call(st, kp, ki, kd, np.array([e]))[0] for e in g["in_pid_err"]]
    np.testing.assert_allclose(u, g["out_pid_u"], rtol=1e-12)
    np.testing.assert_allclose([st["P"][0], st["I"][0], st["D"][0]], g["out_pid_state"], rtol=1e-12)


# ------------------------------------------------------------------------------------------ classic envs, deluca/igpc
def _oracle_envs():
    from oracle import igpc as G
    return G, {"pendulum": G.Pendulum(), "quad": G.PlanarQuadrotor(), "quad_wind": G.PlanarQuadrotor(wind=0.3),
               "quad_const": G.PlanarQuadrotor(wind=0.2, wind_kind="constant"), "reacher": G.Reacher()}


def test_classic_env_steps_equal_reference_files():
    G, envs = _oracle_envs()
    c = _load("classic")
    assert c["out_cartpole_literal_raises_typeerror"] == 1.0      # A-7: Cartpole.__call__ is a TypeError as written
    for name, env in envs.items():
        x0, U = c[f"in_{name}_x0"], c[f"in_{name}_U"]
        for dtype, suf, tol in ((np.float64, "", 1e-13), (np.float32, "_f32", 2e-5)):
            X = c[f"out_{name}_X{suf}"]
            x = x0.astype(dtype)
            for t in range(U.shape[0]):
                x = env.step(x, U[t].astype(dtype))
                assert np.abs(x - X[t + 1]).max() <= tol * max(1.0, np.abs(X).max()), (name, suf, t)


@pytest.mark.parametrize("name", ["quad", "quadc", "pend"])
def test_igpc_oracle_equals_reference_files(name):
    """`deluca/igpc/{rollout,lqr_solver,ilqr,igpc,ilc}.py` executed literally (autodiff Jacobians / Hessians, host NumPy
    Riccati pass, Python accept / reject) vs oracle/igpc.py (analytic Jacobians, masks): float64, rounding only."""
    from oracle import igpc as G
    g = _load("igpc")
    H, T = g[f"out_{name}_X"].shape[0] - 1, 4
    if name == "pend":
        et, es = G.Pendulum(H=H, g=9.0), G.Pendulum(H=H)
        cost = G.QuadCost(es.goal_state, 1.0, 0.1)
    else:
        et = G.PlanarQuadrotor(H=H, wind=0.1, wind_kind="constant" if name == "quadc" else "dissipative")
        es = G.PlanarQuadrotor(H=H)
        cost = G.QuadCost(es.goal_state, 1.0, 0.1, goal_action=es.goal_action)

    def chk(a, key, tol=1e-10):
        b = g[f"out_{name}_{key}"]
        assert np.abs(np.asarray(a) - b).max() <= tol * max(1.0, np.abs(b).max()), key

    U0 = g[f"in_{name}_U0"][None]
    X, U, c0 = G.rollout(es, cost, U0)
    chk(X[0], "X"); chk(c0[0], "c")
    D, F, C = G.compute_ders(es, cost, X, U)
    for a, key in ((D, "D"), (F[0], "Fx"), (F[1], "Fu"), (C[0], "cx"), (C[1], "cu"), (C[2], "cxx"), (C[3], "cuu")):
        chk(a[0], key)
    k, K = G.lqr_backward(F, C)
    chk(k[0], "k"); chk(K[0], "K")
    X1, U1, c1 = G.rollout(es, cost, U, k, K, X, 0.5)
    chk(X1[0], "X1"); chk(U1[0], "U1"); chk(c1[0], "c1")
    Xi, Ui, ki, Ki, ci, _ = G.ilqr(es, cost, U0, T)
    chk(Xi[0], "ilqr_X"); chk(Ui[0], "ilqr_U"); chk(ci[0], "ilqr_c")
    Dg, Fg, Cg = G.compute_ders(es, cost, Xi, Ui)
    kg, Kg = G.lqr_backward(Fg, Cg)
    for w_is in ("de", "dede", "delin"):
        Xg, Ug, cg = G.gpc_rollout(et, es, cost, Ui, kg, Kg, Xi, Dg, Fg, 0.6, w_is, 0.05)
        chk(Xg[0], f"gpc_{w_is}_X"); chk(Ug[0], f"gpc_{w_is}_U"); chk(cg[0], f"gpc_{w_is}_c")
    Xc, Uc, _, _, cc, _ = G.igpc_closed(et, es, cost, U0, T, lr=0.05)
    chk(Xc[0], "igpc_X"); chk(Uc[0], "igpc_U"); chk(cc[0], "igpc_c")
    Xl, Ul, _, _, cl, _ = G.ilc_closed(et, es, cost, U0, T)
    chk(Xl[0], "ilc_X"); chk(Ul[0], "ilc_U"); chk(cl[0], "ilc_c")


# ------------------------------------------------------------------------------------------ output-feedback agents
def test_output_feedback_oracle_equals_reference_files():
    """`GRC` / `SFC` / `DSC` / `DRC` / `LQG` executed literally vs oracle/of.py (filter-bank form, hand adjoints)."""
    from oracle import of as F
    g = _load("of")
    A, B, C, W, x0 = (g[k] for k in ("in_A", "in_B", "in_C", "in_W", "in_x0"))
    N = A.shape[0]

    def chk(a, b, tol=1e-11):
        assert np.abs(np.asarray(a) - b).max() <= tol * max(1.0, np.abs(b).max())

    for name, kw in (("grc", dict(m=5, lr=0.002, decay=True, R_M=10.0)), ("grc_proj", dict(m=5, lr=0.05, R_M=2.0, decay=False))):
        Phi, _ = F.filter_bank("grc", 5)
        r = F.rollout_of(A, B, C, x0, W, "fgrc", dtype=np.float64, Theta0=g[f"in_{name}_M"], Phi=Phi, **kw)
        chk(r["costs"], g[f"out_{name}_costs"]); chk(r["U"], g[f"out_{name}_U"]); chk(r["state"]["Theta"], g[f"out_{name}_Mf"])
        chk(r["x_final"], g[f"out_{name}_x"])
    chk(F.spectral_filters(8, 4), g["in_sfc_filters"][0])
    Phi, _ = F.filter_bank("sfc", 8, 4)
    Theta0 = np.concatenate([g["in_sfc_M_0"][:, None], g["in_sfc_M"]], axis=1)
    r = F.rollout_of(A, B, C, x0, W, "fgrc", dtype=np.float64, Theta0=Theta0, Phi=Phi, m=8, lr=0.002)
    chk(r["costs"], g["out_sfc_costs"]); chk(r["U"], g["out_sfc_U"])
    chk(r["state"]["Theta"][:, 1:], g["out_sfc_Mf"]); chk(r["state"]["Theta"][:, 0], g["out_sfc_M_0f"])
    Phi, _ = F.filter_bank("dsc", 6, 3)
    pack = lambda suf: np.stack([F.dsc_pack_theta(*(g[f"{'out' if suf else 'in'}_dsc_M_{i}{suf}"][e] for i in range(4)))
                                 for e in range(N)])
    r = F.rollout_of(A, B, C, x0, W, "fgrc", dtype=np.float64, Theta0=pack(""), Phi=Phi, m=6, lr=0.002)
    chk(r["costs"], g["out_dsc_costs"]); chk(r["U"], g["out_dsc_U"]); chk(r["state"]["Theta"], pack("f"))
    for name, kw in (("drc", dict(decay=True, RM=1000.0)), ("drc_nodecay", dict(decay=False, RM=0.5))):
        r = F.rollout_of(A, B, C, x0, W, "drc", dtype=np.float64, M0=g[f"in_{name}_M"], K=g["in_drc_K"], h=6,
                         lr_scale=0.03, **kw)
        chk(r["costs"], g[f"out_{name}_costs"]); chk(r["U"], g[f"out_{name}_U"]); chk(r["state"]["M"], g[f"out_{name}_Mf"])
    chk(F.drc_markov(A, B, C, 6), g["in_drc_G"])
    KL = [F.lqg_gains(A[e], B[e], C[e]) for e in range(N)]
    K, Lk = np.stack([k for k, _ in KL]), np.stack([l for _, l in KL])
    chk(K, g["in_lqg_K"], 1e-9); chk(Lk, g["in_lqg_L"], 1e-9)
    r = F.rollout_of(A, B, C, x0, W, "lqg", dtype=np.float64, K=K, Lk=Lk)
    chk(r["costs"], g["out_lqg_costs"], 1e-9); chk(r["U"], g["out_lqg_U"], 1e-9); chk(r["state"]["x_hat"], g["out_lqg_x_hat"], 1e-9)


def test_gae_oracle_equals_reference_file():
    from oracle import ppo as P
    g = _load("drivers")
    adv = P.gae_advantages(g["in_gae_rewards"], g["in_gae_masks"], g["in_gae_values"], 0.99, 0.95)
    np.testing.assert_allclose(adv, g["out_gae_adv"], rtol=1e-13, atol=1e-13)


# ------------------------------------------------------------------------------------------ the fixtures reproduce
def test_fixture_generator_reproduces_committed_files():
    """Where the reference tree is present (the build container), re-running the generator on the shim gives the
    committed numbers back; on the GPU box (no /root/reference) this is skipped and the committed files stand."""
    import subprocess
    import sys
    import tempfile
    ref_root = os.environ.get("DELUCA_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "deluca")):
        pytest.skip("reference tree not present")
    code = ("import sys, numpy as np; sys.argv=['x']; sys.path.insert(0, %r); import make_ref_fixtures as m; "
            "m.loader.install(x64=True); out = m.FAMILIES['drivers'](); np.savez(sys.stdout.buffer, **out)") % GOLD
    raw = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True).stdout
    with tempfile.NamedTemporaryFile(suffix=".npz") as f:
        f.write(raw)
        f.flush()
        new = np.load(f.name)
        old = _load("drivers")
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            np.testing.assert_allclose(new[k], old[k], rtol=1e-12, atol=1e-12, err_msg=k)


def test_env_agent_oracle_equals_reference_apg_rollout():
    """`apg_parallel_rollouts` / `apg_rollout` / `jax.value_and_grad(apg_loss)` (deluca/training/apg.py:39-122) run
    literally on Pendulum x generic PID: losses, actions, the stacked PID states, the mean loss and its gradient
    w.r.t. (K_P, K_I, K_D)."""
    from oracle import env_agent as E, igpc as G
    g = _load("drivers")
    env = G.Pendulum()
    N, T = g["out_apg_losses"].shape
    r = E.rollout(env, "pid", g["in_apg_gains"], g["in_apg_arr0"], T, [1, 1, 1], env.goal_state, [0.1], [0.0],
                  pid_decay=0.03 / (0.03 + 0.5), want_grad=True)
    tm = lambda a: a.transpose(1, 0, 2)
    np.testing.assert_allclose(r["losses"].T, g["out_apg_losses"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(tm(r["actions"]), g["out_apg_actions"], rtol=1e-12, atol=1e-12)
    for k in "PID":
        np.testing.assert_allclose(tm(r[k]), g[f"out_apg_{k}"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r["losses"][:, 0], g["out_apg_single_losses"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r["losses"].mean(), g["out_apg_value"], rtol=1e-12)
    np.testing.assert_allclose(r["grad"].sum(0) / (N * T), g["out_apg_grad"], rtol=1e-9)


def test_env_agent_oracle_tangents_equal_finite_differences():
    """Forward-mode tangents of the oracle against central differences, for every env and both agents."""
    from oracle import env_agent as E, igpc as G
    rng = np.random.default_rng(3)
    N, T = 3, 15
    for env, x0, gu in ((G.Pendulum(), None, [0.0]), (G.Cartpole(), None, [0.0]),
                        (G.PlanarQuadrotor(wind=0.1), None, [0.4905, 0.4905]), (G.Reacher(), None, [0.0, 0.0])):
        d, m = env.d, env.m
        x0 = env.init(N, np.float64) + 0.05 * rng.standard_normal((N, d))
        goal = np.asarray(getattr(env, "goal_state", np.zeros(d)), np.float64)
        for agent, gains in (("pid", np.array([0.05, 0.02, 0.01])), ("linear", 0.03 * rng.standard_normal((m, d)))):
            kw = dict(q=np.ones(d), goal=goal, r=0.1 * np.ones(m), goal_u=gu, pid_decay=0.0566)
            r = E.rollout(env, agent, gains, x0, T, want_grad=True, **kw)
            flat = gains.reshape(-1)
            for k in range(flat.size):
                h = 1e-6
                gp, gm = flat.copy(), flat.copy()
                gp[k] += h; gm[k] -= h
                lp = E.rollout(env, agent, gp.reshape(gains.shape), x0, T, **kw)["losses"].sum(0)
                lm = E.rollout(env, agent, gm.reshape(gains.shape), x0, T, **kw)["losses"].sum(0)
                fd = (lp - lm) / (2 * h)
                np.testing.assert_allclose(r["grad"][:, k], fd, rtol=2e-5, atol=1e-6 * np.abs(fd).max())


LUNG_TRAIN_CASES = [("balloon", "balloon", 29, None), ("balloon", "balloon_leak", 29, dict(leak=True)),
                    ("balloon", "balloon", 60, None), ("delay", "delay", 29, None),
                    ("delay", "delay1", 130, dict(delay=1)), ("delay", "delay", 940, None)]


@pytest.mark.parametrize("sim,name,T,kw", LUNG_TRAIN_CASES)
def test_lung_train_oracle_equals_reference_files(sim, name, T, kw):
    """`train_controller.rollout` and `jax.value_and_grad(rollout_parallel)` (train_controller.py:33-92,151-153)
    executed literally vs oracle/lung_grad.py: per-breath losses and d loss / d (K_P, K_I, K_D), float64 to rounding
    and the reference's float32 run at 1e-4 (both walk the same float32 clocks, so the DelayLung threshold crossings
    agree; the float64 run sees `3 - 1e-8` as a separate waveform knot and is a different trajectory after 1.5 s)."""
    import torch
    from oracle import lung_grad as OG
    g = _load("lung_train")
    gains, pips, noise = g["in_gains"], g["in_pips"], float(g["in_noise_value"])
    K, P = np.repeat(gains, len(pips), 0), np.tile(pips, len(gains))
    tt = g[f"in_tt_T{T}"]
    seen = 0
    for loss in ("l1", "l2"):
        for un in (0.0, 1.0):
            tag = f"{name}_T{T}_{loss}_n{int(un)}"
            if f"out_{tag}_loss" not in g.files:
                continue
            seen += 1
            for dt, suf, tol in ((torch.float64, "", 1e-8), (torch.float32, "_f32", 1e-4)):
                lo, go = OG.value_and_grad(K, P, 5.0, tt, loss=loss, use_noise=un, noise=noise, dtype=dt, sim=sim,
                                           sim_kwargs=kw)
                rl, rg = g[f"out_{tag}_loss{suf}"].reshape(-1), g[f"out_{tag}_loss_grad{suf}"].reshape(-1, 3)
                np.testing.assert_allclose(lo, rl, rtol=tol, atol=tol)
                assert np.abs(go - rg).max() <= tol * max(np.abs(rg).max(), 1.0), (tag, suf, np.abs(go - rg).max())
                # rollout_parallel: the mean over the pips, value and gradient
                np.testing.assert_allclose(lo.reshape(len(gains), -1).mean(1), g[f"out_{tag}_value{suf}"], rtol=tol)
                gm = go.reshape(len(gains), -1, 3).mean(1)
                assert np.abs(gm - g[f"out_{tag}_grad{suf}"]).max() <= tol * max(np.abs(rg).max(), 1.0)
    assert seen >= 1
    if name == "delay" and T == 29:       # the action cannot reach the pressure before 25 s: exact zeros
        assert not g["out_delay_T29_l1_n0_grad"].any() and not g["out_delay_T29_l1_n0_grad_f32"].any()


@pytest.mark.parametrize("case", ["h3", "h10", "d4"])
def test_gpc_weighted_cost_fn_oracle_equals_reference_file(case):
    """`GPC(cost_fn=...)` with a diagonal quadratic cost (deluca/agents/_gpc.py:52,76,125) executed literally
    (fixture family `lds_cost`) vs oracle/lds.py's `cost_q` / `cost_r`: float64 to rounding."""
    g = _load("lds_cost")
    d, m, H, HH, lr, decay, T = g[f"in_{case}_cfg"]
    d, m, H, HH, T = int(d), int(m), int(H), int(HH), int(T)
    A, B, W, K = g[f"in_{case}_A"], g[f"in_{case}_B"], g[f"in_{case}_W"], g[f"out_{case}_K"]
    ref = O.rollout_lds(A, B, K, np.zeros((A.shape[0], d)), W, "gpc", H=H, HH=HH, lr_scale=lr, decay=bool(decay),
                        dtype=np.float64, cost_q=g[f"in_{case}_q"], cost_r=g[f"in_{case}_r"])
    scale = np.abs(g[f"out_{case}_costs"]).max()
    np.testing.assert_allclose(ref["costs"], g[f"out_{case}_costs"], rtol=1e-9, atol=1e-12 * scale)
    np.testing.assert_allclose(ref["X"], g[f"out_{case}_X"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ref["state"]["M"], g[f"out_{case}_M"], rtol=1e-8, atol=1e-15)
    # and the weights matter: the unweighted oracle ends at a different policy tensor
    plain = O.rollout_lds(A, B, K, np.zeros((A.shape[0], d)), W, "gpc", H=H, HH=HH, lr_scale=lr, decay=bool(decay),
                          dtype=np.float64)
    assert np.abs(plain["state"]["M"] - g[f"out_{case}_M"]).max() > 1e-3 * np.abs(g[f"out_{case}_M"]).max()
