"""Lung simulators + PID / Expiratory controllers restated on the CPU (test infrastructure).

Pinned: ``BreathWaveform.at / is_in / elapsed``, ``LungEnv.time``, ``proper_time`` and
``Controller.decay`` against the reference's own known answers
(``deluca/tests/lung/core_test.py:43-68,73-84,97-99,119-127``; tests/test_oracle_lung.py).
Pinned as well, to the reference's own files run here: ``BalloonLung``, ``DelayLung``, lung ``PID``,
``Expiratory``, ``run_controller_scan`` and ``test_controller`` executed literally on tests/refshim
(``tests/golden/ref_lung.npz``: whole episodes; ``ref_lung_steps.npz``: every visited state, one step at a time);
tests/test_ref_fixtures_oracle.py holds this module to them (1e-11 at the reference's horizon T = 29 in float64).  The
reference itself only smoke-tests them (``deluca/tests/lung/utils/scripts/test_controller_test.py:189-208``).

Everything is batched over a leading env axis N and computed in ``dtype`` (float32 = default
JAX; every Python-float constant of the reference is weakly typed and therefore rounded to
float32 at its point of use, which is what ``c()`` does below).
"""
import numpy as np

DEFAULT_XP = [0.0, 1.0, 1.5, 3.0 - 1e-8, 3.0]       # deluca/lung/core.py:26
DEFAULT_DT = 0.03                                    # deluca/lung/core.py:27


# ----------------------------------------------------------------------------- waveform
def waveform_knots(pip, peep, bpm=20, xp=None, fp=None, dtype=np.float32):
    """Knot table of ``jnp.interp(t, xp, fp, period=60/bpm)`` as ``BreathWaveform.at`` builds it
    (``deluca/lung/core.py:47-60``): xp is reduced mod period IN ``dtype`` (float32 turns
    3 - 1e-8 into 3.0, SURVEY A-9), stably sorted, and padded with one periodic knot each side.

    pip, peep: [N] (or scalars).  Returns kx[nk], kf[N, nk], period."""
    pip = np.atleast_1d(np.asarray(pip, dtype=dtype))
    peep = np.atleast_1d(np.asarray(peep, dtype=dtype))
    N = max(pip.shape[0], peep.shape[0])
    pip, peep = np.broadcast_to(pip, (N,)), np.broadcast_to(peep, (N,))
    period = dtype(60 / bpm)
    xp = np.asarray(DEFAULT_XP if xp is None else xp).astype(dtype)
    if fp is None:
        fp = np.stack([pip, pip, peep, peep, pip], axis=1)                  # core.py:48-49
    else:
        fp = np.broadcast_to(np.asarray(fp, dtype=dtype), (N, len(xp)))
    xpm = np.mod(xp, period)
    order = np.argsort(xpm, kind="stable")
    xs, fs = xpm[order], fp[:, order]
    kx = np.concatenate([xs[-1:] - period, xs, xs[:1] + period]).astype(dtype)
    kf = np.concatenate([fs[:, -1:], fs, fs[:, :1]], axis=1).astype(dtype)
    return kx, kf, period


def waveform_at(t, kx, kf, period):
    """``BreathWaveform.at`` (``core.py:54-60``) with jnp.interp's arithmetic:
    i = clip(searchsorted(xp, x, 'right'), 1, n-1); f = fp[i-1] + (x - xp[i-1]) / dx * df,
    and fp[i-1] where dx == 0."""
    dtype = kf.dtype.type
    x = np.mod(np.asarray(t, dtype=dtype), dtype(period))
    i = np.clip(np.searchsorted(kx, x, side="right"), 1, len(kx) - 1)
    n = np.arange(kf.shape[0])
    f0, f1 = kf[n, i - 1], kf[n, i]
    dx = kx[i] - kx[i - 1]
    delta = x - kx[i - 1]
    with np.errstate(all="ignore"):
        val = f0 + (delta / np.where(dx == 0, dtype(1), dx)) * (f1 - f0)
    return np.where(dx == 0, f0, val).astype(dtype)


def waveform_is_in(t, period, xp_in=1.0):
    """``is_in`` (``core.py:62-66``): elapsed(t) <= xp[in_bounds[1]]."""
    t = np.asarray(t)
    return np.mod(t, t.dtype.type(period)) <= t.dtype.type(xp_in)


def proper_time(t):
    """``proper_time`` (``core.py:109-110``)."""
    t = np.asarray(t)
    return np.where(np.isinf(t), t.dtype.type(0), t)


def controller_decay(t, period=3.0, xp=DEFAULT_XP):
    """``Controller.decay`` (``core.py:127-151``), scalar."""
    elapsed = np.float32(t) % np.float32(period)
    if elapsed < np.float32(xp[1]):
        return float("inf")
    if elapsed < np.float32(xp[2]):
        return 0.0
    return np.float32(5) * (np.float32(1) - np.exp(np.float32(5) * (np.float32(xp[2]) - elapsed)))


def lung_env_time(steps, dt=DEFAULT_DT, dtype=np.float32):
    """``LungEnv.time`` (``core.py:75-76``)."""
    return dtype(dt) * np.asarray(steps, dtype=dtype)


# ----------------------------------------------------------------------------- controllers
def pid_init(N, dtype=np.float32):
    """``PIDControllerState`` defaults (``controllers/_pid.py:29-36``)."""
    z = np.zeros(N, dtype=dtype)
    return dict(p=z.copy(), i=z.copy(), d=z.copy(), time=np.full(N, np.inf, dtype=dtype),
                steps=np.zeros(N, dtype=np.int32), dt=np.full(N, DEFAULT_DT, dtype=dtype))


def pid_call(st, K, pressure, t, kx, kf, period, RC=0.5, dt=0.03):
    """Lung ``PID.__call__`` (``controllers/_pid.py:77-102``).  K[N,3] = the bias-free Dense(1)
    kernel (``:39-44``); returns u_in[N] in [0, 100]."""
    dtype = pressure.dtype
    c = dtype.type
    target = waveform_at(t, kx, kf, period)
    err = target - pressure
    decay = c(dt / (dt + RC))
    p, i, d = st["p"], st["i"], st["d"]
    next_p = err
    next_i = i + decay * (err - i)
    next_d = d + decay * (err - p - d)
    u_in = next_p * K[:, 0] + next_i * K[:, 1] + next_d * K[:, 2]
    u_in = np.clip(u_in, c(0), c(100)).astype(dtype)
    new_dt = np.maximum(c(DEFAULT_DT), t - proper_time(st["time"]))
    st.update(p=next_p, i=next_i, d=next_d, time=t.copy(), steps=st["steps"] + 1, dt=new_dt.astype(dtype))
    return u_in


def expiratory_init(N, dtype=np.float32):
    """``ControllerState`` defaults (``core.py:103-106``)."""
    return dict(time=np.full(N, np.inf, dtype=dtype), steps=np.zeros(N, dtype=np.int32),
                dt=np.full(N, DEFAULT_DT, dtype=dtype))


def expiratory_call(st, t, period, xp_in=1.0):
    """``Expiratory.__call__`` (``controllers/_expiratory.py:35-45``)."""
    c = t.dtype.type
    u_out = np.where(waveform_is_in(t, period, xp_in), c(0), c(1)).astype(t.dtype)
    new_dt = np.maximum(c(DEFAULT_DT), t - proper_time(st["time"]))
    st.update(time=t.copy(), steps=st["steps"] + 1, dt=new_dt.astype(t.dtype))
    return u_out


def generic_pid_call(st, K_P, K_I, K_D, obs, RC=0.5, dt=0.03):
    """Generic ``PID.__call__`` (``deluca/agents/_pid.py:34-48``): obs is the error."""
    c = obs.dtype.type
    decay = c(dt / (dt + RC))
    P = obs
    I = st["I"] + decay * (obs - st["I"])
    D = st["D"] + decay * (obs - st["P"] - st["D"])
    st.update(P=P, I=I, D=D)
    return K_P * P + K_I * I + K_D * D


# ----------------------------------------------------------------------------- simulators
def prop_valve(x):
    """``PropValve`` (``envs/_balloon_lung.py:39-43``)."""
    c = x.dtype.type
    y = c(3.0) * x
    f = c(1.0) * (np.tanh(c(0.03) * (y - c(130))) + c(1.0))
    return np.clip(f, c(0.0), c(1.72))


def balloon_reset(N, kx, kf, period, min_volume=1.5, reset_normalized_peep=0.0, dtype=np.float32):
    """``BalloonLung.reset`` (``_balloon_lung.py:87-98``)."""
    z = np.zeros(N, dtype=dtype)
    st = dict(steps=z.copy(), time=z.copy(), volume=np.full(N, min_volume, dtype=dtype),
              pressure=np.full(N, reset_normalized_peep, dtype=dtype), target=waveform_at(z, kx, kf, period))
    obs = dict(pressure=st["pressure"].copy(), time=z.copy())
    return st, obs


def balloon_call(st, u_in, u_out, kx, kf, period, leak=False, peep_valve=5.0, PC=40.0, P0=0.0, R=15.0,
                 dt=0.03, min_volume=1.5):
    """``BalloonLung.__call__`` (``_balloon_lung.py:100-146``).  Mutates ``st``; returns obs."""
    dtype = st["volume"].dtype
    c = dtype.type
    r0 = (3.0 * min_volume / (4.0 * np.pi)) ** (1.0 / 3.0)                   # python float, :76
    volume, pressure = st["volume"], st["pressure"]
    flow = np.clip(prop_valve(u_in) * c(R), c(0.0), c(2.0))                   # :113
    sol = np.clip((u_out > 0).astype(dtype), c(0.0), c(2.0))
    flow = flow - np.where(pressure > c(peep_valve), sol * c(0.05) * pressure, c(0.0))   # :114-119
    volume = volume + flow * c(dt)                                            # :121
    if leak:
        volume = volume + c(dt / (5.0 + dt)) * (c(min_volume) - volume)       # :122-129
    r = np.power(c(3.0) * volume / c(4.0 * np.pi), c(1.0 / 3.0))              # :131
    pressure = c(P0) + c(PC) * (c(1.0) - np.power(c(r0) / r, c(6.0))) / (c(r0 ** 2.0) * r)   # :132
    new_time = st["time"] + c(dt)
    st.update(steps=st["time"] + c(1), time=new_time, volume=volume.astype(dtype),
              pressure=pressure.astype(dtype), target=waveform_at(new_time, kx, kf, period))   # :135-141
    return dict(pressure=st["pressure"].copy(), time=st["time"].copy())       # :143-146


def delay_reset(N, kx, kf, period, min_volume=1.5, delay=25, dtype=np.float32):
    """``DelayLung.reset`` (``_delay_lung.py:61-73``)."""
    z = np.zeros(N, dtype=dtype)
    st = dict(steps=z.copy(), time=z.copy(), volume=np.full(N, min_volume, dtype=dtype), pressure=z.copy(),
              pipe_pressure=z.copy(), target=waveform_at(z, kx, kf, period),
              in_history=np.zeros((N, delay), dtype=dtype), out_history=np.zeros((N, delay), dtype=dtype))
    obs = dict(pressure=z.copy(), time=z.copy())
    return st, obs


def delay_call(st, u_in, u_out, kx, kf, period, min_volume=1.5, R=10, C=6, delay=25, inertia=0.995,
               control_gain=0.02, dt=0.03):
    """``DelayLung.__call__`` + ``dynamics`` (``_delay_lung.py:75-133``).  Mutates ``st``; returns
    obs, which the reference takes from the state BEFORE the dynamics update (``:128-131``)."""
    dtype = st["volume"].dtype
    c = dtype.type
    r0 = (3 * min_volume / (4 * np.pi)) ** (1 / 3)
    u_in = np.where(u_in > c(0.0), u_in, c(0.0))                              # :117-119
    st["in_history"] = np.concatenate([u_in[:, None], st["in_history"][:, :-1]], axis=1)    # :121-124
    st["out_history"] = np.concatenate([u_out[:, None], st["out_history"][:, :-1]], axis=1)
    obs = dict(pressure=st["pressure"].copy(), time=st["time"].copy())
    flow = st["pressure"] / c(R)                                              # :81
    volume = np.maximum(st["volume"] + flow * c(dt), c(min_volume))           # :82-83
    r = np.power(c(3.0) * volume / c(4.0 * np.pi), c(1.0 / 3.0))              # :85
    lung_pressure = c(C) * (c(1) - np.power(c(r0) / r, c(6))) / (c(r0 ** 2) * r)   # :86
    early = st["time"] < c(delay)                                             # :88-96 (seconds vs steps, sic)
    impulse = np.where(early, c(0.0), c(control_gain) * st["in_history"][:, 0])
    peep = np.where(early, c(0.0), st["out_history"][:, 0])
    pipe = c(inertia) * st["pipe_pressure"] + impulse                         # :98
    pressure = np.maximum(c(0), pipe - lung_pressure)                         # :99
    pipe = np.where(peep != 0, pipe * c(0.995), pipe)                         # :101-103
    new_time = st["time"] + c(dt)
    st.update(steps=st["time"] + c(1), time=new_time, volume=volume.astype(dtype),
              pressure=pressure.astype(dtype), pipe_pressure=pipe.astype(dtype),
              target=waveform_at(new_time, kx, kf, period))
    return obs


# ----------------------------------------------------------------------------- episode
def run_controller_scan(K, pip, peep, T=29, sim="balloon", bpm=20, dt=0.03, dtype=np.float32,
                        sim_kwargs=None, env_flow=0.0):
    """``run_controller_scan`` / ``loop_over_tt`` (``lung/utils/scripts/run_controller.py:111-220``)
    for N breaths: PID (gains K[N,3]) + Expiratory driving BalloonLung or DelayLung.

    Returns the reference's ``timeseries`` dict with [T, N] arrays plus the final states."""
    sim_kwargs = dict(sim_kwargs or {})
    K = np.asarray(K, dtype=dtype)
    N = K.shape[0]
    c = dtype
    kx, kf, period = waveform_knots(pip, peep, bpm=bpm, dtype=dtype)
    kf = np.broadcast_to(kf, (N, kf.shape[1]))
    if sim == "balloon":
        st, obs = balloon_reset(N, kx, kf, period, dtype=dtype,
                                **{k: v for k, v in sim_kwargs.items() if k in ("min_volume", "reset_normalized_peep")})
        env = lambda u_in, u_out: balloon_call(st, u_in, u_out, kx, kf, period, **{
            k: v for k, v in sim_kwargs.items() if k != "reset_normalized_peep"})
    else:
        st, obs = delay_reset(N, kx, kf, period, dtype=dtype,
                              **{k: v for k, v in sim_kwargs.items() if k in ("min_volume", "delay")})
        env = lambda u_in, u_out: delay_call(st, u_in, u_out, kx, kf, period, **sim_kwargs)
    cst, est = pid_init(N, dtype), expiratory_init(N, dtype)
    ts = {k: np.zeros((T, N), dtype=dtype) for k in ("timestamp", "pressure", "flow", "u_in", "u_out")}
    for i in range(T):
        pressure = obs["pressure"]                                            # run_controller.py:118
        u_in = pid_call(cst, K, pressure, obs["time"], kx, kf, period)        # :123
        u_out = expiratory_call(est, obs["time"], period)                     # :124
        obs = env(u_in, u_out)                                                # :126
        ts["timestamp"][i] = c(dt) * st["steps"] - c(dt)                      # :128 env.time(state) - dt
        ts["u_in"][i], ts["u_out"][i], ts["pressure"][i] = u_in, u_out, pressure
        ts["flow"][i] = c(env_flow)                                           # :132 env.flow is a static field
    ts["target"] = np.stack([waveform_at(ts["timestamp"][i], kx, kf, period) for i in range(T)])   # :207
    return dict(timeseries=ts, env_state=st, obs=obs, controller_state=cst, expiratory_state=est)


def test_controller_score(ts):
    """``test_controller``'s per-waveform score (``test_controller.py:40-45``): mean |pressure - target|."""
    return np.abs(ts["pressure"] - ts["target"]).mean(axis=0)
