"""Oracle checks for the lung path: the reference's own known answers
(deluca/tests/lung/core_test.py) plus behavioural sanity of the simulators (which are pinned to reference runs in test_ref_fixtures_oracle.py)."""
import numpy as np
import pytest

from oracle import lung as L


def _at(t, pip=35.0, peep=5.0, dtype=np.float32):
    kx, kf, period = L.waveform_knots(pip, peep, dtype=dtype)
    return float(L.waveform_at(np.array([t], dtype=dtype), kx, kf, period)[0])


def test_waveform_at_known_answers():
    # deluca/tests/lung/core_test.py:55-62
    for t, v in [(0.0, 35.0), (0.5, 35.0), (1.0, 35.0), (1.25, 20.0), (1.5, 5.0), (3.0, 35.0)]:
        assert _at(t) == v, (t, _at(t))


def test_waveform_fp32_exhale_is_a_ramp():
    # SURVEY A-9: in float32 3 - 1e-8 == 3, so the PEEP plateau collapses into a 5 -> 35 ramp
    assert _at(2.0) == 15.0 and _at(2.5) == 25.0
    # np.interp on the same float32 knots agrees everywhere
    xp = np.asarray(L.DEFAULT_XP, dtype=np.float32)
    fp = np.array([35, 35, 5, 5, 35], dtype=np.float32)
    for t in np.linspace(0, 9, 181):
        assert abs(_at(t) - np.interp(np.float32(t), xp, fp, period=3.0)) < 2e-5


def test_waveform_is_in_and_elapsed():
    # deluca/tests/lung/core_test.py:43-47, 64-68
    f = lambda t: bool(L.waveform_is_in(np.float32(t), 3.0))
    assert f(0.5) and f(1.0) and not f(1.5)
    assert np.isclose(np.float32(1.3) % np.float32(3), np.float32(1.3 + 3) % np.float32(3))


def test_lung_env_time_proper_time_decay():
    # core_test.py:73-84, 97-99, 119-127
    assert L.lung_env_time(1) == np.float32(0.03)
    assert L.proper_time(np.float32(np.inf)) == 0 and L.proper_time(np.float32(10.0)) == 10.0
    assert L.controller_decay(0.5) == float("inf")
    assert L.controller_decay(1.0) == 0.0
    assert L.controller_decay(1.5) == np.float32(5) * (1 - np.exp(np.float32(5) * (np.float32(1.5) - np.float32(1.5))))


@pytest.mark.parametrize("sim", ["balloon", "delay"])
def test_run_controller_scan_f32_tracks_f64(sim):
    rng = np.random.default_rng(0)
    N, T = 12, 120
    K = rng.uniform(0, 5, (N, 3))
    pip = rng.choice([10, 15, 20, 25, 30, 35], N).astype(np.float64)
    a = L.run_controller_scan(K, pip, 5.0, T=T, sim=sim, dtype=np.float32)
    b = L.run_controller_scan(K, pip, 5.0, T=T, sim=sim, dtype=np.float64)
    # float64 keeps 3 - 1e-8 distinct from 3, so targets differ from the first exhale on
    inh = slice(0, 33)   # env time < 1 s
    assert np.allclose(a["timeseries"]["u_in"][inh], b["timeseries"]["u_in"][inh], rtol=1e-3, atol=1e-3)
    assert np.allclose(a["timeseries"]["pressure"][inh], b["timeseries"]["pressure"][inh], rtol=1e-3, atol=1e-3)
    assert a["timeseries"]["timestamp"].shape == (T, N)
    # env.time(state) = dt * steps with steps = time + 1 (sic, _balloon_lung.py:135): the emitted
    # "timestamp" is dt * (elapsed seconds), i.e. 0.0009 * i
    assert np.allclose(a["timeseries"]["timestamp"][:, 0], 0.03 * 0.03 * np.arange(T), atol=1e-5)


def test_balloon_pid_raises_pressure_and_delay_lung_is_inert_before_25s():
    N, T = 4, 29
    K = np.tile([3.0, 4.0, 0.0], (N, 1))
    r = L.run_controller_scan(K, 35.0, 5.0, T=T, sim="balloon")
    p = r["timeseries"]["pressure"]
    assert p[0].max() == 0.0 and p[-1].min() > 5.0          # the balloon inflates under PID
    assert (r["timeseries"]["flow"] == 0).all()             # env.flow is a static field (run_controller.py:132)
    assert L.test_controller_score(r["timeseries"]).shape == (N,)
    d = L.run_controller_scan(K, 35.0, 5.0, T=T, sim="delay")
    assert (d["timeseries"]["pressure"] == 0).all()         # `time < delay` compares seconds to steps (sic)
    assert d["env_state"]["in_history"][0, 0] == d["timeseries"]["u_in"][-1, 0]


def test_generic_pid():
    st = dict(P=np.zeros(2, np.float32), I=np.zeros(2, np.float32), D=np.zeros(2, np.float32))
    u = L.generic_pid_call(st, np.float32(1), np.float32(2), np.float32(3), np.array([1.0, -1.0], np.float32))
    decay = np.float32(0.03 / 0.53)
    assert np.allclose(u, np.array([1, -1]) * (1 + 2 * decay + 3 * decay))
