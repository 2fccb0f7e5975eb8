"""Committed fixtures (tests/golden/): the reference's own known answers pin the oracle; the frozen oracle
vectors pin both the oracle (CPU) and the CUDA kernels (GPU)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import lds as O
from oracle import lung as OL

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_reference_known_answers():
    k = json.load(open(os.path.join(G, "reference_known_answers.json")))
    kx, kf, period = OL.waveform_knots(35.0, 5.0)
    for t, v in zip(k["waveform_at"]["t"], k["waveform_at"]["value"]):
        assert float(OL.waveform_at(np.array([t], np.float32), kx, kf, period)[0]) == v
    for t, v in zip(k["waveform_is_in"]["t"], k["waveform_is_in"]["value"]):
        assert bool(OL.waveform_is_in(np.float32(t), 3.0)) == v
    assert OL.lung_env_time(k["lung_env_time"]["steps"]) == np.float32(k["lung_env_time"]["value"])
    assert OL.proper_time(np.float32(np.inf)) == 0.0 and OL.proper_time(np.float32(10.0)) == 10.0
    assert OL.controller_decay(0.5) == math.inf and OL.controller_decay(1.0) == 0.0
    assert OL.controller_decay(1.5) == np.float32(5) * (1 - np.exp(np.float32(5) * (np.float32(1.5) - np.float32(1.5))))


def test_oracle_matches_its_frozen_vectors():
    z = np.load(os.path.join(G, "oracle_lds_gpc_small.npz"))
    r = O.rollout_lds(z["A"], z["B"], z["K"], z["x0"], z["W"], "gpc", H=3, HH=10, lr_scale=1e-4, dtype=np.float64)
    np.testing.assert_allclose(r["costs"], z["costs"], rtol=1e-12)
    np.testing.assert_allclose(r["state"]["M"], z["M"], rtol=1e-10, atol=1e-15)
    z = np.load(os.path.join(G, "oracle_lung_small.npz"))
    for sim in ("balloon", "delay"):
        rr = OL.run_controller_scan(z["K"], z["pip"], 5.0, T=29, sim=sim, dtype=np.float32)
        for k, v in rr["timeseries"].items():
            np.testing.assert_allclose(v, z[f"{sim}_{k}"], rtol=1e-6, atol=1e-6)
    from oracle import of as OF
    from oracle import ppo as OP
    z = np.load(os.path.join(G, "oracle_gae_small.npz"))
    adv, ret = OP.ac_tail(z["losses"], z["log_probs"], z["values"], z["masks"], 0.99, 0.95)
    np.testing.assert_array_equal(adv, z["advantages"])
    np.testing.assert_array_equal(ret, z["returns"])
    z = np.load(os.path.join(G, "oracle_bpc_small.npz"))
    N, H, m, d = z["M0"].shape
    r = O.rollout_lds(z["A"], z["B"], z["K"], z["x0"], z["W"], "bpc", H=H,
                      state=O.bpc_init(N, d, m, H, z["M0"], z["eps0"], dtype=np.float64), bpc_eps=z["eps"], decay=False,
                      lr_scale=1e-5, dtype=np.float64)
    np.testing.assert_allclose(r["costs"], z["costs"], rtol=1e-12)
    z = np.load(os.path.join(G, "oracle_grc_small.npz"))
    r = OF.rollout_of(z["A"], z["B"], z["C"], z["x0"], z["W"], "fgrc", dtype=np.float64, Theta0=z["Theta0"], Phi=z["Phi"],
                      m=10, lr=0.005, decay=True, R_M=10.0)
    np.testing.assert_allclose(r["costs"], z["costs"], rtol=1e-10)


@pytest.mark.gpu
def test_kernels_match_the_frozen_vectors():
    from deluca_b200.fused import lds_rollout
    from deluca_b200.lung.controllers import PID
    from deluca_b200.lung.core import BreathWaveform
    from deluca_b200.lung.envs import BalloonLung, DelayLung
    from deluca_b200.lung.utils.scripts.run_controller import run_controller_scan
    z = np.load(os.path.join(G, "oracle_lds_gpc_small.npz"))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
    T = z["W"].shape[0]
    for variant in (0, 1, 5):
        r = lds_rollout(t(z["A"]), t(z["B"]), t(z["K"]), t(z["x0"]), T, policy="gpc", W=t(z["W"]), H=3, HH=10,
                        lr_scale=1e-4, kernel_variant=variant)
        c = r.costs.cpu().numpy()
        assert (np.abs(c - z["costs"]) <= 1e-4 * np.maximum(np.abs(z["costs"]), 1e-2 * z["costs"].mean())).all()
        np.testing.assert_allclose(r.M.cpu().numpy(), z["M"], rtol=1e-3, atol=1e-5 * np.abs(z["M"]).max())
    z = np.load(os.path.join(G, "oracle_lung_small.npz"))
    for sim, Env in (("balloon", BalloonLung), ("delay", DelayLung)):
        res = run_controller_scan(PID.create(params=torch.tensor(z["K"])), T=29, env=Env.create(),
                                  waveform=BreathWaveform.create(pip=torch.tensor(z["pip"]), peep=5.0),
                                  init_controller=True)
        for k in ("u_in", "pressure", "target", "timestamp", "u_out"):
            want = z[f"{sim}_{k}"].T
            got = res["timeseries"][k].cpu().numpy()
            assert (np.abs(got - want) <= 1e-4 * np.maximum(np.abs(want), max(np.abs(want).mean(), 1e-3))).all(), (sim, k)
    # experience kernel: float32, bit for bit
    from deluca_b200 import fused
    z = np.load(os.path.join(G, "oracle_gae_small.npz"))
    adv, ret = fused.gae(t(z["losses"]), t(z["values"]), t(z["log_probs"]), t(z["masks"]), 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy(), z["advantages"])
    np.testing.assert_array_equal(ret.cpu().numpy(), z["returns"])
    # BPC (lane-split kernel and the runtime-dims one)
    z = np.load(os.path.join(G, "oracle_bpc_small.npz"))
    T = z["W"].shape[0]
    for variant in (0, 1):
        r = lds_rollout(t(z["A"]), t(z["B"]), t(z["K"]), t(z["x0"]), T, policy="bpc", W=t(z["W"]), H=z["M0"].shape[1],
                        M=t(0.01 * z["M0"]), eps=t(z["eps0"]), eps_new=t(z["eps"]), decay=False, lr_scale=1e-5,
                        kernel_variant=variant)
        c = r.costs.cpu().numpy()
        assert (np.abs(c - z["costs"]) <= 1e-4 * np.maximum(np.abs(z["costs"]), 1e-2 * z["costs"].mean())).all()
        np.testing.assert_allclose(r.M.cpu().numpy(), z["M"], rtol=1e-3, atol=1e-6)
    # output feedback: GRC on the colab's shape (compile-time-dims instantiation) and the runtime-dims kernel
    z = np.load(os.path.join(G, "oracle_grc_small.npz"))
    T = z["W"].shape[0]
    for lanes, hint in ((0, False), (32, False), (0, True)):   # compile-time dims, runtime dims, register-resident
        r = fused.lds_of_rollout(t(z["A"]), t(z["B"]), t(z["C"]), t(z["x0"]), T, agent="fgrc", W=t(z["W"]), Phi=t(z["Phi"]),
                                 theta=t(z["Theta0"]), m=10, lr=0.005, decay=True, R_M=10.0, lanes_per_env=lanes,
                                 phi_is_identity=hint)
        c = r.costs.cpu().numpy()
        assert (np.abs(c - z["costs"]) <= 1e-4 * np.maximum(np.abs(z["costs"]), 1e-2 * z["costs"].mean())).all()
        np.testing.assert_allclose(r.theta.cpu().numpy(), z["Theta"], rtol=1e-3, atol=1e-4 * np.abs(z["Theta"]).max())
