#!/usr/bin/env python
"""Writes the fixtures in this directory.

  reference_known_answers.json  -- the known answers of the REFERENCE's own tests for this path
      (deluca/tests/lung/core_test.py:43-127), transcribed by hand with their file:line.  These pin the oracle
      (tests/test_oracle_lung.py) and the CUDA interpolator (tests/test_lung_gpu.py).
  oracle_lds_gpc_small.npz, oracle_lung_small.npz, oracle_gae_small.npz, oracle_bpc_small.npz, oracle_grc_small.npz
      -- regression vectors produced by THIS repo's float64
      oracle (oracle/lds.py, oracle/lung.py) on seeded inputs.  They are NOT reference outputs (the reference
      cannot run here: no JAX); they freeze the oracle so that a later edit of it cannot
      silently move the target the kernels are compared with.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lds as O  # noqa: E402
from oracle import lung as OL  # noqa: E402
from oracle import of as OF  # noqa: E402
from oracle import ppo as OP  # noqa: E402

known = {
    "source": "deluca/tests/lung/core_test.py (google/deluca v0.0.17)",
    "waveform_at": {"lines": "55-62", "default": {"pip": 35.0, "peep": 5.0, "bpm": 20},
                    "t": [0.0, 0.5, 1.0, 1.25, 1.5, 3.0], "value": [35.0, 35.0, 35.0, 20.0, 5.0, 35.0]},
    "waveform_is_in": {"lines": "43-47", "t": [0.5, 1.0, 1.5], "value": [True, True, False]},
    "lung_env_time": {"lines": "73-84", "steps": 1, "value": 0.03, "dt": 0.03},
    "proper_time": {"lines": "97-99", "t": ["inf", 10.0], "value": [0.0, 10.0]},
    "controller_decay": {"lines": "119-127", "t": [0.5, 1.0], "value": ["inf", 0.0],
                         "t_formula": 1.5, "formula": "5 * (1 - exp(5 * (xp[2] - elapsed(1.5))))"},
}
json.dump(known, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)

rng = np.random.default_rng(20240)
N, d, m, T = 6, 10, 3, 120
A, B, _ = O.lds_default_params(rng, N, d, m)
K = O.lqr_gain(A, B)
x0 = (0.1 * rng.standard_normal((N, d))).astype(np.float32)
W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
r = O.rollout_lds(A, B, K, x0, W, "gpc", H=3, HH=10, lr_scale=1e-4, dtype=np.float64)
np.savez_compressed(os.path.join(HERE, "oracle_lds_gpc_small.npz"), A=A, B=B, K=K, x0=x0, W=W,
                    costs=r["costs"], x_final=r["x_final"], M=r["state"]["M"])

Kp = rng.uniform(0.5, 4.0, (N, 3)).astype(np.float32)
pip = rng.choice([10, 15, 20, 25, 30, 35], N).astype(np.float32)
out = {}
for sim in ("balloon", "delay"):
    rr = OL.run_controller_scan(Kp, pip, 5.0, T=29, sim=sim, dtype=np.float32)
    for k, v in rr["timeseries"].items():
        out[f"{sim}_{k}"] = v
np.savez_compressed(os.path.join(HERE, "oracle_lung_small.npz"), K=Kp, pip=pip, **out)

# GAE + returns (oracle/ppo.py, float32: the CUDA kernel must reproduce these BIT FOR BIT)
Ng, Tg = 9, 29
losses, values, logp = (rng.standard_normal((Ng, Tg)).astype(np.float32) for _ in range(3))
masks = (rng.random((Ng, Tg)) > 0.15).astype(np.float32)
adv, ret = OP.ac_tail(losses, logp, values, masks, 0.99, 0.95)
np.savez_compressed(os.path.join(HERE, "oracle_gae_small.npz"), losses=losses, values=values, log_probs=logp,
                    masks=masks, advantages=adv, returns=ret)

# LDS + BPC at the reference default H = 5 (oracle/lds.py, float64), explicit perturbation stream
Hb, Tb = 5, 60
M0 = O.unit_norm(rng.standard_normal((N, Hb, m, d)).astype(np.float32))
eps0 = O.unit_norm(rng.standard_normal((N, Hb, Hb, m, d)).astype(np.float32))
eps = O.unit_norm(rng.standard_normal((Tb * N, Hb, m, d)).astype(np.float32)).reshape(Tb, N, Hb, m, d)
rb = O.rollout_lds(A, B, K, x0, W[:Tb], "bpc", H=Hb, state=O.bpc_init(N, d, m, Hb, M0, eps0, dtype=np.float64),
                   bpc_eps=eps, decay=False, lr_scale=1e-5, dtype=np.float64)
np.savez_compressed(os.path.join(HERE, "oracle_bpc_small.npz"), A=A, B=B, K=K, x0=x0, W=W[:Tb], M0=M0, eps0=eps0,
                    eps=eps, costs=rb["costs"], x_final=rb["x_final"], M=rb["state"]["M"])

# partially observed LDS + GRC (m = 10) on the colab's shape (oracle/of.py, float64)
pg, ng, mg, Tq = 2, 3, 10, 80
Ag = np.zeros((N, d, d), np.float32)
for i in range(N):
    Q = np.linalg.qr(rng.standard_normal((d, d)))[0]
    Ag[i] = Q @ np.diag(rng.uniform(0.5, 0.9, d)) @ Q.T
Bg = (rng.standard_normal((N, d, ng)) / np.sqrt(d)).astype(np.float32)
Cg = (rng.standard_normal((N, pg, d)) / np.sqrt(d)).astype(np.float32)
Phi, _ = OF.filter_bank("grc", mg, None)
Th0 = (0.3 * rng.standard_normal((N, Phi.shape[0], ng, pg))).astype(np.float32)
rg = OF.rollout_of(Ag, Bg, Cg, x0, W[:Tq], "fgrc", dtype=np.float64, Theta0=Th0, Phi=Phi, m=mg, lr=0.005, decay=True,
                   R_M=10.0)
np.savez_compressed(os.path.join(HERE, "oracle_grc_small.npz"), A=Ag, B=Bg, C=Cg, x0=x0, W=W[:Tq], Phi=Phi, Theta0=Th0,
                    costs=rg["costs"], U=rg["U"], Theta=rg["state"]["Theta"])
print("wrote", sorted(os.listdir(HERE)))
