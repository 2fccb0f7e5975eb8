"""Lung + PID episode launches for ncu: BalloonLung / DelayLung at the reference's horizon (T = 29) and T = 1000.
usage: run_lung_one.py [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.lung import _launch  # noqa: E402
from deluca_b200.lung.core import BreathWaveform  # noqa: E402
from deluca_b200.lung.envs import BalloonLung, DelayLung  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator().manual_seed(0)
K = 5 * torch.rand(N, 3, generator=g)
pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
wave = BreathWaveform.create(pip=pip, peep=5.0)
for Env in (BalloonLung, DelayLung):
    for T in (29, 1000):
        env = Env.create()
        st, obs = env.reset()
        p, b, series = _launch.prepare(env, wave, N, T, "closed_loop", "cuda", K=K, env_state=env.state_dict(st),
                                       obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time))
        fused.lung_rollout(p, b)
        torch.cuda.synchronize()
        print(Env.__name__, T, float(series["pressure"][-1].mean()))
