"""Throughput of the experience kernels on one B200: dlc_gae_f32 over 2^20 agents at the
PPO lung horizon (T = 29) and at T = 1000, against the measured HBM peak (20 B algorithmic per env-step: losses,
values, log_probs read; advantages, returns written), and dlc_env_collect_f32 (random-policy dataset generation)
on the classic envs.  Prints one JSON object."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.envs.classic import Cartpole, Pendulum, PlanarQuadrotor, Reacher  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
N = args.envs
dev = "cuda"
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm_peak = float(peaks.get("hbm_gbs", 6549.8))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()  # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {"hbm_peak_gbs": hbm_peak, "gae": {}, "collect": {}}
g = torch.Generator(device=dev).manual_seed(0)
for T in (29, 1000):
    Ng = N if T == 29 else N // 4
    losses, values, logp = (torch.randn(Ng, T, device=dev, generator=g) for _ in range(3))
    for layout, tm in (("batch_major", False), ("time_major", True)):
        a = [x.t().contiguous() if tm else x for x in (losses, values, logp)]
        adv, ret = torch.empty_like(a[0]), torch.empty_like(a[0])
        ms = timed(lambda: fused.gae(a[0], a[1], a[2], None, 0.99, 0.95, time_major=tm), args.reps)
        gbs = 20.0 * Ng * T / (ms * 1e-3) / 1e9
        out["gae"][f"T{T}_{layout}"] = {"agents": Ng, "ms": ms, "env_steps_per_s": Ng * T / (ms * 1e-3),
                                        "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    del losses, values, logp
for name, env, T in (("pendulum", Pendulum.create(), 100), ("cartpole", Cartpole.create(), 100),
                     ("quadrotor", PlanarQuadrotor.create(), 100), ("reacher", Reacher.create(), 100)):
    x0 = env.init(N)[0].arr.contiguous()
    ms = timed(lambda: fused.env_collect(env, x0, T, seed=1), args.reps)
    bytes_per_step = 4.0 * (2 * env.state_dim + env.action_dim)
    gbs = bytes_per_step * N * T / (ms * 1e-3) / 1e9
    out["collect"][name] = {"trajectories": N, "T": T, "ms": ms, "env_steps_per_s": N * T / (ms * 1e-3),
                            "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
print(json.dumps(out))
