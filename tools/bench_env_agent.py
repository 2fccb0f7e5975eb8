#!/usr/bin/env python
"""The functional Env / Agent driver (dlc_env_agent_rollout_f32) on one B200: 2^20 episodes x T steps of
each classic env under the generic PID agent and under linear state feedback, (a) with the full `Rollout` stacks
(losses[T,N] + actions[T,N,.]), (b) losses only, (c) loss sums only, (d) value + gradient w.r.t. the gains (forward
mode, loss sums only).  Bytes are the stores the call is asked for (there are no per-step loads).  One JSON object."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.envs.classic import Cartpole, Pendulum, PlanarQuadrotor, Reacher  # noqa: E402
from deluca_b200.envs.classic._pendulum import QuadCost  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--T", type=int, default=200)
ap.add_argument("--only", default="")
args = ap.parse_args()
N, T = args.envs, args.T
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6549.8))
except OSError:
    hbm = 6549.8
g = torch.Generator(device="cuda").manual_seed(0)
envs = [("pendulum", Pendulum.create(), [0.0]), ("cartpole", Cartpole.create(), [0.0]),
        ("quadrotor", PlanarQuadrotor.create(wind=0.1), [0.4905, 0.4905]), ("reacher", Reacher.create(), [0.0, 0.0])]
out = {"workload": f"classic env x (PID | linear feedback), {N} episodes x T={T}, per-episode gains", "hbm_peak_gbs": hbm}


def timed(f):
    f()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[1]


for name, env, gu in envs:
    if args.only and name not in args.only:
        continue
    d, m = env.state_dim, env.action_dim
    x0 = 0.05 * torch.randn(N, d, device="cuda", generator=g)
    cost = QuadCost([0.0] * d, 1.0, 0.1, gu)
    for agent in ("pid", "linear"):
        if agent == "pid":
            gains = torch.tensor([0.05, 0.02, 0.01], device="cuda") * (1 + 0.2 * torch.randn(N, 3, device="cuda", generator=g))
            da = d
        else:
            gains = 0.03 * torch.randn(N, m, d, device="cuda", generator=g)
            da = m
        gains = gains.contiguous()
        # the wrapper allocates its outputs per call: time the launch with a caching allocator that is already warm
        legs = (("full_stacks", dict(), 4 * (1 + da)), ("losses_only", dict(keep_actions=False), 4),
                ("loss_sums_only", dict(keep_actions=False, keep_losses=False), 0),
                ("value_and_grad", dict(keep_actions=False, keep_losses=False, want_grad=True), 0))
        for leg, kw, b in legs:
            ms = timed(lambda: fused.env_agent_rollout(env, cost, agent, gains, x0, T, pid_decay=0.0566, **kw))
            rec = {"ms": ms, "env_steps_per_s": N * T / (ms * 1e-3)}
            if b:
                rec.update(bytes_per_env_step=b, frac_of_hbm_peak=b * N * T / (ms * 1e-3) / 1e9 / hbm)
            out[f"{name}_{agent}_{leg}"] = rec
print(json.dumps(out))
