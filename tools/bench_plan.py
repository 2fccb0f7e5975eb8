"""Config 3 on one B200: Pendulum / Cartpole + iLQR, iGPC_closed, iLC_closed, 4096 trajectories x H = 200 x 20
iterations, whole loop in one launch of dlc_plan_f32.  Prints one JSON object (kernel time from CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.envs.classic import Cartpole, Pendulum, QuadCost  # noqa: E402

dev = "cuda"


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {}
N, H, T = 4096, 200, 20
for name, Env in (("pendulum", Pendulum), ("cartpole", Cartpole)):
    env = Env.create(H=H)
    cost = QuadCost(list(env.goal_state), 1.0, 0.1)
    gg = torch.Generator(device=dev).manual_seed(0)
    x0 = env.init(N)[0].arr + 0.05 * torch.randn(N, env.state_dim, device=dev, generator=gg)
    U0 = torch.zeros(N, H, 1, device=dev)
    for algo, label, kw in ((0, "ilqr_as_written", dict(alpha0=0.5, n_alpha=10)),
                            (0, "ilqr_fixed_backtracking", dict(alpha0=0.5, n_alpha=10, fix_backtracking=True)),
                            (1, "igpc", dict(alpha0=1.0, n_alpha=6, lr=0.001)),
                            (2, "ilc", dict(alpha0=1.0, n_alpha=10))):
        p = fused.plan_params(env, env, cost, N, H, T=T, algo=algo, **kw)
        ms = timeit(lambda: fused.plan(p, U0, x0))
        c = fused.plan(p, U0, x0)[4]
        out[f"{label}_{name}"] = {"trajectories": N, "H": H, "iterations": T, "kernel_ms": ms,
                                  "trajectory_steps_iterations_per_s": N * H * T / (ms * 1e-3),
                                  "mean_final_cost": float(c.mean()), "finite": bool(torch.isfinite(c).all()),
                                  "bound": "instruction latency (serial T x H dependency chain per trajectory)"}
# the building blocks on their own: igpc.rollout and GPC_rollout, one launch each (pendulum)
from deluca_b200.igpc.rollout import rollout as _rollout  # noqa: E402
env = Pendulum.create(H=H)
cost = QuadCost(list(env.goal_state), 1.0, 0.1)
gg = torch.Generator(device=dev).manual_seed(0)
x0 = env.init(N)[0].arr + 0.05 * torch.randn(N, env.state_dim, device=dev, generator=gg)
U0 = 0.1 * torch.randn(N, H, 1, device=dev, generator=gg)
p = fused.plan_params(env, env, cost, N, H)
X, U, c = fused.env_rollout(p, U0, x0)
kz = torch.zeros(N, H, 1, device=dev)
Kz = torch.zeros(N, H, 1, env.state_dim, device=dev)
al = torch.ones(N, device=dev)
out["env_rollout_open_loop_pendulum"] = {"trajectories": N, "H": H, "kernel_ms": timeit(lambda: fused.env_rollout(p, U0, x0))}
out["env_rollout_feedback_pendulum"] = {"trajectories": N, "H": H,
                                        "kernel_ms": timeit(lambda: fused.env_rollout(p, U, x0, kz, Kz, X, al))}
pg = fused.plan_params(env, env, cost, N, H, algo=1, lr=0.001)
out["igpc_rollout_pendulum"] = {"trajectories": N, "H": H,
                                "kernel_ms": timeit(lambda: fused.igpc_rollout(pg, U, kz, Kz, X, al, x0))}
print(json.dumps(out))
