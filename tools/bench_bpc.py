"""LDS + BPC (deluca/agents/_bpc.py, reference defaults H = 5, delta = 0.01) on one B200: 65,536 envs x T = 1000 with
per-env A/B/K, in-kernel Philox disturbances and perturbations; the lane-split kernel next to the runtime-dims one
(kernel_variant = 1).  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as _bench  # noqa: E402
from deluca_b200 import fused  # noqa: E402

N, T, d, m, H = 65536, 1000, 10, 3, 5
dev = torch.device("cuda")
wl = dict(_bench.WORKLOADS["c2"], N=N, T=T)
A, B, K, x0 = _bench.make_problem(torch, wl, dev, 0)   # the LDS defaults + batched DARE gains of bench.py
g = torch.Generator(device=dev).manual_seed(1)
unit = lambda t: t / t.flatten(1).norm(dim=1).reshape(-1, *([1] * (t.dim() - 1)))
M0 = 0.01 * unit(torch.randn(N, H, m, d, device=dev, generator=g))
eps0 = unit(torch.randn(N, H, H, m, d, device=dev, generator=g))


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); r = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], r


out = {"workload": f"LDS(d={d}, m={m}) + BPC(H={H}), {N} envs x T={T}, in-kernel Philox disturbances and perturbations"}
for label, variant, Tc in (("lane_split", 0, T), ("runtime_dims", 1, T // 10)):
    run = lambda: fused.lds_rollout(A, B, K, x0, Tc, policy="bpc", H=H, M=M0, eps=eps0, decay=False, lr_scale=1e-6,
                                    seed=3, keep_costs=False, kernel_variant=variant)
    ms, r = timed(run)
    out[label] = {"T": Tc, "ms": ms, "env_steps_per_s": N * Tc / (ms * 1e-3), "finite_envs": int((r.status == 0).sum()),
                  "normals_per_s": N * Tc * (H * m * d + d) / (ms * 1e-3)}
print(json.dumps(out))
