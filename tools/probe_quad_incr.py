"""Direct (kernel_variant 13) vs incremental-window (17) target kernel: agreement and time.
usage: probe_quad_incr.py [envs] [T]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200.fused import gaussian_disturbance, lds_rollout  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
import bench  # noqa: E402

wl = dict(bench.WORKLOADS["tgt"])
d, m, H = wl["d"], wl["m"], wl["H"]
A, B, K, x0 = bench.make_problem(torch, wl, torch.device("cuda"), 0, N)      # the benchmark's systems (LQR gains)
W = gaussian_disturbance(N, T, d, seed=bench.SEED, std=bench.W_STD)
out = {"envs": N, "T": T}
res = {}
for variant in (13, 17):
    ms = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = lds_rollout(A, B, K, x0, T, policy="gpc", W=W, H=H, HH=H, lr_scale=bench.LR_SCALE, seed=bench.SEED, kernel_variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res[variant] = r
    out[f"ms_v{variant}"] = sorted(ms[1:])[1]
a, b = res[13], res[17]
ok = (a.status == 0) & (b.status == 0)
out["finite_envs"] = int(ok.sum())
rel = ((a.cost_sum - b.cost_sum).abs() / a.cost_sum.abs().clamp_min(1e-30))[ok]
out["cost_sum_rel_max"] = float(rel.max())
out["costs_rel_max"] = float(((a.costs - b.costs).abs()[:, ok].amax(0) / a.costs[:, ok].abs().amax(0)).max())
out["M_rel_max"] = float(((a.M - b.M).abs()[ok].amax() / a.M[ok].abs().amax()))
out["x_rel_max"] = float(((a.x - b.x).abs()[ok].amax() / a.x[ok].abs().amax()))
print(json.dumps(out))
