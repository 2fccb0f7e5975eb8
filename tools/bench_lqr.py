"""LDS(d=10, m=3) + LQR closed loop on one B200: 262,144 envs x T = 1000, per-env A/B/K; register-resident
thread-per-env kernel (default) next to the lane-split kernel (kernel_variant = 5).  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

N, T, d, m = 262144, 1000, 10, 3
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g)).contiguous()
B = torch.randn(N, d, m, device="cuda", generator=g)
K = 0.05 * torch.randn(N, m, d, device="cuda", generator=g)
x0 = torch.zeros(N, d, device="cuda")
W = fused.gaussian_disturbance(N, T, d, seed=1, std=0.5)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = float(peaks.get("hbm_gbs", 6549.8))
out = {"workload": f"LDS(d={d}, m={m}) + LQR, {N} envs x T={T}, per-env A/B/K", "hbm_peak_gbs": hbm}
for label, variant in (("register_resident", 0), ("lane_split", 5)):
    for src, kw in (("W_from_hbm", dict(W=W)), ("in_kernel_w", dict(seed=1))):
        f = lambda: fused.lds_rollout(A, B, K, x0, T, policy="lqr", kernel_variant=variant, **kw)
        f()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        b = 4 * (d + 1) if src == "W_from_hbm" else 4
        out[f"{label}_{src}"] = {"ms": ms, "env_steps_per_s": N * T / (ms * 1e-3), "bytes_per_env_step": b,
                                 "frac_of_hbm_peak": b * N * T / (ms * 1e-3) / 1e9 / hbm}
print(json.dumps(out))
