"""dlc_lds_open_rollout_f32 on one B200: the colab's LDS (d_hidden = 10, d_action = 3, d_obs = 2), per-env A/B/C,
262,144 envs x T = 1000: random policy with in-kernel streams, and given actions + disturbances read from HBM.
Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

N, T, d, n, p = 262144, 1000, 10, 3, 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device=dev, generator=g)).contiguous()
B = torch.randn(N, d, n, device=dev, generator=g)
C = torch.randn(N, p, d, device=dev, generator=g)
x0 = torch.zeros(N, d, device=dev)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
hbm = float(peaks.get("hbm_gbs", 6549.8))
out = {"workload": f"LDS(d={d}, n={n}, p={p}) open loop, {N} envs x T={T}, per-env A/B/C", "hbm_peak_gbs": hbm}
ms = timed(lambda: fused.lds_open_rollout(A, B, C, x0, T, seed=1))
out["random_policy_in_kernel_streams"] = {"ms": ms, "env_steps_per_s": N * T / (ms * 1e-3),
                                          "bytes_per_env_step": 4 * p, "GBs": 4 * p * N * T / (ms * 1e-3) / 1e9}
U = torch.randn(T, N, n, device=dev, generator=g)
W = fused.gaussian_disturbance(N, T, d, seed=1, std=0.5)
ms = timed(lambda: fused.lds_open_rollout(A, B, C, x0, T, u_in=U, W=W))
b = 4 * (n + d + p)
out["given_actions_and_disturbances"] = {"ms": ms, "env_steps_per_s": N * T / (ms * 1e-3), "bytes_per_env_step": b,
                                         "GBs": b * N * T / (ms * 1e-3) / 1e9,
                                         "frac_of_hbm_peak": b * N * T / (ms * 1e-3) / 1e9 / hbm}
print(json.dumps(out))
