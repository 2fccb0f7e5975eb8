"""LDS + GPC on shapes outside the compile-time list: the lane-group kernel (lds_gpc_group.cu) next to the
thread-per-env runtime-dims kernel (kernel_variant = 1).  Default: the reference's own defaults, LDS(d_hidden = 25,
d_in = 1) + GPC(H = 3, HH = 2), 65,536 envs x T = 500, in-kernel disturbances.  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.riccati import dare_gain  # noqa: E402

dev = "cuda"
out = {}
for (d, m, H, HH, N, T) in ((25, 1, 3, 2, 65536, 500), (10, 3, 5, 5, 65536, 500), (40, 5, 3, 6, 32768, 300)):
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device=dev, generator=g)).contiguous()
    B = torch.randn(N, d, m, device=dev, generator=g)
    K = dare_gain(A, B)[0].contiguous()
    x0 = torch.zeros(N, d, device=dev)
    row = {}
    for label, variant, Tc in (("lane_group", 0, T), ("thread_per_env", 1, max(T // 10, 20))):
        run = lambda: fused.lds_rollout(A, B, K, x0, Tc, policy="gpc", H=H, HH=HH, lr_scale=1e-4, seed=1,
                                        keep_costs=False, kernel_variant=variant)
        run()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); r = run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        row[label] = {"T": Tc, "ms": ms, "env_steps_per_s": N * Tc / (ms * 1e-3), "finite_envs": int((r.status == 0).sum())}
    out[f"d{d}_m{m}_H{H}_HH{HH}_N{N}"] = row
print(json.dumps(out))
