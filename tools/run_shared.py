#!/usr/bin/env python
"""Small driver for profiling the shared-dynamics tcgen05 kernel: python tools/run_shared.py [N] [T] [W?]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deluca_b200 import fused
from deluca_b200.riccati import dare_gain

N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
with_w = len(sys.argv) > 3 and sys.argv[3] == "1"
d, m, dev = 256, 64, "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.linalg.qr(torch.randn(d, d, device=dev, generator=g, dtype=torch.float64))[0]
A = ((Q * (0.7 + 0.25 * torch.rand(d, device=dev, generator=g, dtype=torch.float64))) @ Q.T).float().contiguous()
B = (torch.randn(d, m, device=dev, generator=g) / d ** 0.5).contiguous()
K = dare_gain(A, B)[0].contiguous()
G = fused.lds_shared_pack(A, B, K)
x = torch.randn(N, d, device=dev, generator=g)
W = 0.5 * torch.randn(T, N, d, device=dev, generator=g) if with_w else None
for _ in range(3):
    r = fused.lds_shared_lqr_rollout(G, x, T, m, W=W)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = fused.lds_shared_lqr_rollout(G, x, T, m, W=W)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"N={N} T={T} W={with_w}: {ms:.3f} ms, {N * T / ms / 1e3:.1f} M env-steps/s, {ms * 1e3 / T / ((N + 64 * 148 - 1) // (64 * 148)):.2f} us per CTA-step, "
      f"nonfinite={int(r.status.sum())}")
