"""One open-loop LDS rollout (register-resident kernel, given actions + disturbances from HBM) for ncu.
usage: run_lds_open_one.py [T] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
d, n, p = 10, 3, 2
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g)).contiguous()
B = torch.randn(N, d, n, device="cuda", generator=g)
C = torch.randn(N, p, d, device="cuda", generator=g)
x0 = torch.zeros(N, d, device="cuda")
U = torch.randn(T, N, n, device="cuda", generator=g)
W = fused.gaussian_disturbance(N, T, d, seed=1, std=0.5)
x, y, _ = fused.lds_open_rollout(A, B, C, x0, T, u_in=U, W=W)
torch.cuda.synchronize()
print(float(y.abs().mean()))
