"""One LDS + LQR rollout (register-resident kernel, W from HBM) for ncu.  usage: run_lqr_one.py [T] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
d, m = 10, 3
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g)).contiguous()
B = torch.randn(N, d, m, device="cuda", generator=g)
K = 0.05 * torch.randn(N, m, d, device="cuda", generator=g)
x0 = torch.zeros(N, d, device="cuda")
W = fused.gaussian_disturbance(N, T, d, seed=1, std=0.5)
r = fused.lds_rollout(A, B, K, x0, T, policy="lqr", W=W)
torch.cuda.synchronize()
print(int((r.status == 0).sum()))
