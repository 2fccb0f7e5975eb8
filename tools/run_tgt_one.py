"""One LDS+GPC(H=10) rollout launch for ncu.  usage: run_tgt_one.py [envs] [T] [variant] [hbm|kernel]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200.fused import gaussian_disturbance, lds_rollout  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = sys.argv[4] if len(sys.argv) > 4 else "hbm"
d, m, H = 10, 3, 10
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g))
B = torch.randn(N, d, m, device="cuda", generator=g)
K = 0.05 * torch.randn(N, m, d, device="cuda", generator=g)
x0 = torch.zeros(N, d, device="cuda")
W = gaussian_disturbance(N, T, d, seed=1) if src == "hbm" else None
for _ in range(2):
    r = lds_rollout(A, B, K, x0, T, policy="gpc", W=W, H=H, HH=H, lr_scale=1e-4, seed=1, kernel_variant=variant)
torch.cuda.synchronize()
print(variant, float(r.cost_sum.mean()))
