"""One lane-split BPC rollout (reference default shape) for ncu.  usage: run_bpc_one.py [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N, d, m, H = 65536, 10, 3, 5
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g)).contiguous()
B = torch.randn(N, d, m, device="cuda", generator=g)
K = 0.05 * torch.randn(N, m, d, device="cuda", generator=g)
x0 = torch.zeros(N, d, device="cuda")
unit = lambda t: t / t.flatten(1).norm(dim=1).reshape(-1, *([1] * (t.dim() - 1)))
M0 = 0.01 * unit(torch.randn(N, H, m, d, device="cuda", generator=g))
eps0 = unit(torch.randn(N, H, H, m, d, device="cuda", generator=g))
r = fused.lds_rollout(A, B, K, x0, T, policy="bpc", H=H, M=M0, eps=eps0, decay=False, lr_scale=1e-6, seed=3,
                      keep_costs=False)
torch.cuda.synchronize()
print(int((r.status == 0).sum()))
