"""One dlc_gae_f32 launch for ncu.  usage: run_gae_one.py [T] [agents] [time_major]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
tm = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
g = torch.Generator(device="cuda").manual_seed(0)
shape = (T, N) if tm else (N, T)
losses, values, logp = (torch.randn(*shape, device="cuda", generator=g) for _ in range(3))
for _ in range(2):
    adv, ret = fused.gae(losses, values, logp, None, 0.99, 0.95, time_major=tm)
torch.cuda.synchronize()
print(float(adv.sum()))
