"""One dlc_plan_f32 launch of config 3 (pendulum, 4096 x H=200 x 20 iterations) for ncu.
usage: run_plan_one.py [ilqr|ilqr_fixed|igpc|ilc] [iterations]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.envs.classic import Pendulum, QuadCost  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ilqr"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N, H = 4096, 200
kw = {"ilqr": dict(algo=0, alpha0=0.5, n_alpha=10), "ilqr_fixed": dict(algo=0, alpha0=0.5, n_alpha=10, fix_backtracking=True),
      "igpc": dict(algo=1, alpha0=1.0, n_alpha=6, lr=0.001), "ilc": dict(algo=2, alpha0=1.0, n_alpha=10)}[which]
env = Pendulum.create(H=H)
cost = QuadCost(list(env.goal_state), 1.0, 0.1)
g = torch.Generator(device="cuda").manual_seed(0)
x0 = env.init(N)[0].arr + 0.05 * torch.randn(N, env.state_dim, device="cuda", generator=g)
U0 = torch.zeros(N, H, 1, device="cuda")
p = fused.plan_params(env, env, cost, N, H, T=T, **kw)
c = fused.plan(p, U0, x0)[4]
torch.cuda.synchronize()
print(which, float(c.mean()))
