"""value_and_grad of the lung episode w.r.t. the PID gains (2^20 breaths x T = 29): forward-mode kernel (default) vs the
checkpointed reverse-mode kernel (DLC_LUNG_GRAD_REVERSE=1, set before the library loads), and their agreement.
usage: probe_lung_grad.py [balloon|delay] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deluca_b200 import _abi, fused
from deluca_b200.lung import _launch
from deluca_b200.lung.core import DEFAULT_DT, BreathWaveform
from deluca_b200.lung.envs import BalloonLung, DelayLung

SIM = sys.argv[1] if len(sys.argv) > 1 else "balloon"
N, T = 1 << 20, int(sys.argv[2]) if len(sys.argv) > 2 else 29
g = torch.Generator().manual_seed(0)
K = (5 * torch.rand(N, 3, generator=g)).cuda().contiguous()
pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
wave = BreathWaveform.create(pip=pip, peep=5.0)
env = BalloonLung.create() if SIM == "balloon" else DelayLung.create()
pg = _abi.DlcLungParams()
pg.N, pg.T, pg.mode = N, T, 0
kx, kf, period, xp_in = wave.knots(N=N, device="cuda")
pg.nk = len(kx)
for i, v in enumerate(kx):
    pg.kx[i] = v
pg.kf_per_env = int(kf.dim() == 2)
pg.period, pg.xp_in, pg.default_dt, pg.pid_decay = period, xp_in, DEFAULT_DT, 0.03 / 0.53
_launch.sim_params(env, pg)
tt = torch.linspace(0, 0.03 * T, T, device="cuda")
for _ in range(3):
    loss, grad, sums = fused.lung_pid_value_and_grad(pg, K, kf, tt)
torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss, grad, sums = fused.lung_pid_value_and_grad(pg, K, kf, tt); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
mode = "reverse" if os.environ.get("DLC_LUNG_GRAD_REVERSE") == "1" else "forward"
print(f"{SIM} T={T} {mode}: {sorted(ts)[3]:.3f} ms; loss mean {float(loss.mean()):.6f}; grad mean {grad.mean(0).tolist()}")
torch.save((loss.cpu(), grad.cpu()), f"/tmp/lung_grad_{mode}.pt")
if os.path.exists("/tmp/lung_grad_forward.pt") and os.path.exists("/tmp/lung_grad_reverse.pt"):
    lf, gf = torch.load("/tmp/lung_grad_forward.pt"); lr_, gr = torch.load("/tmp/lung_grad_reverse.pt")
    sc = gr.abs().mean(0, keepdim=True)
    rel = ((gf - gr).abs() / torch.maximum(gr.abs(), sc)).amax(1)
    print(f"forward vs reverse: loss max abs diff {float((lf - lr_).abs().max()):.3e}; "
          f"grad rel diff: median {float(rel.median()):.2e}, 99.9th pct {float(rel.quantile(0.999)):.2e}, max {float(rel.max()):.2e}")
