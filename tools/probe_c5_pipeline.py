#!/usr/bin/env python
"""Config 5 (shared-dynamics LDS d=256, m=64 + GPC H=HH=16, per-env M): the serial step loop against the two-half
software pipeline of dlc_lds_shared_gpc_rollout_f32 (DLC_SGPC_PIPE, DLC_SGPC_CHAIN_CTAS), same process, same buffers.
Per-step time = (call with T - call with T_SHORT) / (T - T_SHORT), CUDA events, median of --reps.  One JSON object."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deluca_b200 import fused
from deluca_b200.riccati import dare_gain

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=148 * 64 * 6)
ap.add_argument("--T", type=int, default=12)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ctas", type=str, default="12,16,18,20,24")
args = ap.parse_args()
dev = "cuda:0"
d, m, H = 256, 64, 16
N, T = args.envs, args.T
gg = torch.Generator(device=dev).manual_seed(0)
Qm = torch.linalg.qr(torch.randn(d, d, device=dev, generator=gg, dtype=torch.float64))[0]
A = ((Qm * (0.7 + 0.25 * torch.rand(d, device=dev, generator=gg, dtype=torch.float64))) @ Qm.T).float().contiguous()
B = (torch.randn(d, m, device=dev, generator=gg) / d ** 0.5).contiguous()
K = dare_gain(A, B)[0].contiguous()
x = torch.randn(N, d, device=dev, generator=gg)
W = 0.5 * torch.randn(T, N, d, device=dev, generator=gg)
M = torch.zeros(N, H, m, d, device=dev)
hist = 0.5 * torch.randn(N, 2 * H, d, device=dev, generator=gg)
xprev = torch.zeros(N, d, device=dev)
T_SHORT = max(2, T // 3)
CLOCK = [0]


def run(steps):
    r = fused.lds_shared_gpc_rollout(A, B, K, x, steps, H=H, W=W[:steps], M=M, hist=hist, xprev=xprev,
                                     t0=CLOCK[0], lr_scale=2e-6, keep_costs=False, donate=True)
    CLOCK[0] = r.t
    return r


def timed(steps):
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); r = run(steps); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], r


def per_step():
    run(T_SHORT)
    a, _ = timed(T_SHORT)
    b, r = timed(T)
    ok = bool(torch.isfinite(r.cost_sum).all()) and int(r.status.sum()) == 0
    return {"ms_per_step": (b - a) / (T - T_SHORT), "ms_per_call": b, "finite": ok}


bytes_step = 2 * 4 * H * m * d + 4 * 2 * H * d + 2 * 4 * H * m
peak = 6543.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except (OSError, KeyError):
    pass
out = {"workload": f"shared-dynamics LDS(d={d}, m={m}) + GPC(H=HH={H}), {N} envs x T={T}", "hbm_peak_gbs": peak, "runs": {}}
os.environ["DLC_SGPC_PIPE"] = "0"
out["runs"]["serial"] = per_step()
os.environ["DLC_SGPC_PIPE"] = "1"
for c in [int(v) for v in args.ctas.split(",") if v]:
    os.environ["DLC_SGPC_CHAIN_CTAS"] = str(c)
    out["runs"][f"pipeline_chain_ctas_{c}"] = per_step()
for k, v in out["runs"].items():
    v["frac_of_hbm_peak"] = bytes_step * N / (v["ms_per_step"] * 1e-3) / 1e9 / peak
print(json.dumps(out, indent=1))
