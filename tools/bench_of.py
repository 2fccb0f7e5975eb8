"""Throughput of the fused output-feedback rollout on one B200: env-steps/s for GRC / SFC / DSC /
DRC / LQG on the partially observed LDS at the colab's shape (d_hidden = 10, d_action = 3, d_obs = 2) with the
reference's default history lengths, next to the CPU oracle on a bounded sample.  Prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.agents._of import filter_bank  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=148 * 32 * 8 * 4)
ap.add_argument("--T", type=int, default=500)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cpu", type=int, default=1)
ap.add_argument("--only", default="")
ap.add_argument("--lanes", type=int, default=0, help="lanes per env: 0 = automatic, or 8 / 16 / 32")
args = ap.parse_args()
N, T = args.envs, args.T
d, n, p = 10, 3, 2
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.linalg.qr(torch.randn(N, d, d, device=dev, generator=g))[0]
A = (Q * (0.5 + 0.4 * torch.rand(N, 1, d, device=dev, generator=g))) @ Q.transpose(1, 2)
B = torch.randn(N, d, n, device=dev, generator=g) / d ** 0.5
C = torch.randn(N, p, d, device=dev, generator=g) / d ** 0.5
A, B, C = A.contiguous(), B.contiguous(), C.contiguous()
x0 = 0.1 * torch.randn(N, d, device=dev, generator=g)
W = fused.gaussian_disturbance(N, T, d, seed=1, std=0.5)


def flops_fgrc(F, L, m):
    """Reference arithmetic per env-step (FMA = 2 FLOP): get_action, z / y_nat update, policy_loss forward (m+1
    filtered actions + m evolve steps) and its reverse pass, parameter update + projection, env step, cost."""
    act = 2 * F * L * p + 2 * F * n * p                    # filter the window, contract with Theta
    evolve = 2 * d * d + 2 * d * n
    fwd = (m + 1) * act + m * evolve + 2 * p * d
    return (act + evolve + 2 * p * d + fwd + fwd + 4 * F * n * p + evolve + d + 2 * p * d + 2 * (p + n))


def flops_drc(m, h):
    return (2 * h * p * n) * 2 + (2 * m * n * p + 2 * n * p) * 2 + 4 * m * n * p + 2 * d * d + 2 * d * n + d + 2 * p * d


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); r = fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], r


import bench as _bench  # noqa: E402
fp32_peak = max(_bench.measure_fp32_peak(torch, torch.device(dev)).values())
out = {"workload": f"partially observed LDS(d={d}, n={n}, p={p}), {N} envs x T={T}, per-env A/B/C, W[T,N,d] in HBM",
       "fp32_peak_tflops": fp32_peak, "agents": {}}
# DSC's default controller state is 35 KB per env (121 filters x 61 taps): a smaller batch keeps the tool quick
cases = [("grc", dict(m=10), (N, T)), ("sfc", dict(m=30, h=10), (N, T)), ("dsc", dict(m=30, h=10), (148 * 256, 100)),
         ("drc", dict(m=10, h=50), (N, T)), ("drc_lanes", dict(m=10, h=50), (N, T)), ("lqg", {}, (N, T))]
full = (A, B, C, x0, W, N, T)
for name, kw, (Nc, Tc) in cases:
    if args.only and name != args.only:
        continue
    A, B, C, x0, W, N, T = full
    if (Nc, Tc) != (N, T):
        A, B, C, x0 = A[:Nc], B[:Nc], C[:Nc], x0[:Nc]
        W = fused.gaussian_disturbance(Nc, Tc, d, seed=1, std=0.5)
        N, T = Nc, Tc
    if name in ("grc", "sfc", "dsc"):
        Phi = torch.as_tensor(filter_bank(name, kw["m"], kw.get("h")), dtype=torch.float32, device=dev)
        F, L = Phi.shape
        theta = 0.3 / F ** 0.5 * torch.randn(N, F, n, p, device=dev, generator=g)
        run = lambda: fused.lds_of_rollout(A, B, C, x0, T, agent="fgrc", W=W, Phi=Phi, theta=theta, m=kw["m"], R_M=1.0,
                                           keep_actions=False, lanes_per_env=args.lanes,
                                           phi_is_identity=(name == "grc" and not args.lanes))
        fl = flops_fgrc(F, L, kw["m"])
    elif name in ("drc", "drc_lanes"):   # drc: the register-resident kernel (stable_a); drc_lanes: 8 lanes per env
        theta = 0.03 * torch.randn(N, kw["m"], n, p, device=dev, generator=g)
        K = torch.zeros(N, n, p, device=dev)
        run = lambda: fused.lds_of_rollout(A, B, C, x0, T, agent="drc", W=W, K=K, theta=theta, m=kw["m"], h=kw["h"],
                                           lr=0.03, R_M=1000.0, keep_actions=False, lanes_per_env=args.lanes,
                                           stable_a=(name == "drc"))
        fl = flops_drc(kw["m"], kw["h"])
    else:
        K = 0.03 * torch.randn(N, n, d, device=dev, generator=g)
        Lk = 0.03 * torch.randn(N, d, p, device=dev, generator=g)
        run = lambda: fused.lds_of_rollout(A, B, C, x0, T, agent="lqg", W=W, K=K, Lk=Lk, keep_actions=False, lanes_per_env=args.lanes)
        fl = 2 * n * d + 2 * p * d + p + 2 * d * d + 2 * d * n + 2 * d * p + 2 * d + 2 * d * d + 2 * d * n + d + 2 * p * d + 2 * (p + n)
    ms, r = timed(run, args.reps)
    rate = N * T / (ms * 1e-3)
    out["agents"][name] = {
        "params": kw, "envs": N, "T": T, "ms_per_rollout": ms, "env_steps_per_s": rate,
        "finite_envs": int((r.status == 0).sum()), "flops_per_env_step_reference_arithmetic": fl,
        "algorithmic_tflops": rate * fl / 1e12, "frac_of_fp32_peak": rate * fl / 1e12 / fp32_peak,
        "hbm_gbs": rate * (4 * d + 4) / 1e9,
    }

A, B, C, x0, W, N, T = full
if args.cpu:
    from oracle import of as O  # CPU baseline leg only
    Nc, Tc = 2048, 50
    rng = np.random.default_rng(0)
    Ac, Bc, Cc = (v[:Nc].cpu().numpy() for v in (A, B, C))
    Wc = W[:Tc, :Nc].cpu().numpy()
    Phi, L = O.filter_bank("grc", 10)
    th = (0.1 * rng.standard_normal((Nc, 10, n, p))).astype(np.float32)
    t0 = time.perf_counter()
    O.rollout_of(Ac, Bc, Cc, np.zeros((Nc, d), np.float32), Wc, "fgrc", dtype=np.float32, Theta0=th, Phi=Phi, m=10, R_M=1.0)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"agent": "grc", "value": Nc * Tc / dt, "unit": "env-steps/s", "kind": "port",
                           "cores": os.cpu_count(), "sample": f"{Nc} envs x {Tc} steps, oracle/of.py (batched NumPy, fp32), {dt:.1f} s"}
print(json.dumps(out, indent=1))
