#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 3 and 4) measured on one B200: kernel time by CUDA
events, achieved fraction of the binding roofline, and the CPU oracle on a bounded sample.
Prints one JSON object; results are recorded in profiles/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deluca_b200 import fused
from deluca_b200.envs.classic import Cartpole, Pendulum, QuadCost
from deluca_b200.lung import _launch
from deluca_b200.lung.core import BreathWaveform
from deluca_b200.lung.envs import BalloonLung, DelayLung

dev = "cuda"
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except OSError:
    pass
HBM = peaks.get("hbm_gbs", 6650.0)
out = {"hbm_peak_gbs": HBM, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


# ---- config 4: BalloonLung / DelayLung + PID, 2^20 breaths
N = 1 << 20
g = torch.Generator().manual_seed(0)
K = 5 * torch.rand(N, 3, generator=g)
pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
wave = BreathWaveform.create(pip=pip, peep=5.0)
for name, Env in (("balloon", BalloonLung), ("delay", DelayLung)):
    for T in (29, 1000):
        env = Env.create()
        st, obs = env.reset()
        p, b, series = _launch.prepare(env, wave, N, T, "closed_loop", dev, K=K, env_state=env.state_dict(st),
                                       obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time))
        ms = timeit(lambda: fused.lung_rollout(p, b))
        byts = 24.0 * N * T   # six fp32 series per env-step (20 B + the target series)
        out[f"lung_{name}_T{T}"] = {"envs": N, "T": T, "kernel_ms": ms, "env_steps_per_s": N * T / (ms * 1e-3),
                                    "roofline": {"bound": "hbm", "achieved": byts / (ms * 1e-3) / 1e9, "peak": HBM,
                                                 "unit": "GB/s", "frac": byts / (ms * 1e-3) / 1e9 / HBM,
                                                 "bytes_per_env_step": 24}}
        del p, b, series
        torch.cuda.empty_cache()
if os.environ.get("DLC_EXTRA_ONLY") == "lung":
    print(json.dumps(out, indent=1))
    sys.exit(0)

# ---- 8(f) rank 1: value + gradient of train_controller.rollout w.r.t. the PID gains, 2^20 breaths x T=29
from deluca_b200 import _abi
from deluca_b200.lung.core import DEFAULT_DT
import ctypes
env = BalloonLung.create()
pg = _abi.DlcLungParams()
pg.N, pg.T, pg.mode = N, 29, 0
kx, kf, period, xp_in = wave.knots(N=N, device=dev)
pg.nk = len(kx)
for i, v in enumerate(kx):
    pg.kx[i] = v
pg.kf_per_env = int(kf.dim() == 2)
pg.period, pg.xp_in, pg.default_dt, pg.pid_decay = period, xp_in, DEFAULT_DT, 0.03 / 0.53
_launch.sim_params(env, pg)
Kd = K.to(dev).contiguous()
ttd = torch.linspace(0, 0.87, 29, device=dev)
ms = timeit(lambda: fused.lung_pid_value_and_grad(pg, Kd, kf, ttd))
byts = 2 * 6 * 4.0 * N * 29   # checkpoint written by the forward pass and read by the backward pass
out["lung_balloon_pid_value_and_grad_T29"] = {
    "envs": N, "T": 29, "kernel_ms": ms, "env_steps_per_s": N * 29 / (ms * 1e-3),
    "roofline": {"bound": "hbm", "achieved": byts / (ms * 1e-3) / 1e9, "peak": HBM, "unit": "GB/s",
                 "frac": byts / (ms * 1e-3) / 1e9 / HBM, "bytes_per_env_step": 48}}
del Kd, kf
torch.cuda.empty_cache()

# ---- config 3: Pendulum / Cartpole + iLQR, IGPC: 4096 trajectories x H=200 x 20 iterations
N, H, T = 4096, 200, 20
for name, Env in (("pendulum", Pendulum), ("cartpole", Cartpole)):
    env = Env.create(H=H)
    cost = QuadCost(list(env.goal_state), 1.0, 0.1)
    gg = torch.Generator(device=dev).manual_seed(0)
    x0 = env.init(N)[0].arr + 0.05 * torch.randn(N, env.state_dim, device=dev, generator=gg)
    U0 = torch.zeros(N, H, 1, device=dev)
    for algo, label, kw in ((0, "ilqr_as_written", dict(alpha0=0.5, n_alpha=10)),
                            (0, "ilqr_fixed_backtracking", dict(alpha0=0.5, n_alpha=10, fix_backtracking=True)),
                            (1, "igpc", dict(alpha0=1.0, n_alpha=6, lr=0.001))):
        p = fused.plan_params(env, env, cost, N, H, T=T, algo=algo, **kw)
        ms = timeit(lambda: fused.plan(p, U0, x0), reps=3, warm=1)
        c = fused.plan(p, U0, x0)[4]
        out[f"{label}_{name}"] = {"trajectories": N, "H": H, "iterations": T, "kernel_ms": ms,
                                  "trajectory_steps_iterations_per_s": N * H * T / (ms * 1e-3),
                                  "mean_final_cost": float(c.mean()), "finite": bool(torch.isfinite(c).all()),
                                  "bound": "instruction latency (serial T x H dependency chain per trajectory)"}

# ---- config 5 (dense part): shared-dynamics LDS d=256, m=64 + LQR on tcgen05 (3xTF32)
from deluca_b200.riccati import dare_gain
d5, m5 = 256, 64
gg = torch.Generator(device=dev).manual_seed(0)
Qm = torch.linalg.qr(torch.randn(d5, d5, device=dev, generator=gg, dtype=torch.float64))[0]
A5 = ((Qm * (0.7 + 0.25 * torch.rand(d5, device=dev, generator=gg, dtype=torch.float64))) @ Qm.T).float().contiguous()
B5 = (torch.randn(d5, m5, device=dev, generator=gg) / d5 ** 0.5).contiguous()
K5 = dare_gain(A5, B5)[0].contiguous()
G5 = fused.lds_shared_pack(A5, B5, K5)
F5 = 2 * d5 * d5 + 2 * d5 * m5 + d5 + 2 * m5 * d5 + 2 * d5 + 2 * m5      # env + action + cost FLOPs per env-step
tens_peak = peaks.get("bf16_tflops", 1590.0) / 2.0                        # dense TF32 = half the bf16 rate
for N5, T5, with_w in ((148 * 64 * 8, 50, True), (1 << 20, 20, False)):
    x5 = torch.randn(N5, d5, device=dev, generator=gg)
    W5 = (0.5 * torch.randn(T5, N5, d5, device=dev, generator=gg)) if with_w else None
    ms = timeit(lambda: fused.lds_shared_lqr_rollout(G5, x5, T5, m5, W=W5, keep_costs=True), reps=3, warm=2)
    r5 = fused.lds_shared_lqr_rollout(G5, x5, T5, m5, W=W5)
    ach = F5 * N5 * T5 / (ms * 1e-3) / 1e12
    exe = 3 * 2 * 384 * d5 * N5 * T5 / (ms * 1e-3) / 1e12                 # executed tensor FLOPs: 3 passes, 384 padded rows
    out[f"shared_lqr_d256_m64_N{N5}_{'W' if with_w else 'noW'}"] = {
        "envs": N5, "T": T5, "kernel_ms": ms, "env_steps_per_s": N5 * T5 / (ms * 1e-3), "finite": int(r5.status.sum()) == 0,
        "roofline": {"bound": "tensor", "achieved": ach, "executed_tensor_tflops": exe, "peak": tens_peak,
                     "unit": "TFLOP/s", "frac": exe / tens_peak, "flops_per_env_step": F5,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (dense TF32 rate; not measured directly)"}}
    del x5, W5, r5
    torch.cuda.empty_cache()

# ---- CPU oracle on bounded samples
from oracle import igpc as OG
from oracle import lung as OL
Nc = 4096
t0 = time.perf_counter()
OL.run_controller_scan(np.tile([3.0, 4.0, 0.0], (Nc, 1)), 35.0, 5.0, T=200, sim="balloon")
dt = time.perf_counter() - t0
out["cpu_lung_balloon"] = {"env_steps_per_s": Nc * 200 / dt, "cores": 1, "kind": "port",
                           "sample": f"{Nc} breaths x 200 steps, oracle/lung.py (NumPy), {dt:.1f} s"}
oe = OG.Pendulum(H=50)
oc = OG.QuadCost(oe.goal_state, 1.0, 0.1)
t0 = time.perf_counter()
OG.ilqr(oe, oc, np.zeros((256, 50, 1), np.float32), 5, x0=oe.init(256, np.float32))
dt = time.perf_counter() - t0
out["cpu_ilqr_pendulum"] = {"trajectory_steps_iterations_per_s": 256 * 50 * 5 / dt, "cores": 1, "kind": "port",
                            "sample": f"256 trajectories x H=50 x 5 iterations, oracle/igpc.py (NumPy), {dt:.1f} s"}
print(json.dumps(out, indent=1))
