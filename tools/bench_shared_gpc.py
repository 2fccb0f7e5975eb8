#!/usr/bin/env python
"""BASELINE config 5 with the per-env policy tensor: shared-dynamics LDS (d=256, m=64) + GPC (H=HH=16).
Times dlc_lds_shared_gpc_rollout_f32 with CUDA events (whole call, and each of its two kernels alone through the
DLC_SGPC_DBG profiling switch) and reports the M-stream kernel against the measured HBM peak and the chain kernel
against the measured FP32 peak.  Prints one JSON object; results are recorded in profiles/.

Multi-GPU (weak scaling, one process per GPU, no data-path collective: the envs are independent and every rank holds the
shared operators): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P tools/bench_shared_gpc.py.  Each rank carries --envs envs (56,832 by default: their policy tensors are
59.6 GB of the GPU's 180 GB next to the same-sized staging a caller would keep; BASELINE.json's 2^20 envs are 1.1 TB of M
and need 8 GPUs at 131,072 envs = 137 GB each).  Times are the MAX over ranks; the line reports the whole job."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from deluca_b200 import fused
from deluca_b200.riccati import dare_gain

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=148 * 64 * 6)   # whole waves of the chain kernel (64 envs per CTA)
ap.add_argument("--T", type=int, default=12)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except OSError:
    pass
HBM = peaks.get("hbm_gbs", 6650.0)
d, m, H = 256, 64, 16
N, T = args.envs, args.T
gg = torch.Generator(device=dev).manual_seed(0)          # the shared operators: the same on every rank
Qm = torch.linalg.qr(torch.randn(d, d, device=dev, generator=gg, dtype=torch.float64))[0]
A = ((Qm * (0.7 + 0.25 * torch.rand(d, device=dev, generator=gg, dtype=torch.float64))) @ Qm.T).float().contiguous()
B = (torch.randn(d, m, device=dev, generator=gg) / d ** 0.5).contiguous()
K = dare_gain(A, B)[0].contiguous()
gg.manual_seed(1000 + rank)                               # this rank's envs
x = torch.randn(N, d, device=dev, generator=gg)
W = 0.5 * torch.randn(T, N, d, device=dev, generator=gg)
M = torch.zeros(N, H, m, d, device=dev)
hist = 0.5 * torch.randn(N, 2 * H, d, device=dev, generator=gg)   # a warm noise history: every window is live
xprev = torch.zeros(N, d, device=dev)


T_SHORT = max(2, T // 3)


CLOCK = [0]   # agent step counter carried across calls (lr = lr_scale / (t + 1)), so the timed calls are one long episode


def run(steps):
    r = fused.lds_shared_gpc_rollout(A, B, K, x, steps, H=H, W=W[:steps], M=M, hist=hist, xprev=xprev,
                                     t0=CLOCK[0], lr_scale=2e-6, keep_costs=False, donate=True)
    CLOCK[0] = r.t
    return r


def timed(steps, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(); r = run(steps); e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ts.append(float(ms))
    return sorted(ts)[len(ts) // 2], r


def per_step(reps):
    """ms per step with the per-call setup (workspace memset, operator packing, state transposes) differenced out."""
    run(T_SHORT)
    a, _ = timed(T_SHORT, reps)
    b, r = timed(T, reps)
    return (b - a) / (T - T_SHORT), b, r


ms_step, ms_all, r = per_step(args.reps)
finite = bool(torch.isfinite(r.cost_sum).all()) and int(r.status.sum()) == 0
os.environ["DLC_SGPC_DBG"] = "1"
ms_stream_step, _, _ = per_step(args.reps)
os.environ["DLC_SGPC_DBG"] = "2"
ms_chain_step, _, _ = per_step(args.reps)
del os.environ["DLC_SGPC_DBG"]

bytes_step = 2 * 4 * H * m * d + 4 * 2 * H * d + 2 * 4 * H * m          # M read + write, noise ring, G in, V out
flops_stream = 2 * (H + 1) * H * m * d                                    # rank-H update + the newest window product
flops_chain = (2 * m * d + 2 * d * (d + m)) + 2 * d * (H - 1) * m + 2 * 2 * m * d \
    + 2 * (H - 1) * m * d + 2 * d * 2 * d        # env step, y = Gcat Vcat, du / lam, g = Gcat' lam, the sliding q update
per_stream = ms_stream_step * 1e-3 / N
per_chain = ms_chain_step * 1e-3 / N
import bench as _bench
fp32_peak = max(_bench.measure_fp32_peak(torch, torch.device(dev)).values())
out = {
    "workload": f"shared-dynamics LDS(d={d}, m={m}) + GPC(H=HH={H}), per-env M (1 MiB), {N} envs per GPU x T={T}",
    "n_gpus": world, "scaling": "weak", "envs": N * world, "envs_per_gpu": N, "T": T, "ms_per_call": ms_all,
    "ms_per_step": ms_step, "env_steps_per_s": N * world / (ms_step * 1e-3),
    "env_steps_per_s_whole_call": N * world * T / (ms_all * 1e-3), "finite": finite,
    "timing": f"CUDA events; per-step = (call with T={T} - call with T={T_SHORT}) / {T - T_SHORT}, median of {args.reps}",
    "M_bytes_resident": N * H * m * d * 4,
    "stream_kernel": {"ms_per_launch": ms_stream_step, "ns_per_env_step": per_stream * 1e9,
                      "roofline": {"bound": "hbm", "achieved": bytes_step / per_stream / 1e9, "peak": HBM,
                                   "unit": "GB/s", "frac": bytes_step / per_stream / 1e9 / HBM,
                                   "bytes_per_env_step": bytes_step, "fp32_tflops": flops_stream / per_stream / 1e12}},
    "chain_kernel": {"ms_per_launch": ms_chain_step, "ns_per_env_step": per_chain * 1e9,
                     "roofline": {"bound": "fp32", "achieved": flops_chain / per_chain / 1e12, "peak": fp32_peak,
                                  "unit": "TFLOP/s", "frac": (flops_chain / per_chain / 1e12 / fp32_peak) if fp32_peak else None,
                                  "flops_per_env_step": flops_chain}},
    "whole_step_vs_hbm": {"achieved": bytes_step * N / (ms_step * 1e-3) / 1e9, "peak": HBM,
                          "frac": bytes_step * N / (ms_step * 1e-3) / 1e9 / HBM},
}
if rank == 0:
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
