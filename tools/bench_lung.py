#!/usr/bin/env python
"""BASELINE.json configs[3] (BalloonLung / DelayLung + PID, 2^20 breaths) on N GPUs of one node.

    python tools/bench_lung.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_lung.py --gpus N [--scaling weak|strong]

The breaths are independent: each rank takes a contiguous shard (`shard_range`), there is no data-path collective;
timing is CUDA events around K launches bracketed by a barrier, MAX over ranks.  One JSON line per (simulator, T) from
rank 0.  Roofline: the kernel only stores -- six float32 series per env-step (24 B), or the four that differ
between the breaths of a batch (16 B: `timestamp` depends on the step only and `flow` is the env's static field, so
`run_controller_scan` keeps one column of each) -- plus, once per launch, the carried pytree leaves read and written
back (`state_bytes_per_env`); `frac` is on the per-step figure, `frac_with_state_io` adds the state traffic the
functional contract requires.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from deluca_b200 import fused
from deluca_b200.lung import _launch
from deluca_b200.lung.core import BreathWaveform
from deluca_b200.lung.envs import BalloonLung, DelayLung
from deluca_b200.sharding import shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except OSError:
    pass
HBM = peaks.get("hbm_gbs", 6650.0)
total = args.envs * (world if args.scaling == "weak" else 1)
lo, hi = shard_range(total, rank, world)
n = hi - lo
g = torch.Generator().manual_seed(1234 + rank)
K = 5 * torch.rand(n, 3, generator=g)
pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (n,), generator=g)]
wave = BreathWaveform.create(pip=pip, peep=5.0)
flush = torch.empty(64 << 20, device=dev)                       # 256 MB > L2: written between timed launches
for name, Env in (("balloon", BalloonLung), ("delay", DelayLung)):
  for lean in (True, False):     # four series per breath (what run_controller_scan stores for a batch) / all six
    for T in (29, 1000):
        env = Env.create(device=str(dev))
        st, obs = env.reset()
        p, b, series = _launch.prepare(env, wave, n, T, "closed_loop", dev, K=K, env_state=env.state_dict(st),
                                       obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time), lean=lean)
        for _ in range(args.warmup):
            fused.lung_rollout(p, b)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fused.lung_rollout(p, b); e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= args.steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        # carried leaves: K 12 + kf 28 read; 13 float + 0 int leaves read, 15 float + 2 int (rmw) + status written
        state = 12 + 28 + 13 * 4 + (15 + 2 * 2 + 1) * 4
        if name == "delay":     # + pipe_pressure (rw) and the two [delay] histories (written; read only if T < delay)
            state += 8 + 2 * 4 * env.delay * (1 if T >= env.delay else 2)
        per_step = 16 if lean else 24
        series_b = float(per_step) * n * T
        ach = series_b / (ms * 1e-3) / 1e9
        if rank == 0:
            print(json.dumps({
                "metric": "env_steps_per_s", "workload": f"{name} lung + PID, {total} breaths x T={T}, "
                                                         f"{'four' if lean else 'six'} series stored per breath",
                "value": total * T / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world, "scaling": args.scaling,
                "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "l2": "256 MB flush between launches",
                "roofline": {"bound": "hbm", "achieved": ach, "peak": HBM, "unit": "GB/s", "frac": ach / HBM,
                             "bytes_per_env_step": per_step, "state_bytes_per_env": state,
                             "frac_with_state_io": (series_b + state * n) / (ms * 1e-3) / 1e9 / HBM}}))
        del p, b, series
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
