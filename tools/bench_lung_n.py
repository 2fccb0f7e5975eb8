"""Lung + PID episode kernel time vs the batch size N (row pitch of the [T,N] series): probes whether power-of-two
pitches camp on DRAM partitions.  usage: bench_lung_n.py [T] [kernel_variant] [N,N,...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deluca_b200 import fused  # noqa: E402
from deluca_b200.lung import _launch  # noqa: E402
from deluca_b200.lung.core import BreathWaveform  # noqa: E402
from deluca_b200.lung.envs import BalloonLung, DelayLung  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
VARIANT = int(sys.argv[2]) if len(sys.argv) > 2 else 0
NS = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else (1 << 20, (1 << 20) + 128, (1 << 20) + 4096 + 128, 1000000, 148 * 7 * 1024)
out = {}
for Env in (BalloonLung, DelayLung):
    for N in NS:
        g = torch.Generator().manual_seed(0)
        K = 5 * torch.rand(N, 3, generator=g)
        pip = torch.tensor([10.0, 15, 20, 25, 30, 35])[torch.randint(0, 6, (N,), generator=g)]
        wave = BreathWaveform.create(pip=pip, peep=5.0)
        env = Env.create()
        st, obs = env.reset()
        p, b, series = _launch.prepare(env, wave, N, T, "closed_loop", "cuda", K=K, env_state=env.state_dict(st),
                                       obs=dict(predicted_pressure=obs.predicted_pressure, time=obs.time), kernel_variant=VARIANT)
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fused.lung_rollout(p, b); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[f"{Env.__name__}_N{N}"] = {"ms": ms, "GB_per_s": 24.0 * N * T / (ms * 1e-3) / 1e9}
        del p, b, series
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
