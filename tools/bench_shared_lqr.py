#!/usr/bin/env python
"""Shared-dynamics LDS(d=256, m=64) + LQR on tcgen05 (3xTF32): kernel time per launch shape (DLC_SHARED_VARIANT: 0 = two
32-env CTAs per SM, 1 = one 64-env CTA per SM) and cluster size (DLC_SHARED_CLUSTER).  Prints one JSON object."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import torch
    from deluca_b200 import fused
    from deluca_b200.riccati import dare_gain
    dev = "cuda"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except OSError:
        pass
    d5, m5 = 256, 64
    gg = torch.Generator(device=dev).manual_seed(0)
    Qm = torch.linalg.qr(torch.randn(d5, d5, device=dev, generator=gg, dtype=torch.float64))[0]
    A5 = ((Qm * (0.7 + 0.25 * torch.rand(d5, device=dev, generator=gg, dtype=torch.float64))) @ Qm.T).float().contiguous()
    B5 = (torch.randn(d5, m5, device=dev, generator=gg) / d5 ** 0.5).contiguous()
    K5 = dare_gain(A5, B5)[0].contiguous()
    G5 = fused.lds_shared_pack(A5, B5, K5)
    F5 = 2 * d5 * d5 + 2 * d5 * m5 + d5 + 2 * m5 * d5 + 2 * d5 + 2 * m5
    tens_peak = peaks.get("bf16_tflops", 1590.0) / 2.0
    out = {}
    for N5, T5, with_w in ((148 * 64 * 8, 50, True), (1 << 20, 20, False)):
        x5 = torch.randn(N5, d5, device=dev, generator=gg)
        W5 = (0.5 * torch.randn(T5, N5, d5, device=dev, generator=gg)) if with_w else None
        ts = []
        for i in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r5 = fused.lds_shared_lqr_rollout(G5, x5, T5, m5, W=W5, keep_costs=True); e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        exe = 3 * 2 * 384 * d5 * N5 * T5 / (ms * 1e-3) / 1e12
        out[f"N{N5}_{'W' if with_w else 'noW'}"] = {
            "envs": N5, "T": T5, "kernel_ms": ms, "env_steps_per_s": N5 * T5 / (ms * 1e-3),
            "us_per_64env_step": ms * 1e3 / T5 / (N5 / 64 / 148), "finite": int(r5.status.sum()) == 0,
            "algorithmic_tflops": F5 * N5 * T5 / (ms * 1e-3) / 1e12, "executed_tensor_tflops": exe,
            "frac_of_dense_tf32": exe / tens_peak}
        del x5, W5, r5
        torch.cuda.empty_cache()
    print(json.dumps(out))
    sys.exit(0)

res = {}
for variant in (0, 1):
    for cluster in (1, 2, 4):
        env = dict(os.environ, DLC_SHARED_VARIANT=str(variant), DLC_SHARED_CLUSTER=str(cluster))
        p = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        res[f"variant{variant}_cluster{cluster}"] = json.loads(line[-1]) if line else {"error": p.stderr[-400:]}
print(json.dumps(res, indent=1))
