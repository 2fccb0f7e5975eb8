#!/usr/bin/env python
"""Headline benchmark: LDS + GPC env-steps/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: the fused T-step rollout of all envs of the
workload (default = BASELINE.json configs[1]: 65,536 envs x T=10,000, LDS d=10 m=3, GPC H=3 HH=10).
One JSON line is printed by rank 0; see DESIGN.md section 6 for how each key is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(N=65536, T=10000, d=10, m=3, H=3, HH=10,
               desc="LDS(d=10,m=3)+GPC(H=3,HH=10) 65536 envs x T=10000, per-env A/B/K, Gaussian w (std 0.5)"),
    # BASELINE.json north_star target
    "tgt": dict(N=1 << 20, T=1000, d=10, m=3, H=10, HH=10,
                desc="LDS(d=10,m=3)+GPC(H=10,HH=10) 1M envs x T=1000"),
    "smoke": dict(N=4096, T=200, d=10, m=3, H=3, HH=10, desc="small test workload"),
}
# The reference default lr_scale=0.005 makes GPC-as-written diverge on the default LDS
# (tests/test_oracle_lds.py::test_reference_noise_quirk...); 1e-4 keeps every env finite so that
# the rollout that is timed is also the rollout whose parity is tested.
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed
# `ncu --set full` capture of this same command (profiles/r01_lds_f2_32x9_full.txt): 26.33 + 2.68 GB
# against 28.9 GB of algorithmic bytes.  Other workloads / modes have no capture -> null.
NCU_TRAFFIC_BYTES = {("c2", "hbm"): 26.330129e9 + 2.676375e9}
LR_SCALE = 1e-4
W_STD = 0.5
SEED = 20240


def flops_per_env_step(d, m, H):
    """algorithmic FLOPs of one LDS+GPC env-step (reference arithmetic, FMA = 2)."""
    env = 2 * d * d + 2 * d * m + d
    act = 2 * m * d + 2 * H * m * d
    noise = 2 * d * d + 2 * d * m + 2 * d
    fwd = (H - 1) * (2 * m * d + 2 * H * m * d + m + 2 * d * d + 2 * d * m + d) + (2 * m * d + 2 * H * m * d + 3 * m + 2 * d)
    bwd = (2 * m * d + 2 * H * m * d + m + 2 * d) + (H - 1) * (2 * d * m + 2 * H * m * d + 2 * d * d + 2 * m * d + d)
    return env + act + noise + fwd + bwd + 2 * H * m * d + (2 * d + 2 * m)


def bytes_per_env_step(d, m, H, T, w_from_hbm):
    """algorithmic HBM bytes of one env-step (s = 0: no sampled trajectories)."""
    return (4 * d if w_from_hbm else 0) + 4 + (4 * (d * d + 2 * d * m) + 4 * H * m * d) / T


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def make_problem(torch, wl, dev, rank):
    """Synthetic per-env systems with the reference's default distributions (_lds.py:69-91):
    A = diag(U(.9,.95)) stored and treated as a dense [d,d] matrix, B ~ N(0,1), K = LQR gain."""
    from deluca_b200.riccati import dare_gain
    N, d, m = wl["N"], wl["d"], wl["m"]
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device=dev, generator=g))
    B = torch.randn(N, d, m, device=dev, generator=g)
    K = torch.empty(N, m, d, device=dev)
    for s in range(0, N, 1 << 16):
        K[s:s + (1 << 16)] = dare_gain(A[s:s + (1 << 16)], B[s:s + (1 << 16)])[0]
    x0 = torch.zeros(N, d, device=dev)
    return A.contiguous(), B.contiguous(), K.contiguous(), x0


def measure_fp32_peak(torch, dev):
    """FP32 FMA peak (TFLOP/s) measured live with dlc_microbench_f32 (FFMA and packed FFMA2)."""
    from deluca_b200 import _abi
    lib = _abi.lib()
    sink = torch.zeros(64, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    out = {}
    for name, which, fma_per_iter in (("ffma", 0, 72), ("ffma2", 1, 72)):
        iters, blocks = 20000, sms * 8
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.check(lib.dlc_microbench_f32(which, _abi.ptr(sink), iters, blocks, _abi.current_stream_ptr()),
                       "dlc_microbench_f32")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep:
                best = max(best, 2.0 * fma_per_iter * iters * blocks * 256 / (ms * 1e-3) / 1e12)
        out[name] = best
    return out


def cpu_baseline(wl, seconds_target=12.0):
    """The oracle's C restatement (oracle/lds_c.c, OpenMP over envs, every host thread) timed on a
    bounded sample of the workload: same step, same distributions, fewer envs and steps."""
    import numpy as np
    from oracle.lds_c import gpc_rollout_c, max_threads
    # every host thread this process may use: torchrun exports OMP_NUM_THREADS=1, which omp_get_max_threads() obeys,
    # so count the affinity mask instead and hand the number to omp_set_num_threads() explicitly
    try:
        cores = max(max_threads(), len(os.sched_getaffinity(0)))
    except AttributeError:
        cores = max(max_threads(), os.cpu_count() or 1)
    d, m = wl["d"], wl["m"]
    Nc = 2048 * cores
    rng = np.random.default_rng(0)
    A = np.zeros((Nc, d, d), np.float32)
    A[:, np.arange(d), np.arange(d)] = rng.uniform(0.9, 0.95, (Nc, d))
    B = rng.standard_normal((Nc, d, m)).astype(np.float32)
    K = (0.05 * rng.standard_normal((Nc, m, d))).astype(np.float32)
    x0 = np.zeros((Nc, d), np.float32)

    def run(T):
        W = (0.5 * rng.standard_normal((T, Nc, d))).astype(np.float32)
        t0 = time.perf_counter()
        gpc_rollout_c(A, B, K, x0, W, wl["H"], wl["HH"], LR_SCALE, True, np.float32, cores)
        return time.perf_counter() - t0

    run(8)
    Tc = 500
    dt1 = run(Tc)
    reps = int(max(1, min(40, round(seconds_target / max(dt1, 1e-3)))))
    dt = sum(run(Tc) for _ in range(reps))
    return {"value": Nc * Tc * reps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "seconds": dt, "reps": reps,
            "sample": f"{reps} x ({Nc} envs x {Tc} steps) of the same LDS+GPC step (oracle/lds_c.c: scalar C, OpenMP "
                      f"over envs, fp32, {cores} threads), {dt:.1f} s"}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cb = cpu_baseline(wl, seconds_target=min(60.0, 8.0 * max(1, args.steps)))
    line = {"impl": "reference", "metric": "LDS+GPC env-steps/sec", "value": cb["value"], "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cb["seconds"] / cb["reps"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "note": "reference semantics restated on the CPU (JAX is not "
                       "installable in this image, SURVEY F2); bounded sample, see cpu_baseline.sample"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override env count (testing)")
    ap.add_argument("--horizon", type=int, default=0, help="override T (testing)")
    ap.add_argument("--disturbance", default="hbm", choices=["hbm", "kernel"],
                    help="hbm: W[T,N,d] resident in HBM before timing; kernel: Philox drawn in-register")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="DlcLdsParams.kernel_variant (0 = auto)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.envs:
        wl["N"] = args.envs
    if args.horizon:
        wl["T"] = args.horizon
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    from deluca_b200 import _abi
    from deluca_b200.fused import gaussian_disturbance, lds_rollout
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _abi.check(_abi.lib().dlc_check_device(), "dlc_check_device")
    if world > 1:
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout ahead of the JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    N, T, d, m, H, HH = (wl[k] for k in ("N", "T", "d", "m", "H", "HH"))
    env_offset = rank * N  # weak scaling: every rank owns N envs of a world*N-env job
    A, B, K, x0 = make_problem(torch, wl, dev, rank)
    use_hbm_w = args.disturbance == "hbm"
    W = gaussian_disturbance(N, T, d, seed=SEED, env_offset=env_offset, std=W_STD, device=dev) if use_hbm_w else None
    kw = dict(policy="gpc", H=H, HH=HH, lr_scale=LR_SCALE, decay=True, seed=SEED, env_offset=env_offset, w_std=W_STD,
              kernel_variant=args.variant)
    stats = torch.zeros(4, dtype=torch.float64, device=dev)

    from deluca_b200.sharding import local_stats, reduce_stats, summarize

    def step():
        r = lds_rollout(A, B, K, x0, T, W=W, **kw)
        # final statistics [sum cost, sum cost^2, finite envs, non-finite envs], one tiny all-reduce
        # (NCCL over NVLink) per rollout -- the only collective on this path
        stats.copy_(reduce_stats(local_stats(r.cost_sum, r.status)))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler (nvidia-smi -lms) is started BEFORE the warm-up: its first query on a fresh box
    # takes long enough to stall a kernel, which must not land inside the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 0)):
        step()
    # ---- timed region: K steps; kernel time on the launch stream via CUDA events
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record()
    torch.cuda.nvtx.range_push("dlc_timed")
    r = None
    for i in range(args.steps):
        # drop the previous step's outputs first: holding them while the next rollout allocates its
        # 2.6 GB cost buffer makes torch's caching allocator cudaMalloc a second block inside the timed
        # region (seen as a one-off +25..85 ms "kernel" time on the second step)
        r = None
        kev[i][0].record()
        r = lds_rollout(A, B, K, x0, T, W=W, **kw)
        kev[i][1].record()
        stats.copy_(reduce_stats(local_stats(r.cost_sum, r.status)))
    torch.cuda.nvtx.range_pop()
    t_all1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_all0.elapsed_time(t_all1)
    kern_each = [a.elapsed_time(b) for a, b in kev]
    kern_ms = sum(kern_each) / max(args.steps, 1)
    tm = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(tm[0]), float(tm[1])
    st = stats.cpu().tolist()

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    hA, hB, hK, hx = (t.cpu().pin_memory() for t in (A, B, K, x0))
    h_out = torch.empty(N, dtype=torch.float64).pin_memory()
    kw_e2e = dict(kw)
    def e2e_step():
        dA, dB, dK, dx = (t.to(dev, non_blocking=True) for t in (hA, hB, hK, hx))
        rr = lds_rollout(dA, dB, dK, dx, T, W=None, keep_costs=False, donate=True, **kw_e2e)
        h_out.copy_(rr.cost_sum, non_blocking=True)
    e2e_step(); sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(1, min(args.steps, 3))
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    sync()
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / n_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms[0])
    h2d = sum(t.numel() * t.element_size() for t in (hA, hB, hK, hx))
    d2h = h_out.numel() * h_out.element_size()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        fp32 = measure_fp32_peak(torch, dev)
        fp32_peak = max(fp32.values())
        F = flops_per_env_step(d, m, H)
        Bts = bytes_per_env_step(d, m, H, T, use_hbm_w)
        ach = F * N * T / (kern_ms * 1e-3) / 1e12
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "LDS+GPC env-steps/sec", "value": world * N * T * args.steps / (total_ms * 1e-3),
            "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "envs_per_gpu": N, "T": T, "d": d, "m": m, "H": H, "HH": HH,
                       "lr_scale": LR_SCALE, "decay": True,
                       "disturbance": "W[T,N,d] resident in HBM (%.1f GB, larger than L2)" % (N * T * d * 4 / 1e9)
                       if use_hbm_w else "Philox4x32-10 + Box-Muller drawn in-register",
                       "l2": "per-step inputs exceed L2" if use_hbm_w else "state is on-chip; costs written once",
                       "parallelism": f"env-sharded x{world}"},
            "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": ach / fp32_peak if fp32_peak else None,
                         "traffic": NCU_TRAFFIC_BYTES.get((args.workload, args.disturbance))
                         if not (args.envs or args.horizon) else None,
                         "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_lds_f2_32x9_full.txt)",
                         "algorithmic_bytes": Bts * N * T,
                         "kernel": "lds_f2_kernel" if args.variant in (0, 8, 9, 10, 11, 12) else "lds_split_kernel", "kernel_ms": kern_ms,
                         "kernel_ms_each": [round(v, 3) for v in kern_each],
                         "flops_per_env_step": F, "bytes_per_env_step": Bts,
                         "peak_source": "measured live: dlc_microbench_f32 (FFMA %.1f, FFMA2 %.1f TFLOP/s); "
                                        "MEASURED_PEAKS.json has no FP32 non-tensor peak" % (fp32["ffma"], fp32["ffma2"]),
                         "hbm": {"achieved": Bts * N * T / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}},
            "e2e": {"value": world * N * T / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "note": "host A,B,K,x0 (pinned) -> device, fused rollout with in-kernel disturbances, "
                            "per-env cost sums -> host"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "stats": summarize(stats.cpu(), T),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
