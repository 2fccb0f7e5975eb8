"""Where the end-to-end time of the target rollout goes: chunk count sweep of fused.lds_rollout_to_host, with and
without the device -> host copies.  usage: probe_e2e_chunks.py"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("dlc_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
from deluca_b200 import fused  # noqa: E402

wl = dict(bench.WORKLOADS["tgt"])
dev = torch.device("cuda")
N, T = wl["N"], wl["T"]
A, B, K, x0 = bench.make_problem(torch, wl, dev, 0, N)
kw = dict(policy="gpc", H=wl["H"], HH=wl["HH"], lr_scale=bench.LR_SCALE, decay=True, seed=bench.SEED, env_offset=0,
          w_std=bench.W_STD)
host = torch.empty(T, N).pin_memory()


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def streamed(chunks, mat):
    r = fused.lds_rollout_to_host(A, B, K, x0, T, host, chunks=chunks, materialize_disturbance=mat, **kw)
    torch.cuda.current_stream().wait_event(r.done)


def chunked_no_copy(chunks):
    st = dict(M=None, hist=None, xprev=None, uprev=None)
    xc = x0
    for c in range(chunks):
        lo, hi = round(c * T / chunks), round((c + 1) * T / chunks)
        r = fused.lds_rollout(A, B, K, xc, hi - lo, W=None, t0=lo, env_t0=lo, keep_costs=True, donate=c > 0, **st, **kw)
        xc = r.x
        st = {k: r[k] for k in st}


print("one launch, in-kernel RNG, costs kept:", timeit(lambda: fused.lds_rollout(A, B, K, x0, T, W=None, **kw)))
for c in (1, 2, 4, 8, 16):
    print(f"chunks={c:2d}: no copy {timeit(lambda: chunked_no_copy(c)):7.1f} ms | streamed, RNG in kernel "
          f"{timeit(lambda: streamed(c, False)):7.1f} | streamed, W materialised {timeit(lambda: streamed(c, True)):7.1f}")
