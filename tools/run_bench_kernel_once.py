"""The bench.py headline launch, once, for `ncu --set full` (dram traffic of the dominant kernel per launch).
usage: run_bench_kernel_once.py [tgt|c2]"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("dlc_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
name = sys.argv[1] if len(sys.argv) > 1 else "tgt"
wl = dict(bench.WORKLOADS[name])
dev = torch.device("cuda", 0)
run = bench.Runner(torch, None, wl, dev, 0, 1, "strong", True, 0)
for _ in range(2):
    r = run.rollout(run.W)
torch.cuda.synchronize()
print(name, float(r.cost_sum[torch.isfinite(r.cost_sum)].mean()))
