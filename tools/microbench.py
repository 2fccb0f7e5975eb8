#!/usr/bin/env python
"""Print the pipe rates measured by dlc_microbench_f32 (run on the B200 box)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deluca_b200 import _abi

lib = _abi.lib()
dev = torch.device("cuda", 0)
sink = torch.zeros(64, device=dev)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
# name, which, ops per thread per iter, unit scale
CASES = [("ffma", 0, 72), ("ffma2_packed", 1, 72), ("lds32", 2, 32), ("shfl", 3, 32), ("mix_ffma_lds", 4, 80)]
out = {"sms": sms}
for bps in (2, 4, 8):
    for name, which, ops in CASES:
        iters, blocks = 20000, sms * bps
        best = 1e30
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.check(lib.dlc_microbench_f32(which, _abi.ptr(sink), iters, blocks, _abi.current_stream_ptr()), name)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        rate = ops * iters * blocks * 256 / (best * 1e-3)  # thread-ops per second
        out[f"{name}@{bps * 8}warps/SM"] = {"ms": best, "Gops_per_s": rate / 1e9,
                                            "per_SM_per_clk_at_1.9GHz": rate / sms / 1.9e9}
print(json.dumps(out, indent=1))
