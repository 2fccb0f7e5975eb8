"""What HBM takes from the time-major series store pattern (dlc_microbench_f32 which = 5, 6, 7)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deluca_b200 import _abi

lib = _abi.lib()
out = {}
for T in (29, 1000):
    blocks = 8192
    N = blocks * 128
    buf = torch.empty(6 * T * N, device="cuda")
    for name, which, ns in (("six_series_8fma", 5, 6), ("six_series_100fma", 6, 6), ("one_series_8fma", 7, 1),
                            ("six_vec4", 10, 6), ("six_vec2", 11, 6), ("six_cs", 12, 6), ("six_wt", 13, 6),
                            ("six_vec4_cs", 14, 6), ("six_tpb512", 15, 6), ("six_tpb32", 16, 6),
                            ("six_vec4_tpb32", 17, 6)):
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.check(lib.dlc_microbench_f32(which, _abi.ptr(buf), T, blocks, _abi.current_stream_ptr()), name)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[f"{name}_T{T}"] = {"ms": ms, "GB_per_s": 4.0 * ns * T * N / (ms * 1e-3) / 1e9}
print(json.dumps(out, indent=1))
