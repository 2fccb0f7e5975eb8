"""Write-only HBM bandwidth on this GPU (fill / memset of a buffer far larger than L2), next to the copy figure
MEASURED_PEAKS.json holds: the roofline for kernels whose traffic is all stores (the lung episode's six series)."""
import json

import torch

out = {}
for gb in (1, 4, 16):
    n = gb * (1 << 30) // 4
    x = torch.empty(n, device="cuda")
    for name, fn in (("fill", lambda: x.fill_(1.0)), ("memset", lambda: x.zero_())):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"{name}_{gb}GiB"] = {"ms": best, "GB_per_s": 4.0 * n / (best * 1e-3) / 1e9}
    y = torch.empty_like(x)
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"copy_{gb}GiB"] = {"ms": best, "GB_per_s": 8.0 * n / (best * 1e-3) / 1e9}
    del x, y
print(json.dumps(out, indent=1))
