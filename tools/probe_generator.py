import sys, torch
sys.path.insert(0, '/root/repo')
from deluca_b200.fused import gaussian_disturbance
N, T, d = 1 << 20, 250, 10
W = torch.empty(T, N, d, device="cuda")
for d_ in (10, 7):
    Wd = W.view(-1)[: T * N * d_].view(T, N, d_)
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gaussian_disturbance(N, T, d_, seed=3, std=0.5, out=Wd); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"d={d_}: {ms:.2f} ms for {T*N*d_*4/1e9:.1f} GB = {T*N*d_*4/ms/1e6:.0f} GB/s, {T*N*d_/ms/1e6:.1f} G normals/s")
