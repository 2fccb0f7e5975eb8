import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from deluca_b200.fused import lds_shared_lqr_rollout, lds_shared_pack
d, m, N, T = 64, 16, 64, 1
rng = np.random.default_rng(0)
A = (0.1 * rng.standard_normal((d, d))).astype(np.float32)
B = np.zeros((d, m), np.float32); K = np.zeros((m, d), np.float32)
x0 = np.zeros((N, d), np.float32)
for n in range(N): x0[n, n % d] = 1.0 + n      # env n excites coordinate n: x1[n] = (1+n) * A[:, n]
t = lambda a: torch.as_tensor(a).cuda()
r = lds_shared_lqr_rollout(lds_shared_pack(t(A), t(B), t(K)), t(x0), T, m)
x1 = r.x.cpu().numpy()
ref = x0 @ A.T
print('max err', np.abs(x1 - ref).max())
# which (env, coord) of the reference does each output match?
for n in (0, 1, 2, 3, 4, 5, 8, 9, 17, 33):
    errs = [np.abs(x1[n] - ref[j]).max() for j in range(N)]
    j = int(np.argmin(errs)); print('env', n, 'matches ref env', j, 'err', errs[j])
