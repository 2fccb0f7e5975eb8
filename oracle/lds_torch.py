"""Multi-threaded CPU restatement of the LDS + GPC rollout (test infrastructure / CPU baseline).

Same arithmetic as ``oracle/lds.py`` (which follows ``deluca/agents/_gpc.py:109-183`` and
``deluca/envs/_lds.py:102-111`` line by line) written with batched torch CPU ops so that it uses
every host core (``torch.set_num_threads``); used only by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs and checked against ``oracle/lds.py`` in tests/test_oracle_lds.py.
Parity: pinned through ``oracle/lds.py`` (which reproduces the reference-generated ``tests/golden/ref_lds.npz``).
"""
import torch


def _mv(Mat, v):
    return torch.bmm(Mat, v.unsqueeze(-1)).squeeze(-1)


def _mtv(Mat, v):
    return torch.bmm(Mat.transpose(1, 2), v.unsqueeze(-1)).squeeze(-1)


def gpc_rollout_torch(A, B, K, x0, W, H=3, HH=10, lr_scale=0.005, decay=True, t0=0):
    """A[N,d,d] B[N,d,m] K[N,m,d] x0[N,d] W[T,N,d] (CPU tensors, any float dtype) -> costs[T,N], x.
    Requires HH >= H-1 (no dynamic_slice clamping)."""
    assert HH >= H - 1
    N, d = x0.shape
    m = B.shape[2]
    T = W.shape[0]
    dt = x0.dtype
    M = torch.zeros(N, H, m * 0 + m, d, dtype=dt)
    hist = torch.zeros(N, H + HH, d, dtype=dt)
    xprev = torch.zeros(N, d, dtype=dt)
    x = x0.clone()
    costs = torch.empty(T, N, dtype=dt)

    def mwin(start):  # sum_i M[i] w[start+i]
        return torch.einsum("nhij,nhj->ni", M, hist[:, start:start + H])

    for t in range(T):
        u = -_mv(K, x) + mwin(HH)                                   # _gpc.py:171-183
        noise = x - _mv(A, xprev) - _mv(B, u)                        # _gpc.py:155
        hist = torch.cat([hist[:, 1:], noise.unsqueeze(1)], dim=1)   # _gpc.py:156-157
        # policy_loss forward (_gpc.py:109-125) and hand adjoint (oracle/lds.py: gpc_grad)
        y = torch.zeros(N, d, dtype=dt)
        for h in range(H - 1):
            y = (_mv(A, y) + _mv(B, -_mv(K, y) + mwin(h))) + hist[:, h + H]
        wf = hist[:, HH - 1:HH - 1 + H]
        uf = -_mv(K, y) + torch.einsum("nhij,nhj->ni", M, wf)
        du = 2 * uf
        G = torch.einsum("ni,nhj->nhij", du, wf)
        lam = 2 * y - _mtv(K, du)
        for h in range(H - 2, -1, -1):
            lu = _mtv(B, lam)
            G = G + torch.einsum("ni,nhj->nhij", lu, hist[:, h:h + H])
            lam = _mtv(A, lam) - _mtv(K, lu)
        lr = lr_scale * (1.0 / (t0 + t + 1)) if decay else lr_scale  # _gpc.py:161-162
        M = M - lr * G
        xprev = x
        costs[t] = (x * x).sum(1) + (u * u).sum(1)                   # _gpc.py:29-40
        x = (_mv(A, x) + _mv(B, u)) + W[t]                            # _lds.py:102-111
    return costs, x
