"""Batched discrete algebraic Riccati solve for the LQR gain (setup-time plumbing).

The reference computes ``K = (B'XB + R)^-1 (B'XA)`` with ``X = scipy.linalg.solve_discrete_are``
on the host, once per agent (``deluca/agents/_lqr.py:48-60``).  For a batch of per-env systems the
same fixed point is found here with the structure-preserving doubling algorithm in float64 torch
ops on whatever device the inputs live on (quadratic convergence; ~10 iterations for the default
LDS).  Setup only -- never inside a timed rollout.
"""
import torch


def dare_gain(A, B, Q=None, R=None, iters=40, tol=1e-13):
    """A[...,d,d], B[...,d,m] -> K[...,m,d] (same dtype as A), X[...,d,d] (float64)."""
    dt = A.dtype
    A64, B64 = A.double(), B.double()
    d, m = B64.shape[-2], B64.shape[-1]
    eye_d = torch.eye(d, dtype=torch.float64, device=A.device).expand_as(A64)
    Q64 = eye_d if Q is None else Q.double().expand_as(A64)
    R64 = (torch.eye(m, dtype=torch.float64, device=A.device) if R is None else R.double())
    R64 = R64.expand(*B64.shape[:-2], m, m)
    G = B64 @ torch.linalg.solve(R64, B64.transpose(-1, -2))
    Ak, Gk, Hk = A64, G, Q64
    for _ in range(iters):
        Winv_A = torch.linalg.solve(eye_d + Gk @ Hk, Ak)            # (I + G H)^-1 A
        Winv_G = torch.linalg.solve(eye_d + Gk @ Hk, Gk)            # (I + G H)^-1 G
        A_next = Ak @ Winv_A
        G_next = Gk + Ak @ Winv_G @ Ak.transpose(-1, -2)
        H_next = Hk + Ak.transpose(-1, -2) @ Hk @ Winv_A
        delta = (H_next - Hk).abs().amax()
        Ak, Gk, Hk = A_next, G_next, H_next
        if float(delta) < tol * float(Hk.abs().amax()):
            break
    X = Hk
    BtX = B64.transpose(-1, -2) @ X
    K = torch.linalg.solve(BtX @ B64 + R64, BtX @ A64)
    return K.to(dt), X
