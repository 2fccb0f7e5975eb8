"""Parity of the shared-dynamics tcgen05 (3xTF32) LDS + LQR rollout against the float64 oracle (B200)."""
import numpy as np
import pytest
import torch

from oracle import lds as O

pytestmark = pytest.mark.gpu


def _system(d, m, seed=0):
    rng = np.random.default_rng(seed)
    Q = np.linalg.qr(rng.standard_normal((d, d)))[0]
    A = (Q * rng.uniform(0.7, 0.95, d)) @ Q.T                      # spectral radius < 1
    B = rng.standard_normal((d, m)) / np.sqrt(d)
    K = O.lqr_gain(A[None], B[None], dtype=np.float64)[0]
    return A.astype(np.float32), B.astype(np.float32), K.astype(np.float32)


@pytest.mark.parametrize("d,m,N,T", [(256, 64, 130, 12), (64, 16, 64, 30), (128, 32, 200, 20), (48, 8, 70, 25),
                                     (256, 64, 1, 3)])
def test_shared_lqr_rollout_matches_oracle(d, m, N, T):
    from deluca_b200.fused import lds_shared_lqr_rollout, lds_shared_pack
    A, B, K = _system(d, m, seed=d + m)
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((N, d)).astype(np.float32)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    rep = lambda a: np.broadcast_to(a, (N,) + a.shape)
    ref = O.rollout_lds(rep(A), rep(B), rep(K), x0, W, "lqr", dtype=np.float64, keep_x=False)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    G = lds_shared_pack(t(A), t(B), t(K))
    res = lds_shared_lqr_rollout(G, t(x0), T, m, W=t(W))
    costs = res.costs.cpu().numpy()
    rel = np.abs(costs - ref["costs"]) / np.maximum(np.abs(ref["costs"]), 1e-2 * np.abs(ref["costs"]).mean())
    assert np.isfinite(costs).all() and rel.max() < 1e-4, rel.max()
    xs = np.abs(ref["x_final"]).mean()
    relx = np.abs(res.x.cpu().numpy() - ref["x_final"]) / np.maximum(np.abs(ref["x_final"]), xs)
    assert relx.max() < 1e-4, relx.max()
    np.testing.assert_allclose(res.cost_sum.cpu().numpy(), ref["costs"].sum(0), rtol=1e-4)
    assert int(res.status.sum()) == 0


def test_shared_path_agrees_with_the_per_env_cuda_core_kernel():
    """Same rollout through dlc_lds_rollout_f32 (fp32 CUDA cores, runtime-dims kernel) and through
    the tensor-core path: two independent implementations of the same reference lines."""
    from deluca_b200.fused import lds_rollout, lds_shared_lqr_rollout, lds_shared_pack
    d, m, N, T = 64, 16, 96, 40
    A, B, K = _system(d, m, seed=3)
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn(N, d, device="cuda", generator=g)
    W = 0.5 * torch.randn(T, N, d, device="cuda", generator=g)
    t = lambda a: torch.as_tensor(a).cuda()
    a = lds_rollout(t(A), t(B), t(K), x0, T, policy="lqr", W=W)
    b = lds_shared_lqr_rollout(lds_shared_pack(t(A), t(B), t(K)), x0, T, m, W=W)
    assert torch.allclose(a.costs, b.costs, rtol=1e-4, atol=1e-4)
    assert torch.allclose(a.x, b.x, rtol=1e-4, atol=1e-4)


def test_shared_rejects_unsupported_shapes():
    from deluca_b200 import _abi
    from deluca_b200.fused import lds_shared_pack
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(ValueError):
        lds_shared_pack(z(10, 10), z(10, 3), z(3, 10))      # d % 16 != 0
    assert _abi.lib().dlc_lds_shared_packed_floats(512, 64) > 0   # size helper does not validate the range
    with pytest.raises(_abi.DlcError):
        lds_shared_pack(z(512, 512), z(512, 64), z(64, 512))


# ------------------------------------------------------------------ shared dynamics + per-env GPC (config 5)
def _gpc_case(d, m, H, N, T, lr_scale, seed):
    A, B, K = _system(d, m, seed=d + m + seed)
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal((N, d)).astype(np.float32)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    rep = lambda a: np.broadcast_to(a, (N,) + a.shape)
    ref = O.rollout_lds(rep(A), rep(B), rep(K), x0, W, "gpc", H=H, HH=H, lr_scale=lr_scale, decay=True,
                        dtype=np.float64, keep_x=False)
    return A, B, K, x0, W, ref


@pytest.mark.parametrize("d,m,H,N,T,lr", [(64, 32, 4, 70, 14, 0.002), (256, 64, 16, 12, 20, 0.002),
                                          (64, 32, 4, 1, 3, 0.002)])
def test_shared_gpc_rollout_matches_oracle(d, m, H, N, T, lr):
    """dlc_lds_shared_gpc_rollout_f32 (M-stream + shared-operator chain kernels) against the float64 restatement
    of GPC.__call__ + LDS.__call__ (oracle/lds.py); tolerance 1e-4 relative (north_star), stated per quantity."""
    from deluca_b200.fused import lds_shared_gpc_rollout
    A, B, K, x0, W, ref = _gpc_case(d, m, H, N, T, lr, seed=2)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    res = lds_shared_gpc_rollout(t(A), t(B), t(K), t(x0), T, H=H, W=t(W), lr_scale=lr, decay=True)
    costs = res.costs.cpu().numpy()
    rel = np.abs(costs - ref["costs"]) / np.maximum(np.abs(ref["costs"]), 1e-2 * np.abs(ref["costs"]).mean())
    assert np.isfinite(costs).all() and rel.max() < 1e-4, rel.max()
    xs = np.abs(ref["x_final"]).mean()
    relx = np.abs(res.x.cpu().numpy() - ref["x_final"]) / np.maximum(np.abs(ref["x_final"]), xs)
    assert relx.max() < 1e-4, relx.max()
    Mref = ref["state"]["M"]
    assert np.abs(Mref).max() > 1e-4                      # the policy did learn something in T steps
    assert np.abs(res.M.cpu().numpy() - Mref).max() < 1e-4 * np.abs(Mref).max()
    href = ref["state"]["hist"]
    assert np.abs(res.hist.cpu().numpy() - href).max() < 1e-4 * np.abs(href).max()
    np.testing.assert_allclose(res.xprev.cpu().numpy(), ref["state"]["xprev"], rtol=1e-4, atol=1e-4 * xs)
    np.testing.assert_allclose(res.cost_sum.cpu().numpy(), ref["costs"].sum(0), rtol=1e-4)
    assert int(res.status.sum()) == 0 and res.t == T


def test_shared_gpc_resume_equals_uninterrupted_run():
    from deluca_b200.fused import lds_shared_gpc_rollout
    d, m, H, N, T = 64, 32, 4, 130, 17
    A, B, K = _system(d, m, seed=11)
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randn(N, d, device="cuda", generator=g)
    W = 0.5 * torch.randn(T, N, d, device="cuda", generator=g)
    t = lambda a: torch.as_tensor(a).cuda()
    full = lds_shared_gpc_rollout(t(A), t(B), t(K), x0, T, H=H, W=W, lr_scale=0.002)
    a = lds_shared_gpc_rollout(t(A), t(B), t(K), x0, 7, H=H, W=W[:7].contiguous(), lr_scale=0.002)
    b = lds_shared_gpc_rollout(t(A), t(B), t(K), a.x, T - 7, H=H, W=W[7:].contiguous(), lr_scale=0.002,
                               M=a.M, hist=a.hist, xprev=a.xprev, t0=a.t)
    # the carried A x_{t-1} is recomputed from xprev on entry with a different summation order
    assert torch.allclose(torch.cat([a.costs, b.costs]), full.costs, rtol=1e-5, atol=1e-5)
    assert torch.allclose(b.x, full.x, rtol=1e-5, atol=1e-5)
    assert torch.allclose(b.M, full.M, rtol=1e-4, atol=1e-6)
    assert torch.allclose(b.hist, full.hist, rtol=1e-4, atol=1e-5)


def test_shared_gpc_agrees_with_the_per_env_kernel():
    """Same rollout through dlc_lds_rollout_f32 (runtime-dims kernel, one thread per env, shared_dynamics = 1):
    an independent implementation of the same reference lines, on more envs than the oracle is asked for."""
    from deluca_b200.fused import lds_rollout, lds_shared_gpc_rollout
    d, m, H, N, T = 64, 32, 4, 700, 25
    A, B, K = _system(d, m, seed=4)
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(N, d, device="cuda", generator=g)
    W = 0.5 * torch.randn(T, N, d, device="cuda", generator=g)
    t = lambda a: torch.as_tensor(a).cuda()
    a = lds_rollout(t(A), t(B), t(K), x0, T, policy="gpc", H=H, HH=H, W=W, lr_scale=0.002)
    b = lds_shared_gpc_rollout(t(A), t(B), t(K), x0, T, H=H, W=W, lr_scale=0.002)
    assert torch.allclose(a.costs, b.costs, rtol=1e-4, atol=1e-3)
    assert torch.allclose(a.x, b.x, rtol=1e-4, atol=1e-4)
    assert torch.allclose(a.M, b.M, rtol=1e-3, atol=1e-4 * float(a.M.abs().max()))


@pytest.mark.parametrize("d,m,H,N,T,ctas", [(64, 32, 4, 300, 9, 1), (64, 32, 4, 130, 1, 18), (64, 32, 4, 65, 5, 2),
                                            (256, 64, 16, 200, 4, 1)])
def test_shared_gpc_two_half_pipeline_is_bit_identical(d, m, H, N, T, ctas, monkeypatch):
    """The two-half software pipeline (chain of one half on a few persistent CTAs of a second stream while the other
    half's M stream runs) launches, per env, exactly what the serial loop launches: every output must be bit-equal.
    `ctas` < tiles per half exercises the persistent tile loop of the chain kernel; resuming exercises q_init."""
    from deluca_b200.fused import lds_shared_gpc_rollout
    A, B, K = _system(d, m, seed=21)
    g = torch.Generator(device="cuda").manual_seed(9)
    x0 = torch.randn(N, d, device="cuda", generator=g)
    W = 0.5 * torch.randn(T + 3, N, d, device="cuda", generator=g)
    t = lambda a: torch.as_tensor(a).cuda()

    def run():
        a = lds_shared_gpc_rollout(t(A), t(B), t(K), x0, T, H=H, W=W[:T].contiguous(), lr_scale=0.002)
        b = lds_shared_gpc_rollout(t(A), t(B), t(K), a.x, 3, H=H, W=W[T:].contiguous(), lr_scale=0.002,
                                   M=a.M, hist=a.hist, xprev=a.xprev, t0=a.t)
        torch.cuda.synchronize()
        return a, b

    monkeypatch.setenv("DLC_SGPC_PIPE", "0")
    sa, sb = run()
    monkeypatch.setenv("DLC_SGPC_PIPE", "1")
    monkeypatch.setenv("DLC_SGPC_CHAIN_CTAS", str(ctas))
    pa, pb = run()
    bad = []
    for tag, ser, pip in (("first call", sa, pa), ("resumed call", sb, pb)):
        for name in ("costs", "x", "M", "hist", "xprev", "cost_sum", "status"):
            u, v = getattr(ser, name), getattr(pip, name)
            if not torch.equal(u, v):
                diff = (u.double() - v.double()).abs()
                envs = torch.nonzero(diff.reshape(-1, N).sum(0) if name == "costs" else diff.reshape(N, -1).sum(1)).flatten()
                bad.append((tag, name, float(diff.max()), envs[:8].tolist(), int(envs.numel())))
    assert not bad, bad
    assert float(sb.M.abs().max()) > 0      # (the resumed call: a T = 1 call from a zero history leaves M = 0)


def test_shared_gpc_rejects_unsupported_shapes():
    from deluca_b200.fused import lds_shared_gpc_rollout
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(ValueError):
        lds_shared_gpc_rollout(z(48, 48), z(48, 8), z(8, 48), z(4, 48), 2, H=4)
    with pytest.raises(ValueError):
        lds_shared_gpc_rollout(z(64, 64), z(64, 32), z(32, 64), z(4, 64), 2, H=4, HH=6)
