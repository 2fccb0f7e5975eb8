"""Parity of the fused LDS + LQR/GPC/BPC CUDA rollout against the CPU oracle (B200 only).

Tolerance (BASELINE.json north_star): <= 1e-4 relative on costs and states for stable systems,
checked against the float64 oracle; the float32 oracle's own deviation is asserted to be of the
same order so the bound is meaningful.
"""
import numpy as np
import pytest
import torch

from oracle import lds as O
from oracle import philox as P

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def _problem(N, d, m, T, seed=0, dense=False):
    rng = np.random.default_rng(seed)
    A, B, x0 = O.lds_default_params(rng, N, d, m)
    if dense:  # general (non-diagonal) stable A
        Q = np.linalg.qr(rng.standard_normal((N, d, d)))[0]
        A = (Q @ A @ np.swapaxes(Q, 1, 2)).astype(np.float32)
    K = O.lqr_gain(A, B)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    x0 = (0.1 * rng.standard_normal((N, d))).astype(np.float32)
    return A, B, K, x0, W


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def _check(res, ref, T):
    costs = res.costs.cpu().numpy()
    assert np.isfinite(costs).all()
    assert _rel(costs, ref["costs"], 1e-2 * np.abs(ref["costs"]).mean()).max() < RTOL
    xs = np.abs(ref["x_final"]).mean()
    assert _rel(res.x.cpu().numpy(), ref["x_final"], xs).max() < RTOL
    np.testing.assert_allclose(res.cost_sum.cpu().numpy(), ref["costs"].astype(np.float64).sum(0), rtol=RTOL)
    assert (res.status.cpu().numpy() == 0).all()


@pytest.mark.parametrize("d,m,H,HH,variant", [
    (10, 3, 3, 10, 0), (10, 3, 3, 10, 1), (10, 3, 3, 2, 0), (10, 3, 10, 10, 0), (4, 1, 3, 2, 0),
    (10, 3, 3, 10, 5), (10, 3, 3, 10, 8), (10, 3, 3, 10, 9), (10, 3, 3, 2, 8), (10, 3, 10, 10, 8), (10, 3, 10, 10, 12), (10, 3, 10, 10, 11), (10, 3, 10, 10, 5), (10, 3, 10, 10, 17),
    (4, 2, 5, 7, 0), (4, 2, 5, 7, 1), (6, 2, 4, 5, 0), (3, 1, 1, 1, 0), (5, 2, 4, 2, 0)])
@pytest.mark.parametrize("dense", [False, True])
def test_gpc_matches_oracle(d, m, H, HH, variant, dense):
    from deluca_b200.fused import lds_rollout
    N, T = 70, 300   # N not a multiple of the block / group size
    A, B, K, x0, W = _problem(N, d, m, T, seed=d * 100 + H, dense=dense)
    kw = dict(H=H, HH=HH, lr_scale=1e-4)
    ref64 = O.rollout_lds(A, B, K, x0, W, "gpc", dtype=np.float64, **kw)
    ref32 = O.rollout_lds(A, B, K, x0, W, "gpc", dtype=np.float32, **kw)
    res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W), traj_stride=1,
                      kernel_variant=variant, **kw)
    _check(res, ref64, T)
    # the fp32 restatement is no closer to fp64 than the kernel is (same order of magnitude)
    e_k = np.abs(res.costs.cpu().numpy() - ref64["costs"]).max()
    e_o = np.abs(ref32["costs"] - ref64["costs"]).max()
    assert e_k < 20 * e_o + 1e-6
    # agent state and sampled trajectories
    st = ref64["state"]
    np.testing.assert_allclose(res.M.cpu().numpy(), st["M"], rtol=1e-3, atol=1e-5 * np.abs(st["M"]).max() + 1e-9)
    np.testing.assert_allclose(res.hist.cpu().numpy(), st["hist"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(res.xprev.cpu().numpy(), st["xprev"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(res.X.cpu().numpy(), ref64["X"][:T], rtol=1e-3, atol=1e-4)


def test_gpc_reference_default_lr_short_horizon():
    """Reference defaults (lr_scale=0.005, decay) on a horizon short enough to stay finite."""
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 33, 60, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=11)
    ref = O.rollout_lds(A, B, K, x0, W, "gpc", H=3, HH=10, dtype=np.float64)
    res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W), H=3, HH=10)
    c = res.costs.cpu().numpy()
    ok = np.isfinite(ref["costs"]).all(axis=0) & (np.abs(ref["costs"]).max(axis=0) < 1e4)
    assert ok.sum() > N // 2
    assert _rel(c[:, ok], ref["costs"][:, ok], 1e-2 * np.abs(ref["costs"][:, ok]).mean()).max() < 1e-3


def test_gpc_intended_noise_variant():
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 40, 400, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=5)
    kw = dict(H=3, HH=10, noise_uses_prev_action=True)
    ref = O.rollout_lds(A, B, K, x0, W, "gpc", dtype=np.float64, **kw)
    for variant in (0, 1):
        res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W),
                          kernel_variant=variant, **kw)
        _check(res, ref, T)


@pytest.mark.parametrize("variant", [0, 1])
def test_lqr_matches_oracle_and_zero_lr_gpc(variant):
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 129, 200, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=2, dense=True)
    ref = O.rollout_lds(A, B, K, x0, W, "lqr", dtype=np.float64)
    a = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="lqr", W=_dev(W), kernel_variant=variant)
    _check(a, ref, T)
    b = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W), H=3, HH=10,
                    lr_scale=0.0, kernel_variant=variant)
    # M == 0 and lr == 0 reduce GPC to LQR (SURVEY A-3.6).  The runtime-dims kernel runs the same
    # arithmetic for both (same bits); the templated GPC kernel packs its FMAs (other rounding order).
    if variant == 1:
        assert torch.equal(a.costs, b.costs) and torch.equal(a.x, b.x)
    else:
        assert torch.allclose(a.costs, b.costs, rtol=1e-5, atol=1e-5) and torch.allclose(a.x, b.x, rtol=1e-5, atol=1e-5)


def test_lqr_register_kernel_against_lane_split_and_resume():
    """The register-resident LQR kernel (d = 10, m = 3, kernel_variant 0) against the lane-split kernel (forced launch
    shape 5) with in-kernel disturbances, shared dynamics, odd T, sampled trajectories; resume bit for bit."""
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 77, 91, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=12, dense=True)
    for shared in (False, True):
        mats = tuple(_dev(v[0] if shared else v) for v in (A, B, K))
        kw = dict(policy="lqr", seed=8, env_offset=100, traj_stride=3)
        a = lds_rollout(*mats, _dev(x0), T, **kw)
        b = lds_rollout(*mats, _dev(x0), T, kernel_variant=5, **kw)
        for f in ("costs", "x", "X", "U"):
            np.testing.assert_allclose(getattr(a, f).cpu().numpy(), getattr(b, f).cpu().numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(a.cost_sum.cpu().numpy(), b.cost_sum.cpu().numpy(), rtol=1e-5)
        r1 = lds_rollout(*mats, _dev(x0), 40, **kw)
        r2 = lds_rollout(*mats, r1.x, T - 40, t0=40, **kw)
        assert torch.equal(torch.cat([r1.costs, r2.costs]), a.costs) and torch.equal(r2.x, a.x)


def test_lqr_row_ring_equals_register_buffers(monkeypatch):
    """The cp.async disturbance ring (row_ring.cuh: 4, 8 or 16 stages per warp) against the register double buffer
    (DLC_ROW_RING=0): same bits, for horizons shorter and longer than the ring, a ragged last warp (N = 130: a 2-env
    tile), an odd N (rows not 16-byte tileable: falls back), and against the float64 oracle."""
    from deluca_b200.fused import lds_rollout
    d, m = 10, 3
    for N, T in ((130, 1), (130, 3), (130, 16), (130, 17), (130, 75), (4096 + 62, 40), (129, 20)):
        A, B, K, x0, W = _problem(N, d, m, T, seed=21, dense=True)
        args = tuple(map(_dev, (A, B, K, x0)))
        out = {}
        for ring in ("0", "4", "8", "16"):
            monkeypatch.setenv("DLC_ROW_RING", ring)
            out[ring] = lds_rollout(*args, T, policy="lqr", W=_dev(W), traj_stride=7)
        monkeypatch.delenv("DLC_ROW_RING")
        for ring in ("4", "8", "16"):
            for f in ("costs", "x", "X", "U", "cost_sum", "status"):
                assert torch.equal(getattr(out[ring], f), getattr(out["0"], f)), (N, T, ring, f)
        if N == 130 and T == 75:
            _check(out["16"], O.rollout_lds(A, B, K, x0, W, "lqr", dtype=np.float64), T)


def test_resume_is_identical_and_inputs_untouched():
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 64, 90, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=3)
    dA, dB, dK, dx, dW = map(_dev, (A, B, K, x0, W))
    kw = dict(policy="gpc", H=3, HH=10, lr_scale=5e-4)
    full = lds_rollout(dA, dB, dK, dx, T, W=dW, **kw)
    assert torch.equal(dx, _dev(x0))
    h1 = lds_rollout(dA, dB, dK, dx, 40, W=dW[:40].contiguous(), **kw)
    h2 = lds_rollout(dA, dB, dK, h1.x, T - 40, W=dW[40:].contiguous(), M=h1.M, hist=h1.hist,
                     xprev=h1.xprev, uprev=h1.uprev, t0=h1.t, **kw)
    assert torch.equal(full.costs[40:], h2.costs)
    assert torch.equal(full.x, h2.x) and torch.equal(full.M, h2.M) and torch.equal(full.hist, h2.hist)


@pytest.mark.parametrize("variant", [0, 8, 11, 17])
def test_resume_long_history(variant):
    """H = HH = 10: split rollouts continue bit-identically in the packed kernel (slot-paired noise ring, variants
    8 / 11).  The quad-split kernel (variant 0) keeps the noise part of the counterfactual state as a sliding sum q
    (exact algebra, lds_gpc_quad.cu) that it re-forms from the ring on entry, so a resumed rollout agrees with the
    uninterrupted one to rounding (1e-6 relative), not bit for bit."""
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 45, 70, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=5)
    dA, dB, dK, dx, dW = map(_dev, (A, B, K, x0, W))
    kw = dict(policy="gpc", H=10, HH=10, lr_scale=1e-4, kernel_variant=variant)
    full = lds_rollout(dA, dB, dK, dx, T, W=dW, **kw)
    h1 = lds_rollout(dA, dB, dK, dx, 27, W=dW[:27].contiguous(), **kw)
    h2 = lds_rollout(dA, dB, dK, h1.x, T - 27, W=dW[27:].contiguous(), M=h1.M, hist=h1.hist,
                     xprev=h1.xprev, uprev=h1.uprev, t0=h1.t, **kw)
    if variant in (0, 17):      # (17: incremental window products, re-formed from (M, ring) on entry as well)
        close = lambda a, b: bool(((a - b).abs() <= 1e-6 * b.abs().max()).all())
        assert close(h2.costs, full.costs[27:]) and close(h2.x, full.x) and close(h2.M, full.M) and close(h2.hist, full.hist)
        return
    assert torch.equal(full.costs[27:], h2.costs)
    assert torch.equal(full.x, h2.x) and torch.equal(full.M, h2.M) and torch.equal(full.hist, h2.hist)


def test_shared_dynamics_equals_per_env_copy():
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 50, 100, 10, 3
    A, B, K, x0, W = _problem(1, d, m, T, seed=4)
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal((N, d)).astype(np.float32)
    W = (0.5 * rng.standard_normal((T, N, d))).astype(np.float32)
    kw = dict(policy="gpc", H=3, HH=10, lr_scale=5e-4)
    a = lds_rollout(_dev(A[0]), _dev(B[0]), _dev(K[0]), _dev(x0), T, W=_dev(W), **kw)
    b = lds_rollout(_dev(np.repeat(A, N, 0)), _dev(np.repeat(B, N, 0)), _dev(np.repeat(K, N, 0)), _dev(x0), T,
                    W=_dev(W), **kw)
    assert torch.equal(a.costs, b.costs)


def test_in_kernel_disturbance_equals_materialised_stream_and_oracle():
    from deluca_b200.fused import gaussian_disturbance, lds_rollout
    N, T, d, m = 100, 64, 10, 3
    A, B, K, x0, _ = _problem(N, d, m, T, seed=6)
    W = gaussian_disturbance(N, T, d, seed=1234, env_offset=7, t0=0, std=0.5)
    Wo = P.gaussian_disturbance(1234, np.arange(7, 7 + N), 0, T, d, 0.5)
    # integer stream is bit-exact; the float transform uses MUFU intrinsics on the device, libm here
    assert np.abs(W.cpu().numpy() - Wo).max() < 2e-5, np.abs(W.cpu().numpy() - Wo).max()
    kw = dict(policy="gpc", H=3, HH=10, lr_scale=5e-4)
    for variant in (1, 0):
        a = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, W=None, seed=1234, env_offset=7, w_std=0.5,
                        kernel_variant=variant, **kw)
        b = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, W=W, kernel_variant=variant, **kw)
        assert torch.equal(a.costs, b.costs)
    # sharding invariance: envs [32, 64) on their own draw the same stream
    s = slice(32, 64)
    c = lds_rollout(_dev(A[s]), _dev(B[s]), _dev(K[s]), _dev(x0[s]), T, W=None, seed=1234, env_offset=7 + 32, **kw)
    assert torch.equal(a.costs[:, s], c.costs)


def test_bpc_matches_oracle():
    from deluca_b200.fused import lds_rollout
    rng = np.random.default_rng(6)
    N, d, m, H, T = 20, 10, 3, 5, 120
    A, B, K, x0, W = _problem(N, d, m, T, seed=8)
    M0 = O.unit_norm(rng.standard_normal((N, H, m, d)).astype(np.float32))
    eps0 = O.unit_norm(rng.standard_normal((N, H, H, m, d)).astype(np.float32))
    eps = O.unit_norm(rng.standard_normal((T * N, H, m, d)).astype(np.float32)).reshape(T, N, H, m, d)
    for decay in (False, True):
        st = O.bpc_init(N, d, m, H, M0, eps0, dtype=np.float64)
        ref = O.rollout_lds(A, B, K, x0, W, "bpc", H=H, state=st, bpc_eps=eps, decay=decay, lr_scale=1e-5,
                            dtype=np.float64)
        res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="bpc", W=_dev(W), H=H,
                          M=_dev(0.01 * M0), eps=_dev(eps0), eps_new=_dev(eps), decay=decay, lr_scale=1e-5)
        _check(res, ref, T)
        np.testing.assert_allclose(res.M.cpu().numpy(), ref["state"]["M"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("d,m,H", [(10, 3, 5), (10, 3, 3), (4, 2, 2)])
@pytest.mark.parametrize("decay,prev", [(False, False), (True, True)])
def test_bpc_lane_split_kernel(d, m, H, decay, prev):
    """The lane-split BPC kernel (compile-time shapes) against the float64 oracle with an explicit perturbation
    stream, against the runtime-dims kernel with the in-kernel Philox streams, and resume == uninterrupted."""
    from deluca_b200.fused import lds_rollout
    rng = np.random.default_rng(d + H)
    N, T = 41, 90          # 41: the last CTA has idle lane groups
    A, B, K, x0, W = _problem(N, d, m, T, seed=3, dense=True)
    M0 = (0.01 * O.unit_norm(rng.standard_normal((N, H, m, d)).astype(np.float32))).astype(np.float32)
    eps0 = O.unit_norm(rng.standard_normal((N, H, H, m, d)).astype(np.float32))
    eps = O.unit_norm(rng.standard_normal((T * N, H, m, d)).astype(np.float32)).reshape(T, N, H, m, d)
    kw = dict(policy="bpc", H=H, M=_dev(M0), eps=_dev(eps0), decay=decay, lr_scale=1e-5)
    st = O.bpc_init(N, d, m, H, M0 / 0.01, eps0, dtype=np.float64)
    ref = O.rollout_lds(A, B, K, x0, W, "bpc", H=H, state=st, bpc_eps=eps, decay=decay, lr_scale=1e-5,
                        dtype=np.float64)
    res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, W=_dev(W), eps_new=_dev(eps), **kw)
    _check(res, ref, T)
    kw["noise_uses_prev_action"] = prev
    np.testing.assert_allclose(res.M.cpu().numpy(), ref["state"]["M"], rtol=1e-3, atol=1e-6)
    # in-kernel streams (disturbances and perturbations): same draws as the runtime-dims kernel
    a = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, seed=77, env_offset=5, traj_stride=4, **kw)
    b = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, seed=77, env_offset=5, traj_stride=4, kernel_variant=1, **kw)
    for f in ("costs", "x", "M", "hist", "xprev", "uprev", "eps", "bpc_cost", "X", "U"):
        ta, tb = getattr(a, f).cpu().numpy(), getattr(b, f).cpu().numpy()
        np.testing.assert_allclose(ta, tb, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(tb).max()), err_msg=f)
    # resume: 2 x T/2 with the returned state is the same rollout, bit for bit
    h = T // 2
    r1 = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), h, seed=77, env_offset=5, **kw)
    kw2 = dict(kw, M=r1.M, eps=r1.eps)
    r2 = lds_rollout(_dev(A), _dev(B), _dev(K), r1.x, T - h, seed=77, env_offset=5, hist=r1.hist, xprev=r1.xprev,
                     uprev=r1.uprev, bpc_cost=r1.bpc_cost, t0=r1.t, **kw2)
    assert torch.equal(torch.cat([r1.costs, r2.costs]), a.costs)
    assert torch.equal(r2.M, a.M) and torch.equal(r2.eps, a.eps) and torch.equal(r2.x, a.x)


@pytest.mark.parametrize("d,m,H,HH", [(25, 1, 3, 2), (6, 2, 4, 5), (5, 2, 4, 2), (40, 5, 3, 6), (64, 8, 2, 3), (17, 3, 1, 1)])
def test_gpc_lane_group_kernel(d, m, H, HH):
    """The lane-group GPC kernel (any shape with d, m <= 64; the LDS defaults d_hidden = 25, d_in = 1 among them)
    against the float64 oracle, against the thread-per-env runtime-dims kernel on every state tensor with in-kernel
    disturbances and sampled trajectories, and resume == uninterrupted bit for bit."""
    from deluca_b200.fused import lds_rollout
    N, T = 45, 80
    A, B, K, x0, W = _problem(N, d, m, T, seed=d + H, dense=True)
    ref = O.rollout_lds(A, B, K, x0, W, "gpc", H=H, HH=HH, lr_scale=1e-4, dtype=np.float64)
    res = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W), H=H, HH=HH, lr_scale=1e-4)
    _check(res, ref, T)
    np.testing.assert_allclose(res.M.cpu().numpy(), ref["state"]["M"], rtol=1e-3, atol=1e-5 * max(np.abs(ref["state"]["M"]).max(), 1e-6))
    kw = dict(policy="gpc", H=H, HH=HH, lr_scale=1e-4, seed=31, env_offset=9, traj_stride=5, noise_uses_prev_action=True)
    a = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, **kw)
    b = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, kernel_variant=1, **kw)
    for f in ("costs", "x", "M", "hist", "xprev", "uprev", "X", "U"):
        ta, tb = getattr(a, f).cpu().numpy(), getattr(b, f).cpu().numpy()
        np.testing.assert_allclose(ta, tb, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(tb).max()), err_msg=f)
    h = T // 2
    kw.pop("traj_stride")
    full = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, **kw)
    r1 = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), h, **kw)
    r2 = lds_rollout(_dev(A), _dev(B), _dev(K), r1.x, T - h, M=r1.M, hist=r1.hist, xprev=r1.xprev, uprev=r1.uprev,
                     t0=r1.t, **kw)
    assert torch.equal(torch.cat([r1.costs, r2.costs]), full.costs)
    assert torch.equal(r2.M, full.M) and torch.equal(r2.hist, full.hist) and torch.equal(r2.x, full.x)


def test_empty_batch_and_argument_errors():
    from deluca_b200 import _abi
    from deluca_b200.fused import lds_rollout
    z = lambda *s: torch.zeros(s, device="cuda")
    r = lds_rollout(z(0, 10, 10), z(0, 10, 3), z(0, 3, 10), z(0, 10), 5, policy="lqr", W=z(5, 0, 10))
    assert r.costs.shape == (5, 0)
    with pytest.raises(TypeError):
        lds_rollout(torch.zeros(1, 2, 2), z(1, 2, 1), z(1, 1, 2), z(1, 2), 1, policy="lqr")
    with pytest.raises(_abi.DlcError):  # HH = 0 is rejected by the ABI
        lds_rollout(z(1, 2, 2), z(1, 2, 1), z(1, 1, 2), z(1, 2), 1, policy="gpc", H=2, HH=0)


def test_nonfinite_status_flag():
    from deluca_b200.fused import lds_rollout
    N, T, d, m = 8, 50, 10, 3
    A, B, K, x0, W = _problem(N, d, m, T, seed=9)
    A[3] *= 40.0  # wildly unstable env
    r = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, policy="gpc", W=_dev(W), H=3, HH=10)
    s = r.status.cpu().numpy()
    assert s[3] == 1 and (np.delete(s, 3) == 0).all()


def test_large_batch_properties():
    """BASELINE-sized batch, short horizon: linearity of the LQR rollout in (x0, W) and the
    cost_sum == sum(costs) checksum hold for every env."""
    from deluca_b200.fused import gaussian_disturbance, lds_rollout
    N, T, d, m = 65536, 64, 10, 3
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g))
    B = torch.randn(N, d, m, device="cuda", generator=g)
    from deluca_b200.riccati import dare_gain
    K = dare_gain(A, B)[0].contiguous()
    x0 = torch.randn(N, d, device="cuda", generator=g)
    W = gaussian_disturbance(N, T, d, seed=3)
    a = lds_rollout(A, B, K, x0, T, policy="lqr", W=W, traj_stride=64)
    b = lds_rollout(A, B, K, 2 * x0, T, policy="lqr", W=2 * W, traj_stride=64)
    assert torch.allclose(b.X, 2 * a.X, rtol=1e-5, atol=1e-5)
    assert torch.allclose(b.costs, 4 * a.costs, rtol=1e-5, atol=1e-5)
    assert torch.allclose(a.costs.double().sum(0), a.cost_sum, rtol=1e-6)
    gp = lds_rollout(A, B, K, x0, T, policy="gpc", W=W, H=3, HH=10, lr_scale=1e-4)
    assert torch.allclose(gp.costs.double().sum(0), gp.cost_sum, rtol=1e-6)
    assert int(gp.status.sum()) == 0


@pytest.mark.parametrize("H,HH,N,T", [(3, 2, 300, 41), (10, 10, 256, 60)])
def test_rollout_streamed_to_host_equals_one_launch(H, HH, N, T):
    """`lds_rollout_to_host`: the horizon in resumed chunks, losses copied to pinned host memory under the next chunk's
    kernel.  Same numbers as one launch (to rounding where a kernel re-forms a carried sum on entry), same final
    agent / env state, and the in-kernel disturbance stream continues across the cuts."""
    from deluca_b200.fused import lds_rollout, lds_rollout_to_host
    g = torch.Generator(device="cuda").manual_seed(3)
    d, m = 10, 3
    A = torch.diag_embed(0.9 + 0.05 * torch.rand(N, d, device="cuda", generator=g))
    B = torch.randn(N, d, m, device="cuda", generator=g) / d ** 0.5
    K = 0.05 * torch.randn(N, m, d, device="cuda", generator=g)
    x0 = torch.randn(N, d, device="cuda", generator=g)
    kw = dict(policy="gpc", H=H, HH=HH, lr_scale=1e-5, seed=77, env_offset=5, w_std=0.5)
    one = lds_rollout(A, B, K, x0, T, W=None, **kw)
    assert bool(torch.isfinite(one.costs).all())
    host = torch.empty(T, N).pin_memory()
    r = lds_rollout_to_host(A, B, K, x0, T, host, chunks=5, **kw)
    r.done.synchronize()
    # (the quad-split kernel of the long-history case re-forms its sliding sum from the ring on entry: rounding only)
    tol = dict(rtol=2e-6, atol=1e-6) if H < 10 else dict(rtol=5e-5, atol=5e-5)
    err = (host - one.costs.cpu()).abs().max()
    assert torch.allclose(host, one.costs.cpu(), **tol), float(err)
    assert torch.allclose(r.x, one.x, **tol) and torch.allclose(r.M, one.M, rtol=1e-4, atol=1e-7)
    assert torch.allclose(r.cost_sum, one.cost_sum, rtol=1e-6) and r.t == T and int(r.status.sum()) == 0
    with pytest.raises(TypeError):
        lds_rollout_to_host(A, B, K, x0, T, torch.empty(T, N), **kw)        # not pinned


def test_env_sliced_host_to_host_rollout_is_the_same_job():
    """fused.lds_rollout_host_to_host (pinned host systems in, losses[T,N] in pinned host memory out, env slices whose
    copies and kernels overlap) against ONE lds_rollout call on the whole batch with the in-kernel disturbance stream:
    same bits (the stream is keyed by the global env index), for the target's shape (H = HH = 10) and configs[1]'s;
    and dlc_copy_rows_async in both directions."""
    from deluca_b200.fused import copy_rows_async, lds_rollout, lds_rollout_host_to_host
    src = torch.arange(7 * 40, dtype=torch.float32).reshape(7, 40)
    host = torch.zeros(7, 40).pin_memory()
    dev = src.cuda()
    copy_rows_async(host[:, 8:24], dev[:, 3:19])
    torch.cuda.synchronize()
    assert torch.equal(host[:, 8:24], src[:, 3:19]) and host[:, :8].eq(0).all() and host[:, 24:].eq(0).all()
    back = torch.zeros(7, 40, device="cuda")
    copy_rows_async(back[:, 1:17], host[:, 8:24])
    torch.cuda.synchronize()
    assert torch.equal(back[:, 1:17].cpu(), src[:, 3:19])
    with pytest.raises(ValueError):
        copy_rows_async(host[:, ::2], dev[:, :20])
    d, m = 10, 3
    # (slices=None: the default cut into whole grid waves -- 64 envs per SM -- so 2.x waves make three slices)
    wave = 64 * torch.cuda.get_device_properties(0).multi_processor_count
    for (H, HH), N, T, slices in (((10, 10), 1000, 40, 4), ((3, 10), 700, 60, 16), ((10, 10), 64, 12, 1),
                                  ((10, 10), 2 * wave + 1000, 6, None), ((3, 10), wave + 50, 5, None)):
        A, B, K, x0, _ = _problem(min(N, 500), d, m, T, seed=31 + H, dense=True)
        if N > 500:                                                  # (tiled: the DARE solves dominate otherwise)
            A, B, K, x0 = (np.resize(a, (N,) + a.shape[1:]) for a in (A, B, K, x0))
        kw = dict(policy="gpc", H=H, HH=HH, lr_scale=1e-4, seed=5, w_std=0.5)
        whole = lds_rollout(_dev(A), _dev(B), _dev(K), _dev(x0), T, env_offset=77, **kw)
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).pin_memory()
        out = torch.full((T, N), float("nan")).pin_memory()
        r = lds_rollout_host_to_host(pin(A), pin(B), pin(K), pin(x0), T, out, slices=slices, env_offset=77, **kw)
        r.done.synchronize()
        assert torch.equal(out, whole.costs.cpu()), (H, N)
        assert torch.equal(r.cost_sum, whole.cost_sum) and torch.equal(r.x, whole.x) and torch.equal(r.status, whole.status)
        r2 = lds_rollout_host_to_host(pin(A), pin(B), pin(K), pin(x0), T, None, slices=slices, env_offset=77, **kw)
        assert r2.costs is None and torch.equal(r2.cost_sum, whole.cost_sum) and torch.equal(r2.x, whole.x)
    with pytest.raises(TypeError):
        lds_rollout_host_to_host(torch.zeros(4, 10, 10), pin(B[:4]), pin(K[:4]), pin(x0[:4]), 3, out[:3, :4].contiguous().pin_memory())
