"""Parity of the LDS + GPC rollout AT THE BENCHMARKED HORIZONS and sizes ("a fixed random subset
(>= 256 envs, all T) plus all-env checksums").

* 256 envs x T = 10,000 (H = 3, HH = 10: BASELINE.json configs[1]) and 256 envs x T = 1,000 (H = HH = 10: the
  north-star target) against the float64 C oracle (oracle/lds_c.c), on the systems, learning rate and
  disturbance stream `bench.py` uses -- with the disturbances read from HBM (the `value` variant) and drawn in the
  kernel with `keep_costs=False, donate=True` (the `e2e` variant);
* at the full C2 / target sizes: the first 256 envs of the full job against the oracle, and the all-env cost-sum
  checksum between two independent kernels (different lane layouts and summation orders).

Tolerance: 1e-4 relative (BASELINE.json north_star) on per-step costs, final states and episode cost sums.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle.lds_c import gpc_rollout_c

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-4


def _bench():
    spec = importlib.util.spec_from_file_location("dlc_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def _subset_problem(bench, wl, n, lo=0):
    dev = torch.device("cuda")
    A, B, K, x0 = bench.make_problem(torch, wl, dev, lo, n)
    from deluca_b200.fused import gaussian_disturbance
    W = gaussian_disturbance(n, wl["T"], wl["d"], seed=bench.SEED, env_offset=lo, std=bench.W_STD, device=dev)
    return A, B, K, x0, W


def _oracle(bench, wl, A, B, K, x0, W):
    c = lambda t: t.detach().cpu().numpy()
    return gpc_rollout_c(c(A), c(B), c(K), c(x0), c(W), wl["H"], wl["HH"], bench.LR_SCALE, True, np.float64, 0)


def _compare(costs, x_final, ref_costs, ref_x):
    assert np.isfinite(costs).all()
    floor = 1e-2 * np.abs(ref_costs).mean()
    err = _rel(costs, ref_costs, floor)
    assert err.max() < RTOL, f"per-step costs: max rel err {err.max():.3e}"
    ex = _rel(x_final, ref_x, np.abs(ref_x).mean())
    assert ex.max() < RTOL, f"final state: max rel err {ex.max():.3e}"
    return err.max(), ex.max()


@pytest.mark.parametrize("name", ["c2", "tgt"])
def test_benchmarked_horizon_subset_matches_fp64_oracle(name):
    from deluca_b200.fused import lds_rollout
    bench = _bench()
    wl = dict(bench.WORKLOADS[name])
    n, T = 256, wl["T"]
    A, B, K, x0, W = _subset_problem(bench, wl, n)
    ref_costs, ref_x = _oracle(bench, wl, A, B, K, x0, W)
    kw = dict(policy="gpc", H=wl["H"], HH=wl["HH"], lr_scale=bench.LR_SCALE, decay=True, seed=bench.SEED, env_offset=0,
              w_std=bench.W_STD)
    # (1) the `value` variant: disturbances read from HBM, per-step costs stored
    r = lds_rollout(A, B, K, x0, T, W=W, **kw)
    e_c, e_x = _compare(r.costs.cpu().numpy(), r.x.cpu().numpy(), ref_costs, ref_x)
    assert (r.status.cpu().numpy() == 0).all()
    np.testing.assert_allclose(r.cost_sum.cpu().numpy(), ref_costs.sum(0), rtol=RTOL)
    # (2) the `e2e` variant: disturbances drawn in the kernel, no per-step costs, inputs donated
    r2 = lds_rollout(A.clone(), B.clone(), K.clone(), x0.clone(), T, W=None, keep_costs=False, donate=True, **kw)
    assert r2.costs is None
    np.testing.assert_allclose(r2.cost_sum.cpu().numpy(), ref_costs.sum(0), rtol=RTOL)
    ex = _rel(r2.x.cpu().numpy(), ref_x, np.abs(ref_x).mean())
    assert ex.max() < RTOL
    # (3) in-kernel disturbances with the costs kept (what the e2e leg that returns losses[T,N] runs)
    r3 = lds_rollout(A, B, K, x0, T, W=None, **kw)
    _compare(r3.costs.cpu().numpy(), r3.x.cpu().numpy(), ref_costs, ref_x)
    # the two disturbance sources are the same stream: bit-identical results
    assert torch.equal(r3.costs, r.costs) and torch.equal(r3.x, r.x)
    print(f"{name}: 256 envs x T={T}: max rel err costs {e_c:.2e}, final state {e_x:.2e}")


@pytest.mark.parametrize("name,other_variant", [("c2", 5), ("tgt", 5)])
def test_full_size_checksum_and_subset(name, other_variant):
    """The whole benchmarked job: env subset vs the oracle, and the all-env checksum between two kernels."""
    from deluca_b200.fused import lds_rollout
    bench = _bench()
    wl = dict(bench.WORKLOADS[name])
    N, T = wl["N"], wl["T"]
    A, B, K, x0, W = _subset_problem(bench, wl, N)
    kw = dict(policy="gpc", H=wl["H"], HH=wl["HH"], lr_scale=bench.LR_SCALE, decay=True, seed=bench.SEED, env_offset=0,
              w_std=bench.W_STD)
    r = lds_rollout(A, B, K, x0, T, W=W, **kw)                       # bench.py's `value` launch, exactly
    n = 256
    sel = torch.linspace(0, N - 1, n, device="cuda").long()          # a fixed subset spread over the whole job
    ref_costs, ref_x = _oracle(bench, wl, A[sel], B[sel], K[sel], x0[sel], W[:, sel].contiguous())
    _compare(r.costs[:, sel].cpu().numpy(), r.x[sel].cpu().numpy(), ref_costs, ref_x)
    cs = r.cost_sum.clone()
    status_ok = int((r.status == 0).sum())
    xf = r.x.clone()
    bad = torch.nonzero(r.status != 0).flatten()[:16]
    if bad.numel():
        # GPC-as-written diverges on a few systems even at lr_scale = 1e-4 (tests/test_oracle_lds.py); that must be
        # the algorithm's behaviour, not the kernel's: the float64 oracle blows up on the same envs
        bc, _ = _oracle(bench, wl, A[bad], B[bad], K[bad], x0[bad], W[:, bad].contiguous())
        assert (~np.isfinite(bc.sum(0)) | (np.abs(bc).max(0) > 1e30)).all(), "kernel-only divergence"
    del r
    torch.cuda.empty_cache()
    # an independent kernel (scalar lane-split layout, other summation order) on the same job, in-kernel stream
    r2 = lds_rollout(A, B, K, x0, T, W=None, keep_costs=False, kernel_variant=other_variant, **kw)
    ok = (r2.status == 0) & torch.isfinite(cs)
    assert int(ok.sum()) >= status_ok - 8 and status_ok >= N - 64     # (a handful of envs may diverge: lr quirk)
    rel = ((r2.cost_sum - cs).abs() / cs.abs().clamp_min(1e-30))[ok]
    assert float(rel.max()) < 2 * RTOL, f"per-env episode cost, kernel vs kernel: {float(rel.max()):.3e}"
    tot_a, tot_b = float(cs[ok].sum()), float(r2.cost_sum[ok].sum())
    assert abs(tot_a - tot_b) <= 1e-6 * abs(tot_a), "all-env checksum"
    # final states: each kernel is within RTOL of the float64 truth, so two kernels are within 2 RTOL of each other
    exf = ((r2.x - xf).abs().amax(1) / xf[ok].abs().mean())[ok]
    assert float(exf.max()) < 2 * RTOL
    print(f"{name}: full size {N} x {T}: checksum {tot_a:.9e} vs {tot_b:.9e}; finite envs {status_ok}")
