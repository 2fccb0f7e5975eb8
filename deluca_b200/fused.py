"""Fused-rollout entry points: thin, allocation-only wrappers over the C ABI.

These are what the scan x vmap drivers (``deluca_b200.training.apg`` etc.) dispatch to.  All
tensors are torch CUDA tensors used purely as device-memory handles; every number is produced
by ``libdeluca_b200.so``.
"""
import ctypes
import os
from typing import NamedTuple, Optional

import torch

from . import _abi


def _f32(t, name, shape=None):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def gaussian_disturbance(N, T, d, seed, env_offset=0, t0=0, std=0.5, device="cuda", out=None):
    """W[T,N,d] of the counter-based stream the kernels draw in-register (include/deluca_b200.h)."""
    W = out if out is not None else torch.empty((T, N, d), dtype=torch.float32, device=device)
    _f32(W, "out", (T, N, d))
    with torch.cuda.device(W.device):
        rc = _abi.lib().dlc_gaussian_disturbance_f32(_abi.ptr(W), N, T, d, seed, env_offset, t0, std,
                                                    _abi.current_stream_ptr())
    _abi.check(rc, "dlc_gaussian_disturbance_f32")
    return W


def sinus_disturbance(N, T, phases, amplitude=0.5, t0=0):
    """W[T,N,d] = sin((t0 + t) * phases) * amplitude (``SinusDisturbance``, ``deluca/envs/_lds.py:44-52``)."""
    _f32(phases, "phases")
    d = phases.numel()
    W = torch.empty((T, N, d), dtype=torch.float32, device=phases.device)
    rc = _launch_on(phases, _abi.lib().dlc_sinus_disturbance_f32, _abi.ptr(W), N, int(T), d, _abi.ptr(phases),
                    float(amplitude), int(t0))
    _abi.check(rc, "dlc_sinus_disturbance_f32")
    return W


class LdsRolloutResult(dict):
    __getattr__ = dict.__getitem__


def lds_rollout(A, B, K, x, T, *, policy="gpc", W=None, H=3, HH=2, M=None, hist=None, xprev=None,
                uprev=None, eps=None, eps_new=None, bpc_cost=None, t0=0, lr_scale=0.005, decay=True,
                noise_uses_prev_action=False, w_std=0.5, seed=0, env_offset=0, bpc_delta=0.01,
                keep_costs=True, traj_stride=0, kernel_variant=0, donate=False, u_in=None, agent_only=False,
                env_t0=None, cost_q=None, cost_r=None):
    """T fused steps of LDS + LQR/GPC/BPC for N envs (``dlc_lds_rollout_f32``).

    A[N,d,d] B[N,d,m] K[N,m,d] (or unbatched = shared by all envs), x[N,d].  State tensors
    (``M, hist, xprev, uprev, eps, bpc_cost``) default to the agents' initial values
    (_gpc.py:82-101, _bpc.py:79-90 with explicit M/eps).  Unless ``donate`` the inputs are left
    untouched and updated copies are returned (the reference's functional convention).
    Returns costs[T,N], cost_sum[N] (fp64), x (final), the new agent state, status[N] and, when
    ``traj_stride`` > 0, sampled trajectories X[T,Ns,d], U[T,Ns,m] of every ``traj_stride``-th env.
    ``t0`` is the AGENT's step counter (lr decay, BPC perturbation counter), ``env_t0`` the ENV's (``LDS.t``: the
    counter of the in-kernel disturbance stream; defaults to ``t0``).
    ``cost_q[d]`` / ``cost_r[m]``: GPC's ``cost_fn`` as a diagonal quadratic ``sum(q x^2) + sum(r u^2)``
    (``deluca/agents/_gpc.py:52,76,125``; ``dlc_lds_rollout_cost_f32``); the reported costs stay ``quad_loss``.
    """
    lib = _abi.lib()
    _f32(x, "x")
    N, d = x.shape
    shared = A.dim() == 2
    m = B.shape[-1]
    _f32(A, "A", (d, d) if shared else (N, d, d))
    _f32(B, "B", (d, m) if shared else (N, d, m))
    _f32(K, "K", (m, d) if shared else (N, m, d))
    dev = x.device
    pol = _abi.POLICY[policy]
    learn = pol != 0
    P = (H + HH) if policy == "gpc" else H
    cl = (lambda t: t) if donate else (lambda t: None if t is None else t.clone())
    x_io = cl(x)
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    M_io = hist_io = xprev_io = uprev_io = eps_io = cost_io = None
    if learn:
        M_io = cl(_f32(M, "M", (N, H, m, d))) if M is not None else z(N, H, m, d)
        hist_io = cl(_f32(hist, "hist", (N, P, d))) if hist is not None else z(N, P, d)
        xprev_io = cl(_f32(xprev, "xprev", (N, d))) if xprev is not None else z(N, d)
        uprev_io = cl(_f32(uprev, "uprev", (N, m))) if uprev is not None else z(N, m)
    if policy == "bpc":
        if M is None or eps is None:
            raise ValueError("BPC needs explicit M and eps (the reference draws them from an unseeded RNG)")
        eps_io = cl(_f32(eps, "eps", (N, H, H, m, d)))
        cost_io = cl(_f32(bpc_cost, "bpc_cost", (N,))) if bpc_cost is not None else z(N)
        _f32(eps_new, "eps_new", (T, N, H, m, d))
    _f32(W, "W", (T, N, d))
    if policy == "open":
        _f32(u_in, "u_in", (T, N, m))
    costs = torch.empty((T, N), dtype=torch.float32, device=dev) if keep_costs else None
    cost_sum = torch.empty((N,), dtype=torch.float64, device=dev)
    status = torch.empty((N,), dtype=torch.int32, device=dev)
    X = U = None
    if traj_stride and traj_stride > 0:
        Ns = (N + traj_stride - 1) // traj_stride
        X = torch.empty((T, Ns, d), dtype=torch.float32, device=dev)
        U = torch.empty((T, Ns, m), dtype=torch.float32, device=dev)
    p = _abi.DlcLdsParams(N=N, env_offset=env_offset, T=T, d=d, m=m, H=H, HH=HH, policy=pol, t0=t0,
                          decay=int(bool(decay)), noise_uses_prev_action=int(bool(noise_uses_prev_action)),
                          shared_dynamics=int(shared), traj_stride=max(1, int(traj_stride or 1)),
                          lr_scale=lr_scale, w_std=w_std, bpc_delta=bpc_delta,
                          kernel_variant=kernel_variant, mode=int(bool(agent_only)), seed=seed,
                          env_t0=int(t0 if env_t0 is None else env_t0))
    ws_bytes = lib.dlc_lds_workspace_bytes(ctypes.byref(p))
    ws = torch.empty((max(ws_bytes, 4) // 4,), dtype=torch.float32, device=dev)
    if cost_q is not None or cost_r is not None:
        if policy != "gpc":
            raise ValueError("cost_q / cost_r weight GPC's cost_fn")
        cq = None if cost_q is None else _f32(torch.as_tensor(cost_q, dtype=torch.float32, device=dev).contiguous(),
                                              "cost_q", (d,))
        cr = None if cost_r is None else _f32(torch.as_tensor(cost_r, dtype=torch.float32, device=dev).contiguous(),
                                              "cost_r", (m,))
        sized = _abi.DlcLdsParams.from_buffer_copy(p)             # (workspace as for the runtime-dims kernel)
        sized.kernel_variant = 1
        nb = lib.dlc_lds_workspace_bytes(ctypes.byref(sized))
        if nb > ws.numel() * 4:
            ws = torch.empty((nb // 4 + 1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dlc_lds_rollout_cost_f32(ctypes.byref(p), _abi.ptr(A), _abi.ptr(B), _abi.ptr(K), _abi.ptr(W),
                                              _abi.ptr(x_io), _abi.ptr(M_io), _abi.ptr(hist_io), _abi.ptr(xprev_io),
                                              _abi.ptr(uprev_io), _abi.ptr(cq), _abi.ptr(cr), _abi.ptr(costs),
                                              _abi.ptr(cost_sum), _abi.ptr(X), _abi.ptr(U), _abi.ptr(status),
                                              _abi.ptr(ws), ws.numel() * 4, _abi.current_stream_ptr())
        _abi.check(rc, "dlc_lds_rollout_cost_f32")
        return LdsRolloutResult(costs=costs, cost_sum=cost_sum, x=x_io, M=M_io, hist=hist_io, xprev=xprev_io,
                                uprev=uprev_io, eps=None, bpc_cost=None, status=status, X=X, U=U, t=t0 + T)
    with torch.cuda.device(dev):
        rc = lib.dlc_lds_rollout_f32(ctypes.byref(p), _abi.ptr(A), _abi.ptr(B), _abi.ptr(K), _abi.ptr(W), _abi.ptr(u_in),
                                     _abi.ptr(x_io), _abi.ptr(M_io), _abi.ptr(hist_io), _abi.ptr(xprev_io),
                                     _abi.ptr(uprev_io), _abi.ptr(eps_io), _abi.ptr(eps_new),
                                     _abi.ptr(cost_io), _abi.ptr(costs), _abi.ptr(cost_sum), _abi.ptr(X),
                                     _abi.ptr(U), _abi.ptr(status), _abi.ptr(ws), ws.numel() * 4,
                                     _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_rollout_f32")
    return LdsRolloutResult(costs=costs, cost_sum=cost_sum, x=x_io, M=M_io, hist=hist_io, xprev=xprev_io,
                            uprev=uprev_io, eps=eps_io, bpc_cost=cost_io, status=status, X=X, U=U,
                            t=t0 + T)


_COPY_STREAMS = {}


def _copy_stream(dev):
    dev = torch.device(dev)
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[dev]


def lds_rollout_to_host(A, B, K, x, T, host_costs, *, chunks=8, W=None, t0=0, env_t0=None, donate=False,
                        materialize_disturbance=False, **kw):
    """``lds_rollout`` whose losses[T,N] (``Rollout.losses``, ``deluca/training/apg.py:68``) land in PINNED host memory
    while the rollout is still running: the horizon is cut into ``chunks`` launches resumed from the carried agent / env
    state (same arithmetic; the in-kernel disturbance stream is keyed by the env step counter), and each chunk's costs
    travel device -> host on a side stream under the next chunk's kernel.  Only the last chunk's copy is exposed.
    ``materialize_disturbance`` (off): with ``W=None``, write each chunk's slice of the in-kernel Gaussian stream to a
    reusable HBM buffer with ``dlc_gaussian_disturbance_f32`` first (the SAME stream, bit for bit) and let the rollout
    read it.  Measured on the target (tools/probe_e2e_chunks.py): no gain -- the generator pass costs what drawing the
    normals inside the rollout kernel does (391 vs 392 ms at 8 chunks); a chunk costs ~3 ms of state I/O and set-up.
    Returns the ``lds_rollout`` result of the whole horizon (``costs`` = ``host_costs``) plus ``done``, a CUDA event that
    completes when the last copy has; the caller must wait on it (or synchronise) before reading ``host_costs``."""
    if not (isinstance(host_costs, torch.Tensor) and not host_costs.is_cuda and host_costs.is_pinned()):
        raise TypeError("host_costs must be a pinned CPU tensor")
    N = x.shape[0]
    if tuple(host_costs.shape) != (T, N) or host_costs.dtype != torch.float32:
        raise ValueError(f"host_costs must be float32 [{T},{N}]")
    for k in ("keep_costs", "traj_stride"):
        if kw.get(k):
            raise ValueError(f"{k} is not available on the streamed path")
    kw.pop("keep_costs", None)
    dev = x.device
    main, side = torch.cuda.current_stream(dev), _copy_stream(dev)
    chunks = max(1, min(int(chunks), T))
    bounds = [round(i * T / chunks) for i in range(chunks + 1)]
    env_t0 = t0 if env_t0 is None else env_t0
    state = {k: kw.pop(k, None) for k in ("M", "hist", "xprev", "uprev", "eps", "bpc_cost")}
    eps_new = kw.pop("eps_new", None)
    cost_sum = status = r = None
    xc = x
    w_std = kw.get("w_std", 0.5)
    gen = W is None and materialize_disturbance and w_std > 0 and kw.get("policy", "gpc") != "open"
    wbuf = None
    if gen:
        wbuf = torch.empty((max(bounds[c + 1] - bounds[c] for c in range(chunks)), N, x.shape[1]), device=dev)
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi == lo:
            continue
        Wc = None if W is None else W[lo:hi]
        if gen:
            Wc = gaussian_disturbance(N, hi - lo, x.shape[1], kw.get("seed", 0), env_offset=kw.get("env_offset", 0),
                                      t0=env_t0 + lo, std=w_std, out=wbuf[:hi - lo])
        r = lds_rollout(A, B, K, xc, hi - lo, W=Wc, t0=t0 + lo, env_t0=env_t0 + lo,
                        eps_new=None if eps_new is None else eps_new[lo:hi], keep_costs=True,
                        donate=donate or c > 0, **state, **kw)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            host_costs[lo:hi].copy_(r.costs, non_blocking=True)
        r.costs.record_stream(side)
        cost_sum = r.cost_sum if cost_sum is None else cost_sum + r.cost_sum
        status = r.status if status is None else status | r.status
        xc = r.x
        state = {k: r[k] for k in state}
    done = torch.cuda.Event()
    done.record(side)
    out = LdsRolloutResult(r)
    out.update(costs=host_costs, cost_sum=cost_sum, status=status, t=t0 + T, done=done)
    return out


_SLICE_STREAMS = {}


def _slice_streams(dev):
    dev = torch.device(dev)
    if dev not in _SLICE_STREAMS:
        _SLICE_STREAMS[dev] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev),
                               (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)))
    return _SLICE_STREAMS[dev]


def copy_rows_async(dst, src, stream=None):
    """``dlc_copy_rows_async``: dst[r, :] = src[r, :] for 2-D tensors whose rows are contiguous (row pitches free), host
    or device on either side, as ONE asynchronous 2-D copy on ``stream`` (default: the current stream)."""
    if dst.dim() != 2 or src.dim() != 2 or dst.shape != src.shape or dst.dtype != src.dtype:
        raise ValueError("copy_rows_async needs two 2-D tensors of one shape and dtype")
    if dst.stride(1) != 1 or src.stride(1) != 1:
        raise ValueError("copy_rows_async: rows must be contiguous")
    for t in (dst, src):
        if not t.is_cuda and not t.is_pinned():
            raise TypeError("copy_rows_async: host tensors must be pinned")
    kind = 3 if (dst.is_cuda and src.is_cuda) else 2 if src.is_cuda else 1
    if not (dst.is_cuda or src.is_cuda):
        raise TypeError("copy_rows_async: at least one side must be a CUDA tensor")
    es = dst.element_size()
    dev = dst.device if dst.is_cuda else src.device
    st = torch.cuda.current_stream(dev) if stream is None else stream
    with torch.cuda.device(dev):
        rc = _abi.lib().dlc_copy_rows_async(ctypes.c_void_p(dst.data_ptr()), dst.stride(0) * es,
                                            ctypes.c_void_p(src.data_ptr()), src.stride(0) * es, dst.shape[1] * es,
                                            dst.shape[0], kind, ctypes.c_void_p(st.cuda_stream))
    _abi.check(rc, "dlc_copy_rows_async")


def lds_rollout_host_to_host(A, B, K, x, T, host_costs, *, device=None, slices=None, env_offset=0, **kw):
    """The whole end-to-end job of a vmapped rollout whose systems live on the HOST: per-env ``A[N,d,d] B[N,d,m]
    K[N,m,d] x[N,d]`` in pinned host memory in, ``losses[T,N]`` (``Rollout.losses``, ``deluca/training/apg.py:68``) in
    pinned host memory out.  Episodes never interact (``apg.py:91``), so the batch is cut into ``slices`` env ranges
    (default: up to 16, each a whole number of full waves of the rollout kernels' grids):
    slice k + 1's systems travel host -> device and slice k - 1's losses device -> host (one 2-D copy into columns
    [lo, hi) of ``host_costs``) while slice k's rollout runs -- every slice is ONE launch over the full horizon, so there
    is no carried agent state to write and re-read between launches (``lds_rollout_to_host`` cuts the horizon instead
    and pays ~3 ms per cut on the target job).  Slices alternate between two compute streams so that the tail of one
    slice's grid overlaps the head of the next.  Disturbances are drawn in the kernel, keyed by the GLOBAL env index
    (``env_offset`` + position): the result is the same job as one ``lds_rollout`` call on the whole batch.
    ``host_costs=None`` stores no per-step losses (per-env cost sums only).
    Returns ``cost_sum[N]`` (fp64), ``x[N,d]`` (final states), ``status[N]`` on the device, ``costs = host_costs`` and
    ``done``, a CUDA event that completes when the last copy has."""
    for name, t in (("A", A), ("B", B), ("K", K), ("x", x), ("host_costs", host_costs)):
        if name == "host_costs" and t is None:
            continue
        if not (isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned() and t.dtype == torch.float32
                and t.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous pinned float32 CPU tensor")
    if A.dim() != 3:
        raise ValueError("lds_rollout_host_to_host takes per-env systems A[N,d,d] (shared dynamics need no slicing)")
    N, d = x.shape
    if host_costs is not None and tuple(host_costs.shape) != (T, N):
        raise ValueError(f"host_costs must be float32 [{T},{N}]")
    for k in ("W", "keep_costs", "traj_stride", "donate", "M", "hist", "xprev", "uprev", "eps", "eps_new", "bpc_cost"):
        if kw.get(k) is not None and kw.get(k) is not False:
            raise ValueError(f"{k} is not available on the env-sliced path")
        kw.pop(k, None)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if slices is None:
        # whole waves of the rollout kernels' grids per slice (a CTA holds 32 envs, two CTAs per SM), at most 16
        # slices: a slice that ends mid-wave leaves SMs idle at its tail, and short slices start the downloads early
        # and leave little of the last one exposed (which is what a rank of an 8-GPU job, 14 waves, is bound by)
        wave = 64 * torch.cuda.get_device_properties(dev).multi_processor_count
        per = max(wave, -(-N // (int(os.environ.get("DLC_E2E_MAX_SLICES", "16")) * wave)) * wave)
        bounds = list(range(0, N, per)) + [N]
        if len(bounds) > 2 and N - bounds[-2] < wave:
            del bounds[-2]
    else:
        slices = max(1, min(int(slices), N)) if N else 1
        # slice boundaries on multiples of 32 envs (whole CTAs of every LDS kernel)
        bounds = sorted({min(N, (round(i * N / slices) + 31) // 32 * 32) for i in range(slices)} | {0, N})
    cost_sum = torch.empty((N,), dtype=torch.float64, device=dev)
    status = torch.empty((N,), dtype=torch.int32, device=dev)
    x_out = torch.empty((N, d), dtype=torch.float32, device=dev)
    up, down, comp = _slice_streams(dev)
    caller = torch.cuda.current_stream(dev)
    start = torch.cuda.Event()
    start.record(caller)
    for st in (up, down) + comp:
        st.wait_event(start)
    for i, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
        cs = comp[i % 2]
        with torch.cuda.stream(up):
            dA, dB, dK, dx = (t[lo:hi].to(dev, non_blocking=True) for t in (A, B, K, x))
            ready = torch.cuda.Event()
            ready.record(up)
        with torch.cuda.stream(cs):
            cs.wait_event(ready)
            for t in (dA, dB, dK, dx):
                t.record_stream(cs)
            r = lds_rollout(dA, dB, dK, dx, T, keep_costs=host_costs is not None, donate=True,
                            env_offset=env_offset + lo, **kw)
            cost_sum[lo:hi].copy_(r.cost_sum)
            status[lo:hi].copy_(r.status)
            x_out[lo:hi].copy_(r.x)
            finished = torch.cuda.Event()
            finished.record(cs)
        if host_costs is not None:
            down.wait_event(finished)
            copy_rows_async(host_costs[:, lo:hi], r.costs, stream=down)
            r.costs.record_stream(down)
    done = torch.cuda.Event()
    done.record(down)
    for cs in comp:
        caller.wait_stream(cs)
    return LdsRolloutResult(costs=host_costs, cost_sum=cost_sum, x=x_out, status=status, t=kw.get("t0", 0) + T,
                            done=done)


# ------------------------------------------------------------------------------------------ lung
def lung_rollout(params, bufs):
    """``dlc_lung_pid_rollout_f32`` on already-prepared ctypes structs (see deluca_b200.lung)."""
    any_t = next(v for v in bufs.values() if v is not None)
    b = _abi.DlcLungBuffers(**{k: (None if v is None else v.data_ptr()) for k, v in bufs.items()})
    with torch.cuda.device(any_t.device):
        rc = _abi.lib().dlc_lung_pid_rollout_f32(ctypes.byref(params), ctypes.byref(b), _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lung_pid_rollout_f32")


def pid_rollout(K_P, K_I, K_D, err, P, I, D, decay):
    """``dlc_pid_f32``: generic PID over err[T,N]; returns u[T,N] and the new (P, I, D)."""
    T, N = err.shape
    for name, t in (("K_P", K_P), ("K_I", K_I), ("K_D", K_D), ("P", P), ("I", I), ("D", D)):
        _f32(t, name, (N,))
    _f32(err, "err")
    P, I, D = P.clone(), I.clone(), D.clone()
    u = torch.empty_like(err)
    with torch.cuda.device(err.device):
        rc = _abi.lib().dlc_pid_f32(N, T, decay, _abi.ptr(K_P), _abi.ptr(K_I), _abi.ptr(K_D), _abi.ptr(err),
                                    _abi.ptr(P), _abi.ptr(I), _abi.ptr(D), _abi.ptr(u), _abi.current_stream_ptr())
    _abi.check(rc, "dlc_pid_f32")
    return u, P, I, D


# ------------------------------------------------------------------------------------------ igpc
def plan_params(env_true, env_sim, cost, N, H, T=0, algo=0, backtracking=True, fix_backtracking=False,
                n_alpha=10, w_is=0, alpha0=1.0, lr=0.1):
    """``DlcPlanParams`` from env/cost mirror objects (``env.spec()`` -> (kind, [static fields]))."""
    p = _abi.DlcPlanParams()
    p.N, p.H, p.T = N, H, T
    for dst, env in ((p.env_true, env_true), (p.env_sim, env_sim)):
        kind, vals = env.spec()
        dst.kind = kind
        for i, v in enumerate(vals):
            dst.p[i] = v
    d, m = env_true.state_dim, env_true.action_dim
    q, r, goal = cost.vectors(d, m)
    goal_u = cost.action_goal(m)
    for i in range(d):
        p.q[i], p.goal[i] = q[i], goal[i]
    for i in range(m):
        p.r[i], p.goal_u[i] = r[i], goal_u[i]
    p.algo, p.backtracking, p.fix_backtracking, p.n_alpha = algo, int(backtracking), int(fix_backtracking), n_alpha
    p.w_is, p.alpha0, p.lr = w_is, alpha0, lr
    return p


def _launch_on(t, fn, *args):
    with torch.cuda.device(t.device):
        return fn(*args, _abi.current_stream_ptr())


def env_rollout(p, U_old, x0, k=None, K=None, X_old=None, alpha=None, use_sim=False):
    N, H, m = U_old.shape
    d = x0.shape[1]
    for name, t in (("U_old", U_old), ("x0", x0), ("k", k), ("K", K), ("X_old", X_old), ("alpha", alpha)):
        _f32(t, name)
    X = torch.empty(N, H + 1, d, device=x0.device)
    U = torch.empty(N, H, m, device=x0.device)
    c = torch.empty(N, device=x0.device)
    rc = _launch_on(x0, _abi.lib().dlc_env_rollout_f32, ctypes.byref(p), int(use_sim), _abi.ptr(U_old), _abi.ptr(k),
                    _abi.ptr(K), _abi.ptr(X_old), _abi.ptr(alpha), _abi.ptr(x0), _abi.ptr(X), _abi.ptr(U), _abi.ptr(c))
    _abi.check(rc, "dlc_env_rollout_f32")
    return X, U, c


def igpc_rollout(p, U_old, k, K, X_old, alpha, x0):
    N, H, m = U_old.shape
    d = x0.shape[1]
    for name, t in (("U_old", U_old), ("x0", x0), ("k", k), ("K", K), ("X_old", X_old), ("alpha", alpha)):
        _f32(t, name)
    X = torch.empty(N, H + 1, d, device=x0.device)
    U = torch.empty(N, H, m, device=x0.device)
    c = torch.empty(N, device=x0.device)
    rc = _launch_on(x0, _abi.lib().dlc_igpc_rollout_f32, ctypes.byref(p), _abi.ptr(U_old), _abi.ptr(k), _abi.ptr(K),
                    _abi.ptr(X_old), _abi.ptr(alpha), _abi.ptr(x0), _abi.ptr(X), _abi.ptr(U), _abi.ptr(c))
    _abi.check(rc, "dlc_igpc_rollout_f32")
    return X, U, c


def linearize(p, X, U, use_sim=True):
    N, H, m = U.shape
    d = X.shape[2]
    _f32(X, "X", (N, H + 1, d)); _f32(U, "U")
    e = lambda *s: torch.empty(N, H, *s, device=X.device)
    D, Fx, Fu, cx, cu, cxx, cuu = e(d), e(d, d), e(d, m), e(d), e(m), e(d, d), e(m, m)
    rc = _launch_on(X, _abi.lib().dlc_linearize_f32, ctypes.byref(p), int(use_sim), _abi.ptr(X), _abi.ptr(U),
                    _abi.ptr(D), _abi.ptr(Fx), _abi.ptr(Fu), _abi.ptr(cx), _abi.ptr(cu), _abi.ptr(cxx), _abi.ptr(cuu))
    _abi.check(rc, "dlc_linearize_f32")
    return D, (Fx, Fu), (cx, cu, cxx, cuu)


def riccati_backward(F, C):
    Fx, Fu = F
    cx, cu, cxx, cuu = C
    N, H, d, m = Fu.shape
    for t in (Fx, Fu, cx, cu, cxx, cuu):
        _f32(t, "F/C leaf")
    k = torch.empty(N, H, m, device=Fx.device)
    K = torch.empty(N, H, m, d, device=Fx.device)
    status = torch.empty(N, dtype=torch.int32, device=Fx.device)
    rc = _launch_on(Fx, _abi.lib().dlc_riccati_backward_f32, N, H, d, m, _abi.ptr(Fx), _abi.ptr(Fu), _abi.ptr(cx),
                    _abi.ptr(cu), _abi.ptr(cxx), _abi.ptr(cuu), _abi.ptr(k), _abi.ptr(K), _abi.ptr(status))
    _abi.check(rc, "dlc_riccati_backward_f32")
    return k, K, status


def lqg_backward(F, C, reg_lambda, multiplier, factor, lambda_min, reg_type):
    """``dlc_lqg_backward_f32``: regularised Riccati pass with the lambda schedule on the device.
    reg_lambda / multiplier are [N] tensors updated in place.  Returns k, K, gains[N,2], status[N]."""
    Fx, Fu = F
    cx, cu, cxx, cuu, cux = C
    N, H, d, m = Fu.shape
    for t in (Fx, Fu, cx, cu, cxx, cuu, cux, reg_lambda, multiplier):
        _f32(t, "F/C leaf")
    k = torch.zeros(N, H, m, device=Fx.device)
    K = torch.zeros(N, H, m, d, device=Fx.device)
    gains = torch.empty(N, 2, device=Fx.device)
    status = torch.empty(N, dtype=torch.int32, device=Fx.device)
    with torch.cuda.device(Fx.device):
        rc = _abi.lib().dlc_lqg_backward_f32(N, H, d, m, _abi.ptr(Fx), _abi.ptr(Fu), _abi.ptr(cx), _abi.ptr(cu),
                                             _abi.ptr(cxx), _abi.ptr(cuu), _abi.ptr(cux), _abi.ptr(reg_lambda),
                                             _abi.ptr(multiplier), float(factor), float(lambda_min), int(reg_type),
                                             _abi.ptr(k), _abi.ptr(K), _abi.ptr(gains), _abi.ptr(status),
                                             _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lqg_backward_f32")
    return k, K, gains, status


def plan(p, U, x0, trace=False):
    """``dlc_plan_f32``: the whole iLQR / iGPC / iLC loop.  Returns X, U, k, K, cost, alpha, status; with ``trace`` also
    (decisions[T,N] int32 -- the accepted backtracking candidate per outer iteration, -1 = none -- and costs[T,N])."""
    N, H, m = U.shape
    d = x0.shape[1]
    _f32(U, "U"); _f32(x0, "x0", (N, d))
    dev = x0.device
    U_io = U.clone()
    X = torch.empty(N, H + 1, d, device=dev)
    k = torch.zeros(N, H, m, device=dev)
    K = torch.zeros(N, H, m, d, device=dev)
    c = torch.empty(N, device=dev)
    alpha = torch.empty(N, device=dev)
    status = torch.empty(N, dtype=torch.int32, device=dev)
    nbytes = _abi.lib().dlc_plan_workspace_bytes(ctypes.byref(p))
    ws = torch.empty(max(nbytes, 4) // 4, device=dev)
    with torch.cuda.device(dev):
        dec = torch.empty(p.T, N, dtype=torch.int32, device=dev) if trace else None
        ctr = torch.empty(p.T, N, device=dev) if trace else None
        rc = _abi.lib().dlc_plan_f32(ctypes.byref(p), _abi.ptr(U_io), _abi.ptr(x0), _abi.ptr(X), _abi.ptr(k),
                                     _abi.ptr(K), _abi.ptr(c), _abi.ptr(alpha), _abi.ptr(status), _abi.ptr(dec),
                                     _abi.ptr(ctr), _abi.ptr(ws), ws.numel() * 4, _abi.current_stream_ptr())
    _abi.check(rc, "dlc_plan_f32")
    if trace:
        return X, U_io, k, K, c, alpha, status, dec, ctr
    return X, U_io, k, K, c, alpha, status


# ------------------------------------------------------------------------------------------ shared dynamics
def lds_shared_pack(A, B, K):
    """``dlc_lds_shared_pack_f32``: [A - B K ; -K] as TF32 hi/lo in the tensor-core staging layout."""
    d, m = B.shape
    _f32(A, "A", (d, d)); _f32(B, "B", (d, m)); _f32(K, "K", (m, d))
    n = _abi.lib().dlc_lds_shared_packed_floats(d, m)
    if n == 0:
        raise ValueError("shared-dynamics tensor-core path needs d % 16 == 0, 16 <= d <= 256, d + m <= 384")
    G = torch.empty(n, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        rc = _abi.lib().dlc_lds_shared_pack_f32(_abi.ptr(A), _abi.ptr(B), _abi.ptr(K), d, m, _abi.ptr(G),
                                                _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_shared_pack_f32")
    return G


def lds_shared_lqr_rollout(G_packed, x, T, m, W=None, keep_costs=True):
    """``dlc_lds_shared_lqr_rollout_f32``: T steps of shared-dynamics LDS + LQR for N envs on tcgen05.
    Returns costs[T,N] (or None), cost_sum[N] (fp64), x_final[N,d], status[N]."""
    _f32(x, "x")
    N, d = x.shape
    _f32(W, "W", (T, N, d))
    x_io = x.clone()
    costs = torch.empty(T, N, device=x.device) if keep_costs else None
    cost_sum = torch.empty(N, dtype=torch.float64, device=x.device)
    status = torch.empty(N, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _abi.lib().dlc_lds_shared_lqr_rollout_f32(N, T, d, m, _abi.ptr(G_packed), _abi.ptr(W), _abi.ptr(x_io),
                                                       _abi.ptr(costs), _abi.ptr(cost_sum), _abi.ptr(status),
                                                       _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_shared_lqr_rollout_f32")
    return LdsRolloutResult(costs=costs, cost_sum=cost_sum, x=x_io, status=status)


def lds_shared_gpc_rollout(A, B, K, x, T, *, H=16, HH=None, W=None, M=None, hist=None, xprev=None, t0=0,
                           lr_scale=0.005, decay=True, keep_costs=True, donate=False):
    """``dlc_lds_shared_gpc_rollout_f32``: T steps of LDS + GPC with shared A[d,d], B[d,m], K[m,d] and a per-env
    policy tensor M[N,H,m,d] (``deluca/agents/_gpc.py:109-183`` under vmap with ``in_axes=None`` dynamics).
    ``M``/``hist``/``xprev`` default to the ``GPC.__init__`` state (zeros); pass the returned ones (and ``t0``) to
    resume.  ``donate=True`` updates the given state tensors in place instead of cloning them."""
    _f32(x, "x")
    N, d = x.shape
    m = B.shape[1]
    HH = H if HH is None else HH
    _f32(A, "A", (d, d)); _f32(B, "B", (d, m)); _f32(K, "K", (m, d)); _f32(W, "W", (T, N, d))
    dev = x.device
    own = lambda t, shape, name: (torch.zeros(shape, device=dev) if t is None
                                  else (_f32(t, name, shape) if donate else _f32(t, name, shape).clone()))
    x_io = x if donate else x.clone()
    M_io = own(M, (N, H, m, d), "M")
    hist_io = own(hist, (N, H + HH, d), "hist")
    xprev_io = own(xprev, (N, d), "xprev")
    p = _abi.DlcSharedGpcParams(N=N, T=T, d=d, m=m, H=H, HH=HH, t0=int(t0), decay=int(bool(decay)),
                                lr_scale=float(lr_scale), reserved=0)
    nbytes = _abi.lib().dlc_lds_shared_gpc_workspace_bytes(ctypes.byref(p))
    if nbytes == 0 and N > 0:
        raise ValueError("shared-dynamics GPC path is built for (d, m, H = HH) in {(256, 64, 16), (64, 32, 4)}")
    ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device=dev)
    costs = torch.empty(T, N, device=dev) if keep_costs else None
    cost_sum = torch.zeros(N, dtype=torch.float64, device=dev)
    status = torch.zeros(N, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _abi.lib().dlc_lds_shared_gpc_rollout_f32(ctypes.byref(p), _abi.ptr(A), _abi.ptr(B), _abi.ptr(K),
                                                       _abi.ptr(W), _abi.ptr(x_io), _abi.ptr(M_io),
                                                       _abi.ptr(hist_io), _abi.ptr(xprev_io), _abi.ptr(costs),
                                                       _abi.ptr(cost_sum), _abi.ptr(status), _abi.ptr(ws),
                                                       ws.numel() * 4, _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_shared_gpc_rollout_f32")
    return LdsRolloutResult(costs=costs, cost_sum=cost_sum, x=x_io, M=M_io, hist=hist_io, xprev=xprev_io,
                            t=t0 + T, status=status)


def lung_pid_value_and_grad(params, K, kf, tt, loss_kind=0, noise=0.0, reset_pressure=0.0, want_sums=True):
    """``dlc_lung_pid_grad_f32``: per-breath loss[N], dloss/dK[N,3] and their batch sums [4] (fp64)."""
    N, T = params.N, params.T
    dev = K.device
    _f32(K, "K", (N, 3)); _f32(kf, "kf"); _f32(tt, "tt", (T,))
    loss = torch.empty(N, device=dev)
    grad = torch.empty(N, 3, device=dev)
    sums = torch.zeros(4, dtype=torch.float64, device=dev) if want_sums else None
    nbytes = _abi.lib().dlc_lung_pid_grad_workspace_bytes(ctypes.byref(params))
    ws = torch.empty(max(nbytes, 4) // 4, device=dev)
    with torch.cuda.device(dev):
        rc = _abi.lib().dlc_lung_pid_grad_f32(ctypes.byref(params), _abi.ptr(K), _abi.ptr(kf), _abi.ptr(tt),
                                              int(loss_kind), float(noise), float(reset_pressure), _abi.ptr(loss),
                                              _abi.ptr(grad), _abi.ptr(sums), _abi.ptr(ws), ws.numel() * 4,
                                              _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lung_pid_grad_f32")
    return loss, grad, sums


def lds_lqr_value_and_grad(A, B, K, x0, T, W=None, seed=0, env_offset=0, t0=0, w_std=0.5, want_x0_grad=False):
    """``dlc_lds_lqr_grad_f32``: loss[N] = sum_t quad_loss(x_t, -K x_t), dloss/dK[N,m,d] (per env), optional
    dloss/dx0[N,d], and the batch sums [1 + m*d] (fp64)."""
    _f32(x0, "x0")
    N, d = x0.shape
    shared = A.dim() == 2
    m = B.shape[-1]
    _f32(A, "A", (d, d) if shared else (N, d, d)); _f32(B, "B", (d, m) if shared else (N, d, m))
    _f32(K, "K", (m, d) if shared else (N, m, d)); _f32(W, "W", (T, N, d))
    dev = x0.device
    loss = torch.empty(N, device=dev)
    gK = torch.empty(N, m, d, device=dev)
    gx = torch.empty(N, d, device=dev) if want_x0_grad else None
    sums = torch.zeros(1 + m * d, dtype=torch.float64, device=dev)
    nbytes = _abi.lib().dlc_lds_lqr_grad_workspace_bytes(N, T, d)
    ws = torch.empty(max(nbytes, 4) // 4, device=dev)
    with torch.cuda.device(dev):
        rc = _abi.lib().dlc_lds_lqr_grad_f32(N, T, d, m, int(shared), _abi.ptr(A), _abi.ptr(B), _abi.ptr(K),
                                             _abi.ptr(x0), _abi.ptr(W), seed, env_offset, t0, w_std, _abi.ptr(loss),
                                             _abi.ptr(gK), _abi.ptr(gx), _abi.ptr(sums), _abi.ptr(ws), ws.numel() * 4,
                                             _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_lqr_grad_f32")
    return loss, gK, gx, sums


def lds_of_rollout(A, B, C, x, T, *, agent, W=None, y_in=None, K=None, Lk=None, Phi=None, theta=None, z=None,
                   ynat=None, us=None, m=0, h=0, t0=0, lr=0.005, decay=True, R_M=10.0, seed=0, env_offset=0,
                   w_std=0.5, agent_only=False, keep_costs=True, keep_actions=True, donate=False, lanes_per_env=0, phi_is_identity=False,
                   stable_a=False, env_t0=None):
    """``dlc_lds_of_rollout_f32``: T fused steps of the partially observed LDS (``deluca/envs/_lds.py:102-111``,
    y = C x) under an output-feedback controller: ``agent`` = "lqg" (``_lqg.py:83-98``; K[.,n,d], Lk[.,d,p],
    z = x_hat), "drc" (``_drc.py:131-203``; theta = M[N,m,n,p], ynat[N,m,p] and us[N,h,n] newest first,
    K[.,n,p] or None) or "fgrc" (GRC/SFC/DSC, ``_grc.py:110-169``; Phi[F,L], theta[N,F,n,p], z[N,d],
    ynat[N,L+m,p] oldest -> newest).  State tensors default to the agents' construction-time zeros; the returned
    ones (and ``t``) resume the episode.  ``agent_only``: ``Agent.__call__(y)`` with y = ``y_in[T,N,p]``.
    ``lanes_per_env``: 0 (automatic) or 8 / 16 / 32 lanes of a warp per env.  ``phi_is_identity``: the caller's
    promise that ``Phi`` is the identity (GRC); ``stable_a``: the promise that the truncated response decays,
    ``|A^h B|_F <= |B|_F`` (DRC: ``DLC_OF_FLAG_STABLE_A``); both verified on the device.  ``t0``: the agent's step counter (lr decay);
    ``env_t0``: the env's (counter of the in-kernel disturbance stream; defaults to ``t0``)."""
    shared = A.dim() == 2
    d, n, p = A.shape[-1], B.shape[-1], C.shape[-2]
    if agent_only:
        _f32(y_in, "y_in")
        N = y_in.shape[1]
        _f32(y_in, "y_in", (T, N, p))
        dev = y_in.device
        x_io = None
    else:
        _f32(x, "x")
        N = x.shape[0]
        _f32(x, "x", (N, d))
        dev = x.device
        x_io = x if donate else x.clone()
    lead = () if shared else (N,)
    _f32(A, "A", lead + (d, d)); _f32(B, "B", lead + (d, n)); _f32(C, "C", lead + (p, d)); _f32(W, "W", (T, N, d))
    own = lambda t, shape, name: (torch.zeros(shape, device=dev) if t is None
                                  else (_f32(t, name, shape) if donate else _f32(t, name, shape).clone()))
    code = _abi.OF_AGENT[agent]
    F = L = 0
    theta_io = z_io = ynat_io = us_io = None
    if agent == "lqg":
        _f32(K, "K", lead + (n, d)); _f32(Lk, "Lk", lead + (d, p))
        z_io = own(z, (N, d), "z")
    elif agent == "drc":
        _f32(K, "K", lead + (n, p))
        if theta is None:
            raise ValueError("drc: theta (M[N,m,n,p]) is required -- the reference draws it from PRNGKey(0) (_drc.py:105)")
        theta_io = own(theta, (N, m, n, p), "theta")
        ynat_io = own(ynat, (N, m, p), "ynat")
        us_io = own(us, (N, h, n), "us")
    else:
        _f32(Phi, "Phi")
        F, L = Phi.shape
        if theta is None:
            raise ValueError("fgrc: theta[N,F,n,p] is required -- the reference draws it from PRNGKey(0) (_grc.py:77)")
        theta_io = own(theta, (N, F, n, p), "theta")
        z_io = own(z, (N, d), "z")
        ynat_io = own(ynat, (N, L + m, p), "ynat")
    prm = _abi.DlcOfParams(N=N, env_offset=int(env_offset), T=T, d=d, n=n, p=p, agent=code, m=int(m), h=int(h), F=F,
                           L=L, t0=int(t0), decay=int(bool(decay)), shared_dynamics=int(shared),
                           mode=1 if agent_only else 0, lr=float(lr), R_M=float(R_M), w_std=float(w_std),
                           lanes_per_env=int(lanes_per_env), flags=int(bool(phi_is_identity)) | (2 * int(bool(stable_a))), seed=int(seed),
                           env_t0=int(t0 if env_t0 is None else env_t0))
    costs = torch.empty(T, N, device=dev) if keep_costs else None
    U = torch.empty(T, N, n, device=dev) if keep_actions else None
    cost_sum = torch.zeros(N, dtype=torch.float64, device=dev)
    status = torch.zeros(N, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _abi.lib().dlc_lds_of_rollout_f32(ctypes.byref(prm), _abi.ptr(A), _abi.ptr(B), _abi.ptr(C), _abi.ptr(K),
                                               _abi.ptr(Lk), _abi.ptr(Phi), _abi.ptr(W), _abi.ptr(y_in),
                                               _abi.ptr(x_io), _abi.ptr(theta_io), _abi.ptr(z_io), _abi.ptr(ynat_io),
                                               _abi.ptr(us_io), _abi.ptr(costs), _abi.ptr(cost_sum), _abi.ptr(U),
                                               _abi.ptr(status), _abi.current_stream_ptr())
    _abi.check(rc, "dlc_lds_of_rollout_f32")
    return LdsRolloutResult(costs=costs, cost_sum=cost_sum, U=U, x=x_io, theta=theta_io, z=z_io, ynat=ynat_io,
                            us=us_io, t=t0 + T, status=status)


def gae(losses, values, log_probs=None, done_masks=None, discount=0.99, gae_param=0.95, time_major=False,
        returns_use_values=False, want_returns=True):
    """``gae_advantages`` + the returns line of ``ac_rollout`` (``deluca/training/ppo.py:44-84,131-134``)
    over [N,T] (or [T,N] with ``time_major``) series -> (advantages, returns | None)."""
    _f32(losses, "losses")
    if losses.dim() != 2:
        raise ValueError("losses must be [N,T] (or [T,N] with time_major)")
    for name, t in (("values", values), ("log_probs", log_probs), ("done_masks", done_masks)):
        _f32(t, name, losses.shape)
    if want_returns and not returns_use_values and log_probs is None:
        raise ValueError("returns = advantages + log_probs (ppo.py:133, as written) needs log_probs")
    T, N = losses.shape if time_major else losses.shape[::-1]
    adv = torch.empty_like(losses)
    ret = torch.empty_like(losses) if want_returns else None
    rc = _launch_on(losses, _abi.lib().dlc_gae_f32, N, T, int(time_major), _abi.ptr(losses), _abi.ptr(values),
                    _abi.ptr(log_probs), _abi.ptr(done_masks), float(discount), float(gae_param),
                    int(returns_use_values), _abi.ptr(adv), _abi.ptr(ret))
    _abi.check(rc, "dlc_gae_f32")
    return adv, ret


class EnvAgentResult(NamedTuple):
    x: torch.Tensor                 # [N,d] final env state
    agent_state: Optional[tuple]    # PID: (P, I, D) each [N,d]
    losses: Optional[torch.Tensor]  # [T,N]
    actions: Optional[torch.Tensor] # [T,N,da]
    loss_sum: torch.Tensor          # [N] float64
    grad: Optional[torch.Tensor]    # [N,ng] d(sum_t loss)/d(gains)
    agent_stack: Optional[torch.Tensor]  # [T // state_stride, N, 3, d]
    x_stack: Optional[torch.Tensor]      # [T // state_stride, N, d]
    status: torch.Tensor


def env_agent_rollout(env, cost, agent, gains, x0, T, *, pid_state=None, pid_decay=0.0, keep_losses=True,
                      keep_actions=True, state_stride=0, want_grad=False, donate=False):
    """``dlc_env_agent_rollout_f32``: ``apg_rollout`` x ``vmap`` (``deluca/training/apg.py:39-100``) of a classic env
    under the generic PID agent (``agent="pid"``, ``gains`` [3] or [N,3]) or linear state feedback (``"linear"``,
    ``gains`` [m,d] or [N,m,d], u = -K x) with the quadratic loss ``cost`` (an object with ``vectors(d,m)`` /
    ``action_goal(m)``: sum q (x - goal)^2 + sum r (u - goal_u)^2)."""
    N, d = x0.shape
    m = env.action_dim
    _f32(x0, "x0", (N, env.state_dim))
    pid = agent == "pid"
    if not pid and agent != "linear":
        raise ValueError("agent must be 'pid' or 'linear'")
    ng, da = (3, d) if pid else (m * d, m)
    _f32(gains, "gains")
    shp = (3,) if pid else (m, d)
    if tuple(gains.shape) not in (shp, (N,) + shp):
        raise ValueError(f"gains must be {shp} (shared) or {(N,) + shp} (one set per env), got {tuple(gains.shape)}")
    p = _abi.DlcEnvAgentParams()
    p.N, p.T, p.agent = N, int(T), 0 if pid else 1
    p.env.kind, vals = env.spec()
    for i, v in enumerate(vals):
        p.env.p[i] = v
    p.gains_per_env = int(gains.dim() == (2 if pid else 3))
    p.state_stride, p.want_grad, p.pid_decay = int(state_stride), int(bool(want_grad)), float(pid_decay)
    q, r, goal = cost.vectors(d, m)
    gu = cost.action_goal(m)
    for i in range(d):
        p.q[i], p.goal[i] = q[i], goal[i]
    for i in range(m):
        p.r[i], p.goal_u[i] = r[i], gu[i]
    dev = x0.device
    x_io = x0 if donate else x0.clone()
    a_io = None
    if pid:
        if pid_state is None:
            a_io = torch.zeros(N, 3, d, device=dev)
        else:
            a_io = torch.stack([torch.as_tensor(s, dtype=torch.float32, device=dev).expand(N, d) for s in pid_state], 1)
            a_io = a_io.contiguous()
    losses = torch.empty(T, N, device=dev) if keep_losses else None
    actions = torch.empty(T, N, da, device=dev) if keep_actions else None
    ns = T // state_stride if state_stride > 0 else 0
    a_stack = torch.empty(ns, N, 3, d, device=dev) if (ns and pid) else None
    x_stack = torch.empty(ns, N, d, device=dev) if ns else None
    lsum = torch.empty(N, dtype=torch.float64, device=dev)
    grad = torch.empty(N, ng, device=dev) if want_grad else None
    status = torch.empty(N, dtype=torch.int32, device=dev)
    rc = _launch_on(x0, _abi.lib().dlc_env_agent_rollout_f32, ctypes.byref(p), _abi.ptr(gains), _abi.ptr(x_io),
                    _abi.ptr(a_io), _abi.ptr(losses), _abi.ptr(actions), _abi.ptr(a_stack), _abi.ptr(x_stack),
                    _abi.ptr(lsum), _abi.ptr(grad), _abi.ptr(status))
    _abi.check(rc, "dlc_env_agent_rollout_f32")
    st = (a_io[:, 0], a_io[:, 1], a_io[:, 2]) if pid else None
    return EnvAgentResult(x_io, st, losses, actions, lsum, grad, a_stack, x_stack, status)


def env_collect(env, x0, T, seed=0, env_offset=0, action_std=1.0):
    """``Learner.generate_trajectories`` (``deluca/learners/core.py:17-41``) on a classic env:
    (obs[N,T,d], action[N,T,m], next_obs[N,T,d]) under the N(0, action_std^2) random policy."""
    N, d = x0.shape
    _f32(x0, "x0", (N, env.state_dim))
    spec = _abi.DlcEnvSpec()
    spec.kind, vals = env.spec()
    for i, v in enumerate(vals):
        spec.p[i] = v
    obs = torch.empty(N, T, d, device=x0.device)
    act = torch.empty(N, T, env.action_dim, device=x0.device)
    nxt = torch.empty(N, T, d, device=x0.device)
    rc = _launch_on(x0, _abi.lib().dlc_env_collect_f32, ctypes.byref(spec), N, int(T), int(seed), int(env_offset),
                    float(action_std), _abi.ptr(x0), _abi.ptr(obs), _abi.ptr(act), _abi.ptr(nxt))
    _abi.check(rc, "dlc_env_collect_f32")
    return obs, act, nxt


def lds_open_rollout(A, B, C, x, T, *, u_in=None, W=None, seed=0, env_offset=0, t0=0, w_std=0.5, action_std=1.0,
                     observe0=False, keep_actions=False, want_obs=True, donate=False):
    """``dlc_lds_open_rollout_f32``: T steps of ``LDS.__call__`` (``deluca/envs/_lds.py:102-111``) for N envs under
    the given actions ``u_in[T,N,n]`` or, when None, N(0, action_std^2) actions (``generate_random_trajectory``,
    ``:117-140``).  Returns (x_final[N,d], y[T (+1 with observe0),N,p] or None, u[T,N,n] or None)."""
    _f32(x, "x")
    N, d = x.shape
    shared = A.dim() == 2
    n = B.shape[-1]
    lead = () if shared else (N,)
    _f32(A, "A", lead + (d, d)); _f32(B, "B", lead + (d, n))
    p = 1
    if C is not None:
        p = C.shape[-2]
        _f32(C, "C", lead + (p, d))
    elif want_obs:
        raise ValueError("observations need C")
    _f32(u_in, "u_in", (T, N, n)); _f32(W, "W", (T, N, d))
    x_io = x if donate else x.clone()
    y = torch.empty(T + int(bool(observe0)), N, p, device=x.device) if want_obs else None
    u = torch.empty(T, N, n, device=x.device) if keep_actions else None
    rc = _launch_on(x, _abi.lib().dlc_lds_open_rollout_f32, N, int(T), d, n, p, int(shared), _abi.ptr(A), _abi.ptr(B),
                    _abi.ptr(C), _abi.ptr(x_io), _abi.ptr(u_in), _abi.ptr(W), int(seed), int(env_offset), int(t0),
                    float(w_std), float(action_std), int(bool(observe0)), _abi.ptr(y), _abi.ptr(u))
    _abi.check(rc, "dlc_lds_open_rollout_f32")
    return x_io, y, u
