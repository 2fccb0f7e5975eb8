"""LDS + GPC / BPC / LQR restated on the CPU (test infrastructure; see oracle/__init__.py).

Parity pinned to the reference's own files run here: ``tests/golden/ref_lds.npz`` (``deluca/envs/_lds.py``,
``deluca/agents/_gpc.py``, ``_bpc.py``, ``_lqr.py`` executed literally by tests/golden/make_ref_fixtures.py on
tests/refshim; the reference ships no test or golden vector of its own for them) -- this module reproduces those runs
to <= 1e-9 relative in float64 (tests/test_ref_fixtures_oracle.py).

Everything is batched over a leading env axis N (the reference's vmap axis,
``deluca/training/apg.py:91``); vectors are stored flat ``[N, d]`` instead of the
reference's ``[d, 1]`` columns.  ``dtype`` is the arithmetic type (float32 = default JAX).
"""
import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------- LDS
def lds_default_params(rng, N, d, m, dtype=np.float32):
    """Per-env draws with the distributions of ``LDS.init`` (``deluca/envs/_lds.py:69-91``):
    A = diag(U(0.9, 0.95)), B ~ N(0, 1), x0 = 0.  (The reference draws from JAX Threefry;
    the distributions, not the bit streams, are what is reproducible here.)"""
    A = np.zeros((N, d, d), dtype=dtype)
    idx = np.arange(d)
    A[:, idx, idx] = rng.uniform(0.9, 0.95, size=(N, d)).astype(dtype)
    B = rng.standard_normal((N, d, m)).astype(dtype)
    x0 = np.zeros((N, d), dtype=dtype)
    return A, B, x0


def lds_step(A, B, x, u, w):
    """``LDS.__call__`` (``deluca/envs/_lds.py:102-111``): x <- A x + B u + w_t (left to right)."""
    return (np.einsum("nij,nj->ni", A, x) + np.einsum("nij,nj->ni", B, u)) + w


def quad_loss(x, u):
    """``quad_loss`` (``deluca/agents/_gpc.py:29-40``): x'x + u'u."""
    return np.einsum("ni,ni->n", x, x) + np.einsum("ni,ni->n", u, u)


# --------------------------------------------------------------------------- LQR
def lqr_gain(A, B, Q=None, R=None, dtype=np.float32):
    """``LQR.__init__`` (``deluca/agents/_lqr.py:29-60``): X = scipy DARE(A,B,Q,R);
    K = (B'XB + R)^-1 (B'XA).  Solved per env in float64 and cast (the reference hands
    fp32 jnp arrays to scipy, which computes in double)."""
    N, d, m = B.shape
    K = np.zeros((N, m, d), dtype=dtype)
    for n in range(N):
        An, Bn = A[n].astype(np.float64), B[n].astype(np.float64)
        Qn = np.eye(d) if Q is None else np.asarray(Q, dtype=np.float64)
        Rn = np.eye(m) if R is None else np.asarray(R, dtype=np.float64)
        X = scipy.linalg.solve_discrete_are(An, Bn, Qn, Rn)
        K[n] = (np.linalg.inv(Bn.T @ X @ Bn + Rn) @ (Bn.T @ X @ An)).astype(dtype)
    return K


def lqr_action(K, x):
    """``LQR.__call__`` (``deluca/agents/_lqr.py:62-72``): u = -K x."""
    return -np.einsum("nij,nj->ni", K, x)


# --------------------------------------------------------------------------- GPC
def _window(w, start, size):
    """``jax.lax.dynamic_slice_in_dim(w, start, size)``: out-of-range starts are clamped to
    [0, n - size] (JAX semantics as documented; not executable here -- SURVEY A-3.5).
    Never exercised when HH >= H - 1."""
    n = w.shape[1]
    start = min(max(start, 0), n - size)
    return w[:, start:start + size]


def gpc_policy_loss(M, w, A, B, K, H, HH, cost_q=None, cost_r=None):
    """``policy_loss`` (``deluca/agents/_gpc.py:109-125``), literal forward restatement.

    M[N,H,m,d], w[N,H+HH,d] oldest -> newest.  Returns loss[N].  ``cost_q[d]`` / ``cost_r[m]`` stand for a caller's
    diagonal quadratic ``cost_fn(x, u) = sum(q x^2) + sum(r u^2)`` (``_gpc.py:52,76``); None is ``quad_loss``."""
    N, P, d = w.shape

    def action(state, h):                                     # _gpc.py:112-116
        return -np.einsum("nij,nj->ni", K, state) + np.einsum(
            "nhij,nhj->ni", M, _window(w, h, H))

    y = np.zeros((N, d), dtype=M.dtype)                       # _gpc.py:123
    for h in range(H - 1):                                    # _gpc.py:118-124
        idx = min(h + H, P - 1)                               # traced scalar index clamps
        y = (np.einsum("nij,nj->ni", A, y)
             + np.einsum("nij,nj->ni", B, action(y, h))) + w[:, idx]
    u = action(y, HH - 1)
    if cost_q is None and cost_r is None:
        return quad_loss(y, u)                                # _gpc.py:125
    q = np.ones(d, M.dtype) if cost_q is None else np.asarray(cost_q, M.dtype)
    r = np.ones(u.shape[1], M.dtype) if cost_r is None else np.asarray(cost_r, M.dtype)
    return (q * y * y).sum(1) + (r * u * u).sum(1)


def gpc_grad(M, w, A, B, K, H, HH, cost_q=None, cost_r=None):
    """dM of ``policy_loss`` -- the hand-derived adjoint the kernels implement
    (SURVEY A-3.5; checked against torch.autograd and finite differences in
    tests/test_oracle_lds.py).  Replaces ``jit(grad(policy_loss, (0, 1)))``
    (``_gpc.py:128``); the second output (dW) only feeds the dead ``bias`` (``:164``)."""
    N, P, d = w.shape
    m = K.shape[1]
    mv = lambda Mat, v: np.einsum("nij,nj->ni", Mat, v)
    mtv = lambda Mat, v: np.einsum("nij,ni->nj", Mat, v)

    # forward, keeping nothing but y_f (the adjoint needs no intermediate states)
    y = np.zeros((N, d), dtype=M.dtype)
    for h in range(H - 1):
        v = np.einsum("nhij,nhj->ni", M, _window(w, h, H))
        idx = min(h + H, P - 1)
        y = (mv(A, y) + mv(B, -mv(K, y) + v)) + w[:, idx]
    wf = _window(w, HH - 1, H)
    u_f = -mv(K, y) + np.einsum("nhij,nhj->ni", M, wf)

    G = np.zeros_like(M)
    du = 2 * u_f if cost_r is None else 2 * np.asarray(cost_r, M.dtype) * u_f
    G += np.einsum("ni,nhj->nhij", du, wf)
    lam = (2 * y if cost_q is None else 2 * np.asarray(cost_q, M.dtype) * y) - mtv(K, du)
    for h in range(H - 2, -1, -1):
        lam_u = mtv(B, lam)
        G += np.einsum("ni,nhj->nhij", lam_u, _window(w, h, H))
        lam = mtv(A, lam) - mtv(K, lam_u)
    return G


def gpc_init(N, d, m, H, HH, dtype=np.float32):
    """``GPC.__init__`` state (``deluca/agents/_gpc.py:82-101``): M = 0, noise_history = 0,
    state = 0, t = 0.  (``bias`` and ``action`` are dead state and dropped, SURVEY A-1.)"""
    return dict(M=np.zeros((N, H, m, d), dtype=dtype),
                hist=np.zeros((N, H + HH, d), dtype=dtype),
                xprev=np.zeros((N, d), dtype=dtype),
                uprev=np.zeros((N, m), dtype=dtype),
                t=0)


def gpc_call(st, x, A, B, K, H, HH, lr_scale=0.005, decay=True, noise_uses_prev_action=False, cost_q=None,
             cost_r=None):
    """``GPC.__call__`` = ``get_action`` then ``update`` (``_gpc.py:130-183``).  Mutates ``st``.

    ``noise_uses_prev_action=False`` is the reference as written: the disturbance estimate
    subtracts ``B @ u_t`` with the action *just computed* (``_gpc.py:141-142,155``).  With
    the default ``lr_scale=0.005`` that recursion diverges on most default LDS draws
    (measured: 62/64 envs NaN within 2000 steps at H=3; tests/test_oracle_lds.py).  ``True``
    selects the mathematically intended ``B @ u_{t-1}`` (non-default, SURVEY A-3.2)."""
    dtype = x.dtype
    M, hist = st["M"], st["hist"]
    # get_action (_gpc.py:171-183): -K x + tensordot(M, last H noises)
    u = -np.einsum("nij,nj->ni", K, x) + np.einsum("nhij,nhj->ni", M, hist[:, HH:HH + H])
    # update (_gpc.py:145-169); note B @ u uses the action just computed (SURVEY A-3.2)
    u_n = st.get("uprev", np.zeros_like(u)) if noise_uses_prev_action else u
    noise = x - np.einsum("nij,nj->ni", A, st["xprev"]) - np.einsum("nij,nj->ni", B, u_n)
    hist = np.concatenate([hist[:, 1:], noise[:, None, :]], axis=1)   # .at[0].set; roll(-1)
    G = gpc_grad(M, hist, A, B, K, H, HH, cost_q, cost_r)
    lr = dtype.type(lr_scale)
    if decay:
        lr = dtype.type(lr_scale * (1.0 / (st["t"] + 1)))
    st["M"] = (M - lr * G).astype(dtype)
    st["hist"] = hist
    st["xprev"] = x.copy()
    st["uprev"] = u.copy()
    st["t"] += 1
    return u


# --------------------------------------------------------------------------- BPC
def unit_norm(v):
    """``generate_uniform`` normalisation (``deluca/agents/_bpc.py:26-30``): the whole
    trailing [H, m, d] tensor is scaled to Frobenius norm 1."""
    nrm = np.sqrt(np.sum(v.astype(np.float64) ** 2, axis=tuple(range(1, v.ndim)), keepdims=True))
    return (v / nrm).astype(v.dtype)


def bpc_init(N, d, m, H, M0_dir, eps0, delta=0.01, dtype=np.float32):
    """``BPC.__init__`` (``_bpc.py:69-90``).  The reference draws ``M0_dir`` (unit-norm
    [H,m,d]) and ``eps0`` (unit-norm [H,H,m,d], normalised as one tensor) from the
    *unseeded* global numpy RNG; here they are explicit inputs (SURVEY A-4)."""
    return dict(M=(dtype(delta) * M0_dir).astype(dtype),
                hist=np.zeros((N, H, d), dtype=dtype),
                xprev=np.zeros((N, d), dtype=dtype),
                eps=eps0.astype(dtype).copy(), t=0)


def bpc_call(st, x, cost, eps_new, A, B, K, H, lr_scale=0.005, decay=False, delta=0.01):
    """``BPC.__call__`` (``_bpc.py:97-158``).  ``eps_new`` [N,H,m,d] is this step's fresh
    unit-norm perturbation (``:134-136``), ``cost`` [N] the realised scalar cost."""
    dtype = x.dtype
    M, hist = st["M"], st["hist"]
    u = -np.einsum("nij,nj->ni", K, x) + np.einsum("nhij,nhj->ni", M, hist)   # :156-158
    noise = x - np.einsum("nij,nj->ni", A, st["xprev"]) - np.einsum("nij,nj->ni", B, u)
    hist = np.concatenate([hist[:, 1:], noise[:, None, :]], axis=1)           # :125-126
    lr = lr_scale
    if decay:
        lr = lr_scale * (1.0 / (st["t"] ** (3 / 4) + 1))                       # :128-129
    lr = dtype.type(lr)
    dM = cost[:, None, None, None] * st["eps"].sum(axis=1)                    # :92-93,131
    M = (M - lr * dM).astype(dtype)
    eps = np.concatenate([st["eps"][:, 1:], eps_new[:, None]], axis=1)        # :134-137
    M = (M + dtype.type(delta) * eps[:, -1]).astype(dtype)                    # :139
    st.update(M=M, hist=hist, eps=eps, xprev=x.copy(), t=st["t"] + 1)
    return u


# --------------------------------------------------------------------------- rollouts
def rollout_lds(A, B, K, x0, W, policy="gpc", H=3, HH=10, lr_scale=0.005, decay=True,
                state=None, bpc_eps=None, bpc_delta=0.01, dtype=np.float32, keep_x=True,
                noise_uses_prev_action=False, cost_q=None, cost_r=None):
    """The episode loop of SURVEY S2 (driver notebook missing from the snapshot; restated
    from the class APIs): per step  u = agent(x);  cost_t = quad_loss(x, u);
    x <- A x + B u + w_t.

    policy: "gpc" | "lqr" (u = -K x, no learning) | "bpc" (``bpc_eps`` [T,N,H,m,d] gives
    the fresh perturbations; the cost handed to BPC at step t is the cost realised at
    step t-1, 0 at the first step -- a stated choice, the reference driver is missing).
    W[T,N,d].  ``cost_q`` / ``cost_r``: GPC's ``cost_fn`` as a diagonal quadratic (see ``gpc_policy_loss``); the
    realised ``costs`` stay ``quad_loss``.  Returns dict(costs[T,N], X[T+1,N,d] if keep_x, x_final, state)."""
    A, B, K = A.astype(dtype), B.astype(dtype), K.astype(dtype)
    x = x0.astype(dtype).copy()
    W = W.astype(dtype)
    T, N, d = W.shape
    m = B.shape[2]
    if state is None:
        if policy == "gpc":
            state = gpc_init(N, d, m, H, HH, dtype)
        elif policy == "bpc":
            raise ValueError("bpc needs an explicit initial state (bpc_init)")
        else:
            state = dict(t=0)
    costs = np.zeros((T, N), dtype=dtype)
    X = np.zeros((T + 1, N, d), dtype=dtype) if keep_x else None
    if keep_x:
        X[0] = x
    prev_cost = np.zeros(N, dtype=dtype)
    for t in range(T):
        if policy == "gpc":
            u = gpc_call(state, x, A, B, K, H, HH, lr_scale, decay, noise_uses_prev_action, cost_q, cost_r)
        elif policy == "bpc":
            u = bpc_call(state, x, prev_cost, bpc_eps[t].astype(dtype), A, B, K, H,
                         lr_scale, decay, bpc_delta)
        else:
            u = lqr_action(K, x)
            state["t"] += 1
        costs[t] = quad_loss(x, u)
        prev_cost = costs[t]
        x = lds_step(A, B, x, u, W[t]).astype(dtype)
        if keep_x:
            X[t + 1] = x
    return dict(costs=costs, X=X, x_final=x, state=state)


def random_actions_stream3(seed, env_ids, t0, T, n, std=1.0):
    """Actions of dlc_lds_open_rollout_f32 when none are given: U[T,N,n] = std * z with z the normals of the
    Philox counter (env, t0 + t, c // 4, 3) -- stream 3 (disturbances are stream 0)."""
    from . import philox
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    N = env_ids.shape[0]
    ngrp = (n + 3) // 4
    ctr = np.zeros((T, N, ngrp, 4), dtype=np.uint32)
    ctr[..., 0] = env_ids[None, :, None]
    ctr[..., 1] = (np.arange(T, dtype=np.uint64) + np.uint64(t0)).astype(np.uint32)[:, None, None]
    ctr[..., 2] = np.arange(ngrp, dtype=np.uint32)[None, None, :]
    ctr[..., 3] = 3
    key = np.zeros((T, N, ngrp, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox.philox4x32(ctr, key)
    z0, z1 = philox.box_muller_f32(r[..., 0], r[..., 1])
    z2, z3 = philox.box_muller_f32(r[..., 2], r[..., 3])
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(T, N, ngrp * 4)[..., :n]
    return (np.float32(std) * z).astype(np.float32)


def open_rollout(A, B, C, x0, U, W, dtype=np.float64):
    """``LDS.__call__`` T times (``deluca/envs/_lds.py:102-111``) for N envs: x <- A x + B u_t + w_t, y_t = C x.
    A[N,d,d] B[N,d,n] C[N,p,d] (or unbatched), U[T,N,n], W[T,N,d] -> (Y[T,N,p], x_final[N,d])."""
    A, B, C = (np.asarray(v, dtype=dtype) for v in (A, B, C))
    bat = A.ndim == 3
    x = np.asarray(x0, dtype=dtype).copy()
    T = U.shape[0]
    Y = np.empty((T, x.shape[0], C.shape[-2]), dtype=dtype)
    mv = (lambda Mx, v: np.einsum("nij,nj->ni", Mx, v)) if bat else (lambda Mx, v: v @ Mx.T)
    for t in range(T):
        x = mv(A, x) + mv(B, U[t].astype(dtype)) + W[t].astype(dtype)                   # :107
        Y[t] = mv(C, x)                                                                  # :108
    return Y, x
