"""CPU restatement (NumPy, batched over envs) of deluca's output-feedback controllers on a
partially observed LDS  (SURVEY 8(f) rank 3):

    LQG  ``deluca/agents/_lqg.py:27-101``     Kalman filter + LQR gain
    DRC  ``deluca/agents/_drc.py:43-203``     disturbance-response controller on the Markov operator G
    GRC  ``deluca/agents/_grc.py:42-169``     gradient response controller (windows of y_nat)
    SFC  ``deluca/agents/_sfc.py:42-214``     spectral filtering controller
    DSC  ``deluca/agents/_dsc.py:42-292``     double spectral controller

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and ``__graft_entry__.smoke()``); the product
never imports it.  PARITY PINNED to the reference's own files run here: ``tests/golden/ref_of.npz`` holds closed
loops of the reference's ``GRC`` / ``SFC`` / ``DSC`` / ``DRC`` / ``LQG`` classes executed literally on tests/refshim
(tests/golden/make_ref_fixtures.py, family ``of``; the reference ships no test of its own for them), and
this module reproduces them to rounding in float64 (tests/test_ref_fixtures_oracle.py).  In addition the hand
adjoints are checked against ``torch.autograd`` on literal restatements of the reference ``policy_loss`` einsums
(``tests/test_oracle_of.py``).

GRC, SFC and DSC share one structure: the action is linear in the parameters and in a fixed bank of
linear filters of the y_nat history,
    a(k) = sum_f Theta[f] phi_f(k),     phi_f(k) = sum_o Phi[f, o] ynat[k + o]        (o < L)
with (F, L, Phi) = (m, m, I) for GRC, (h + 1, m + 1, [e_m ; filters]) for SFC and
(1 + 2h + h^2, 2m + 1, ...) for DSC (``filter_bank``).  ``fgrc_*`` implements that family once;
``grc/sfc/dsc_policy_loss_literal`` restate the reference einsums literally for the tests.

Random initial parameters (``jax.random.normal`` under ``PRNGKey(0)``, e.g. ``_grc.py:77``) cannot be
reproduced without JAX; every function takes the initial parameters as an argument.
"""
import numpy as np


def quad_loss(y, u):
    """``quad_loss`` (``_grc.py:27-39``): sum(y'y + u'u) per env."""
    return np.sum(y * y, axis=-1) + np.sum(u * u, axis=-1)


def _mv(Mat, v):
    return np.einsum("...ij,...j->...i", Mat, v)


def _mtv(Mat, v):
    return np.einsum("...ij,...i->...j", Mat, v)


# --------------------------------------------------------------------------- spectral filters
def spectral_filters(m, h, gamma=0.2):
    """``_sfc.py:98-112`` / ``_dsc.py:111-125``: top-h eigenpairs of the Hankel matrix
    H[i,j] = (1-gamma)^(i+j-1)/(i+j-1), scaled by sigma^(1/4), flipped along the window axis.
    Returns filters[h, m] (float64)."""
    idx = np.arange(1, m + 1)
    I, J = np.meshgrid(idx, idx, indexing="ij")
    Hm = ((1 - gamma) ** (I + J - 1)) / (I + J - 1)
    eigvals, eigvecs = np.linalg.eigh(Hm)
    top_vals = eigvals[-h:]
    top_vecs = eigvecs[:, -h:]
    scaling = top_vals ** 0.25
    return np.flip(scaling[:, None] * top_vecs.T, axis=1).copy()


def filter_bank(kind, m, h=None, gamma=0.2):
    """(Phi[F, L], names) such that the reference action at window offset k is
    sum_f Theta[f] @ (sum_o Phi[f, o] ynat[k + o]); history length is L + m."""
    if kind == "grc":                                   # _grc.py:92-95: window = ynats[k : k+m]
        return np.eye(m), m
    filt = spectral_filters(m, h, gamma)
    if kind == "sfc":                                   # _sfc.py:131-140
        L = m + 1
        Phi = np.zeros((h + 1, L))
        Phi[0, m] = 1.0                                 # M_0 @ ynats[k + m]
        Phi[1:, :m] = filt                              # M[j] @ sum_i filters[j, i] ynats[k + i]
        return Phi, L
    if kind == "dsc":                                   # _dsc.py:151-196
        L = 2 * m + 1
        Phi = np.zeros((1 + 2 * h + h * h, L))
        Phi[0, 2 * m] = 1.0                             # M_0 @ ynats[k + 2m]
        Phi[1:1 + h, m:2 * m] = filt                    # M_1: window at k + m
        Phi[1 + h:1 + 2 * h, m + 1:2 * m + 1] = filt    # M_2: window at k + m + 1
        for a in range(h):                              # M_3[a, b]: filters[a, i] filters[b, q] ynats[k+1+i+q]
            for b in range(h):
                Phi[1 + 2 * h + a * h + b, 1:2 * m] = np.convolve(filt[a], filt[b])
        return Phi, L
    raise ValueError(kind)


# --------------------------------------------------------------------------- FGRC family
def fgrc_actions(Theta, Phi, ynats, m):
    """a(k) for k = 0..m.  Theta[N,F,n,p], Phi[F,L], ynats[N,L+m,p] -> [N, m+1, n]."""
    L = Phi.shape[1]
    win = np.stack([ynats[:, k:k + L] for k in range(m + 1)], axis=1)          # [N, m+1, L, p]
    feat = np.einsum("fo,nkop->nkfp", Phi, win)
    return np.einsum("nfap,nkfp->nka", Theta, feat)


def fgrc_policy_loss(Theta, Phi, ynats, A, B, C, m):
    """``policy_loss`` of ``_grc.py:90-105`` / ``_sfc.py:130-152`` / ``_dsc.py:150-208`` in the common form."""
    act = fgrc_actions(Theta, Phi, ynats, m)
    delta = np.zeros(ynats.shape[:1] + (A.shape[-1],), dtype=Theta.dtype)
    for k in range(m):                                                          # evolve: scan over arange(m)
        delta = _mv(A, delta) + _mv(B, act[:, k])
    final_y = _mv(C, delta) + ynats[:, -1]
    return quad_loss(final_y, act[:, m])


def fgrc_grad(Theta, Phi, ynats, A, B, C, m):
    """dTheta of ``fgrc_policy_loss`` (hand adjoint; replaces ``jit(grad(policy_loss))``)."""
    L = Phi.shape[1]
    act = fgrc_actions(Theta, Phi, ynats, m)
    delta = np.zeros(ynats.shape[:1] + (A.shape[-1],), dtype=Theta.dtype)
    for k in range(m):
        delta = _mv(A, delta) + _mv(B, act[:, k])
    final_y = _mv(C, delta) + ynats[:, -1]
    g = np.zeros_like(act)                                                      # adjoint of a(k)
    g[:, m] = 2 * act[:, m]
    lam = _mtv(C, 2 * final_y)
    for k in range(m - 1, -1, -1):
        g[:, k] = _mtv(B, lam)
        lam = _mtv(A, lam)
    win = np.stack([ynats[:, k:k + L] for k in range(m + 1)], axis=1)
    feat = np.einsum("fo,nkop->nkfp", Phi, win)
    return np.einsum("nka,nkfp->nfap", g, feat)


def fgrc_init(Theta0, d, L, m):
    """Agent state at construction (``_grc.py:76-80``): given parameters, z = 0, ynat_history = 0, t = 0."""
    N, p = Theta0.shape[0], Theta0.shape[3]
    return dict(Theta=Theta0.copy(), z=np.zeros((N, d), Theta0.dtype), ynats=np.zeros((N, L + m, p), Theta0.dtype), t=0)


def fgrc_call(st, y, Phi, A, B, C, m, lr=0.005, decay=True, R_M=10.0):
    """``GRC/SFC/DSC.__call__`` = get_action then update (``_grc.py:110-169``).  Mutates ``st``."""
    dtype = y.dtype
    L = Phi.shape[1]
    Phi = Phi.astype(dtype)
    # get_action: windows of the history *before* this observation enters it (_grc.py:157-169)
    win = st["ynats"][:, m:m + L]
    u = np.einsum("nfap,nfp->na", st["Theta"], np.einsum("fo,nop->nfp", Phi, win))
    # update (_grc.py:124-155)
    st["z"] = _mv(A, st["z"]) + _mv(B, u)
    y_nat = y - _mv(C, st["z"])
    st["ynats"] = np.concatenate([st["ynats"][:, 1:], y_nat[:, None]], axis=1)   # .at[0].set; roll(-1)
    G = fgrc_grad(st["Theta"], Phi, st["ynats"], A, B, C, m)
    lr_t = dtype.type(lr * (1 / (st["t"] + 1) if decay else 1))
    Theta = st["Theta"] - lr_t * G
    norm = np.sqrt(np.sum(Theta * Theta, axis=(1, 2, 3)))
    scale = np.minimum(dtype.type(1.0), dtype.type(R_M) / (norm + dtype.type(1e-8)))
    st["Theta"] = (scale[:, None, None, None] * Theta).astype(dtype)
    st["t"] += 1
    return u


# --------------------------------------------------------------------------- literal reference einsums (tests)
def grc_policy_loss_literal(M, ynats, A, B, C, m):
    """``_grc.py:90-105`` for one env, einsum by einsum.  M[m,n,p], ynats[2m,p,1]."""
    def action(k):
        window = ynats[k:k + m]
        contribs = np.einsum("mnp,mpi->mni", M, window)
        return np.sum(contribs, axis=0)
    delta = np.zeros((A.shape[0], 1), dtype=M.dtype)
    for k in range(m):
        delta = A @ delta + B @ action(k)
    final_y = C @ delta + ynats[-1]
    return np.sum(final_y.T @ final_y + action(m).T @ action(m))


def sfc_policy_loss_literal(M_0, M, filters, ynats, A, B, C, m):
    """``_sfc.py:130-152`` for one env.  M_0[n,p], M[h,n,p], filters[h,m], ynats[2m+1,p,1]."""
    def action(k):
        window = ynats[k:k + m]
        last_ynat = ynats[k + m]
        proj = np.einsum("hm,mpi->hpi", filters, window)
        spectral_term = np.einsum("hnp,hpi->ni", M, proj)
        return M_0 @ last_ynat + spectral_term
    delta = np.zeros((A.shape[0], 1), dtype=M.dtype)
    for k in range(m):
        delta = A @ delta + B @ action(k)
    final_y = C @ delta + ynats[-1]
    return np.sum(final_y.T @ final_y + action(m).T @ action(m))


def dsc_policy_loss_literal(M_0, M_1, M_2, M_3, filters, ynats, A, B, C, m):
    """``_dsc.py:150-208`` for one env.  M_3[h,h,n,p], ynats[3m+1,p,1]."""
    def action(k):
        last_ynat = ynats[k + 2 * m]
        prev_m = ynats[k + m:k + 2 * m]
        t1 = np.einsum("hnp,hpi->ni", M_1, np.einsum("hm,mpi->hpi", filters, prev_m))
        last_m = ynats[k + m + 1:k + 2 * m + 1]
        t2 = np.einsum("hnp,hpi->ni", M_2, np.einsum("hm,mpi->hpi", filters, last_m))
        mat = np.stack([ynats[k + 1 + i:k + 1 + i + m] for i in range(m)])      # [i, m, p, 1]
        temp = np.einsum("hi,impq->hmpq", filters, mat)
        filtered = np.einsum("jm,hmpq->hjpq", filters, temp)
        contrib = np.einsum("hjnp,hjpq->hjnq", M_3, filtered)
        return M_0 @ last_ynat + t1 + t2 + np.sum(contrib, axis=(0, 1))
    delta = np.zeros((A.shape[0], 1), dtype=M_0.dtype)
    for k in range(m):
        delta = A @ delta + B @ action(k)
    final_y = C @ delta + ynats[-1]
    return np.sum(final_y.T @ final_y + action(m).T @ action(m))


def dsc_pack_theta(M_0, M_1, M_2, M_3):
    """Stack the four DSC parameter blocks in ``filter_bank('dsc')`` order -> Theta[F,n,p]."""
    h = M_1.shape[0]
    return np.concatenate([M_0[None], M_1, M_2, M_3.reshape(h * h, *M_3.shape[2:])], axis=0)


# --------------------------------------------------------------------------- DRC
def drc_markov(A, B, C, h):
    """``_drc.py:97-102``: G[i] = C A^i B, i < h.  Batched or not."""
    G = []
    Ap = np.broadcast_to(np.eye(A.shape[-1], dtype=A.dtype), A.shape)
    for _ in range(h):
        G.append(C @ Ap @ B)
        Ap = Ap @ A
    return np.stack(G, axis=-3)


def drc_init(M0, h, n):
    """``_drc.py:104-114``: given M (the reference draws lr_scale * normal), y_nat = 0, us = 0, t = 0."""
    N, m, _, p = M0.shape
    return dict(M=M0.copy(), y_nat=np.zeros((N, m, p), M0.dtype), us=np.zeros((N, h, n), M0.dtype), t=0)


def drc_call(st, obs, G, K, lr_scale=0.03, decay=True, RM=1000.0):
    """``DRC.__call__`` (``_drc.py:131-203``): update_noise, get_action, update_params.  Mutates ``st``.
    G[N,h,p,n], K[N,n,p].  Note ``_drc.py:184``: without decay the learning rate is 1.0, not lr_scale."""
    dtype = obs.dtype
    # update_noise (_drc.py:168-171): y_nat[0] is the most recent
    y_nat = obs - np.einsum("nhpa,nha->np", G, st["us"])
    st["y_nat"] = np.concatenate([y_nat[:, None], st["y_nat"][:, :-1]], axis=1)
    # get_action (_drc.py:152-162)
    u = -_mv(K, obs) + np.einsum("nmap,nmp->na", st["M"], st["y_nat"])
    # update_params (_drc.py:173-203); grad(policy_loss) is with respect to M only (_drc.py:129)
    final_state = st["y_nat"][:, 0] + np.einsum("nhpa,nha->np", G, st["us"])
    act = -_mv(K, final_state) + np.einsum("nmap,nmp->na", st["M"], st["y_nat"])
    delta_M = np.einsum("na,nmp->nmap", 2 * act, st["y_nat"])
    lr = dtype.type(lr_scale * (1 / (st["t"] + 1))) if decay else dtype.type(1.0)
    M = st["M"] - lr * delta_M
    norm = np.sqrt(np.sum(M * M, axis=(1, 2, 3)))
    big = norm > RM
    M[big] = M[big] * (dtype.type(RM) / norm[big])[:, None, None, None]
    st["M"] = M.astype(dtype)
    st["us"] = np.concatenate([u[:, None], st["us"][:, :-1]], axis=1)
    st["t"] += 1
    return u


def drc_policy_loss_literal(M, G, y_nat, us, K):
    """``_drc.py:116-126`` for one env.  M[m,n,p], G[h,p,n], y_nat[m,p,1], us[h,n,1]."""
    def action(obs):
        return -K @ obs + np.tensordot(M, y_nat, axes=([0, 2], [0, 1]))
    final_state = y_nat[0] + np.tensordot(G, us, axes=([0, 2], [0, 1]))
    a = action(final_state)
    return np.sum(final_state.T @ final_state + a.T @ a)


# --------------------------------------------------------------------------- LQG
def lqg_gains(A, B, C, Q=None, R=None, Wc=None, Vc=None):
    """``_lqg.py:61-81`` for one system (scipy DARE, fp64): returns (K[n,d], L[d,p])."""
    from scipy.linalg import solve_discrete_are
    d, n, p = A.shape[0], B.shape[1], C.shape[0]
    Q = np.eye(d) if Q is None else Q
    R = np.eye(n) if R is None else R
    Wc = np.eye(d) if Wc is None else Wc
    Vc = np.eye(p) if Vc is None else Vc
    P = solve_discrete_are(A, B, Q, R)
    K = np.linalg.inv(B.T @ P @ B + R) @ (B.T @ P @ A)
    Sigma = solve_discrete_are(A.T, C.T, Wc, Vc)
    Lk = Sigma @ C.T @ np.linalg.inv(C @ Sigma @ C.T + Vc)
    return K, Lk


def lqg_call(st, y, A, B, C, K, Lk):
    """``LQG.__call__`` (``_lqg.py:83-98``)."""
    u = -_mv(K, st["x_hat"])
    residual = y - _mv(C, st["x_hat"])
    st["x_hat"] = _mv(A, st["x_hat"]) + _mv(B, u) + _mv(Lk, residual)
    return u


# --------------------------------------------------------------------------- episode driver
def rollout_of(A, B, C, x0, W, agent, dtype=np.float32, **kw):
    """Episode loop on the partially observed LDS (``_lds.py:102-111``: x <- A x + B u + w, y = C x):
    per step  y = C x;  u = agent(y);  cost_t = quad_loss(y, u);  x <- A x + B u + w_t.
    A[N,d,d], B[N,d,n], C[N,p,d], x0[N,d], W[T,N,d].  agent: "lqg" (K, Lk), "drc" (M0, K, h, lr_scale,
    decay, RM), "fgrc" (Theta0, Phi, m, lr, decay, R_M).  Returns costs[T,N], U[T,N,n], x_final, state."""
    A, B, C = A.astype(dtype), B.astype(dtype), C.astype(dtype)
    x = x0.astype(dtype).copy()
    W = W.astype(dtype)
    T, N, d = W.shape
    n = B.shape[-1]
    if agent == "lqg":
        K, Lk = kw["K"].astype(dtype), kw["Lk"].astype(dtype)
        st = kw.get("state") or dict(x_hat=np.zeros((N, d), dtype))
    elif agent == "drc":
        K = kw["K"].astype(dtype)
        G = drc_markov(A, B, C, kw["h"]).astype(dtype)
        st = kw.get("state") or drc_init(kw["M0"].astype(dtype), kw["h"], n)
    elif agent == "fgrc":
        Phi = kw["Phi"]
        st = kw.get("state") or fgrc_init(kw["Theta0"].astype(dtype), d, Phi.shape[1], kw["m"])
    else:
        raise ValueError(agent)
    costs = np.zeros((T, N), dtype)
    U = np.zeros((T, N, n), dtype)
    for t in range(T):
        y = _mv(C, x)
        if agent == "lqg":
            u = lqg_call(st, y, A, B, C, K, Lk)
        elif agent == "drc":
            u = drc_call(st, y, G, K, kw.get("lr_scale", 0.03), kw.get("decay", True), kw.get("RM", 1000.0))
        else:
            u = fgrc_call(st, y, Phi, A, B, C, kw["m"], kw.get("lr", 0.005), kw.get("decay", True),
                          kw.get("R_M", 10.0))
        costs[t] = quad_loss(y, u)
        U[t] = u
        x = (_mv(A, x) + _mv(B, u) + W[t]).astype(dtype)
    return dict(costs=costs, U=U, x_final=x, state=st)
