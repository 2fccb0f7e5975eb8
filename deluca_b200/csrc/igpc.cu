// deluca/igpc on the device: trajectory rollouts of the classic envs, closed-form
// linearisation, the Riccati backward pass, GPC_rollout with a hand adjoint of counterfact_loss, and
// the iLQR / iGPC_closed / iLC_closed outer loops with their backtracking, all batched over N
// independent trajectories, one thread per trajectory.  Per-trajectory data (X, U, k, K: 6.4 KB at
// d=3, H=200) stays in the caller's env-major arrays, which are L2-resident for the configured sizes;
// the time axis is a serial dependency chain, so these kernels are latency- not bandwidth-bound.
#include <stdlib.h>

#include "common.cuh"
#include "env_models.cuh"

namespace dlc {

template <int D, int M>
__device__ __forceinline__ float quad_cost(const DlcPlanParams& p, const float* x, const float* u) {
  float cx = 0.f, cu = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float e = x[i] - p.goal[i];
    cx = fmaf(p.q[i] * e, e, cx);
  }
#pragma unroll
  for (int c = 0; c < M; ++c) {
    const float e = u[c] - p.goal_u[c];
    cu = fmaf(p.r[c] * e, e, cu);
  }
  return cx + cu;
}

// u = U_old[h] + alpha k[h] + K[h] (x - X_old[h])                        rollout.py:55-57
template <int D, int M>
__device__ __forceinline__ void feedback_u(const float* Uo, const float* kk, const float* KK, const float* Xo,
                                           float alpha, const float* x, float* u) {
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) s = fmaf(KK[c * D + j], x[j] - Xo[j], s);
    u[c] = (Uo[c] + alpha * kk[c]) + s;
  }
}

// igpc.rollout for one trajectory.  Returns the cost; X/U written to Xw/Uw.
template <class Mdl>
__device__ float rollout_one(const Mdl& env, const DlcPlanParams& p, const float* Uo, const float* kk,
                             const float* KK, const float* Xo, float alpha, const float* x0, float* Xw, float* Uw) {
  constexpr int D = Mdl::D, M = Mdl::M;
  float x[D], u[M], xn[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { x[i] = x0[i]; Xw[i] = x[i]; }
  float cost = 0.f;
  for (int h = 0; h < p.H; ++h) {
    if (kk == nullptr) {
#pragma unroll
      for (int c = 0; c < M; ++c) u[c] = Uo[h * M + c];
    } else {
      feedback_u<D, M>(Uo + h * M, kk + h * M, KK + h * M * D, Xo + h * D, alpha, x, u);
    }
    env.step(x, u, xn);
    cost += quad_cost<D, M>(p, x, u);
#pragma unroll
    for (int c = 0; c < M; ++c) Uw[h * M + c] = u[c];
#pragma unroll
    for (int i = 0; i < D; ++i) { x[i] = xn[i]; Xw[(h + 1) * D + i] = xn[i]; }
  }
  return cost;
}

// The per-step operands of the feedback law, held in registers.  The planning loops are one long
// dependency chain per trajectory, so a global load issued where its value is needed stalls the
// chain for an L1/L2 round trip every step (ncu: long_scoreboard was 45 % of all stall cycles);
// loading step h+1's operands while step h computes takes the memory latency off the chain.
template <int D, int M>
struct StepOps {
  float Uo[M], kk[M], KK[M * D], Xo[D];
  // h < H: everything; h == H: only X_old[H] (read by the dede / delin disturbance models)
  __device__ __forceinline__ void load(const float* U_, const float* k_, const float* K_, const float* X_, int h, int H) {
    if (h < H) {
#pragma unroll
      for (int c = 0; c < M; ++c) Uo[c] = U_[h * M + c];
      if (k_) {
#pragma unroll
        for (int c = 0; c < M; ++c) kk[c] = k_[h * M + c];
#pragma unroll
        for (int i = 0; i < M * D; ++i) KK[i] = K_[h * M * D + i];
      }
    }
    if (X_ && h <= H) {
#pragma unroll
      for (int i = 0; i < D; ++i) Xo[i] = X_[h * D + i];
    }
  }
  __device__ __forceinline__ void load_x(const float* X_, int h) {
#pragma unroll
    for (int i = 0; i < D; ++i) Xo[i] = X_[h * D + i];
  }
  __device__ __forceinline__ void feedback(float alpha, const float* x, float* u) const {
    feedback_u<D, M>(Uo, kk, KK, Xo, alpha, x, u);
  }
};

// rollout_one with the operands of step h+1 loaded while step h computes
template <class Mdl, bool PF>
__device__ float rollout_one_pf(const Mdl& env, const DlcPlanParams& p, const float* Uo, const float* kk,
                             const float* KK, const float* Xo, float alpha, const float* x0, float* Xw, float* Uw) {
  constexpr int D = Mdl::D, M = Mdl::M;
  float x[D], u[M], xn[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { x[i] = x0[i]; Xw[i] = x[i]; }
  float cost = 0.f;
  StepOps<D, M> cur, nxt;
  if (PF) cur.load(Uo, kk, KK, kk ? Xo : nullptr, 0, p.H);
  for (int h = 0; h < p.H; ++h) {
    if (PF) {
      if (h + 1 < p.H) nxt.load(Uo, kk, KK, kk ? Xo : nullptr, h + 1, p.H);
    } else {
      cur.load(Uo, kk, KK, kk ? Xo : nullptr, h, p.H);
    }
    if (kk == nullptr) {
#pragma unroll
      for (int c = 0; c < M; ++c) u[c] = cur.Uo[c];
    } else {
      cur.feedback(alpha, x, u);
    }
    env.step(x, u, xn);
    cost += quad_cost<D, M>(p, x, u);
#pragma unroll
    for (int c = 0; c < M; ++c) Uw[h * M + c] = u[c];
#pragma unroll
    for (int i = 0; i < D; ++i) { x[i] = xn[i]; Xw[(h + 1) * D + i] = xn[i]; }
    if (PF) cur = nxt;
  }
  return cost;
}

// small dense inverse (Gauss-Jordan, partial pivoting); returns false if singular / non-finite
template <int M>
__device__ bool inv_small(const float* A, float* Ai) {
  if (M == 1) {
    Ai[0] = 1.0f / A[0];
    return A[0] != 0.f && isfinite(Ai[0]);
  }
  float a[M][2 * M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) { a[i][j] = A[i * M + j]; a[i][M + j] = i == j ? 1.f : 0.f; }
  for (int c = 0; c < M; ++c) {
    int piv = c;
    for (int r = c + 1; r < M; ++r) if (fabsf(a[r][c]) > fabsf(a[piv][c])) piv = r;
    if (a[piv][c] == 0.f) return false;
    for (int j = 0; j < 2 * M; ++j) { const float t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    const float inv = 1.0f / a[c][c];
    for (int j = 0; j < 2 * M; ++j) a[c][j] *= inv;
    for (int r = 0; r < M; ++r) {
      if (r == c) continue;
      const float f = a[r][c];
      for (int j = 0; j < 2 * M; ++j) a[r][j] -= f * a[c][j];
    }
  }
  bool ok = true;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) { Ai[i * M + j] = a[i][M + j]; ok &= isfinite(a[i][M + j]); }
  return ok;
}

// One Riccati step (lqr_solver.py:68-85); updates V_x, V_xx, writes k[h], K[h].  Returns Q_uu status.
template <int D, int M>
__device__ int riccati_step(const float* fx, const float* fu, const float* cx, const float* cu, const float* cxx,
                            const float* cuu, float* Vx, float* Vxx, float* kh, float* Kh) {
  float Qx[D], Qu[M], Qxx[D * D], Qux[M * D], Quu[M * M], VF[D * D], VFu[D * M];
  // VF = V_xx f_x, VFu = V_xx f_u
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < D; ++l) s = fmaf(Vxx[i * D + l], fx[l * D + j], s);
      VF[i * D + j] = s;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < D; ++l) s = fmaf(Vxx[i * D + l], fu[l * M + c], s);
      VFu[i * M + c] = s;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < D; ++l) s = fmaf(fx[l * D + i], Vx[l], s);
    Qx[i] = cx[i] + s;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < D; ++l) t = fmaf(fx[l * D + i], VF[l * D + j], t);
      Qxx[i * D + j] = cxx[i * D + j] + t;
    }
  }
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < D; ++l) s = fmaf(fu[l * M + c], Vx[l], s);
    Qu[c] = cu[c] + s;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < D; ++l) t = fmaf(fu[l * M + c], VF[l * D + j], t);
      Qux[c * D + j] = t;
    }
#pragma unroll
    for (int e = 0; e < M; ++e) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < D; ++l) t = fmaf(fu[l * M + c], VFu[l * M + e], t);
      Quu[c * M + e] = cuu[c * M + e] + t;
    }
  }
  float Qi[M * M];
  int status = 0;
  if (!inv_small<M>(Quu, Qi)) status |= DLC_STATUS_NOT_PD;
  if (M == 1 && !(Quu[0] > 0.f)) status |= DLC_STATUS_NOT_PD;
  float Kl[M * D], kl[M];
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < M; ++e) s = fmaf(Qi[c * M + e], Qu[e], s);
    kl[c] = -s;
    kh[c] = -s;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < M; ++e) t = fmaf(Qi[c * M + e], Qux[e * D + j], t);
      Kl[c * D + j] = -t;
      Kh[c * D + j] = -t;
    }
  }
  // V_x = Q_x - K' Q_uu k ; V_xx = Q_xx - K' Q_uu K ; symmetrise
  float Qk[M], QK[M * D];
#pragma unroll
  for (int c = 0; c < M; ++c) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < M; ++e) s = fmaf(Quu[c * M + e], kl[e], s);
    Qk[c] = s;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < M; ++e) t = fmaf(Quu[c * M + e], Kl[e * D + j], t);
      QK[c * D + j] = t;
    }
  }
  float Vn[D * D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < M; ++c) s = fmaf(Kl[c * D + i], Qk[c], s);
    Vx[i] = Qx[i] - s;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) t = fmaf(Kl[c * D + i], QK[c * D + j], t);
      Vn[i * D + j] = Qxx[i * D + j] - t;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) Vxx[i * D + j] = (Vn[i * D + j] + Vn[j * D + i]) * 0.5f;
  return status;
}

template <int D, int M>
__device__ __forceinline__ void cost_ders(const DlcPlanParams& p, const float* x, const float* u, float* cx, float* cu,
                                          float* cxx, float* cuu) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    cx[i] = (2.0f * p.q[i]) * (x[i] - p.goal[i]);
#pragma unroll
    for (int j = 0; j < D; ++j) cxx[i * D + j] = i == j ? 2.0f * p.q[i] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < M; ++c) {
    cu[c] = (2.0f * p.r[c]) * (u[c] - p.goal_u[c]);
#pragma unroll
    for (int e = 0; e < M; ++e) cuu[c * M + e] = c == e ? 2.0f * p.r[c] : 0.f;
  }
}

// compute_ders + LQR fused: linearise env_sim along (X, U) on the fly, going backwards.
template <class Mdl>
__device__ int backward_pass(const Mdl& sim, const DlcPlanParams& p, const float* X, const float* U, float* kk,
                             float* KK) {
  constexpr int D = Mdl::D, M = Mdl::M;
  float Vx[D], Vxx[D * D];
#pragma unroll
  for (int i = 0; i < D; ++i) Vx[i] = 0.f;
#pragma unroll
  for (int i = 0; i < D * D; ++i) Vxx[i] = 0.f;
  int status = 0;
  for (int h = p.H - 1; h >= 0; --h) {
    float x[D], u[M], fx[D * D], fu[D * M], cx[D], cu[M], cxx[D * D], cuu[M * M];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = X[h * D + i];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = U[h * M + c];
    sim.jac(x, u, fx, fu);
    cost_ders<D, M>(p, x, u, cx, cu, cxx, cuu);
    status |= riccati_step<D, M>(fx, fu, cx, cu, cxx, cuu, Vx, Vxx, kk + h * M, KK + h * M * D);
  }
  return status;
}

// backward_pass with the next step's (x, u) loaded one step early
template <class Mdl, bool PF>
__device__ int backward_pass_pf(const Mdl& sim, const DlcPlanParams& p, const float* X, const float* U, float* kk,
                             float* KK) {
  constexpr int D = Mdl::D, M = Mdl::M;
  float Vx[D], Vxx[D * D];
#pragma unroll
  for (int i = 0; i < D; ++i) Vx[i] = 0.f;
#pragma unroll
  for (int i = 0; i < D * D; ++i) Vxx[i] = 0.f;
  int status = 0;
  float xq[D], uq[M];  // operands of the next step down, loaded one step early (see StepOps)
#pragma unroll
  for (int i = 0; i < D; ++i) xq[i] = X[(p.H - 1) * D + i];
#pragma unroll
  for (int c = 0; c < M; ++c) uq[c] = U[(p.H - 1) * M + c];
  for (int h = p.H - 1; h >= 0; --h) {
    float x[D], u[M], fx[D * D], fu[D * M], cx[D], cu[M], cxx[D * D], cuu[M * M];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xq[i];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = uq[c];
    if (PF && h > 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) xq[i] = X[(h - 1) * D + i];
#pragma unroll
      for (int c = 0; c < M; ++c) uq[c] = U[(h - 1) * M + c];
    }
    if (!PF) {
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = X[h * D + i];
#pragma unroll
      for (int c = 0; c < M; ++c) u[c] = U[h * M + c];
    }
    sim.jac(x, u, fx, fu);
    cost_ders<D, M>(p, x, u, cx, cu, cxx, cuu);
    status |= riccati_step<D, M>(fx, fu, cx, cu, cxx, cuu, Vx, Vxx, kk + h * M, KK + h * M * D);
  }
  return status;
}

// GPC_rollout (igpc.py:63-126) for one trajectory; counterfact_loss gradient by hand adjoint
// (only the last step's cost survives in the reference, igpc.py:41).
template <class Mdl>
__device__ float igpc_rollout_one(const Mdl& tru, const Mdl& sim, const DlcPlanParams& p, const float* Uo,
                                  const float* kk, const float* KK, const float* Xo, float alpha, const float* x0,
                                  float* Xw, float* Uw) {
  constexpr int D = Mdl::D, M = Mdl::M, HG = 3, MG = 3;  // H, M of GPC_rollout (:67)
  const float Cc = 1.0f;
  float W[HG + MG][D], E[HG][M][D], off[M];
#pragma unroll
  for (int i = 0; i < HG + MG; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) W[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < HG; ++i)
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) E[i][c][j] = 0.f;
#pragma unroll
  for (int c = 0; c < M; ++c) off[c] = 0.f;
  float x[D], u[M], xn[D], xs[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { x[i] = x0[i]; Xw[i] = x[i]; }
  float cost = 0.f;
  for (int h = 0; h < p.H; ++h) {
    // U_delta = tensordot(E, W[-M:]) + off                                :80
    feedback_u<D, M>(Uo + h * M, kk + h * M, KK + h * M * D, Xo + h * D, alpha, x, u);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < HG; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(E[i][c][j], W[HG + i][j], s);
      u[c] += Cc * (s + off[c]);
    }
    tru.step(x, u, xn);                                                    // :87
    cost += quad_cost<D, M>(p, x, u);                                      // :88
    sim.step(x, u, xs);                                                    // :91
    float w[D];
    if (p.w_is == DLC_W_DE) {
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = xn[i] - xs[i];
    } else if (p.w_is == DLC_W_DEDE) {   // D[h] = X_old[h+1] - sim(X_old[h], U_old[h])  (compute_ders)
      float xo[D], uo[M], so[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xo[i] = Xo[h * D + i];
#pragma unroll
      for (int c = 0; c < M; ++c) uo[c] = Uo[h * M + c];
      sim.step(xo, uo, so);
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = (xn[i] - xs[i]) - (Xo[(h + 1) * D + i] - so[i]);
    } else {                             // F[h] = jac of sim at (X_old[h], U_old[h])
      float xo[D], uo[M], fx[D * D], fu[D * M];
#pragma unroll
      for (int i = 0; i < D; ++i) xo[i] = Xo[h * D + i];
#pragma unroll
      for (int c = 0; c < M; ++c) uo[c] = Uo[h * M + c];
      sim.jac(xo, uo, fx, fu);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float s = 0.f, t = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(fx[i * D + j], x[j] - xo[j], s);
#pragma unroll
        for (int c = 0; c < M; ++c) t = fmaf(fu[i * M + c], u[c] - uo[c], t);
        w[i] = xn[i] - ((Xo[(h + 1) * D + i] + s) + t);
      }
    }
#pragma unroll
    for (int i = 0; i < HG + MG - 1; ++i)                                  // W[0] = w; roll(-1)   :103-104
#pragma unroll
      for (int j = 0; j < D; ++j) W[i][j] = W[i + 1][j];
#pragma unroll
    for (int j = 0; j < D; ++j) W[HG + MG - 1][j] = w[j];
#pragma unroll
    for (int c = 0; c < M; ++c) Uw[h * M + c] = u[c];
#pragma unroll
    for (int i = 0; i < D; ++i) Xw[(h + 1) * D + i] = xn[i];

    if (h >= HG) {                                                         // :105-125
      // forward: y_0 = X[h-H]; y_{j+1} = sim(y_j, u_j) + W[j+M]
      const int b = h - HG;
      float y[HG][D], uu[HG][M], Fx[HG - 1][D * D], Fu[HG - 1][D * M];
#pragma unroll
      for (int i = 0; i < D; ++i) y[0][i] = Xw[b * D + i];
#pragma unroll
      for (int j = 0; j < HG; ++j) {
        feedback_u<D, M>(Uo + (b + j) * M, kk + (b + j) * M, KK + (b + j) * M * D, Xo + (b + j) * D, alpha, y[j],
                         uu[j]);
#pragma unroll
        for (int c = 0; c < M; ++c) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < HG; ++i)
#pragma unroll
            for (int l = 0; l < D; ++l) s = fmaf(E[i][c][l], W[j + i][l], s);
          uu[j][c] += Cc * (s + off[c]);
        }
        if (j < HG - 1) {
          float yn[D];
          sim.jac(y[j], uu[j], Fx[j], Fu[j]);
          sim.step(y[j], uu[j], yn);
#pragma unroll
          for (int i = 0; i < D; ++i) y[j + 1][i] = yn[i] + W[j + MG][i];
        }
      }
      // adjoint of c(y_{H-1}, u_{H-1})
      float ly[D], lu[M], cxx[D * D], cuu[M * M];
      cost_ders<D, M>(p, y[HG - 1], uu[HG - 1], ly, lu, cxx, cuu);
      const float sc = p.lr / (float)h;
#pragma unroll
      for (int j = HG - 1; j >= 0; --j) {
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const float g = Cc * lu[c];
          off[c] -= sc * g;
#pragma unroll
          for (int i = 0; i < HG; ++i)
#pragma unroll
            for (int l = 0; l < D; ++l) E[i][c][l] -= sc * (g * W[j + i][l]);
        }
        // note: E/off of this very update must not feed the remaining adjoint steps -- they do not:
        // the adjoint only reads W, K, Fx, Fu.
        const float* Kj = KK + (b + j) * M * D;
#pragma unroll
        for (int l = 0; l < D; ++l) {
          float s = ly[l];
#pragma unroll
          for (int c = 0; c < M; ++c) s = fmaf(Kj[c * D + l], lu[c], s);
          ly[l] = s;
        }
        if (j > 0) {
          float nu[M], ny[D];
#pragma unroll
          for (int c = 0; c < M; ++c) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(Fu[j - 1][i * M + c], ly[i], s);
            nu[c] = s;
          }
#pragma unroll
          for (int l = 0; l < D; ++l) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(Fx[j - 1][i * D + l], ly[i], s);
            ny[l] = s;
          }
#pragma unroll
          for (int c = 0; c < M; ++c) lu[c] = nu[c];
#pragma unroll
          for (int l = 0; l < D; ++l) ly[l] = ny[l];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
  }
  return cost;
}

// igpc_rollout_one with the counterfactual window's operands and states kept in a register ring
template <class Mdl, bool PF>
__device__ float igpc_rollout_one_pf(const Mdl& tru, const Mdl& sim, const DlcPlanParams& p, const float* Uo,
                                  const float* kk, const float* KK, const float* Xo, float alpha, const float* x0,
                                  float* Xw, float* Uw) {
  constexpr int D = Mdl::D, M = Mdl::M, HG = 3, MG = 3;  // H, M of GPC_rollout (:67)
  const float Cc = 1.0f;
  float W[HG + MG][D], E[HG][M][D], off[M];
#pragma unroll
  for (int i = 0; i < HG + MG; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) W[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < HG; ++i)
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int j = 0; j < D; ++j) E[i][c][j] = 0.f;
#pragma unroll
  for (int c = 0; c < M; ++c) off[c] = 0.f;
  float x[D], u[M], xn[D], xs[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { x[i] = x0[i]; Xw[i] = x[i]; }
  float cost = 0.f;
  // operands of steps h-HG .. h-1 (what the counterfactual window replays), of step h and of step h+1
  StepOps<D, M> ring[HG], cur, nxt;
  float xring[HG][D];  // X[h-HG .. h-1] of THIS rollout
  if (PF) cur.load(Uo, kk, KK, Xo, 0, p.H);
  for (int h = 0; h < p.H; ++h) {
    if (PF) {
      nxt.load(Uo, kk, KK, Xo, h + 1, p.H);
    } else {
      cur.load(Uo, kk, KK, Xo, h, p.H);
      if (p.w_is != DLC_W_DE) nxt.load_x(Xo, h + 1);
    }
    // U_delta = tensordot(E, W[-M:]) + off                                :80
    cur.feedback(alpha, x, u);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < HG; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(E[i][c][j], W[HG + i][j], s);
      u[c] += Cc * (s + off[c]);
    }
    tru.step(x, u, xn);                                                    // :87
    cost += quad_cost<D, M>(p, x, u);                                      // :88
    sim.step(x, u, xs);                                                    // :91
    float w[D];
    if (p.w_is == DLC_W_DE) {
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = xn[i] - xs[i];
    } else if (p.w_is == DLC_W_DEDE) {   // D[h] = X_old[h+1] - sim(X_old[h], U_old[h])  (compute_ders)
      float xo[D], uo[M], so[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xo[i] = cur.Xo[i];
#pragma unroll
      for (int c = 0; c < M; ++c) uo[c] = cur.Uo[c];
      sim.step(xo, uo, so);
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = (xn[i] - xs[i]) - (nxt.Xo[i] - so[i]);
    } else {                             // F[h] = jac of sim at (X_old[h], U_old[h])
      float xo[D], uo[M], fx[D * D], fu[D * M];
#pragma unroll
      for (int i = 0; i < D; ++i) xo[i] = cur.Xo[i];
#pragma unroll
      for (int c = 0; c < M; ++c) uo[c] = cur.Uo[c];
      sim.jac(xo, uo, fx, fu);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float s = 0.f, t = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(fx[i * D + j], x[j] - xo[j], s);
#pragma unroll
        for (int c = 0; c < M; ++c) t = fmaf(fu[i * M + c], u[c] - uo[c], t);
        w[i] = xn[i] - ((nxt.Xo[i] + s) + t);
      }
    }
#pragma unroll
    for (int i = 0; i < HG + MG - 1; ++i)                                  // W[0] = w; roll(-1)   :103-104
#pragma unroll
      for (int j = 0; j < D; ++j) W[i][j] = W[i + 1][j];
#pragma unroll
    for (int j = 0; j < D; ++j) W[HG + MG - 1][j] = w[j];
#pragma unroll
    for (int c = 0; c < M; ++c) Uw[h * M + c] = u[c];
#pragma unroll
    for (int i = 0; i < D; ++i) Xw[(h + 1) * D + i] = xn[i];

    if (h >= HG) {                                                         // :105-125
      // forward: y_0 = X[h-H]; y_{j+1} = sim(y_j, u_j) + W[j+M]
      float y[HG][D], uu[HG][M], Fx[HG - 1][D * D], Fu[HG - 1][D * M];
#pragma unroll
      for (int i = 0; i < D; ++i) y[0][i] = xring[0][i];   // X[h-H] of this rollout
#pragma unroll
      for (int j = 0; j < HG; ++j) {
        ring[j].feedback(alpha, y[j], uu[j]);              // step h-H+j
#pragma unroll
        for (int c = 0; c < M; ++c) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < HG; ++i)
#pragma unroll
            for (int l = 0; l < D; ++l) s = fmaf(E[i][c][l], W[j + i][l], s);
          uu[j][c] += Cc * (s + off[c]);
        }
        if (j < HG - 1) {
          float yn[D];
          sim.jac(y[j], uu[j], Fx[j], Fu[j]);
          sim.step(y[j], uu[j], yn);
#pragma unroll
          for (int i = 0; i < D; ++i) y[j + 1][i] = yn[i] + W[j + MG][i];
        }
      }
      // adjoint of c(y_{H-1}, u_{H-1})
      float ly[D], lu[M], cxx[D * D], cuu[M * M];
      cost_ders<D, M>(p, y[HG - 1], uu[HG - 1], ly, lu, cxx, cuu);
      const float sc = p.lr / (float)h;
#pragma unroll
      for (int j = HG - 1; j >= 0; --j) {
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const float g = Cc * lu[c];
          off[c] -= sc * g;
#pragma unroll
          for (int i = 0; i < HG; ++i)
#pragma unroll
            for (int l = 0; l < D; ++l) E[i][c][l] -= sc * (g * W[j + i][l]);
        }
        // note: E/off of this very update must not feed the remaining adjoint steps -- they do not:
        // the adjoint only reads W, K, Fx, Fu.
        const float* Kj = ring[j].KK;
#pragma unroll
        for (int l = 0; l < D; ++l) {
          float s = ly[l];
#pragma unroll
          for (int c = 0; c < M; ++c) s = fmaf(Kj[c * D + l], lu[c], s);
          ly[l] = s;
        }
        if (j > 0) {
          float nu[M], ny[D];
#pragma unroll
          for (int c = 0; c < M; ++c) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(Fu[j - 1][i * M + c], ly[i], s);
            nu[c] = s;
          }
#pragma unroll
          for (int l = 0; l < D; ++l) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(Fx[j - 1][i * D + l], ly[i], s);
            ny[l] = s;
          }
#pragma unroll
          for (int c = 0; c < M; ++c) lu[c] = nu[c];
#pragma unroll
          for (int l = 0; l < D; ++l) ly[l] = ny[l];
        }
      }
    }
    // slide the window: step h joins the ring, h+1 becomes current
#pragma unroll
    for (int j = 0; j < HG - 1; ++j) {
      ring[j] = ring[j + 1];
#pragma unroll
      for (int i = 0; i < D; ++i) xring[j][i] = xring[j + 1][i];
    }
    ring[HG - 1] = cur;
#pragma unroll
    for (int i = 0; i < D; ++i) { xring[HG - 1][i] = x[i]; x[i] = xn[i]; }
    if (PF) cur = nxt;
  }
  return cost;
}

// the register ring of igpc_rollout_one_pf costs 5 x (2m + md + d) registers: worth it for d <= 4, spills beyond
template <class Mdl>
__device__ __forceinline__ float igpc_rollout_best(const Mdl& tru, const Mdl& sim, const DlcPlanParams& p,
                                                   const float* Uo, const float* kk, const float* KK, const float* Xo,
                                                   float alpha, const float* x0, float* Xw, float* Uw) {
  if constexpr (Mdl::D <= 4) return igpc_rollout_one_pf<Mdl, true>(tru, sim, p, Uo, kk, KK, Xo, alpha, x0, Xw, Uw);
  else return igpc_rollout_one(tru, sim, p, Uo, kk, KK, Xo, alpha, x0, Xw, Uw);
}

// ----------------------------------------------------------------------------------- kernels
struct PlanArgs {
  DlcPlanParams p;
  const float *U_old, *k, *K, *X_old, *alpha, *x0;
  float *X_out, *U_out, *cost_out, *k_out, *K_out, *alpha_out;
  int32_t* status_out;
  int32_t* decision_out;   // [T,N]: accepted backtracking candidate per outer iteration, -1 = none (or NULL)
  float* cost_trace_out;   // [T,N]: cost after each outer iteration (or NULL)
  float* ws;
  int use_sim;
};

template <int KIND>
__global__ void __launch_bounds__(128) env_rollout_kernel(const PlanArgs a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.p.N) return;
  const int H = a.p.H;
  const Mdl env(a.use_sim ? a.p.env_sim : a.p.env_true);
  const float c = rollout_one_pf<Mdl, true>(env, a.p, a.U_old + n * H * M, a.k ? a.k + n * H * M : nullptr,
                              a.K ? a.K + n * H * M * D : nullptr, a.X_old ? a.X_old + n * (H + 1) * D : nullptr,
                              a.alpha ? a.alpha[n] : 1.0f, a.x0 + n * D, a.X_out + n * (H + 1) * D,
                              a.U_out + n * H * M);
  a.cost_out[n] = c;
}

// Learner.generate_trajectories (deluca/learners/core.py:17-41): N random-policy episodes of T steps,
// action ~ N(0, 1) per step (SimpleRandom, deluca/agents/_random.py:39-53), emitting the dataset
// triple (obs, action, next_obs)[N,T,.].  Actions come from the library's Philox stream
// (normal number t*m + c of counter stream (env, (t*m + c) / 4, 0, 2)); jax.random streams are not reproducible.
// One thread per trajectory; each warp parks TC = 96/d steps of its 32 trajectories in shared
// memory and writes them out row-wise, so the [N,T,.] batch-major stores are contiguous runs of
// TC*D floats per trajectory instead of D-float fragments.
struct CollectArgs {
  DlcEnvSpec env;
  int64_t N, env_offset;
  int T;
  uint64_t seed;
  float action_std;
  const float* x0;
  float *obs, *act, *next_obs;
};

template <int KIND>
struct CollectLayout {
  static constexpr int D = Model<KIND>::D, M = Model<KIND>::M;
  static constexpr int TC = 96 / D;                                  // 32 / 24 / 16 steps parked per trajectory
  static constexpr int XS = ((TC + 1) * D) | 1, US = (TC * M) | 1;   // odd row strides: conflict-free columns
  static constexpr size_t kSmemBytes = (size_t)4 * 32 * (XS + US) * sizeof(float);
};

template <int KIND>
__global__ void __launch_bounds__(128) env_collect_kernel(const CollectArgs a) {
  using Mdl = Model<KIND>;
  using CL = CollectLayout<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M, TC = CL::TC, XS = CL::XS, US = CL::US;
  extern __shared__ float csm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* const sxw = csm + warp * (32 * (XS + US));
  float* const suw = sxw + 32 * XS;
  const int64_t nw = (int64_t)blockIdx.x * blockDim.x + warp * 32;  // first trajectory of this warp
  const int64_t n = nw + lane;
  const bool live = n < a.N;
  const Mdl env(a.env);
  float x[D], u[M], xn[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = live ? a.x0[n * D + i] : 0.f;
  float* mx = sxw + lane * XS;
  float* mu = suw + lane * US;
  const int rows = (int)min((int64_t)32, a.N - nw);
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t zblk = 0xffffffffu;   // Philox block held in z
  for (int t0 = 0; t0 < a.T; t0 += TC) {
    const int tc = min(TC, a.T - t0);
#pragma unroll
    for (int i = 0; i < D; ++i) mx[i] = x[i];
    for (int j = 0; j < tc; ++j) {
      // action component c of step t is normal number t*M + c of the trajectory's stream: one Philox block
      // serves 4 / M steps
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const uint32_t f = (uint32_t)(t0 + j) * M + c;
        if ((f >> 2) != zblk) {
          zblk = f >> 2;
          normal4(a.seed, (uint32_t)(a.env_offset + n), zblk, 0u, 2u, z);
        }
        const uint32_t k = f & 3u;
        u[c] = __fmul_rn(a.action_std, k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3])));
      }
      env.step(x, u, xn);
#pragma unroll
      for (int c = 0; c < M; ++c) mu[j * M + c] = u[c];
#pragma unroll
      for (int i = 0; i < D; ++i) { x[i] = xn[i]; mx[(j + 1) * D + i] = xn[i]; }
    }
    __syncwarp();
    for (int r = 0; r < rows; ++r) {
      const int64_t base = (nw + r) * a.T + t0;
      const float* rx = sxw + r * XS;
      const float* ru = suw + r * US;
      for (int q = lane; q < tc * D; q += 32) {
        a.obs[base * D + q] = rx[q];
        a.next_obs[base * D + q] = rx[q + D];
      }
      for (int q = lane; q < tc * M; q += 32) a.act[base * M + q] = ru[q];
    }
    __syncwarp();
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) igpc_rollout_kernel(const PlanArgs a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.p.N) return;
  const int H = a.p.H;
  const Mdl tru(a.p.env_true), sim(a.p.env_sim);
  a.cost_out[n] = igpc_rollout_best(tru, sim, a.p, a.U_old + n * H * M, a.k + n * H * M, a.K + n * H * M * D,
                                   a.X_old + n * (H + 1) * D, a.alpha ? a.alpha[n] : 1.0f, a.x0 + n * D,
                                   a.X_out + n * (H + 1) * D, a.U_out + n * H * M);
}

struct LinArgs {
  DlcPlanParams p;
  const float *X, *U;
  float *D_out, *Fx, *Fu, *cx, *cu, *cxx, *cuu;
  int use_sim;
};

template <int KIND>
__global__ void __launch_bounds__(128) linearize_kernel(const LinArgs a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (n, h)
  const int H = a.p.H;
  if (idx >= a.p.N * H) return;
  const int64_t n = idx / H;
  const int h = (int)(idx % H);
  const Mdl env(a.use_sim ? a.p.env_sim : a.p.env_true);
  float x[D], u[M], xn[D], fx[D * D], fu[D * M], cx[D], cu[M], cxx[D * D], cuu[M * M];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = a.X[(n * (H + 1) + h) * D + i];
#pragma unroll
  for (int c = 0; c < M; ++c) u[c] = a.U[(n * H + h) * M + c];
  env.step(x, u, xn);
  env.jac(x, u, fx, fu);
  cost_ders<D, M>(a.p, x, u, cx, cu, cxx, cuu);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (a.D_out) a.D_out[idx * D + i] = a.X[(n * (H + 1) + h + 1) * D + i] - xn[i];
    a.cx[idx * D + i] = cx[i];
  }
#pragma unroll
  for (int i = 0; i < D * D; ++i) { a.Fx[idx * D * D + i] = fx[i]; a.cxx[idx * D * D + i] = cxx[i]; }
#pragma unroll
  for (int i = 0; i < D * M; ++i) a.Fu[idx * D * M + i] = fu[i];
#pragma unroll
  for (int c = 0; c < M; ++c) a.cu[idx * M + c] = cu[c];
#pragma unroll
  for (int i = 0; i < M * M; ++i) a.cuu[idx * M * M + i] = cuu[i];
}

template <int D, int M>
__global__ void __launch_bounds__(128) riccati_kernel(int64_t N, int H, const float* Fx, const float* Fu,
                                                      const float* cx, const float* cu, const float* cxx,
                                                      const float* cuu, float* k_out, float* K_out,
                                                      int32_t* status_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float Vx[D], Vxx[D * D];
#pragma unroll
  for (int i = 0; i < D; ++i) Vx[i] = 0.f;
#pragma unroll
  for (int i = 0; i < D * D; ++i) Vxx[i] = 0.f;
  int status = 0;
  for (int h = H - 1; h >= 0; --h) {
    const int64_t o = n * H + h;
    float fx[D * D], fu[D * M], lcx[D], lcu[M], lcxx[D * D], lcuu[M * M];
#pragma unroll
    for (int i = 0; i < D * D; ++i) { fx[i] = Fx[o * D * D + i]; lcxx[i] = cxx[o * D * D + i]; }
#pragma unroll
    for (int i = 0; i < D * M; ++i) fu[i] = Fu[o * D * M + i];
#pragma unroll
    for (int i = 0; i < D; ++i) lcx[i] = cx[o * D + i];
#pragma unroll
    for (int c = 0; c < M; ++c) lcu[c] = cu[o * M + c];
#pragma unroll
    for (int i = 0; i < M * M; ++i) lcuu[i] = cuu[o * M * M + i];
    status |= riccati_step<D, M>(fx, fu, lcx, lcu, lcxx, lcuu, Vx, Vxx, k_out + o * M, K_out + o * M * D);
  }
  if (status_out) status_out[n] = status;
}

// Cholesky PD test + inverse of a small SPD matrix (lqr_solver.py:157-163 isPD_and_invert: inv = L^-T L^-1).
// A non-positive or non-finite pivot reports "not PD" (NumPy raises / yields NaN there; both mean "raise lambda").
template <int M>
__device__ bool chol_inv(const float* A, float* Ai) {
  float L[M][M], Li[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) { L[i][j] = 0.f; Li[i][j] = 0.f; }
  for (int j = 0; j < M; ++j) {
    float s = A[j * M + j];
    for (int l = 0; l < j; ++l) s -= L[j][l] * L[j][l];
    if (!(s > 0.f) || !isfinite(s)) return false;
    L[j][j] = sqrtf(s);
    for (int i = j + 1; i < M; ++i) {
      float t = A[i * M + j];
      for (int l = 0; l < j; ++l) t -= L[i][l] * L[j][l];
      L[i][j] = t / L[j][j];
    }
  }
  for (int c = 0; c < M; ++c)          // forward substitution L Li = I
    for (int i = c; i < M; ++i) {
      float t = i == c ? 1.f : 0.f;
      for (int l = c; l < i; ++l) t -= L[i][l] * Li[l][c];
      Li[i][c] = t / L[i][i];
    }
  bool ok = true;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      float t = 0.f;
      for (int l = 0; l < M; ++l) t = fmaf(Li[l][i], Li[l][j], t);
      Ai[i * M + j] = t;
      ok &= isfinite(t);
    }
  return ok;
}

// LQG (lqr_solver.py:89-154): regularised backward pass with c_ux, Cholesky PD test and the lambda schedule.
template <int D, int M>
__global__ void __launch_bounds__(128) lqg_kernel(int64_t N, int H, const float* Fx, const float* Fu, const float* cx,
                                                  const float* cu, const float* cxx, const float* cuu,
                                                  const float* cux, float* lambda_io, float* mult_io, float factor,
                                                  float lambda_min, int reg_type, float* k_out, float* K_out,
                                                  float* gains_out, int32_t* status_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float lam = lambda_io[n], mult = mult_io[n];
  float gl = 0.f, gq = 0.f;
  bool ok = false;
  for (int attempt = 1; attempt < 10 && !ok; ++attempt) {          // `counter >= 10` gives up            :100-105
    float Vx[D], Vxx[D * D];
#pragma unroll
    for (int i = 0; i < D; ++i) Vx[i] = 0.f;
#pragma unroll
    for (int i = 0; i < D * D; ++i) Vxx[i] = 0.f;
    gl = 0.f;
    gq = 0.f;
    ok = true;
    for (int h = H - 1; h >= 0 && ok; --h) {
      const int64_t o = n * H + h;
      const float* fx = Fx + o * D * D;
      const float* fu = Fu + o * D * M;
      float VF[D * D], VFu[D * M], Qx[D], Qu[M], Qxx[D * D], Qux[M * D], Quu[M * M], Qur[M * M], Qxr[M * D];
      for (int i = 0; i < D; ++i) {
        for (int j = 0; j < D; ++j) {
          float t = 0.f;
          for (int l = 0; l < D; ++l) t = fmaf(Vxx[i * D + l], fx[l * D + j], t);
          VF[i * D + j] = t;
        }
        for (int c = 0; c < M; ++c) {
          float t = 0.f;
          for (int l = 0; l < D; ++l) t = fmaf(Vxx[i * D + l], fu[l * M + c], t);
          VFu[i * M + c] = t;
        }
      }
      for (int i = 0; i < D; ++i) {
        float t = 0.f;
        for (int l = 0; l < D; ++l) t = fmaf(fx[l * D + i], Vx[l], t);
        Qx[i] = cx[o * D + i] + t;
        for (int j = 0; j < D; ++j) {
          float v = 0.f;
          for (int l = 0; l < D; ++l) v = fmaf(fx[l * D + i], VF[l * D + j], v);
          Qxx[i * D + j] = cxx[o * D * D + i * D + j] + v;
        }
      }
      for (int c = 0; c < M; ++c) {
        float t = 0.f;
        for (int l = 0; l < D; ++l) t = fmaf(fu[l * M + c], Vx[l], t);
        Qu[c] = cu[o * M + c] + t;
        for (int j = 0; j < D; ++j) {
          float v = 0.f, r = 0.f;
          for (int l = 0; l < D; ++l) { v = fmaf(fu[l * M + c], VF[l * D + j], v); r = fmaf(fu[l * M + c], fx[l * D + j], r); }
          Qux[c * D + j] = cux[o * M * D + c * D + j] + v;
          Qxr[c * D + j] = reg_type == 1 ? Qux[c * D + j] + lam * r : Qux[c * D + j];   // type V   :127-129
        }
        for (int e = 0; e < M; ++e) {
          float v = 0.f, r = 0.f;
          for (int l = 0; l < D; ++l) { v = fmaf(fu[l * M + c], VFu[l * M + e], v); r = fmaf(fu[l * M + c], fu[l * M + e], r); }
          Quu[c * M + e] = cuu[o * M * M + c * M + e] + v;
          Qur[c * M + e] = reg_type == 1 ? Quu[c * M + e] + lam * r                      // type V
                                         : Quu[c * M + e] + (c == e ? lam : 0.f);        // type Q   :122-125
        }
      }
      float Qi[M * M];
      if (!chol_inv<M>(Qur, Qi)) { ok = false; break; }
      float Kl[M * D], kl[M];
      for (int c = 0; c < M; ++c) {
        float t = 0.f;
        for (int e = 0; e < M; ++e) t = fmaf(Qi[c * M + e], Qu[e], t);
        kl[c] = -t;
        k_out[o * M + c] = -t;
        for (int j = 0; j < D; ++j) {
          float v = 0.f;
          for (int e = 0; e < M; ++e) v = fmaf(Qi[c * M + e], Qxr[e * D + j], v);
          Kl[c * D + j] = -v;
          K_out[o * M * D + c * D + j] = -v;
        }
      }
      // V_x = Q_x + K'Q_uu k + K'Q_u + Q_ux'k ; V_xx = Q_xx + K'Q_uu K + K'Q_ux + Q_ux'K (unregularised Q)  :134-137
      float Qk[M], QK[M * D];
      for (int c = 0; c < M; ++c) {
        float t = 0.f;
        for (int e = 0; e < M; ++e) t = fmaf(Quu[c * M + e], kl[e], t);
        Qk[c] = t;
        for (int j = 0; j < D; ++j) {
          float v = 0.f;
          for (int e = 0; e < M; ++e) v = fmaf(Quu[c * M + e], Kl[e * D + j], v);
          QK[c * D + j] = v;
        }
      }
      float Vn[D * D];
      for (int i = 0; i < D; ++i) {
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int c = 0; c < M; ++c) {
          a1 = fmaf(Kl[c * D + i], Qk[c], a1);
          a2 = fmaf(Kl[c * D + i], Qu[c], a2);
          a3 = fmaf(Qux[c * D + i], kl[c], a3);
        }
        Vx[i] = ((Qx[i] + a1) + a2) + a3;
        for (int j = 0; j < D; ++j) {
          float b1 = 0.f, b2 = 0.f, b3 = 0.f;
          for (int c = 0; c < M; ++c) {
            b1 = fmaf(Kl[c * D + i], QK[c * D + j], b1);
            b2 = fmaf(Kl[c * D + i], Qux[c * D + j], b2);
            b3 = fmaf(Qux[c * D + i], Kl[c * D + j], b3);
          }
          Vn[i * D + j] = ((Qxx[i * D + j] + b1) + b2) + b3;
        }
      }
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Vxx[i * D + j] = (Vn[i * D + j] + Vn[j * D + i]) * 0.5f;
      for (int c = 0; c < M; ++c) {
        gl = fmaf(kl[c], Qu[c], gl);                                // gains_lin  += k'Q_u       :143
        gq = fmaf(kl[c], Qk[c], gq);                                // gains_quad += k'Q_uu k    :144
      }
    }
    if (!ok) {                                                      // lambda_adjust(..., increase=True)  :21-37
      mult = fmaxf(mult * factor, factor);
      lam = fmaxf(lam * mult, lambda_min);
      lam = fmaxf(lam, lambda_min);
      if (lam > 1e10f) lam = 1e10f;
    }
  }
  lambda_io[n] = lam;
  mult_io[n] = mult;
  if (gains_out) {
    gains_out[2 * n] = ok ? gl : nanf("");
    gains_out[2 * n + 1] = ok ? gq : nanf("");
  }
  if (status_out) status_out[n] = ok ? 0 : DLC_STATUS_NOT_PD;
}

// iLQR / iGPC_closed / iLC_closed, whole outer loop per trajectory (ilqr.py:27-93, igpc.py:129-179,
// ilc.py:23-69).  Nominal (X, U) live in X_out / U_io, the candidate in the workspace.
template <int KIND>
__global__ void __launch_bounds__(128) plan_kernel(const PlanArgs a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  const DlcPlanParams& p = a.p;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int H = p.H;
  const Mdl tru(p.env_true), sim(p.env_sim);
  float* X = a.X_out + n * (H + 1) * D;
  float* U = a.U_out + n * H * M;
  float* kk = a.k_out + n * H * M;
  float* KK = a.K_out + n * H * M * D;
  float* XC = a.ws + n * ((H + 1) * D + H * M);
  float* UC = XC + (H + 1) * D;
  const float* x0 = a.x0 + n * D;
  // initial open-loop rollout into the candidate buffers, then adopt it (U may alias U_old)
  float c = rollout_one_pf<Mdl, true>(tru, p, U, nullptr, nullptr, nullptr, 1.0f, x0, X, UC);
  float alpha = p.alpha0;
  int status = 0;
  for (int t = 0; t < p.T; ++t) {
    status |= backward_pass_pf<Mdl, true>(sim, p, X, U, kk, KK);   // compute_ders(env_sim) + LQR
    const bool ilqr = p.algo == DLC_ALGO_ILQR;
    int dec = -1;
    if (p.backtracking) {
      // iLQR as written passes `alpha` (not alphaC) to every candidate: they are identical, so only
      // the first can be accepted (SURVEY A-10)
      const int ncand = (ilqr && !p.fix_backtracking) ? 1 : p.n_alpha;
      for (int j = 0; j < ncand; ++j) {
        const float alphaC = (alpha * 1.1f) * powf(1.1f, -(float)(j * j));
        const float a_use = (ilqr && !p.fix_backtracking) ? alpha : alphaC;
        float cC;
        if (p.algo == DLC_ALGO_IGPC) cC = igpc_rollout_best(tru, sim, p, U, kk, KK, X, a_use, x0, XC, UC);
        else cC = rollout_one_pf<Mdl, true>(tru, p, U, kk, KK, X, a_use, x0, XC, UC);
        const bool accept = p.algo == DLC_ALGO_IGPC ? (cC < c) : (cC <= c);
        if (accept) {
          for (int i = 0; i < (H + 1) * D; ++i) X[i] = XC[i];
          for (int i = 0; i < H * M; ++i) U[i] = UC[i];
          c = cC;
          alpha = alphaC;
          dec = j;
          break;
        }
      }
    } else {
      if (p.algo == DLC_ALGO_IGPC) c = igpc_rollout_best(tru, sim, p, U, kk, KK, X, alpha, x0, XC, UC);
      else c = rollout_one_pf<Mdl, true>(tru, p, U, kk, KK, X, alpha, x0, XC, UC);
      for (int i = 0; i < (H + 1) * D; ++i) X[i] = XC[i];
      for (int i = 0; i < H * M; ++i) U[i] = UC[i];
      dec = 0;
    }
    if (a.decision_out) a.decision_out[(int64_t)t * p.N + n] = dec;
    if (a.cost_trace_out) a.cost_trace_out[(int64_t)t * p.N + n] = c;
  }
  a.cost_out[n] = c;
  if (a.alpha_out) a.alpha_out[n] = alpha;
  if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
  if (a.status_out) a.status_out[n] = status;
}

// The same outer loops with the backtracking candidates of an iteration evaluated SIDE BY SIDE: a
// group of PLAN_G = 8 lanes owns a trajectory, lane j rolls out candidate j (alphaC = 1.1 alpha 1.1^(-j^2))
// into its own workspace slot, and the group adopts the FIRST accepted candidate in the reference's
// order -- exactly what the sequential `for alphaC in ...: if cC <= c: break` keeps (ilqr.py:55-73,
// igpc.py:159-170, ilc.py:49-60); every candidate is the same arithmetic as in plan_kernel.  More
// than 8 candidates (iLQR / iLC: 10) go in rounds of 8; a later round runs only if no earlier candidate
// passed.  The serial chain per iteration drops from (1 + tried candidates) rollouts to 2; the planning
// loops are latency-bound with most issue slots idle, so the redundant candidates are free.
// Eight lanes, not sixteen: 4096 trajectories x 8 lanes are 256 CTAs, one wave even at 255 registers, which is
// what lets the candidates use the operand prefetch and the register ring of the serial kernel (a 16-lane
// variant had to stay under 128 registers to remain one wave and was 12-50 % slower on every loop).
constexpr int PLAN_G = 8;

__host__ __device__ inline bool plan_parallel_candidates(const DlcPlanParams& p) {
  const bool identical = p.algo == DLC_ALGO_ILQR && !p.fix_backtracking;  // SURVEY A-10: one distinct candidate
  // beyond ~64 K trajectories the thread-per-trajectory kernel already fills the machine, and 16 candidate buffers
  // per trajectory (51 KB at d = 3, H = 200) would be the larger cost
  return p.backtracking && !identical && p.n_alpha > 1 && p.n_alpha <= 2 * PLAN_G && p.N <= 65536;
}

__host__ __device__ inline int plan_group(const DlcPlanParams&) { return PLAN_G; }

template <int KIND, int G, bool PF>
__device__ __forceinline__ void plan_par_body(const PlanArgs& a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  const DlcPlanParams& p = a.p;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (n >= p.N) return;  // whole groups leave together
  const int j = threadIdx.x & (G - 1);
  const int gshift = threadIdx.x & 31 & ~(G - 1);
  const unsigned gmask = ((1u << G) - 1u) << gshift;
  const int H = p.H;
  const int WS = (H + 1) * D + H * M;
  const Mdl tru(p.env_true), sim(p.env_sim);
  float* X = a.X_out + n * (H + 1) * D;
  float* U = a.U_out + n * H * M;
  float* kk = a.k_out + n * H * M;
  float* KK = a.K_out + n * H * M * D;
  float* XC = a.ws + (n * G + j) * (int64_t)WS;
  float* UC = XC + (H + 1) * D;
  const float* x0 = a.x0 + n * D;
  float c = 0.f;
  if (j == 0) c = PF ? rollout_one_pf<Mdl, true>(tru, p, U, nullptr, nullptr, nullptr, 1.0f, x0, X, UC)
                     : rollout_one(tru, p, U, nullptr, nullptr, nullptr, 1.0f, x0, X, UC);
  c = __shfl_sync(gmask, c, gshift);
  float alpha = p.alpha0;
  int status = 0;
  for (int t = 0; t < p.T; ++t) {
    if (j == 0) status |= PF ? backward_pass_pf<Mdl, true>(sim, p, X, U, kk, KK) : backward_pass(sim, p, X, U, kk, KK);
    __syncwarp(gmask);
    int dec = -1;
    // candidates in rounds of G (one round unless n_alpha > G); a later round runs only if no earlier candidate passed
    for (int base = 0; base < p.n_alpha; base += G) {
      const int jc = base + j;
      const float alphaC = (alpha * 1.1f) * powf(1.1f, -(float)(jc * jc));
      float cC = 0.f;
      bool accept = false;
      if (jc < p.n_alpha) {
        if (PF) {
          if (p.algo == DLC_ALGO_IGPC) cC = igpc_rollout_best(tru, sim, p, U, kk, KK, X, alphaC, x0, XC, UC);
          else cC = rollout_one_pf<Mdl, true>(tru, p, U, kk, KK, X, alphaC, x0, XC, UC);
        } else {
          if (p.algo == DLC_ALGO_IGPC) cC = igpc_rollout_one(tru, sim, p, U, kk, KK, X, alphaC, x0, XC, UC);
          else cC = rollout_one(tru, p, U, kk, KK, X, alphaC, x0, XC, UC);
        }
        accept = p.algo == DLC_ALGO_IGPC ? (cC < c) : (cC <= c);
      }
      const unsigned votes = (__ballot_sync(gmask, accept) >> gshift) & ((1u << G) - 1u);
      __syncwarp(gmask);  // every candidate buffer is complete and nobody reads X / U any more
      if (votes) {
        const int js = __ffs(votes) - 1;  // first accepted in the reference's order
        const float* XS = a.ws + (n * G + js) * (int64_t)WS;
        const float* US = XS + (H + 1) * D;
        for (int i = j; i < (H + 1) * D; i += G) X[i] = XS[i];
        for (int i = j; i < H * M; i += G) U[i] = US[i];
        c = __shfl_sync(gmask, cC, gshift + js);
        alpha = __shfl_sync(gmask, alphaC, gshift + js);
        dec = base + js;
        break;   // group-uniform
      }
      __syncwarp(gmask);
    }
    __syncwarp(gmask);
    if (j == 0) {
      if (a.decision_out) a.decision_out[(int64_t)t * p.N + n] = dec;
      if (a.cost_trace_out) a.cost_trace_out[(int64_t)t * p.N + n] = c;
    }
  }
  if (j == 0) {
    a.cost_out[n] = c;
    if (a.alpha_out) a.alpha_out[n] = alpha;
    if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
    if (a.status_out) a.status_out[n] = status;
  }
}

template <int KIND>
__global__ void __launch_bounds__(128, 2) plan_par_kernel(const PlanArgs a) {
  plan_par_body<KIND, PLAN_G, true>(a);
}

static int32_t check_plan(const DlcPlanParams* p, const char* who) {
  if (!p) return fail(DLC_ERR_ARG, "NULL params");
  if (p->N < 0 || p->H < 1 || p->T < 0) return fail(DLC_ERR_ARG, "bad N / H / T");
  if (p->env_true.kind != p->env_sim.kind) return fail(DLC_ERR_SHAPE, "env_true and env_sim must be the same kind of env");
  if (p->env_true.kind < DLC_ENV_PENDULUM || p->env_true.kind > DLC_ENV_REACHER)
    return fail(DLC_ERR_SHAPE, "unknown env kind");
  (void)who;
  return DLC_OK;
}

}  // namespace dlc

using namespace dlc;

// thread-per-trajectory kernels: one-warp CTAs while the batch is small enough that spreading the warps over all SMs
// (L1 residency of each trajectory's ~10 KB working set) matters more than CTA count; 128 threads beyond that
static inline int plan_block(int64_t N) { return N <= 148 * 32 * 8 ? 32 : 128; }

#define DLC_DISPATCH_ENV_B(kind, KERNEL, grid, block, st, args)                            \
  do {                                                                                    \
    if ((kind) == DLC_ENV_PENDULUM) KERNEL<DLC_ENV_PENDULUM><<<grid, block, 0, st>>>(args); \
    else if ((kind) == DLC_ENV_CARTPOLE) KERNEL<DLC_ENV_CARTPOLE><<<grid, block, 0, st>>>(args); \
    else if ((kind) == DLC_ENV_QUADROTOR) KERNEL<DLC_ENV_QUADROTOR><<<grid, block, 0, st>>>(args); \
    else KERNEL<DLC_ENV_REACHER><<<grid, block, 0, st>>>(args);                           \
  } while (0)

#define DLC_DISPATCH_ENV(kind, KERNEL, grid, st, args)                                   \
  do {                                                                                    \
    if ((kind) == DLC_ENV_PENDULUM) KERNEL<DLC_ENV_PENDULUM><<<grid, 128, 0, st>>>(args); \
    else if ((kind) == DLC_ENV_CARTPOLE) KERNEL<DLC_ENV_CARTPOLE><<<grid, 128, 0, st>>>(args); \
    else if ((kind) == DLC_ENV_QUADROTOR) KERNEL<DLC_ENV_QUADROTOR><<<grid, 128, 0, st>>>(args); \
    else KERNEL<DLC_ENV_REACHER><<<grid, 128, 0, st>>>(args);                             \
  } while (0)

extern "C" int32_t dlc_env_rollout_f32(const DlcPlanParams* p, int32_t use_sim_env, const float* U_old,
                                       const float* k, const float* K, const float* X_old, const float* alpha,
                                       const float* x0, float* X_out, float* U_out, float* cost_out, void* stream) {
  int32_t rc = check_plan(p, "dlc_env_rollout_f32");
  if (rc) return rc;
  if (p->N == 0) return DLC_OK;
  if (!U_old || !x0 || !X_out || !U_out || !cost_out) return fail(DLC_ERR_ARG, "dlc_env_rollout_f32: NULL buffer");
  if (k && (!K || !X_old)) return fail(DLC_ERR_ARG, "dlc_env_rollout_f32: k needs K and X_old");
  PlanArgs a{};
  a.p = *p; a.U_old = U_old; a.k = k; a.K = K; a.X_old = X_old; a.alpha = alpha; a.x0 = x0;
  a.X_out = X_out; a.U_out = U_out; a.cost_out = cost_out; a.use_sim = use_sim_env;
  const int block = plan_block(p->N);
  const unsigned grid = (unsigned)((p->N + block - 1) / block);
  DLC_DISPATCH_ENV_B(p->env_true.kind, env_rollout_kernel, grid, block, (cudaStream_t)stream, a);
  return cuda_status(cudaGetLastError(), "env_rollout_kernel launch");
}

extern "C" int32_t dlc_env_collect_f32(const DlcEnvSpec* env, int64_t N, int32_t T, uint64_t seed,
                                      int64_t env_offset, float action_std, const float* x0, float* obs_out,
                                      float* action_out, float* next_obs_out, void* stream) {
  if (!env) return fail(DLC_ERR_ARG, "dlc_env_collect_f32: NULL env");
  if (N < 0 || T < 1) return fail(DLC_ERR_ARG, "dlc_env_collect_f32: bad N / T");
  if (env->kind < DLC_ENV_PENDULUM || env->kind > DLC_ENV_REACHER)
    return fail(DLC_ERR_SHAPE, "dlc_env_collect_f32: unknown env kind");
  if (N == 0) return DLC_OK;
  if (!x0 || !obs_out || !action_out || !next_obs_out) return fail(DLC_ERR_ARG, "dlc_env_collect_f32: NULL buffer");
  CollectArgs a{};
  a.env = *env; a.N = N; a.env_offset = env_offset; a.T = T; a.seed = seed; a.action_std = action_std;
  a.x0 = x0; a.obs = obs_out; a.act = action_out; a.next_obs = next_obs_out;
  const unsigned grid = (unsigned)((N + 127) / 128);
#define COLLECT(KIND_)                                                                                          \
  if (env->kind == KIND_) {                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(env_collect_kernel<KIND_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)CollectLayout<KIND_>::kSmemBytes);                                \
    if (e != cudaSuccess) return cuda_status(e, "env_collect_kernel smem opt-in");                              \
    env_collect_kernel<KIND_><<<grid, 128, CollectLayout<KIND_>::kSmemBytes, (cudaStream_t)stream>>>(a);        \
  }
  COLLECT(DLC_ENV_PENDULUM) COLLECT(DLC_ENV_CARTPOLE) COLLECT(DLC_ENV_QUADROTOR) COLLECT(DLC_ENV_REACHER)
#undef COLLECT
  return cuda_status(cudaGetLastError(), "env_collect_kernel launch");
}

extern "C" int32_t dlc_igpc_rollout_f32(const DlcPlanParams* p, const float* U_old, const float* k, const float* K,
                                        const float* X_old, const float* alpha, const float* x0, float* X_out,
                                        float* U_out, float* cost_out, void* stream) {
  int32_t rc = check_plan(p, "dlc_igpc_rollout_f32");
  if (rc) return rc;
  if (p->N == 0) return DLC_OK;
  if (!U_old || !k || !K || !X_old || !x0 || !X_out || !U_out || !cost_out)
    return fail(DLC_ERR_ARG, "dlc_igpc_rollout_f32: NULL buffer");
  if (p->w_is < DLC_W_DE || p->w_is > DLC_W_DELIN) return fail(DLC_ERR_ARG, "dlc_igpc_rollout_f32: bad w_is");
  PlanArgs a{};
  a.p = *p; a.U_old = U_old; a.k = k; a.K = K; a.X_old = X_old; a.alpha = alpha; a.x0 = x0;
  a.X_out = X_out; a.U_out = U_out; a.cost_out = cost_out;
  const int block = plan_block(p->N);
  const unsigned grid = (unsigned)((p->N + block - 1) / block);
  DLC_DISPATCH_ENV_B(p->env_true.kind, igpc_rollout_kernel, grid, block, (cudaStream_t)stream, a);
  return cuda_status(cudaGetLastError(), "igpc_rollout_kernel launch");
}

extern "C" int32_t dlc_linearize_f32(const DlcPlanParams* p, int32_t use_sim_env, const float* X, const float* U,
                                     float* D_out, float* Fx_out, float* Fu_out, float* cx_out, float* cu_out,
                                     float* cxx_out, float* cuu_out, void* stream) {
  int32_t rc = check_plan(p, "dlc_linearize_f32");
  if (rc) return rc;
  if (p->N == 0) return DLC_OK;
  if (!X || !U || !Fx_out || !Fu_out || !cx_out || !cu_out || !cxx_out || !cuu_out)
    return fail(DLC_ERR_ARG, "dlc_linearize_f32: NULL buffer");
  LinArgs a{};
  a.p = *p; a.X = X; a.U = U; a.D_out = D_out; a.Fx = Fx_out; a.Fu = Fu_out; a.cx = cx_out; a.cu = cu_out;
  a.cxx = cxx_out; a.cuu = cuu_out; a.use_sim = use_sim_env;
  const unsigned grid = (unsigned)((p->N * p->H + 127) / 128);
  DLC_DISPATCH_ENV(p->env_true.kind, linearize_kernel, grid, (cudaStream_t)stream, a);
  return cuda_status(cudaGetLastError(), "linearize_kernel launch");
}

extern "C" int32_t dlc_riccati_backward_f32(int64_t N, int32_t H, int32_t d, int32_t m, const float* Fx,
                                            const float* Fu, const float* cx, const float* cu, const float* cxx,
                                            const float* cuu, float* k_out, float* K_out, int32_t* status_out,
                                            void* stream) {
  if (N < 0 || H < 1) return fail(DLC_ERR_ARG, "dlc_riccati_backward_f32: bad N / H");
  if (N == 0) return DLC_OK;
  if (!Fx || !Fu || !cx || !cu || !cxx || !cuu || !k_out || !K_out)
    return fail(DLC_ERR_ARG, "dlc_riccati_backward_f32: NULL buffer");
  const unsigned grid = (unsigned)((N + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
#define RIC(D_, M_)                                                                                         \
  if (d == D_ && m == M_) {                                                                                 \
    riccati_kernel<D_, M_><<<grid, 128, 0, st>>>(N, H, Fx, Fu, cx, cu, cxx, cuu, k_out, K_out, status_out); \
    return cuda_status(cudaGetLastError(), "riccati_kernel launch");                                        \
  }
  RIC(3, 1) RIC(4, 1) RIC(2, 1) RIC(4, 2) RIC(6, 2) RIC(6, 3) RIC(8, 2)
#undef RIC
  return fail(DLC_ERR_SHAPE, "dlc_riccati_backward_f32: (d, m) not instantiated: (2,1) (3,1) (4,1) (4,2) (6,2) (6,3) (8,2)");
}

extern "C" int32_t dlc_lqg_backward_f32(int64_t N, int32_t H, int32_t d, int32_t m, const float* Fx, const float* Fu,
                                        const float* cx, const float* cu, const float* cxx, const float* cuu,
                                        const float* cux, float* lambda_io, float* multiplier_io, float factor,
                                        float lambda_min, int32_t reg_type, float* k_out, float* K_out,
                                        float* gains_out, int32_t* status_out, void* stream) {
  if (N < 0 || H < 1) return fail(DLC_ERR_ARG, "dlc_lqg_backward_f32: bad N / H");
  if (reg_type != 0 && reg_type != 1) return fail(DLC_ERR_ARG, "dlc_lqg_backward_f32: reg_type must be 0 (Q) or 1 (V)");
  if (N == 0) return DLC_OK;
  if (!Fx || !Fu || !cx || !cu || !cxx || !cuu || !cux || !lambda_io || !multiplier_io || !k_out || !K_out)
    return fail(DLC_ERR_ARG, "dlc_lqg_backward_f32: NULL buffer");
  const unsigned grid = (unsigned)((N + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
#define LQGK(D_, M_)                                                                                           \
  if (d == D_ && m == M_) {                                                                                    \
    lqg_kernel<D_, M_><<<grid, 128, 0, st>>>(N, H, Fx, Fu, cx, cu, cxx, cuu, cux, lambda_io, multiplier_io,    \
                                             factor, lambda_min, reg_type, k_out, K_out, gains_out, status_out); \
    return cuda_status(cudaGetLastError(), "lqg_kernel launch");                                               \
  }
  LQGK(3, 1) LQGK(4, 1) LQGK(2, 1) LQGK(4, 2) LQGK(6, 2) LQGK(6, 3) LQGK(8, 2)
#undef LQGK
  return fail(DLC_ERR_SHAPE, "dlc_lqg_backward_f32: (d, m) not instantiated: (2,1) (3,1) (4,1) (4,2) (6,2) (6,3) (8,2)");
}

extern "C" size_t dlc_plan_workspace_bytes(const DlcPlanParams* p) {
  if (!p || p->N <= 0 || p->H < 1) return 0;
  const int d = p->env_true.kind == DLC_ENV_PENDULUM ? 3 : (p->env_true.kind == DLC_ENV_CARTPOLE ? 4 : 6);
  const int m = p->env_true.kind >= DLC_ENV_QUADROTOR ? 2 : 1;
  const size_t slots = plan_parallel_candidates(*p) ? plan_group(*p) : 1;  // one candidate buffer per lane of a group
  return (size_t)p->N * slots * ((size_t)(p->H + 1) * d + (size_t)p->H * m) * sizeof(float);
}

extern "C" int32_t dlc_plan_f32(const DlcPlanParams* p, float* U_io, const float* x0, float* X_out, float* k_out,
                                float* K_out, float* cost_out, float* alpha_out, int32_t* status_out,
                                int32_t* decision_out, float* cost_trace_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  int32_t rc = check_plan(p, "dlc_plan_f32");
  if (rc) return rc;
  if (p->algo < DLC_ALGO_ILQR || p->algo > DLC_ALGO_ILC) return fail(DLC_ERR_ARG, "dlc_plan_f32: unknown algo");
  if (p->w_is < DLC_W_DE || p->w_is > DLC_W_DELIN) return fail(DLC_ERR_ARG, "dlc_plan_f32: bad w_is");
  if (p->backtracking && p->n_alpha < 1) return fail(DLC_ERR_ARG, "dlc_plan_f32: n_alpha must be >= 1");
  if (p->N == 0) return DLC_OK;
  if (!U_io || !x0 || !X_out || !k_out || !K_out || !cost_out)
    return fail(DLC_ERR_ARG, "dlc_plan_f32: NULL buffer");
  if (!workspace || workspace_bytes < dlc_plan_workspace_bytes(p))
    return fail(DLC_ERR_WORKSPACE, "dlc_plan_f32: workspace smaller than dlc_plan_workspace_bytes()");
  PlanArgs a{};
  a.p = *p; a.U_out = U_io; a.x0 = x0; a.X_out = X_out; a.k_out = k_out; a.K_out = K_out; a.cost_out = cost_out;
  a.alpha_out = alpha_out; a.status_out = status_out; a.ws = (float*)workspace;
  a.decision_out = decision_out; a.cost_trace_out = cost_trace_out;
  if (plan_parallel_candidates(*p)) {
    const unsigned grid = (unsigned)((p->N * plan_group(*p) + 127) / 128);
    DLC_DISPATCH_ENV(p->env_true.kind, plan_par_kernel, grid, (cudaStream_t)stream, a);
    return cuda_status(cudaGetLastError(), "plan_par_kernel launch");
  }
  // one-warp CTAs: a trajectory's working set is ~10 KB, so spreading the warps over all SMs (instead of four
  // per SM on a quarter of them) is what keeps it L1-resident
  const unsigned grid = (unsigned)((p->N + 31) / 32);
  DLC_DISPATCH_ENV_B(p->env_true.kind, plan_kernel, grid, 32, (cudaStream_t)stream, a);
  return cuda_status(cudaGetLastError(), "plan_kernel launch");
}
