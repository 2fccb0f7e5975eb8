// Fused LDS + {LQR, GPC, BPC} rollout kernels.  See include/deluca_b200.h for
// the reference lines each piece stands for; DESIGN.md section 3 for the layout/roofline.
//
// Two kernels:
//   lds_tpe_kernel<D,M,H,HH,POLICY>  "thread per env", dims compile-time.  A and GPC's M live in
//       registers, B, K and the disturbance-history ring live in shared memory laid out
//       [element][thread] (bank = thread, so every access is conflict-free); the T-step loop is
//       fused, only costs / sampled trajectories / final state go back to HBM.
//   lds_generic_kernel               any (d,m,H,HH,policy incl. BPC), runtime dims, per-env state in
//       a global-memory workspace laid out [element][env] (coalesced).  Correct for every shape
//       the reference accepts (including the dynamic_slice clamping of HH < H-1); not tuned.
#include "common.cuh"
#include "lds_args.cuh"

namespace dlc {

// LdsArgs lives in lds_args.cuh (shared with lds_rollout_f2.cu)

// ------------------------------------------------------------------------------------------
// lane-split kernel: L adjacent lanes of a warp own one env.
//
// Lane q of a group owns the state coordinates S_q = [q*DL, q*DL+DL) (DL = ceil(D/L); coordinates
// >= D are zero padding) and keeps, in REGISTERS, everything indexed by them:
//     rows S_q of A and B, columns S_q of K and of every M[i], and x/y/lambda restricted to S_q.
// The disturbance-history ring (columns S_q only) is the one structure indexed dynamically; it
// lives in shared memory as [slot][k][thread] (bank = thread: conflict-free) and is MIRRORED: every
// entry is stored at physical slots i and i + P (2P - 1 slots), so logical index k is always at head + k with no
// wrap test and every access of a step is base-register + immediate.
// Vectors that a lane needs in full are exchanged with xor-shuffles and kept in LANE-RELATIVE
// order: chunk c of a gathered vector holds the coordinates of lane q^c, so every register index
// is a compile-time constant.  A's rows are stored with their columns permuted the same way.
//     A y      : all-gather y, then own rows            ((L-1)*DL shuffles)
//     A' lam   : partial sums over own rows, reduce-scatter ((L-1)*DL shuffles)
//     K y, B' lam, M-window products: partial sums over own coordinates, all-reduce (m*log2 L)
//     K' du, B v, rank-1 updates of M: local.
// ------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ float group_allreduce(float v) {
#pragma unroll
  for (int o = 1; o < L; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int D, int M, int H, int HH, int POLICY, int L, int TPB>
struct SplitLayout {
  static constexpr int DL = (D + L - 1) / L;
  static constexpr int P = H + HH;
  static constexpr int kFloats = POLICY == DLC_POLICY_GPC ? (2 * P - 1) * DL : 0;  // per thread (mirrored ring)
  static constexpr size_t kSmemBytes = (size_t)kFloats * TPB * sizeof(float);
};

template <int D, int M, int H, int HH, int POLICY, int L, int TPB, int MINB>
__global__ void __launch_bounds__(TPB) __maxnreg__((65536 / (TPB * MINB)) / 8 * 8 > 255 ? 255 : (65536 / (TPB * MINB)) / 8 * 8)
    lds_split_kernel(const LdsArgs a) {
  using Lay = SplitLayout<D, M, H, HH, POLICY, L, TPB>;
  constexpr int DL = Lay::DL, DF = DL * L, P = Lay::P;
  constexpr bool GPC = POLICY == DLC_POLICY_GPC;
  static_assert(!GPC || HH >= H - 1, "templated path needs HH >= H-1 (no index clamping)");
  static_assert(L == 1 || L == 2 || L == 4, "group size");
  extern __shared__ float sm[];
  const DlcLdsParams& p = a.p;
  const int tid = threadIdx.x;
  const int q = tid % L;  // lane within the group
  const int64_t n_raw = ((int64_t)blockIdx.x * TPB + tid) / L;
  const bool live = n_raw < p.N;
  const int64_t n = live ? n_raw : p.N - 1;  // idle groups shadow the last env (keeps shuffles whole)
  float* const sH = sm + tid;                // ring element (slot, k): sH[(slot*DL + k) * TPB]

  // ---- parameters -> registers
  const int64_t np = p.shared_dynamics ? 0 : n;
  float Ar[DL][DF], Br[DL][M], Kc[M][DL];
#pragma unroll
  for (int i = 0; i < DL; ++i) {
    const int gi = q * DL + i;
#pragma unroll
    for (int c = 0; c < L; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        const int gj = (q ^ c) * DL + k;
        Ar[i][c * DL + k] = (gi < D && gj < D) ? __ldg(a.A + np * (D * D) + gi * D + gj) : 0.f;
      }
#pragma unroll
    for (int c = 0; c < M; ++c) {
      Br[i][c] = gi < D ? __ldg(a.B + np * (D * M) + gi * M + c) : 0.f;
      Kc[c][i] = gi < D ? __ldg(a.K + np * (M * D) + c * D + gi) : 0.f;
    }
  }
  if (POLICY == DLC_POLICY_GPC) {
    // GPC keeps the closed-loop matrix A - B K instead of A: the counterfactual recursion
    // y <- A y + B(-K y + v) + w is (A - BK) y + B v + w and its adjoint lam <- A'lam - K'B'lam is
    // (A - BK)' lam, which removes one K mat-vec and its cross-lane reduction per recursion step;
    // the env step recovers A x = (A - BK) x + B (K x) from the K x it needs for the action anyway.
#pragma unroll
    for (int i = 0; i < DL; ++i) {
#pragma unroll
      for (int c = 0; c < L; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k) {
          const int gj = (q ^ c) * DL + k;
          float s = Ar[i][c * DL + k];
#pragma unroll
          for (int cc = 0; cc < M; ++cc)
            s = fmaf(-Br[i][cc], gj < D ? __ldg(a.K + np * (M * D) + cc * D + gj) : 0.f, s);
          Ar[i][c * DL + k] = s;
        }
    }
  }
  float x[DL], ax[DL], uprev[M];
#pragma unroll
  for (int k = 0; k < DL; ++k) x[k] = (q * DL + k < D) ? a.x_io[n * D + q * DL + k] : 0.f;
  float Mr[GPC ? H : 1][M][DL];
  int head = 0;
  if (GPC) {
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          Mr[i][c][k] = (q * DL + k < D) ? a.M_io[n * (H * M * D) + (i * M + c) * D + q * DL + k] : 0.f;
#pragma unroll
    for (int s = 0; s < P; ++s)
#pragma unroll
      for (int k = 0; k < DL; ++k)
      {
        const float v = (q * DL + k < D) ? a.hist_io[n * (P * D) + s * D + q * DL + k] : 0.f;
        sH[(s * DL + k) * TPB] = v;
        if (s + P < 2 * P - 1) sH[((s + P) * DL + k) * TPB] = v;
      }
    // ax = A * (previous observed state).  The reference recomputes it every step (_gpc.py:155);
    // here the env step, which forms the same product, hands it to the next step.
    float xf[DF];
#pragma unroll
    for (int k = 0; k < DL; ++k) xf[k] = (q * DL + k < D) ? a.xprev_io[n * D + q * DL + k] : 0.f;
#pragma unroll
    for (int c = 1; c < L; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k) xf[c * DL + k] = __shfl_xor_sync(0xffffffffu, xf[k], c);
    float kxp[M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DL; ++k) s = fmaf(Kc[c][k], xf[k], s);
      kxp[c] = group_allreduce<L>(s);
    }
#pragma unroll
    for (int i = 0; i < DL; ++i) {   // A x_prev = (A - BK) x_prev + B (K x_prev)
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < DF; ++r) s = fmaf(Ar[i][r], xf[r], s);
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], kxp[c], s);
      ax[i] = s;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) uprev[c] = a.uprev_io ? a.uprev_io[n * M + c] : 0.f;
  }

  const bool sampled = live && (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;

  // disturbance for own coordinates at step t
  auto load_w = [&](int t, float w[DL]) {
    if (a.W) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * D + q * DL;
#pragma unroll
      for (int k = 0; k < DL; ++k) w[k] = (q * DL + k < D) ? __ldg(gw + k) : 0.f;
    } else {
      // own coordinates span at most (DL+6)/4 Philox groups
      constexpr int NG = L == 1 ? (DL + 3) / 4 : (DL + 6) / 4;
      const int g0 = (q * DL) / 4;
      float z[NG * 4];
#pragma unroll
      for (int g = 0; g < NG; ++g) normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), g0 + g, 0u, z + 4 * g);
      const int off = q * DL - g0 * 4;  // 0..3
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        float zz = z[k];
#pragma unroll
        for (int o = 1; o < 4; ++o)
          if (k + o < NG * 4) zz = (off == o) ? z[k + o] : zz;
        w[k] = (q * DL + k < D) ? __fmul_rn(p.w_std, zz) : 0.f;  // never contracted into the add below
      }
    }
  };
  float wnext[DL];
  if (p.T > 0) load_w(0, wnext);

  const float* ring = sH;   // = sH + head * DL * TPB: logical index k lives at ring + k * DL * TPB
  auto slot = [&](int k) { return ring + k * (DL * TPB); };

  for (int t = 0; t < p.T; ++t) {
    float w_t[DL];
#pragma unroll
    for (int k = 0; k < DL; ++k) w_t[k] = wnext[k];
    if (t + 1 < p.T) load_w(t + 1, wnext);  // issued early: HBM latency / MUFU work overlaps the step

    // ---- partial sums over own coordinates, reduced together:
    //      red[0..M)            K x                                   _lqr.py:72
    //      red[M..2M)           mw  = sum_i M[i] hist[HH+i]           _gpc.py:171-183
    //      red[2M + h*M + c]    v_h = sum_i M[i] w[h+i], h < H-1      _gpc.py:112-116
    // For short histories all M-window products are formed up front and share one batched
    // reduction; for long ones (H > 4) that would keep M*(H+1) partial sums live, so v_h is formed
    // inside the recursion instead (BATCH_V = false) and only K x and mw are reduced here.
    constexpr bool BATCH_V = H <= 4;
    constexpr int NR = GPC ? (BATCH_V ? M * (H + 1) : 2 * M) : M;
    float red[NR];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DL; ++k) s = fmaf(Kc[c][k], x[k], s);
      red[c] = s;
    }
    auto window_dot = [&](int base, float acc[M]) {   // partial sum_i M[i] w[base+i] over own coordinates
#pragma unroll
      for (int c = 0; c < M; ++c) acc[c] = 0.f;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float* w = slot(base + i);
#pragma unroll
        for (int k = 0; k < DL; ++k) {
          const float wk = w[k * TPB];
#pragma unroll
          for (int c = 0; c < M; ++c) acc[c] = fmaf(Mr[i][c][k], wk, acc[c]);
        }
      }
    };
    if (GPC) {
      // After this step's ring shift every stored noise drops one logical index, so the action
      // window hist[HH..HH+H) of the old ring is window [HH-1..HH+H-1) of the new one -- the final
      // window of policy_loss: mw serves both.  No M-window product depends on the simulated state
      // y nor on the noise appended below (their indices are <= P-2).
      head = (head + 1 == P) ? 0 : head + 1;  // roll(-1)               _gpc.py:156-157
      ring = sH + head * (DL * TPB);
#pragma unroll
      for (int hh = 0; hh < (BATCH_V ? H : 1); ++hh) {   // hh = 0: final/action window; hh >= 1: v_{hh-1}
        float acc[M];
        window_dot(hh == 0 ? HH - 1 : hh - 1, acc);
#pragma unroll
        for (int c = 0; c < M; ++c) red[M + hh * M + c] = acc[c];
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) red[r] = group_allreduce<L>(red[r]);
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = GPC ? red[M + c] - red[c] : -red[c];

    if (GPC) {
      // ---- noise = x - A x_prev - B u ; append                   _gpc.py:155-157
      {
        // newest entry: physical slot head + P - 1 and its mirror head - 1 (the copy at 2P - 1 that
        // head == 0 would need is never read: that entry stays reachable at slot P - 1 until it expires)
        float* wn = sH + (head + P - 1) * (DL * TPB);
        float* wm = wn - P * (DL * TPB);
#pragma unroll
        for (int i = 0; i < DL; ++i) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < M; ++c)
            s = fmaf(Br[i][c], p.noise_uses_prev_action ? uprev[c] : u[c], s);
          const float nz = (x[i] - ax[i]) - s;
          wn[i * TPB] = nz;
          if (head != 0) wm[i * TPB] = nz;
        }
      }
      // ---- policy_loss forward: y <- A y + B(-K y + v_h) + w[h+H]      _gpc.py:118-124
      float y[DL];
#pragma unroll
      for (int k = 0; k < DL; ++k) y[k] = 0.f;
#pragma unroll
      for (int h = 0; h < H - 1; ++h) {
        const float* wh = slot(h + H);
        float v[M];
        if (BATCH_V) {
#pragma unroll
          for (int c = 0; c < M; ++c) v[c] = red[(BATCH_V ? 2 * M + h * M : 0) + c];
        } else {
          window_dot(h, v);
#pragma unroll
          for (int c = 0; c < M; ++c) v[c] = group_allreduce<L>(v[c]);
        }
        float yn[DL];
        if (h == 0) {  // y == 0: A y and K y vanish
#pragma unroll
          for (int i = 0; i < DL; ++i) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], v[c], s);
            yn[i] = s + wh[i * TPB];
          }
        } else {
          float yf[DF];
#pragma unroll
          for (int k = 0; k < DL; ++k) yf[k] = y[k];
#pragma unroll
          for (int c = 1; c < L; ++c)
#pragma unroll
            for (int k = 0; k < DL; ++k) yf[c * DL + k] = __shfl_xor_sync(0xffffffffu, y[k], c);
          // (A - BK) y + B v_h: the -K y term lives in the closed-loop matrix
#pragma unroll
          for (int i = 0; i < DL; ++i) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < DF; ++r) s = fmaf(Ar[i][r], yf[r], s);
#pragma unroll
            for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], v[c], s);
            yn[i] = s + wh[i * TPB];
          }
        }
#pragma unroll
        for (int k = 0; k < DL; ++k) y[k] = yn[k];
      }
      // ---- u_f = -K y + mw ; adjoint seeds (SURVEY A-3.5): du = 2 u_f, lam = 2 y - K' du
      const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
      float du[M];
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float s = 0.f;
        if (H > 1) {
#pragma unroll
          for (int k = 0; k < DL; ++k) s = fmaf(Kc[c][k], y[k], s);
          s = group_allreduce<L>(s);
        }
        du[c] = 2.0f * (red[M + c] - s);
      }
      float lam[DL];
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        float s = 2.0f * y[k];
#pragma unroll
        for (int c = 0; c < M; ++c) s = fmaf(-Kc[c][k], du[c], s);
        lam[k] = s;
      }
      // ---- M[i] -= lr * (dL/du_h) (x) w[.]   (rank-1, in place; the adjoint never reads M)  _gpc.py:163
      {
        float sc[M];
#pragma unroll
        for (int c = 0; c < M; ++c) sc[c] = -lr * du[c];
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float* w = slot(HH - 1 + i);
#pragma unroll
          for (int k = 0; k < DL; ++k) {
            const float wk = w[k * TPB];
#pragma unroll
            for (int c = 0; c < M; ++c) Mr[i][c][k] = fmaf(sc[c], wk, Mr[i][c][k]);
          }
        }
      }
#pragma unroll
      for (int h = H - 2; h >= 0; --h) {
        float sc[M];  // -lr * B' lam
#pragma unroll
        for (int c = 0; c < M; ++c) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < DL; ++i) s = fmaf(Br[i][c], lam[i], s);
          sc[c] = group_allreduce<L>(s);
        }
        if (h > 0) {  // lam <- A' lam - K' lam_u = (A - BK)' lam
          float part[DF];
#pragma unroll
          for (int r = 0; r < DF; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < DL; ++i) s = fmaf(Ar[i][r], lam[i], s);
            part[r] = s;
          }
#pragma unroll
          for (int k = 0; k < DL; ++k) {
            float s = part[k];
#pragma unroll
            for (int c = 1; c < L; ++c) s += __shfl_xor_sync(0xffffffffu, part[c * DL + k], c);
            lam[k] = s;   // (A - BK)' lam

          }
        }
#pragma unroll
        for (int c = 0; c < M; ++c) sc[c] *= -lr;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float* w = slot(h + i);
#pragma unroll
          for (int k = 0; k < DL; ++k) {
            const float wk = w[k * TPB];
#pragma unroll
            for (int c = 0; c < M; ++c) Mr[i][c][k] = fmaf(sc[c], wk, Mr[i][c][k]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < M; ++c) uprev[c] = u[c];
    }

    // ---- realised cost  x'x + u'u                                   _gpc.py:29-40
    float cost;
    {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DL; ++k) s = fmaf(x[k], x[k], s);
      cost = group_allreduce<L>(s);
      float cu = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
      cost += cu;
    }
    csum += (double)cost;
    if (live && q == 0 && a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (GPC && live && t == p.T - 1) {  // the agent's copy of the last state it saw (_gpc.py:167)
#pragma unroll
      for (int k = 0; k < DL; ++k)
        if (q * DL + k < D) a.xprev_io[n * D + q * DL + k] = x[k];
    }
    if (sampled) {
      if (a.x_out) {
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (q * DL + k < D) a.x_out[((int64_t)t * Ns + ns) * D + q * DL + k] = x[k];
      }
      if (a.u_out && q == 0) {
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
      }
    }

    // ---- env: x <- (A x + B u) + w_t                                _lds.py:102-111
    {
      float xf[DF];
#pragma unroll
      for (int k = 0; k < DL; ++k) xf[k] = x[k];
#pragma unroll
      for (int c = 1; c < L; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k) xf[c * DL + k] = __shfl_xor_sync(0xffffffffu, x[k], c);
#pragma unroll
      for (int i = 0; i < DL; ++i) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < DF; ++r) s = fmaf(Ar[i][r], xf[r], s);
        if (GPC) {
          // Ar holds A - BK: x' = (A - BK) x + B mw + w (u = mw - K x), and A x for the next step's
          // noise estimate is (A - BK) x + B (K x)
          float sa = s;
#pragma unroll
          for (int c = 0; c < M; ++c) {
            sa = fmaf(Br[i][c], red[c], sa);
            s = fmaf(Br[i][c], red[GPC ? M + c : c], s);
          }
          ax[i] = sa;
        } else {
          ax[i] = s;
#pragma unroll
          for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], u[c], s);
        }
        x[i] = s + w_t[i];
      }
    }
  }

  // ---- write back
  if (!live) return;
#pragma unroll
  for (int k = 0; k < DL; ++k)
    if (q * DL + k < D) a.x_io[n * D + q * DL + k] = x[k];
  if (q == 0) {
    if (a.cost_sum_out) a.cost_sum_out[n] = csum;
    if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
  }
  if (GPC) {
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (q * DL + k < D) a.M_io[n * (H * M * D) + (i * M + c) * D + q * DL + k] = Mr[i][c][k];
#pragma unroll
    for (int s = 0; s < P; ++s)
#pragma unroll
      for (int k = 0; k < DL; ++k)
        if (q * DL + k < D) a.hist_io[n * (P * D) + s * D + q * DL + k] = slot(s)[k * TPB];
    if (a.uprev_io && q == 0) {
#pragma unroll
      for (int c = 0; c < M; ++c) a.uprev_io[n * M + c] = uprev[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// runtime-dims kernel: one thread per env; every per-env array (A, B, K, x, M, the histories, the BPC ring) is staged
// into an [element][env] workspace copy by tiled transposes before the rollout and back after it, next to the
// scratch vectors, so that the 32 envs of a warp always touch 32 adjacent floats.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int generic_ws_floats_per_env(int d, int m) { return 6 * d + 6 * m; }

// Per-env arrays of the runtime-dims kernel, element i of env n at base[i * stride]: stride = N for the staged
// [element][env] copies (a warp's 32 envs read 32 adjacent floats), 1 for dynamics shared by all envs.
struct Col {
  float* p;
  int64_t s;
  __device__ __forceinline__ float& operator[](int i) const { return p[(int64_t)i * s]; }
  __device__ __forceinline__ Col operator+(int off) const { return Col{p + (int64_t)off * s, s}; }
};

// where each per-env array lives in the workspace, in floats per env after the scratch vectors
struct GenericStage {
  int oA, oB, oK, oX, oM, oH, oXP, oE, total;
};
__host__ __device__ inline GenericStage generic_stage(const DlcLdsParams& p) {
  const bool gpc = p.policy == DLC_POLICY_GPC, bpc = p.policy == DLC_POLICY_BPC;
  const int d = p.d, m = p.m, P = gpc ? p.H + p.HH : p.H, HMD = p.H * m * d;
  GenericStage g;
  int o = generic_ws_floats_per_env(d, m) + 1;
  auto take = [&](int cnt) { const int at = o; o += cnt; return at; };
  g.oA = take(p.shared_dynamics ? 0 : d * d);
  g.oB = take(p.shared_dynamics ? 0 : d * m);
  g.oK = take(p.shared_dynamics ? 0 : m * d);
  g.oX = take(d);
  g.oM = take((gpc || bpc) ? HMD : 0);
  g.oH = take((gpc || bpc) ? P * d : 0);
  g.oXP = take((gpc || bpc) ? d : 0);
  g.oE = take(bpc ? p.H * HMD : 0);
  g.total = o;
  return g;
}

// src[N][S] (env-major, the caller's layout) <-> dst[S][N] (element-major) through a 32 x 33 shared tile
__global__ void __launch_bounds__(256) stage_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int64_t N, int S, int to_cols) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  if (to_cols) {   // read rows of src (contiguous in s), write rows of dst (contiguous in n)
    for (int r = ty; r < 32; r += 8)
      if (n0 + r < N && s0 + tx < S) tile[r][tx] = src[(n0 + r) * S + s0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
      if (s0 + r < S && n0 + tx < N) dst[(int64_t)(s0 + r) * N + n0 + tx] = tile[tx][r];
  } else {         // dst is the caller's [N][S], src the staged [S][N]
    for (int r = ty; r < 32; r += 8)
      if (s0 + r < S && n0 + tx < N) tile[r][tx] = src[(int64_t)(s0 + r) * N + n0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
      if (n0 + r < N && s0 + tx < S) dst[(n0 + r) * S + s0 + tx] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(128) lds_generic_kernel(const LdsArgs a) {
  const DlcLdsParams& p = a.p;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int d = p.d, m = p.m, H = p.H;
  const bool gpc = p.policy == DLC_POLICY_GPC, bpc = p.policy == DLC_POLICY_BPC;
  const bool open_loop = p.policy == DLC_POLICY_OPEN, agent_only = p.mode == DLC_MODE_AGENT_ONLY;
  const int P = gpc ? p.H + p.HH : p.H;  // history length (_gpc.py:98 / _bpc.py:82)
  // every per-env array is read from its staged [element][env] copy (dlc_lds_rollout_f32 transposes the caller's
  // env-major buffers in and out): with the env-major originals each lane of a warp walked its own row and every
  // load or store touched 32 cache lines
  const int64_t N = p.N;
  const GenericStage sg = generic_stage(p);
  float* ws = a.ws + n;
  auto staged = [&](int off) { return Col{ws + (int64_t)off * N, N}; };
  const Col A = p.shared_dynamics ? Col{const_cast<float*>(a.A), 1} : staged(sg.oA);
  const Col B = p.shared_dynamics ? Col{const_cast<float*>(a.B), 1} : staged(sg.oB);
  const Col K = p.shared_dynamics ? Col{const_cast<float*>(a.K), 1} : staged(sg.oK);   // unused for DLC_POLICY_OPEN
  const Col x = staged(sg.oX);
  const Col Mv = staged(sg.oM), hist = staged(sg.oH), xprev = staged(sg.oXP), eps = staged(sg.oE);
  const int HMD = H * m * d;
  // scratch vectors, element e at ws[(off + e) * N + n]
  auto V = [&](int off, int e) -> float& { return ws[(int64_t)(off + e) * N]; };
  const int oU = 0, oUN = m, oV = 2 * m, oMW = 3 * m, oLU = 4 * m, oDU = 5 * m;
  const int oY = 6 * m, oYN = oY + d, oLAM = oYN + d, oLN = oLAM + d, oW = oLN + d, oNZ = oW + d;
  for (int c = 0; c < m; ++c) V(oUN, c) = a.uprev_io ? a.uprev_io[n * m + c] : 0.f;

  const bool sampled = (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;
  float prev_cost = bpc ? V(oNZ, 0) : 0.f;  // staged by bpc_cost_stage_kernel (see below)

  auto win_dot = [&](int start, int off_out) {  // out[c] = sum_i sum_j M[i][c][j] hist[start+i][j]
    for (int c = 0; c < m; ++c) {
      float s = 0.f;
      for (int i = 0; i < H; ++i)
        for (int j = 0; j < d; ++j) s = fmaf(Mv[(i * m + c) * d + j], hist[(start + i) * d + j], s);
      V(off_out, c) = s;
    }
  };
  auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };

  for (int t = 0; t < p.T; ++t) {
    // ---- action
    if (gpc || bpc) win_dot(P - H, oMW);
    for (int c = 0; c < m; ++c) {
      float s = 0.f;
      if (open_loop) {
        V(oU, c) = a.U_in[((int64_t)t * p.N + n) * m + c];
        continue;
      }
      for (int j = 0; j < d; ++j) s = fmaf(K[c * d + j], x[j], s);
      V(oU, c) = (gpc || bpc) ? V(oMW, c) - s : -s;
    }
    if (gpc || bpc) {
      // ---- noise, ring shift                                    _gpc.py:155-157 / _bpc.py:124-126
      for (int i = 0; i < d; ++i) {
        float s = 0.f, sb = 0.f;
        for (int j = 0; j < d; ++j) s = fmaf(A[i * d + j], xprev[j], s);
        for (int c = 0; c < m; ++c)
          sb = fmaf(B[i * m + c], p.noise_uses_prev_action ? V(oUN, c) : V(oU, c), sb);
        V(oW, i) = (x[i] - s) - sb;
      }
      for (int k = 0; k < (P - 1) * d; ++k) hist[k] = hist[k + d];
      for (int i = 0; i < d; ++i) hist[(P - 1) * d + i] = V(oW, i);
    }
    if (gpc) {
      const int HH = p.HH;
      // ---- policy_loss forward (windows clamp like lax.dynamic_slice)   _gpc.py:109-125
      for (int i = 0; i < d; ++i) V(oY, i) = 0.f;
      for (int h = 0; h < H - 1; ++h) {
        win_dot(clampi(h, 0, P - H), oV);
        for (int c = 0; c < m; ++c) {
          float s = 0.f;
          for (int j = 0; j < d; ++j) s = fmaf(K[c * d + j], V(oY, j), s);
          V(oV, c) -= s;
        }
        const int idx = (h + H < P - 1) ? h + H : P - 1;
        for (int i = 0; i < d; ++i) {
          float s = 0.f;
          for (int j = 0; j < d; ++j) s = fmaf(A[i * d + j], V(oY, j), s);
          for (int c = 0; c < m; ++c) s = fmaf(B[i * m + c], V(oV, c), s);
          V(oYN, i) = s + hist[idx * d + i];
        }
        for (int i = 0; i < d; ++i) V(oY, i) = V(oYN, i);
      }
      const int fstart = clampi(HH - 1, 0, P - H);
      win_dot(fstart, oV);
      const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
      for (int c = 0; c < m; ++c) {
        float s = 0.f;
        for (int j = 0; j < d; ++j) s = fmaf(K[c * d + j], V(oY, j), s);
        V(oDU, c) = (a.cost_r ? 2.0f * __ldg(a.cost_r + c) : 2.0f) * (V(oV, c) - s);  // dL/du_f = 2 r (.) u_f
      }
      for (int i = 0; i < d; ++i) {
        float s = (a.cost_q ? 2.0f * __ldg(a.cost_q + i) : 2.0f) * V(oY, i);             // 2 q (.) y_f
        for (int c = 0; c < m; ++c) s = fmaf(-K[c * d + i], V(oDU, c), s);
        V(oLAM, i) = s;
      }
      // adjoint, M updated in place (the backward pass never reads M)
      for (int i = 0; i < H; ++i)
        for (int c = 0; c < m; ++c)
          for (int j = 0; j < d; ++j)
            Mv[(i * m + c) * d + j] =
                fmaf(-lr * V(oDU, c), hist[(fstart + i) * d + j], Mv[(i * m + c) * d + j]);
      for (int h = H - 2; h >= 0; --h) {
        for (int c = 0; c < m; ++c) {
          float s = 0.f;
          for (int i = 0; i < d; ++i) s = fmaf(B[i * m + c], V(oLAM, i), s);
          V(oLU, c) = s;
        }
        const int st = clampi(h, 0, P - H);
        for (int i = 0; i < H; ++i)
          for (int c = 0; c < m; ++c)
            for (int j = 0; j < d; ++j)
              Mv[(i * m + c) * d + j] =
                  fmaf(-lr * V(oLU, c), hist[(st + i) * d + j], Mv[(i * m + c) * d + j]);
        for (int j = 0; j < d; ++j) {
          float s = 0.f;
          for (int i = 0; i < d; ++i) s = fmaf(A[i * d + j], V(oLAM, i), s);
          for (int c = 0; c < m; ++c) s = fmaf(-K[c * d + j], V(oLU, c), s);
          V(oLN, j) = s;
        }
        for (int j = 0; j < d; ++j) V(oLAM, j) = V(oLN, j);
      }
    }
    if (bpc) {
      // ---- bandit update                                        _bpc.py:128-139
      float lr = p.lr_scale;
      if (p.decay) lr = p.lr_scale * (1.0f / (powf((float)(p.t0 + t), 0.75f) + 1.0f));
      for (int k = 0; k < HMD; ++k) {
        float s = 0.f;
        for (int q = 0; q < H; ++q) s += eps[q * HMD + k];
        Mv[k] -= lr * (prev_cost * s);
      }
      for (int k = 0; k < (H - 1) * HMD; ++k) eps[k] = eps[k + HMD];
      const Col en = eps + (H - 1) * HMD;
      if (a.eps_new) {
        const float* src = a.eps_new + ((int64_t)t * p.N + n) * HMD;
        for (int k = 0; k < HMD; ++k) en[k] = src[k];
      } else {  // unit-norm Gaussian direction from the Philox stream 1 (_bpc.py:26-30)
        float ss = 0.f;
        for (int g = 0; g < (HMD + 3) / 4; ++g) {
          float z[4];
          normal4(p.seed, env_id, (uint32_t)(p.t0 + t), g, 1u, z);
          for (int k = 0; k < 4 && g * 4 + k < HMD; ++k) {
            en[g * 4 + k] = z[k];
            ss = fmaf(z[k], z[k], ss);
          }
        }
        const float inv = 1.0f / sqrtf(ss);
        for (int k = 0; k < HMD; ++k) en[k] *= inv;
      }
      for (int k = 0; k < HMD; ++k) Mv[k] = fmaf(p.bpc_delta, en[k], Mv[k]);
    }
    if (gpc || bpc) {
      for (int i = 0; i < d; ++i) xprev[i] = x[i];
      for (int c = 0; c < m; ++c) V(oUN, c) = V(oU, c);
    }
    // ---- realised cost
    float cost = 0.f, cu = 0.f;
    for (int i = 0; i < d; ++i) cost = fmaf(x[i], x[i], cost);
    for (int c = 0; c < m; ++c) cu = fmaf(V(oU, c), V(oU, c), cu);
    cost += cu;
    prev_cost = cost;
    csum += (double)cost;
    if (a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (sampled) {
      if (a.x_out)
        for (int i = 0; i < d; ++i) a.x_out[((int64_t)t * Ns + ns) * d + i] = x[i];
      if (a.u_out)
        for (int c = 0; c < m; ++c) a.u_out[((int64_t)t * Ns + ns) * m + c] = V(oU, c);
    }
    // ---- env step
    if (agent_only) continue;
    if (a.W) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * d;
      for (int i = 0; i < d; ++i) V(oW, i) = gw[i];
    } else {
      for (int g = 0; g < (d + 3) / 4; ++g) {
        float z[4];
        normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), g, 0u, z);
        for (int k = 0; k < 4 && g * 4 + k < d; ++k) V(oW, g * 4 + k) = __fmul_rn(p.w_std, z[k]);
      }
    }
    for (int i = 0; i < d; ++i) {
      float s = 0.f;
      for (int j = 0; j < d; ++j) s = fmaf(A[i * d + j], x[j], s);
      for (int c = 0; c < m; ++c) s = fmaf(B[i * m + c], V(oU, c), s);
      V(oYN, i) = s + V(oW, i);
    }
    for (int i = 0; i < d; ++i) x[i] = V(oYN, i);
  }
  if (a.uprev_io && (gpc || bpc))
    for (int c = 0; c < m; ++c) a.uprev_io[n * m + c] = V(oUN, c);
  if (bpc) V(oNZ, 0) = prev_cost;
  if (a.cost_sum_out) a.cost_sum_out[n] = csum;
  if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
}

// stage / unstage the BPC carried cost through workspace slot oNZ (keeps LdsArgs small)
__global__ void bpc_cost_stage_kernel(float* ws, float* cost_io, int64_t N, int off, int to_ws) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (to_ws) ws[(int64_t)off * N + n] = cost_io ? cost_io[n] : 0.f;
  else if (cost_io) cost_io[n] = ws[(int64_t)off * N + n];
}

// One thread per (t, env) row: all ceil(d / 4) Philox blocks of the row, stored as one contiguous run (rows of a warp are
// adjacent in memory).  blockIdx.y walks the steps, so no index is ever divided (the first version spent more time on
// three 64-bit divisions per block than on the generator: 64 ms for the target's 42 GB, now at the speed of the ALUs).
template <int DC>   // compile-time d (0 = runtime)
__global__ void __launch_bounds__(256) gaussian_disturbance_kernel(float* W, int64_t N, int T, int d_rt, uint64_t seed,
                                                                   int64_t env_offset, int t0, float std) {
  const int d = DC ? DC : d_rt;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const uint32_t env = (uint32_t)(env_offset + n);
  for (int t = blockIdx.y; t < T; t += gridDim.y) {
    float* dst = W + ((int64_t)t * N + n) * d;
    if (DC) {
      float row[DC ? ((DC + 3) / 4) * 4 : 4];
#pragma unroll
      for (int g = 0; g < (DC + 3) / 4; ++g) normal4(seed, env, (uint32_t)(t0 + t), g, 0u, row + 4 * g);
      if (DC % 2 == 0) {   // rows start on 8-byte boundaries
#pragma unroll
        for (int k = 0; k < DC; k += 2)
          *reinterpret_cast<float2*>(dst + k) = make_float2(__fmul_rn(std, row[k]), __fmul_rn(std, row[k + 1]));
      } else {
#pragma unroll
        for (int k = 0; k < DC; ++k) dst[k] = __fmul_rn(std, row[k]);
      }
    } else {
      for (int g = 0; g * 4 < d; ++g) {
        float z[4];
        normal4(seed, env, (uint32_t)(t0 + t), g, 0u, z);
        for (int k = 0; k < 4 && g * 4 + k < d; ++k) dst[g * 4 + k] = __fmul_rn(std, z[k]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int D, int M, int H, int HH, int POLICY, int L, int TPB, int MINB>
static int32_t launch_split(const LdsArgs& a, cudaStream_t st) {
  using Lay = SplitLayout<D, M, H, HH, POLICY, L, TPB>;
  auto kern = lds_split_kernel<D, M, H, HH, POLICY, L, TPB, MINB>;
  static_assert((Lay::kSmemBytes + 1024) * MINB <= 228 * 1024, "ring does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Lay::kSmemBytes);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_split)");
  const int64_t threads = a.p.N * L;
  const unsigned grid = (unsigned)((threads + TPB - 1) / TPB);
  kern<<<grid, TPB, Lay::kSmemBytes, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_split_kernel launch");
}

// Launch shapes of the GPC kernel: (threads per CTA, CTAs per SM) -> warps per SM.  Rollouts are
// equal-length jobs, so the grid runs in ceil(warps / (SMs * warps_per_SM)) rounds and a partially
// filled last round is pure loss (65,536 envs at L = 2 are 27.7 warps per SM: 12 warps/SM leaves the
// third round 31 % full, 10 warps/SM fills three rounds to 92 %, 14 warps/SM two rounds to 99 %).
// The launcher picks the shape with the best round efficiency x measured per-shape throughput.
struct LaunchShape { int tpb, minb; float rel_throughput; };
// rel_throughput: steady-state throughput relative to shape 3, measured on B200 (profiles/r01_*):
// register spills cost far more than occupancy buys, so the no-spill shapes win.
static constexpr LaunchShape kShapes[] = {
    {128, 3, 0.48f},  // kernel_variant 2: 12 warps/SM, 168 regs (spills)
    {64, 5, 0.82f},   // kernel_variant 3: 10 warps/SM, 200 regs (few spills)
    {64, 7, 0.35f},   // kernel_variant 4: 14 warps/SM, 144 regs (spills)
    {128, 2, 1.00f},  // kernel_variant 5:  8 warps/SM, 255 regs (no spills)
    {32, 7, 0.93f},   // kernel_variant 6:  7 warps/SM, 255 regs, one-warp CTAs
    {32, 8, 1.00f},   // kernel_variant 7:  8 warps/SM, 255 regs, one-warp CTAs
};
static constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);

static int pick_shape(const LdsArgs& a, int L, int sms) {
  const int v = a.p.kernel_variant;
  if (v >= 2 && v < 2 + kNumShapes) return v - 2;
  const double warps = (double)((a.p.N * L + 31) / 32);
  int best = 3;
  double best_score = -1.0;
  for (int i = 0; i < kNumShapes; ++i) {
    const double cap = (double)sms * (kShapes[i].tpb / 32) * kShapes[i].minb;
    const double rounds = ceil(warps / cap);
    const double score = warps / (rounds * cap) * kShapes[i].rel_throughput;
    if (score > best_score + 1e-9) { best_score = score; best = i; }
  }
  return best;
}

template <int D, int M, int H, int HH, int L>
static int32_t launch_gpc(const LdsArgs& a, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  switch (pick_shape(a, L, sms)) {
    case 0: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 128, 3>(a, st);
    case 1: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 64, 5>(a, st);
    case 2: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 64, 7>(a, st);
    case 4: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 32, 7>(a, st);
    case 5: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 32, 8>(a, st);
    default: return launch_split<D, M, H, HH, DLC_POLICY_GPC, L, 128, 2>(a, st);
  }
}

// (d, m, H, HH, lanes per env)
#define DLC_SPLIT_SHAPES(X) \
  X(10, 3, 3, 10, 2)        \
  X(10, 3, 3, 2, 2)         \
  X(10, 3, 10, 10, 4)       \
  X(4, 1, 3, 2, 1)          \
  X(4, 2, 5, 7, 1)

// does a compile-time-shape kernel take this call?  (mirrors try_split without launching)
static bool split_supported(const DlcLdsParams& p) {
  if (p.kernel_variant == 1 || p.policy == DLC_POLICY_BPC || p.policy == DLC_POLICY_OPEN ||
      p.mode != DLC_MODE_CLOSED_LOOP)
    return false;
#define X(D_, M_, H_, HH_, L_) \
  if (p.d == D_ && p.m == M_ && (p.policy == DLC_POLICY_LQR || (p.H == H_ && p.HH == HH_))) return true;
  DLC_SPLIT_SHAPES(X)
#undef X
  return false;
}

static bool try_split(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  if (p.kernel_variant == 1 || p.policy == DLC_POLICY_BPC || p.policy == DLC_POLICY_OPEN ||
      p.mode != DLC_MODE_CLOSED_LOOP)
    return false;
#define X(D_, M_, H_, HH_, L_)                                                        \
  if (p.d == D_ && p.m == M_) {                                                       \
    if (p.policy == DLC_POLICY_LQR) {                                                 \
      *rc = launch_split<D_, M_, 1, 1, DLC_POLICY_LQR, L_, 128, 4>(a, st);            \
      return true;                                                                    \
    }                                                                                 \
    if (p.H == H_ && p.HH == HH_) {                                                   \
      /* packed-FFMA2 kernel: default for short histories (6-10 % faster on B200), forced by 8 / 9 */ \
      if (((p.kernel_variant == 0 && (H_ <= 4 || (H_ % 2 == 0 && HH_ == H_))) || (p.kernel_variant >= 8 && p.kernel_variant <= 12)) && \
          launch_gpc_f2(a, st, rc))                                                   \
        return true;                                                                  \
      *rc = launch_gpc<D_, M_, H_, HH_, L_>(a, st);                                   \
      return true;                                                                    \
    }                                                                                 \
  }
  DLC_SPLIT_SHAPES(X)
#undef X
  return false;
}

}  // namespace dlc

using namespace dlc;

extern "C" size_t dlc_lds_workspace_bytes(const DlcLdsParams* p) {
  if (!p || p->N <= 0 || p->d <= 0 || p->m <= 0) return 0;
  if (split_supported(*p) || bpc_supported(*p) || lqr_reg_supported(*p)) return 0;   // the compile-time-shape kernels need no workspace
  // runtime-dims kernel: scratch vectors plus an [element][env] copy of every per-env array
  return (size_t)generic_stage(*p).total * (size_t)p->N * sizeof(float);
}

extern "C" int32_t dlc_gaussian_disturbance_f32(float* W, int64_t N, int32_t T, int32_t d,
                                                uint64_t seed, int64_t env_offset, int32_t t0,
                                                float std, void* stream) {
  if (!W || N < 0 || T < 0 || d <= 0) return fail(DLC_ERR_ARG, "dlc_gaussian_disturbance_f32: bad argument");
  if (N == 0 || T == 0) return DLC_OK;
  const dim3 grid((unsigned)((N + 255) / 256), (unsigned)(T < 1024 ? T : 1024));
  cudaStream_t st = (cudaStream_t)stream;
  const bool al8 = (reinterpret_cast<uintptr_t>(W) & 7u) == 0;
  if (d == 10 && al8) gaussian_disturbance_kernel<10><<<grid, 256, 0, st>>>(W, N, T, d, seed, env_offset, t0, std);
  else gaussian_disturbance_kernel<0><<<grid, 256, 0, st>>>(W, N, T, d, seed, env_offset, t0, std);
  return cuda_status(cudaGetLastError(), "gaussian_disturbance_kernel launch");
}

static int32_t lds_rollout_impl(const DlcLdsParams* p, const float* A, const float* B,
                                const float* K, const float* W, const float* u_in, float* x_io,
                                float* M_io,
                                float* hist_io, float* xprev_io, float* uprev_io,
                                float* bpc_eps_io, const float* bpc_eps_new,
                                float* bpc_cost_io, float* cost_out, double* cost_sum_out,
                                float* x_out, float* u_out, int32_t* status_out,
                                void* workspace, size_t workspace_bytes, void* stream,
                                const float* cost_q, const float* cost_r) {
  if (!p) return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: params is NULL");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: negative N or T");
  if (p->d <= 0 || p->m <= 0) return fail(DLC_ERR_SHAPE, "dlc_lds_rollout_f32: d and m must be positive");
  if (p->policy < DLC_POLICY_LQR || p->policy > DLC_POLICY_OPEN)
    return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: unknown policy");
  if (p->mode != DLC_MODE_CLOSED_LOOP && p->mode != DLC_MODE_AGENT_ONLY)
    return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: unknown mode");
  if (p->policy == DLC_POLICY_OPEN && (!u_in || p->mode != DLC_MODE_CLOSED_LOOP))
    return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: DLC_POLICY_OPEN needs u_in and closed-loop mode");
  if (p->traj_stride < 1) return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: traj_stride must be >= 1");
  if (p->N == 0) return DLC_OK;  // empty batch: nothing to do
  if (!A || !B || !x_io || (!K && p->policy != DLC_POLICY_OPEN))
    return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: A, B, K and x_io are required");
  const bool learn = p->policy == DLC_POLICY_GPC || p->policy == DLC_POLICY_BPC;
  if (learn) {
    if (p->H < 1 || (p->policy == DLC_POLICY_GPC && p->HH < 1))
      return fail(DLC_ERR_SHAPE, "dlc_lds_rollout_f32: H >= 1 and HH >= 1 required");
    if (!M_io || !hist_io || !xprev_io)
      return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: M_io, hist_io and xprev_io are required for GPC/BPC");
    if (p->noise_uses_prev_action && !uprev_io)
      return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: uprev_io is required when noise_uses_prev_action is set");
    if (p->policy == DLC_POLICY_BPC && !bpc_eps_io)
      return fail(DLC_ERR_ARG, "dlc_lds_rollout_f32: bpc_eps_io is required for BPC");
  }
  if ((uint64_t)(p->env_offset + p->N) > 0xFFFFFFFFull)
    return fail(DLC_ERR_SHAPE, "dlc_lds_rollout_f32: env_offset + N exceeds the 32-bit stream counter");
  LdsArgs a;
  a.p = *p;
  a.A = A; a.B = B; a.K = K; a.W = W; a.U_in = u_in;
  a.x_io = x_io; a.M_io = M_io; a.hist_io = hist_io; a.xprev_io = xprev_io; a.uprev_io = uprev_io;
  a.eps_io = bpc_eps_io; a.eps_new = bpc_eps_new;
  a.cost_out = cost_out; a.cost_sum_out = cost_sum_out; a.x_out = x_out; a.u_out = u_out;
  a.status_out = status_out; a.ws = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t rc = DLC_OK;
  const bool weighted = cost_q || cost_r;
  if (weighted) {
    // a caller's cost_fn only enters GPC's update (_gpc.py:125,128): the lane-group kernel (closed loop, any d, m <= 64)
    // and the runtime-dims kernel (everything else, incl. single Agent.__call__ steps) read the weights
    if (p->policy != DLC_POLICY_GPC)
      return fail(DLC_ERR_ARG, "dlc_lds_rollout_cost_f32: cost weights apply to GPC only");
    a.cost_q = cost_q; a.cost_r = cost_r;
  }
  if (!weighted) {
    if (launch_bpc(a, bpc_cost_io, st, &rc)) return rc;
    if (launch_lqr_reg(a, st, &rc)) return rc;
    if (launch_gpc_quad(a, st, &rc)) return rc;
    if (try_split(a, st, &rc)) return rc;
  }
  if (launch_gpc_group(a, st, &rc)) return rc;
  DlcLdsParams sized = *p;
  if (weighted) sized.kernel_variant = 1;   // (a weighted call never runs the compile-time-shape kernels)
  if (!workspace || workspace_bytes < dlc_lds_workspace_bytes(&sized))
    return fail(DLC_ERR_WORKSPACE, "dlc_lds_rollout_f32: this shape runs the runtime-dims kernel and needs "
                                   "dlc_lds_workspace_bytes() of workspace");
  const unsigned grid = (unsigned)((p->N + 127) / 128);
  const int oNZ = 6 * p->m + 5 * p->d;
  const GenericStage sg = generic_stage(*p);
  const int P = p->policy == DLC_POLICY_GPC ? p->H + p->HH : p->H, HMD = p->H * p->m * p->d;
  auto stage = [&](const float* env_major, int off, int S, int to_cols) {   // [N][S] <-> ws[(off + s) * N + n]
    if (!env_major || S <= 0) return;
    const dim3 g((unsigned)((p->N + 31) / 32), (unsigned)((S + 31) / 32));
    float* cols = a.ws + (size_t)off * p->N;
    if (to_cols) stage_transpose_kernel<<<g, 256, 0, st>>>(env_major, cols, p->N, S, 1);
    else stage_transpose_kernel<<<g, 256, 0, st>>>(cols, const_cast<float*>(env_major), p->N, S, 0);
  };
  for (int dir = 1; dir >= 0; --dir) {   // 1: stage in, run; 0: stage the updated state back out
    if (dir == 1 && !p->shared_dynamics) {
      stage(A, sg.oA, p->d * p->d, 1);
      stage(B, sg.oB, p->d * p->m, 1);
      stage(K, sg.oK, p->m * p->d, 1);
    }
    stage(x_io, sg.oX, p->d, dir);
    if (learn) {
      stage(M_io, sg.oM, HMD, dir);
      stage(hist_io, sg.oH, P * p->d, dir);
      stage(xprev_io, sg.oXP, p->d, dir);
      if (p->policy == DLC_POLICY_BPC) stage(bpc_eps_io, sg.oE, p->H * HMD, dir);
    }
    if (dir == 1) {
      if (p->policy == DLC_POLICY_BPC)
        bpc_cost_stage_kernel<<<grid, 128, 0, st>>>(a.ws, bpc_cost_io, p->N, oNZ, 1);
      lds_generic_kernel<<<grid, 128, 0, st>>>(a);
      if (p->policy == DLC_POLICY_BPC)
        bpc_cost_stage_kernel<<<grid, 128, 0, st>>>(a.ws, bpc_cost_io, p->N, oNZ, 0);
    }
  }
  return cuda_status(cudaGetLastError(), "lds_generic_kernel launch");
}

extern "C" int32_t dlc_lds_rollout_f32(const DlcLdsParams* p, const float* A, const float* B,
                                       const float* K, const float* W, const float* u_in, float* x_io,
                                       float* M_io,
                                       float* hist_io, float* xprev_io, float* uprev_io,
                                       float* bpc_eps_io, const float* bpc_eps_new,
                                       float* bpc_cost_io, float* cost_out, double* cost_sum_out,
                                       float* x_out, float* u_out, int32_t* status_out,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  return lds_rollout_impl(p, A, B, K, W, u_in, x_io, M_io, hist_io, xprev_io, uprev_io, bpc_eps_io, bpc_eps_new,
                          bpc_cost_io, cost_out, cost_sum_out, x_out, u_out, status_out, workspace, workspace_bytes,
                          stream, nullptr, nullptr);
}

extern "C" int32_t dlc_lds_rollout_cost_f32(const DlcLdsParams* p, const float* A, const float* B, const float* K,
                                            const float* W, float* x_io, float* M_io, float* hist_io,
                                            float* xprev_io, float* uprev_io, const float* cost_q,
                                            const float* cost_r, float* cost_out, double* cost_sum_out,
                                            float* x_out, float* u_out, int32_t* status_out, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  return lds_rollout_impl(p, A, B, K, W, nullptr, x_io, M_io, hist_io, xprev_io, uprev_io, nullptr, nullptr, nullptr,
                          cost_out, cost_sum_out, x_out, u_out, status_out, workspace, workspace_bytes, stream, cost_q,
                          cost_r);
}
