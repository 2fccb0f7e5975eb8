// Packed-FP32 (FFMA2) version of the lane-split LDS + GPC rollout kernel (see lds_rollout.cu for the
// mapping: L adjacent lanes own one env and split it by state coordinate).
//
// Blackwell's `fma.rn.f32x2` performs two FMAs per issued instruction on 64-bit register pairs and
// accepts a scalar register broadcast as one operand.  Every per-coordinate quantity a lane owns
// (its DL = ceil(D/L) coordinates) is therefore kept as NP = DL/2 pairs (+ one scalar tail when DL is
// odd), and every mat-vec is arranged so that the pair index is the lane's own coordinate index k:
//     sum_k M[i][c][k] w[k]      pairs over k, two partial sums per accumulator     (window products)
//     M[i][c][k] += s_c w[k]     pairs over k, s_c broadcast                        (rank-1 updates)
//     sum_k K[c][k] y[k], sum_k B[k][c] lam[k], sum_k A[k][r] lam[k]   pairs over k (reductions)
//     sum_c B[k][c] v_c, sum_c K[c][k] du_c, sum_r A[k][r] y_r        rows k paired, operand broadcast
// The arithmetic, its order within each dot product pair-lane, and all data layouts seen by the
// caller are those of lds_split_kernel; only the register packing and the shared-memory ring layout
// (pairs as float2, tails separate) differ.  Issued FMA instructions per warp-step drop from ~610 to
// ~370, and each accumulation chain is half as long.
#include "lds_args.cuh"

namespace dlc {

typedef unsigned long long f2;  // two packed floats in a 64-bit register pair

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 dup(float a) { return pk(a, a); }  // folds into a .F32 broadcast operand
__device__ __forceinline__ float lo(f2 p) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
  return a;
}
__device__ __forceinline__ float hi(f2 p) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
  return b;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int L>
__device__ __forceinline__ float group_allreduce2(float v) {
#pragma unroll
  for (int o = 1; o < L; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int D, int H, int HH, int L, int TPB>
struct F2Layout {
  static constexpr int DL = (D + L - 1) / L, NP = DL / 2, TAIL = DL & 1, P = H + HH, SLOTS = 2 * P - 1;
  static constexpr size_t kPairBytes = (size_t)SLOTS * NP * TPB * sizeof(float2);
  static constexpr size_t kSmemBytes = kPairBytes + (size_t)SLOTS * TAIL * TPB * sizeof(float);
};

template <int D, int M, int H, int HH, int L, int TPB, int MINB>
__global__ void __launch_bounds__(TPB)
    __maxnreg__((65536 / (TPB * MINB)) / 8 * 8 > 255 ? 255 : (65536 / (TPB * MINB)) / 8 * 8)
        lds_f2_kernel(const LdsArgs a) {
  using Lay = F2Layout<D, H, HH, L, TPB>;
  constexpr int DL = Lay::DL, NP = Lay::NP, TAIL = Lay::TAIL, DF = DL * L, P = Lay::P;
  constexpr int NPA = NP > 0 ? NP : 1;
  static_assert(HH >= H - 1, "templated path needs HH >= H-1 (no index clamping)");
  extern __shared__ __align__(16) unsigned char smraw[];
  const DlcLdsParams& p = a.p;
  const int tid = threadIdx.x;
  const int q = tid % L;
  const int64_t n_raw = ((int64_t)blockIdx.x * TPB + tid) / L;
  const bool live = n_raw < p.N;
  const int64_t n = live ? n_raw : p.N - 1;
  // ring: pair (slot, kk) at sP[(slot*NP + kk) * TPB], tail (slot) at sT[slot * TPB]
  float2* const sP = reinterpret_cast<float2*>(smraw) + tid;
  float* const sT = reinterpret_cast<float*>(smraw + Lay::kPairBytes) + tid;

  auto coord = [&](int k) { return q * DL + k; };
  auto ld = [&](const float* g, int k) { return coord(k) < D ? g[coord(k)] : 0.f; };

  // ---- parameters -> packed registers
  const int64_t np = p.shared_dynamics ? 0 : n;
  const float* gA = a.A + np * (D * D);
  const float* gB = a.B + np * (D * M);
  const float* gK = a.K + np * (M * D);
  f2 Ar2[NPA][DF], Br2[NPA][M], Kc2[M][NPA];
  float Art[DF], Brt[M], Kct[M];
  // The kernel keeps the closed-loop matrix A - B K instead of A: the counterfactual recursion
  // y <- A y + B(-K y + v) + w is (A - BK) y + B v + w and its adjoint is (A - BK)' lam, which removes one K
  // mat-vec and its cross-lane reduction per recursion step; the env step recovers A x = (A - BK) x + B (K x).
  auto Aelem = [&](int k, int r) {  // row = own coordinate k, column = lane-relative index r
    const int gi = coord(k), gj = (q ^ (r / DL)) * DL + (r % DL);
    if (!(gi < D && gj < D)) return 0.f;
    float s = __ldg(gA + gi * D + gj);
#pragma unroll
    for (int c = 0; c < M; ++c) s = fmaf(-__ldg(gB + gi * M + c), __ldg(gK + c * D + gj), s);
    return s;
  };
#pragma unroll
  for (int kk = 0; kk < NP; ++kk) {
#pragma unroll
    for (int r = 0; r < DF; ++r) Ar2[kk][r] = pk(Aelem(2 * kk, r), Aelem(2 * kk + 1, r));
#pragma unroll
    for (int c = 0; c < M; ++c) {
      Br2[kk][c] = pk(coord(2 * kk) < D ? __ldg(gB + coord(2 * kk) * M + c) : 0.f,
                      coord(2 * kk + 1) < D ? __ldg(gB + coord(2 * kk + 1) * M + c) : 0.f);
      Kc2[c][kk] = pk(coord(2 * kk) < D ? __ldg(gK + c * D + coord(2 * kk)) : 0.f,
                      coord(2 * kk + 1) < D ? __ldg(gK + c * D + coord(2 * kk + 1)) : 0.f);
    }
  }
  if (TAIL) {
#pragma unroll
    for (int r = 0; r < DF; ++r) Art[r] = Aelem(DL - 1, r);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      Brt[c] = coord(DL - 1) < D ? __ldg(gB + coord(DL - 1) * M + c) : 0.f;
      Kct[c] = coord(DL - 1) < D ? __ldg(gK + c * D + coord(DL - 1)) : 0.f;
    }
  }
  f2 Mr2[H][M][NPA];
  float Mrt[H][M];
  {
    const float* gM = a.M_io + n * (H * M * D);
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int c = 0; c < M; ++c) {
#pragma unroll
        for (int kk = 0; kk < NP; ++kk)
          Mr2[i][c][kk] = pk(ld(gM + (i * M + c) * D, 2 * kk), ld(gM + (i * M + c) * D, 2 * kk + 1));
        if (TAIL) Mrt[i][c] = ld(gM + (i * M + c) * D, DL - 1);
      }
  }
  {
    const float* gh = a.hist_io + n * (P * D);
#pragma unroll
    for (int s = 0; s < P; ++s) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const float2 v = make_float2(ld(gh + s * D, 2 * kk), ld(gh + s * D, 2 * kk + 1));
        sP[(s * NP + kk) * TPB] = v;
        if (s + P < 2 * P - 1) sP[((s + P) * NP + kk) * TPB] = v;
      }
      if (TAIL) {
        const float v = ld(gh + s * D, DL - 1);
        sT[s * TPB] = v;
        if (s + P < 2 * P - 1) sT[(s + P) * TPB] = v;
      }
    }
  }

  // ---- helpers on vectors split as pairs + tail
  struct Vk { f2 p[NPA]; float t; };
  auto gather = [&](const Vk& v, float out[DF]) {  // lane-relative all-gather
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) { out[2 * kk] = lo(v.p[kk]); out[2 * kk + 1] = hi(v.p[kk]); }
    if (TAIL) out[DL - 1] = v.t;
#pragma unroll
    for (int c = 1; c < L; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k) out[c * DL + k] = __shfl_xor_sync(0xffffffffu, out[k], c);
  };
  auto Amul = [&](const float vf[DF], Vk& out) {  // out_k = sum_r A[k][r] vf[r]
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) {
      f2 s = mul2(Ar2[kk][0], dup(vf[0]));
#pragma unroll
      for (int r = 1; r < DF; ++r) s = fma2(Ar2[kk][r], dup(vf[r]), s);
      out.p[kk] = s;
    }
    if (TAIL) {
      float s = Art[0] * vf[0];
#pragma unroll
      for (int r = 1; r < DF; ++r) s = fmaf(Art[r], vf[r], s);
      out.t = s;
    }
  };
  auto Bacc = [&](const float v[M], Vk& acc) {  // acc_k += sum_c B[k][c] v_c
#pragma unroll
    for (int c = 0; c < M; ++c) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) acc.p[kk] = fma2(Br2[kk][c], dup(v[c]), acc.p[kk]);
      if (TAIL) acc.t = fmaf(Brt[c], v[c], acc.t);
    }
  };
  auto Kdot = [&](const Vk& v, float out[M]) {  // partial out_c = sum_k K[c][k] v_k (own coordinates)
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = TAIL ? Kct[c] * v.t : 0.f;
      if (NP > 0) {
        f2 s2 = mul2(Kc2[c][0], v.p[0]);
#pragma unroll
        for (int kk = 1; kk < NP; ++kk) s2 = fma2(Kc2[c][kk], v.p[kk], s2);
        s += lo(s2) + hi(s2);
      }
      out[c] = s;
    }
  };
  auto BTdot = [&](const Vk& v, float out[M]) {  // partial out_c = sum_k B[k][c] v_k
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = TAIL ? Brt[c] * v.t : 0.f;
      if (NP > 0) {
        f2 s2 = mul2(Br2[0][c], v.p[0]);
#pragma unroll
        for (int kk = 1; kk < NP; ++kk) s2 = fma2(Br2[kk][c], v.p[kk], s2);
        s += lo(s2) + hi(s2);
      }
      out[c] = s;
    }
  };
  auto KTsub = [&](const float v[M], Vk& acc) {  // acc_k -= sum_c K[c][k] v_c
#pragma unroll
    for (int c = 0; c < M; ++c) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) acc.p[kk] = fma2(Kc2[c][kk], dup(-v[c]), acc.p[kk]);
      if (TAIL) acc.t = fmaf(-Kct[c], v[c], acc.t);
    }
  };

  Vk x, ax;
#pragma unroll
  for (int kk = 0; kk < NP; ++kk) x.p[kk] = pk(ld(a.x_io + n * D, 2 * kk), ld(a.x_io + n * D, 2 * kk + 1));
  x.t = TAIL ? ld(a.x_io + n * D, DL - 1) : 0.f;
  float uprev[M];
  {
    Vk xp;
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) xp.p[kk] = pk(ld(a.xprev_io + n * D, 2 * kk), ld(a.xprev_io + n * D, 2 * kk + 1));
    xp.t = TAIL ? ld(a.xprev_io + n * D, DL - 1) : 0.f;
    float xf[DF], kxp[M];
    gather(xp, xf);
    Amul(xf, ax);                                              // (A - BK) x_prev ...
    Kdot(xp, kxp);
#pragma unroll
    for (int c = 0; c < M; ++c) kxp[c] = group_allreduce2<L>(kxp[c]);
    Bacc(kxp, ax);                                             // ... + B (K x_prev) = A x_prev
#pragma unroll
    for (int c = 0; c < M; ++c) uprev[c] = a.uprev_io ? a.uprev_io[n * M + c] : 0.f;
  }

  const bool sampled = live && (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;

  auto load_w = [&](int t, Vk& w) {
    float ws[DL];
    if (a.W) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * D + q * DL;
#pragma unroll
      for (int k = 0; k < DL; ++k) ws[k] = coord(k) < D ? __ldg(gw + k) : 0.f;
    } else {
      constexpr int NG = L == 1 ? (DL + 3) / 4 : (DL + 6) / 4;
      const int g0 = (q * DL) / 4;
      float z[NG * 4];
#pragma unroll
      for (int g = 0; g < NG; ++g) normal4(p.seed, env_id, (uint32_t)(p.t0 + t), g0 + g, 0u, z + 4 * g);
      const int off = q * DL - g0 * 4;
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        float zz = z[k];
#pragma unroll
        for (int o = 1; o < 4; ++o)
          if (k + o < NG * 4) zz = (off == o) ? z[k + o] : zz;
        ws[k] = coord(k) < D ? __fmul_rn(p.w_std, zz) : 0.f;
      }
    }
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) w.p[kk] = pk(ws[2 * kk], ws[2 * kk + 1]);
    w.t = TAIL ? ws[DL - 1] : 0.f;
  };
  Vk wnext;
  if (p.T > 0) load_w(0, wnext);

  int head = 0;
  const float2* ringP = sP;
  const float* ringT = sT;
  // partial (own coordinates) sum_i M[i] w[base + i]
  auto window_dot = [&](int base, float out[M]) {
    f2 acc[M];
    float acct[M];
#pragma unroll
    for (int c = 0; c < M; ++c) { acc[c] = pk(0.f, 0.f); acct[c] = 0.f; }
#pragma unroll
    for (int i = 0; i < H; ++i) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const float2 wv = ringP[((base + i) * NP + kk) * TPB];
        const f2 w2 = pk(wv.x, wv.y);
#pragma unroll
        for (int c = 0; c < M; ++c) acc[c] = fma2(Mr2[i][c][kk], w2, acc[c]);
      }
      if (TAIL) {
        const float wt = ringT[(base + i) * TPB];
#pragma unroll
        for (int c = 0; c < M; ++c) acct[c] = fmaf(Mrt[i][c], wt, acct[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < M; ++c) out[c] = (lo(acc[c]) + hi(acc[c])) + acct[c];
  };
  // M[i] += s_c (x) w[base + i]
  auto rank1 = [&](int base, const float s[M]) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const float2 wv = ringP[((base + i) * NP + kk) * TPB];
        const f2 w2 = pk(wv.x, wv.y);
#pragma unroll
        for (int c = 0; c < M; ++c) Mr2[i][c][kk] = fma2(dup(s[c]), w2, Mr2[i][c][kk]);
      }
      if (TAIL) {
        const float wt = ringT[(base + i) * TPB];
#pragma unroll
        for (int c = 0; c < M; ++c) Mrt[i][c] = fmaf(s[c], wt, Mrt[i][c]);
      }
    }
  };

  for (int t = 0; t < p.T; ++t) {
    const Vk w_t = wnext;
    if (t + 1 < p.T) load_w(t + 1, wnext);

    constexpr bool BATCH_V = H <= 4;
    constexpr int NR = BATCH_V ? M * (H + 1) : 2 * M;
    float red[NR];
    Kdot(x, red);                                              // K x                    _lqr.py:72
    head = (head + 1 == P) ? 0 : head + 1;                     // roll(-1)               _gpc.py:156-157
    ringP = sP + head * (NP * TPB);
    ringT = sT + head * TPB;
    window_dot(HH - 1, red + M);                               // mw (action == final window)
    if (BATCH_V) {
#pragma unroll
      for (int h = 0; h < H - 1; ++h) window_dot(h, red + 2 * M + h * M);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) red[r] = group_allreduce2<L>(red[r]);
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = red[M + c] - red[c];

    // ---- noise = x - A x_prev - B u ; append (slot head+P-1 and its mirror head-1)   _gpc.py:155-157
    {
      float un[M];
#pragma unroll
      for (int c = 0; c < M; ++c) un[c] = p.noise_uses_prev_action ? uprev[c] : u[c];
      Vk s;
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) s.p[kk] = pk(0.f, 0.f);
      s.t = 0.f;
      Bacc(un, s);
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const f2 nz = sub2(sub2(x.p[kk], ax.p[kk]), s.p[kk]);
        const float2 v = make_float2(lo(nz), hi(nz));
        float2* wn = sP + ((head + P - 1) * NP + kk) * TPB;
        *wn = v;
        if (head != 0) *(wn - P * NP * TPB) = v;
      }
      if (TAIL) {
        const float nz = (x.t - ax.t) - s.t;
        float* wn = sT + (head + P - 1) * TPB;
        *wn = nz;
        if (head != 0) *(wn - P * TPB) = nz;
      }
    }
    // ---- policy_loss forward: y <- A y + B(-K y + v_h) + w[h+H]          _gpc.py:118-124
    Vk y;
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) y.p[kk] = pk(0.f, 0.f);
    y.t = 0.f;
#pragma unroll
    for (int h = 0; h < H - 1; ++h) {
      float v[M];
      if (BATCH_V) {
#pragma unroll
        for (int c = 0; c < M; ++c) v[c] = red[(BATCH_V ? 2 * M + h * M : 0) + c];
      } else {
        window_dot(h, v);
#pragma unroll
        for (int c = 0; c < M; ++c) v[c] = group_allreduce2<L>(v[c]);
      }
      Vk yn;
      if (h == 0) {
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) yn.p[kk] = pk(0.f, 0.f);
        yn.t = 0.f;
      } else {
        float yf[DF];
        gather(y, yf);
        Amul(yf, yn);                                            // (A - BK) y: the -K y term is in the matrix
      }
      Bacc(v, yn);
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const float2 wv = ringP[((h + H) * NP + kk) * TPB];
        y.p[kk] = add2(yn.p[kk], pk(wv.x, wv.y));
      }
      if (TAIL) y.t = yn.t + ringT[(h + H) * TPB];
    }
    // ---- u_f = -K y + mw ; du = 2 u_f ; lam = 2 y - K' du                 SURVEY A-3.5
    const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
    float du[M];
    {
      float ky[M];
      if (H > 1) Kdot(y, ky);
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const float s = H > 1 ? group_allreduce2<L>(ky[c]) : 0.f;
        du[c] = 2.0f * (red[M + c] - s);
      }
    }
    Vk lam;
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) lam.p[kk] = add2(y.p[kk], y.p[kk]);
    lam.t = 2.0f * y.t;
    KTsub(du, lam);
    {
      float sc[M];
#pragma unroll
      for (int c = 0; c < M; ++c) sc[c] = -lr * du[c];
      rank1(HH - 1, sc);                                       // M -= lr du (x) w[HH-1+i]   _gpc.py:163
    }
#pragma unroll
    for (int h = H - 2; h >= 0; --h) {
      float lu[M];
      BTdot(lam, lu);
#pragma unroll
      for (int c = 0; c < M; ++c) lu[c] = group_allreduce2<L>(lu[c]);
      if (h > 0) {  // lam <- A' lam - K' lam_u = (A - BK)' lam
        float part[DF];
#pragma unroll
        for (int r = 0; r < DF; ++r) {
          float s = TAIL ? Art[r] * lam.t : 0.f;
          if (NP > 0) {
            f2 s2 = mul2(Ar2[0][r], lam.p[0]);
#pragma unroll
            for (int kk = 1; kk < NP; ++kk) s2 = fma2(Ar2[kk][r], lam.p[kk], s2);
            s += lo(s2) + hi(s2);
          }
          part[r] = s;
        }
        float own[DL];
#pragma unroll
        for (int k = 0; k < DL; ++k) {
          float s = part[k];
#pragma unroll
          for (int c = 1; c < L; ++c) s += __shfl_xor_sync(0xffffffffu, part[c * DL + k], c);
          own[k] = s;
        }
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) lam.p[kk] = pk(own[2 * kk], own[2 * kk + 1]);
        if (TAIL) lam.t = own[DL - 1];                           // (A - BK)' lam
      }
      float sc[M];
#pragma unroll
      for (int c = 0; c < M; ++c) sc[c] = -lr * lu[c];
      rank1(h, sc);
    }
#pragma unroll
    for (int c = 0; c < M; ++c) uprev[c] = u[c];

    // ---- realised cost x'x + u'u                                           _gpc.py:29-40
    float cost;
    {
      float s = TAIL ? x.t * x.t : 0.f;
      if (NP > 0) {
        f2 s2 = mul2(x.p[0], x.p[0]);
#pragma unroll
        for (int kk = 1; kk < NP; ++kk) s2 = fma2(x.p[kk], x.p[kk], s2);
        s += lo(s2) + hi(s2);
      }
      cost = group_allreduce2<L>(s);
      float cu = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
      cost += cu;
    }
    csum += (double)cost;
    if (live && q == 0 && a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    float xs[DL];
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) { xs[2 * kk] = lo(x.p[kk]); xs[2 * kk + 1] = hi(x.p[kk]); }
    if (TAIL) xs[DL - 1] = x.t;
    if (live && t == p.T - 1) {
#pragma unroll
      for (int k = 0; k < DL; ++k)
        if (coord(k) < D) a.xprev_io[n * D + coord(k)] = xs[k];
    }
    if (sampled) {
      if (a.x_out) {
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (coord(k) < D) a.x_out[((int64_t)t * Ns + ns) * D + coord(k)] = xs[k];
      }
      if (a.u_out && q == 0) {
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
      }
    }

    // ---- env: x <- (A x + B u) + w_t                                        _lds.py:102-111
    {
      // x' = (A - BK) x + B mw + w (u = mw - K x); A x for the next noise estimate = (A - BK) x + B (K x)
      float xf[DF];
      gather(x, xf);
      Amul(xf, ax);
      Vk s = ax;
      Bacc(red + M, s);
      Bacc(red, ax);
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) x.p[kk] = add2(s.p[kk], w_t.p[kk]);
      if (TAIL) x.t = s.t + w_t.t;
    }
  }

  // ---- write back
  if (!live) return;
#pragma unroll
  for (int kk = 0; kk < NP; ++kk) {
    if (coord(2 * kk) < D) a.x_io[n * D + coord(2 * kk)] = lo(x.p[kk]);
    if (coord(2 * kk + 1) < D) a.x_io[n * D + coord(2 * kk + 1)] = hi(x.p[kk]);
  }
  if (TAIL && coord(DL - 1) < D) a.x_io[n * D + coord(DL - 1)] = x.t;
  if (q == 0) {
    if (a.cost_sum_out) a.cost_sum_out[n] = csum;
    if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
  }
  float* gM = a.M_io + n * (H * M * D);
#pragma unroll
  for (int i = 0; i < H; ++i)
#pragma unroll
    for (int c = 0; c < M; ++c) {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        if (coord(2 * kk) < D) gM[(i * M + c) * D + coord(2 * kk)] = lo(Mr2[i][c][kk]);
        if (coord(2 * kk + 1) < D) gM[(i * M + c) * D + coord(2 * kk + 1)] = hi(Mr2[i][c][kk]);
      }
      if (TAIL && coord(DL - 1) < D) gM[(i * M + c) * D + coord(DL - 1)] = Mrt[i][c];
    }
  float* gh = a.hist_io + n * (P * D);
#pragma unroll
  for (int s = 0; s < P; ++s) {
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) {
      const float2 v = ringP[(s * NP + kk) * TPB];
      if (coord(2 * kk) < D) gh[s * D + coord(2 * kk)] = v.x;
      if (coord(2 * kk + 1) < D) gh[s * D + coord(2 * kk + 1)] = v.y;
    }
    if (TAIL && coord(DL - 1) < D) gh[s * D + coord(DL - 1)] = ringT[s * TPB];
  }
  if (a.uprev_io && q == 0) {
#pragma unroll
    for (int c = 0; c < M; ++c) a.uprev_io[n * M + c] = uprev[c];
  }
}

template <int D, int M, int H, int HH, int L, int TPB, int MINB>
static int32_t launch_f2(const LdsArgs& a, cudaStream_t st) {
  using Lay = F2Layout<D, H, HH, L, TPB>;
  auto kern = lds_f2_kernel<D, M, H, HH, L, TPB, MINB>;
  static_assert((Lay::kSmemBytes + 1024) * MINB <= 228 * 1024, "ring does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::kSmemBytes);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_f2)");
  const int64_t threads = a.p.N * L;
  const unsigned grid = (unsigned)((threads + TPB - 1) / TPB);
  kern<<<grid, TPB, Lay::kSmemBytes, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_f2_kernel launch");
}

bool launch_gpc_f2(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  // Launch shape (measured on B200, config 2): one-warp CTAs x 9 per SM at 217 registers run 39.7 ms, 128 x 2
  // at 243 registers 42.5 ms, 64 x 5 at 200 registers 57 ms (ptxas trades ILP for registers), 128 x 3 at 168
  // registers 62 ms (spills).  Default = 32 x 9; kernel_variant 8: 128 x 2, 9: 128 x 3, 10: 64 x 5, 11: 32 x 9.
#define F2(D_, M_, H_, HH_, L_)                                                          \
  if (p.d == D_ && p.m == M_ && p.H == H_ && p.HH == HH_) {                              \
    switch (p.kernel_variant) {                                                          \
      case 9: *rc = launch_f2<D_, M_, H_, HH_, L_, 128, 3>(a, st); break;                \
      case 10: *rc = launch_f2<D_, M_, H_, HH_, L_, 64, 5>(a, st); break;                \
      case 8: *rc = launch_f2<D_, M_, H_, HH_, L_, 128, 2>(a, st); break;                \
      default: *rc = launch_f2<D_, M_, H_, HH_, L_, 32, 9>(a, st); break;                \
    }                                                                                    \
    return true;                                                                         \
  }
  F2(10, 3, 3, 10, 2)
  F2(10, 3, 3, 2, 2)
  F2(10, 3, 10, 10, 4)
#undef F2
  return false;
}

}  // namespace dlc
