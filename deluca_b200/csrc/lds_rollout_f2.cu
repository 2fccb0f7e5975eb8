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
#include "f2.cuh"

#pragma nv_diag_suppress 549  // arrays of the mode that an instantiation does not use (k-paired vs slot-paired)
#pragma nv_diag_suppress 550  // unused half of a mov.b64 unpack

namespace dlc {

template <int L>
__device__ __forceinline__ float group_allreduce2(float v) {
#pragma unroll
  for (int o = 1; o < L; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int D, int H, int HH, int L, int TPB>
struct F2Layout {
  static constexpr int DL = (D + L - 1) / L, NP = DL / 2, TAIL = DL & 1, P = H + HH, SLOTS = 2 * P - 1;
  static constexpr bool IPAIR = (H > 4) && (H % 2 == 0) && (HH == H);  // slot-paired ring, see lds_f2_kernel
  static constexpr size_t kPairBytes = (size_t)SLOTS * NP * TPB * sizeof(float2);
  static constexpr size_t kSmemBytes =
      IPAIR ? (size_t)P * DL * TPB * sizeof(float2) + (size_t)H * 4 * TPB * sizeof(float) /* M <= 4 scratch */
            : kPairBytes + (size_t)SLOTS * TAIL * TPB * sizeof(float);
};

// WSRC: 0 = disturbances from a.W when given, else from the Philox stream (runtime test); 1 = a.W only;
// 2 = Philox only.  The slot-paired instantiations are specialised (1 / 2) to keep the step body small.
template <int D, int M, int H, int HH, int L, int TPB, int MINB, int WSRC>
__global__ void __launch_bounds__(TPB)
    __maxnreg__((65536 / (TPB * MINB)) / 8 * 8 > 255 ? 255 : (65536 / (TPB * MINB)) / 8 * 8)
        lds_f2_kernel(const LdsArgs a) {
  using Lay = F2Layout<D, H, HH, L, TPB>;
  constexpr int DL = Lay::DL, NP = Lay::NP, TAIL = Lay::TAIL, DF = DL * L, P = Lay::P;
  constexpr int NPA = NP > 0 ? NP : 1;
  // Long even histories (H = HH = 10): the H window products of one step form a correlation of M (taps i) with
  // the noise ring (slots j = h + i), so they are paired along the slot axis instead of along k:
  //   forward   (V[2a], V[2a+1])[c]   += M[i][c][k] (scalar broadcast) * (w[2a+i][k], w[2a+i+1][k])
  //   backward  (M[2b], M[2b+1])[c][k] += s[h][c]   (scalar broadcast) * (w[h+2b][k], w[h+2b+1][k])
  // M lives as i-pairs whose halves double as the forward pass's broadcast scalars.  Every M-side FMA is a full
  // FFMA2 for any DL (no tail, no lo+hi adds) and each ring slot is read once per pass.
  // In this mode the ring holds, per slot s and own coordinate k, the float2 (w[s][k], w[s+1][k]) so that any
  // slot pair is one aligned LDS.64; it is not mirrored (P slots, wrap resolved by one select per slot).
  constexpr bool IPAIR = Lay::IPAIR;
  constexpr int HP = H / 2;
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
  float2* const sW = reinterpret_cast<float2*>(smraw) + tid;  // IPAIR ring: (slot, k) at sW[(slot*DL + k) * TPB]
  float* const sS = reinterpret_cast<float*>(smraw + (size_t)P * DL * TPB * sizeof(float2)) + tid;  // IPAIR scratch

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
  f2 Mr2[IPAIR ? 1 : H][M][NPA];
  float Mrt[IPAIR ? 1 : H][M];
  f2 Mi2[IPAIR ? HP : 1][M][DL];
  if constexpr (IPAIR) {
    const float* gM = a.M_io + n * (H * M * D);
#pragma unroll
    for (int b = 0; b < HP; ++b)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          Mi2[b][c][k] = pk(ld(gM + ((2 * b) * M + c) * D, k), ld(gM + ((2 * b + 1) * M + c) * D, k));
  } else {
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
  if constexpr (IPAIR) {
    const float* gh = a.hist_io + n * (P * D);
#pragma unroll
    for (int s = 0; s < P; ++s)
#pragma unroll
      for (int k = 0; k < DL; ++k)
        sW[(s * DL + k) * TPB] = make_float2(ld(gh + s * D, k), s + 1 < P ? ld(gh + (s + 1) * D, k) : 0.f);
  } else {
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
    if (WSRC == 1 || (WSRC == 0 && a.W)) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * D + q * DL;
#pragma unroll
      for (int k = 0; k < DL; ++k) ws[k] = coord(k) < D ? __ldg(gw + k) : 0.f;
    } else {
      constexpr int NG = L == 1 ? (DL + 3) / 4 : (DL + 6) / 4;
      const int g0 = (q * DL) / 4;
      float z[NG * 4];
#pragma unroll
      for (int g = 0; g < NG; ++g) normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), g0 + g, 0u, z + 4 * g);
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
    if constexpr (!IPAIR) {
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
    }
#pragma unroll
    for (int c = 0; c < M; ++c) out[c] = (lo(acc[c]) + hi(acc[c])) + acct[c];
  };
  // M[i] += s_c (x) w[base + i]
  auto rank1 = [&](int base, const float s[M]) {
    if constexpr (!IPAIR) {
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
    }
  };
  // ---- slot-paired forms (IPAIR).  wpair: (w[j][k], w[j+1][k]) of the rolled ring, logical slot j
  const float2* ringLo = sW;   // logical slot j lives at ringLo + j*DL*TPB for j < jwrap, else at ringHi + j*DL*TPB
  const float2* ringHi = sW;
  int jwrap = P;
  auto wslot = [&](int j) -> const float2* { return (j < jwrap ? ringLo : ringHi) + j * (DL * TPB); };
  auto wpair = [&](const float2* slot, int k) -> f2 {
    const float2 v = slot[k * TPB];
    return pk(v.x, v.y);
  };
  auto Mhalf = [&](int i, int c, int k) -> float { return (i & 1) ? hi(Mi2[i / 2][c][k]) : lo(Mi2[i / 2][c][k]); };
  // partial (own coordinates) V[h][c] = sum_i M[i] w[h + i] for h = 0..H-1, out[h * M + c]
  auto windows_all = [&](float out[H * M]) {
    f2 acc[HP][M];
#pragma unroll
    for (int b = 0; b < HP; ++b)
#pragma unroll
      for (int c = 0; c < M; ++c) acc[b][c] = pk(0.f, 0.f);
#pragma unroll
    for (int j = 0; j <= 2 * H - 3; ++j) {
      const float2* slot = wslot(j);
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        const f2 w2 = wpair(slot, k);
#pragma unroll
        for (int b = 0; b < HP; ++b) {
          const int i = j - 2 * b;
          if (i >= 0 && i < H) {
#pragma unroll
            for (int c = 0; c < M; ++c) acc[b][c] = fma2(dup(Mhalf(i, c, k)), w2, acc[b][c]);
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < HP; ++b)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        out[(2 * b) * M + c] = lo(acc[b][c]);
        out[(2 * b + 1) * M + c] = hi(acc[b][c]);
      }
  };
  // M[i] += sum_h s[h] (x) w[h + i]
  auto rank1_all = [&](const float sc[H * M]) {
#pragma unroll
    for (int j = 0; j <= 2 * H - 3; ++j) {
      const float2* slot = wslot(j);
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        const f2 w2 = wpair(slot, k);
#pragma unroll
        for (int b = 0; b < HP; ++b) {
          const int h = j - 2 * b;
          if (h >= 0 && h < H) {
#pragma unroll
            for (int c = 0; c < M; ++c) Mi2[b][c][k] = fma2(dup(sc[h * M + c]), w2, Mi2[b][c][k]);
          }
        }
      }
    }
  };

  for (int t = 0; t < p.T; ++t) {
    // The unrolled step body (~44 KB of SASS in slot-paired mode) exceeds the 32 KB L1.5 instruction cache; a
    // per-step barrier keeps the CTA's warps on the same cache lines so that one fetch from L2 serves all of them.
    if constexpr (IPAIR && TPB > 128) __syncthreads();
    const Vk w_t = wnext;
    if (t + 1 < p.T) load_w(t + 1, wnext);

    constexpr bool BATCH_V = H <= 4 || IPAIR;
    constexpr int NR = BATCH_V ? M * (H + 1) : 2 * M;
    float red[NR];
    Kdot(x, red);                                              // K x                    _lqr.py:72
    head = (head + 1 == P) ? 0 : head + 1;                     // roll(-1)               _gpc.py:156-157
    ringP = sP + head * (NP * TPB);
    ringT = sT + head * TPB;
    if constexpr (IPAIR) {
      ringLo = sW + head * (DL * TPB);
      ringHi = ringLo - P * (DL * TPB);
      jwrap = P - head;
      float V[H * M];
      windows_all(V);
#pragma unroll
      for (int c = 0; c < M; ++c) red[M + c] = V[(H - 1) * M + c];  // mw: base HH-1 == H-1
#pragma unroll
      for (int r = 0; r < (H - 1) * M; ++r) red[2 * M + r] = V[r];
    } else {
      window_dot(HH - 1, red + M);                             // mw (action == final window)
      if (BATCH_V) {
#pragma unroll
        for (int h = 0; h < H - 1; ++h) window_dot(h, red + 2 * M + h * M);
      }
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
      if constexpr (IPAIR) {
        float nz[DL];
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) {
          const f2 d2 = sub2(sub2(x.p[kk], ax.p[kk]), s.p[kk]);
          nz[2 * kk] = lo(d2);
          nz[2 * kk + 1] = hi(d2);
        }
        if (TAIL) nz[DL - 1] = (x.t - ax.t) - s.t;
        float* newest = reinterpret_cast<float*>(const_cast<float2*>(wslot(P - 1)));      // .x of slot P-1
        float* second = reinterpret_cast<float*>(const_cast<float2*>(wslot(P - 2))) + 1;  // .y of slot P-2
#pragma unroll
        for (int k = 0; k < DL; ++k) {
          newest[2 * k * TPB] = nz[k];
          second[2 * k * TPB] = nz[k];
        }
      } else {
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
    }
    // ---- policy_loss forward: y <- A y + B(-K y + v_h) + w[h+H]          _gpc.py:118-124
    Vk y;
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) y.p[kk] = pk(0.f, 0.f);
    y.t = 0.f;
    auto add_w = [&](int j, const Vk& yn) {                      // y = yn + w[j]
      if constexpr (IPAIR) {
        const float* ws = reinterpret_cast<const float*>(wslot(j));
#pragma unroll
        for (int kk = 0; kk < NP; ++kk)
          y.p[kk] = add2(yn.p[kk], pk(ws[2 * (2 * kk) * TPB], ws[2 * (2 * kk + 1) * TPB]));
        if (TAIL) y.t = yn.t + ws[2 * (DL - 1) * TPB];
      } else {
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) {
          const float2 wv = ringP[(j * NP + kk) * TPB];
          y.p[kk] = add2(yn.p[kk], pk(wv.x, wv.y));
        }
        if (TAIL) y.t = yn.t + ringT[j * TPB];
      }
    };
    if constexpr (IPAIR) {
      // The two recursions are rolled (their operands staged in a thread-private shared scratch) so that the
      // step body stays inside the 32 KB L1.5 instruction cache; everything M-sided remains unrolled.
      {
        Vk yn;
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) yn.p[kk] = pk(0.f, 0.f);
        yn.t = 0.f;
        Bacc(red + 2 * M, yn);
        add_w(H, yn);
      }
#pragma unroll
      for (int r = M; r < (H - 1) * M; ++r) sS[r * TPB] = red[2 * M + r];
#pragma unroll 1
      for (int h = 1; h < H - 1; ++h) {
        float v[M], yf[DF];
#pragma unroll
        for (int c = 0; c < M; ++c) v[c] = sS[(h * M + c) * TPB];
        Vk yn;
        gather(y, yf);
        Amul(yf, yn);
        Bacc(v, yn);
        add_w(h + H, yn);
      }
    } else {
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
      add_w(h + H, yn);
    }
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
    float sall[IPAIR ? H * M : 1];
    if constexpr (IPAIR) {
#pragma unroll
      for (int c = 0; c < M; ++c) sall[(H - 1) * M + c] = -lr * du[c];
    } else {
      float sc[M];
#pragma unroll
      for (int c = 0; c < M; ++c) sc[c] = -lr * du[c];
      rank1(HH - 1, sc);                                       // M -= lr du (x) w[HH-1+i]   _gpc.py:163
    }
    auto lam_back = [&]() {  // lam <- A' lam - K' lam_u = (A - BK)' lam
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
      if (TAIL) lam.t = own[DL - 1];
    };
    if constexpr (IPAIR) {
#pragma unroll 1
      for (int h = H - 2; h >= 1; --h) {
        float lu[M];
        BTdot(lam, lu);
#pragma unroll
        for (int c = 0; c < M; ++c) sS[(h * M + c) * TPB] = -lr * group_allreduce2<L>(lu[c]);
        lam_back();
      }
      float lu[M];
      BTdot(lam, lu);
#pragma unroll
      for (int c = 0; c < M; ++c) sall[c] = -lr * group_allreduce2<L>(lu[c]);
#pragma unroll
      for (int r = M; r < (H - 1) * M; ++r) sall[r] = sS[r * TPB];
      rank1_all(sall);                                         // M -= lr sum_h lam_u[h] (x) w[h+i]   _gpc.py:163
    } else {
#pragma unroll
    for (int h = H - 2; h >= 0; --h) {
      float lu[M];
      BTdot(lam, lu);
#pragma unroll
      for (int c = 0; c < M; ++c) lu[c] = group_allreduce2<L>(lu[c]);
      if (h > 0) lam_back();
      float sc[M];
#pragma unroll
      for (int c = 0; c < M; ++c) sc[c] = -lr * lu[c];
      rank1(h, sc);
    }
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
  if constexpr (IPAIR) {
#pragma unroll
    for (int b = 0; b < HP; ++b)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (coord(k) < D) {
            gM[((2 * b) * M + c) * D + coord(k)] = lo(Mi2[b][c][k]);
            gM[((2 * b + 1) * M + c) * D + coord(k)] = hi(Mi2[b][c][k]);
          }
  } else
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
  if constexpr (IPAIR) {
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const float* ws = reinterpret_cast<const float*>(wslot(s));
#pragma unroll
      for (int k = 0; k < DL; ++k)
        if (coord(k) < D) gh[s * D + coord(k)] = ws[2 * k * TPB];
    }
  } else
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

template <int D, int M, int H, int HH, int L, int TPB, int MINB, int WSRC>
static int32_t launch_f2w(const LdsArgs& a, cudaStream_t st) {
  using Lay = F2Layout<D, H, HH, L, TPB>;
  auto kern = lds_f2_kernel<D, M, H, HH, L, TPB, MINB, WSRC>;
  if constexpr ((Lay::kSmemBytes + 1024) * MINB > 228 * 1024) {
    return fail(DLC_ERR_ARG, "lds_f2_kernel: this launch shape does not fit shared memory for the given history");
  } else {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::kSmemBytes);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_f2)");
  const int64_t threads = a.p.N * L;
  const unsigned grid = (unsigned)((threads + TPB - 1) / TPB);
  kern<<<grid, TPB, Lay::kSmemBytes, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_f2_kernel launch");
  }
}

template <int D, int M, int H, int HH, int L, int TPB, int MINB>
static int32_t launch_f2(const LdsArgs& a, cudaStream_t st) {
  if constexpr (F2Layout<D, H, HH, L, TPB>::IPAIR) {
    return a.W ? launch_f2w<D, M, H, HH, L, TPB, MINB, 1>(a, st) : launch_f2w<D, M, H, HH, L, TPB, MINB, 2>(a, st);
  } else {
    return launch_f2w<D, M, H, HH, L, TPB, MINB, 0>(a, st);
  }
}

bool launch_gpc_f2(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  // Launch shape (measured on B200, config 2): one-warp CTAs x 9 per SM at 217 registers run 39.7 ms, 128 x 2
  // at 243 registers 42.5 ms, 64 x 5 at 200 registers 57 ms (ptxas trades ILP for registers), 128 x 3 at 168
  // registers 62 ms (spills).  Default = 32 x 9; kernel_variant 8: 128 x 2, 9: 128 x 3, 10: 64 x 5, 11: 32 x 9,
  // 12: 256 x 1.  Slot-paired mode (H = HH = 10, the 2^20-env target): 128 x 2 499 ms, 32 x 9 507 ms, 256 x 1 523 ms,
  // 64 x 5 582 ms (scalar lds_split_kernel: 545 ms); default there = 128 x 2.
#define F2(D_, M_, H_, HH_, L_)                                                          \
  if (p.d == D_ && p.m == M_ && p.H == H_ && p.HH == HH_) {                              \
    switch (p.kernel_variant) {                                                          \
      case 9: *rc = launch_f2<D_, M_, H_, HH_, L_, 128, 3>(a, st); break;                \
      case 10: *rc = launch_f2<D_, M_, H_, HH_, L_, 64, 5>(a, st); break;                \
      case 8: *rc = launch_f2<D_, M_, H_, HH_, L_, 128, 2>(a, st); break;                \
      case 12: *rc = launch_f2<D_, M_, H_, HH_, L_, 256, 1>(a, st); break;               \
      case 11: *rc = launch_f2<D_, M_, H_, HH_, L_, 32, 9>(a, st); break;                \
      default:                                                                           \
        *rc = (H_ > 4) ? launch_f2<D_, M_, H_, HH_, L_, 128, 2>(a, st)                   \
                       : launch_f2<D_, M_, H_, HH_, L_, 32, 9>(a, st);                   \
        break;                                                                           \
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
