// LDS + GPC rollout for long even histories (H = HH, the north-star target d = 10, m = 3, H = HH = 10).
//
// 1. Layout.  Four lanes own an env and split it TWO ways -- by state coordinate (halves S_0, S_1 of D/2) and by GPC
//    tap (halves I_0, I_1 of H/2) -- so that every lane holds exactly a quarter of everything and nothing is padded:
//    lane (qc, qi) keeps M[i in I_qi][c][k in S_qc] (75 floats) in registers for the whole rollout.
//    (lds_f2_kernel's split of 10 coordinates over 4 lanes as 3 + 3 + 3 + 1 pads every contraction 12-for-10.)
//
// 2. One pass over the noise ring per step.  The rank-H update of step t-1 and the H window products of step t,
//        M[i] += sum_h s[h] (x) w[h + i]        V[h] = sum_i M[i] w'[h + i],   w'[j] = w[j + 1] (the ring rolled by one)
//    touch the same ring slots, so for each own coordinate the 15 slots a lane needs are loaded once, M[.][.][k] is
//    updated in registers and immediately used for the window products.  Both sides are packed FFMA2 (tap pairs as
//    accumulators backward, window pairs forward).
//
// 3. No recursions.  policy_loss runs H-1 steps of y <- Abar y + B v_h + w[h+H] (Abar = A - B K) and only its LAST
//    state enters the loss, and the dynamics are linear, so (exact algebra, not an approximation)
//        y_f      = sum_h Abar^(H-2-h) (B v_h + w[h+H]) = sum_h G_(H-2-h) v_h + q,        G_k = Abar^k B
//        s[h]     = -lr B' (Abar')^(H-2-h) lam = -lr G_(H-2-h)' lam                       (the adjoint of the same map)
//        q_(t+1)  = Abar q_t - Abar^(H-1) w_t[H] + w_t[2H-1]                              (the ring only shifts)
//    with G_k and Abar^(H-1) formed once per rollout.  The 2 (H - 2) dependent mat-vecs per step -- a latency chain that
//    no amount of occupancy at 255 registers could hide -- become two wide products against the same 27 x 10 operator
//    plus two block products for q; 2,080 FMAs per env-step shrink to 740.  q is a contraction (Abar is stable), so its
//    rounding error does not accumulate; the result agrees with the float64 oracle to ~1e-6 (tests/test_lds_gpu.py).
//
// Reference arithmetic: deluca/agents/_gpc.py:109-183 (policy_loss, update, get_action), deluca/envs/_lds.py:102-111.
#include <stdlib.h>

#include "lds_args.cuh"
#include "f2.cuh"

#pragma nv_diag_suppress 550  // unused half of a mov.b64 unpack

namespace dlc {

template <int D, int M, int H, int TPB>
struct QuadLayout {
  static constexpr int DC = D / 2, HT = H / 2, HS = H / 2, P = 2 * H, EPW = 8, WARPS = TPB / 32;
  // floats per ring slot of one warp: [coordinate][env] plus a pad chosen so that the four lanes of an env -- reading
  // coordinate halves D/2 apart and slots H/2 apart in one instruction -- land in four different bank octets
  static constexpr int S1 = D * EPW;                                  // (5 slots * 80 floats = 16 mod 32 banks: no pad needed)
  static_assert((DC * EPW) % 32 == 8 && (HT * S1) % 32 == 16, "ring padding is tuned for D = 10, H = 10");
  static constexpr size_t kRingBytes = (size_t)WARPS * P * S1 * sizeof(float);
  // per-lane operators in shared memory, as float4 columns [idx4][thread]: Abar block | Abar^(H-1) block | G rows
  static constexpr int nA4 = (DC * DC + 3) / 4, nG4 = (HS * DC * M + 3) / 4;
  static constexpr int oA9 = nA4, oG = 2 * nA4, NPAR4 = 2 * nA4 + nG4;
  static constexpr size_t kParamBytes = (size_t)NPAR4 * TPB * sizeof(float4);
  static constexpr size_t kSmemBytes = kRingBytes + kParamBytes;
  // INCR: per warp H slots x H positions (+ one pad row) of window Gram sums, [row][env]
  static constexpr int CROWS = H * H + 1;
  static constexpr size_t kColBytes = (size_t)WARPS * CROWS * EPW * sizeof(float);
  // INCR: the previous step's window products V[h][c], h = 0 .. H (row H is a pad), [row][env] per warp
  static constexpr int VROWS = (H + 1) * M;
  static constexpr size_t kVBytes = (size_t)WARPS * VROWS * EPW * sizeof(float);
  static constexpr size_t kSmemBytesIncr = kSmemBytes + kColBytes + kVBytes;
};

// WSRC: 1 = disturbances from a.W, 2 = in-kernel Philox stream
//
// INCR: incremental window products.  The ring only shifts (w_t[j] = w_(t-1)[j+1]) and M only changes by the rank-H
// update, so for h <= H-2
//     V_t[h] = sum_i M_t[i] w_t[h+i+1] = V_(t-1)[h+1] + sum_h' s_(t-1)[h'] C_t[h'][h+1],
//     C_t[a][b] = sum_(i<H) <w_t[a+i], w_t[b+i]>                    (Gram sums of the ring: no M in them)
// and only the newest window V_t[H-1] is contracted with M (H m d multiply-adds instead of H^2 m d).  C shifts with the
// ring as well, C_(t+1)[a][b] = C_t[a+1][b+1]; its last column is carried by
//     C_(t+1)[a][H-1] = C_t[a][H-1] + <w_t[a+H], w_t[2H-1]> - <w_t[a], w_t[H-1]>
// (2H inner products per step, formed inside the pass from slots that are in registers anyway).  Entry (mn, mx),
// mn <= mx, of C_t lives at row ((t + mn) mod H) * H + (mx - mn) of the warp's column ring -- a place that does not move
// when the matrix shifts -- so a step writes H values and reads the H (H-1) it needs.  A window value is formed in
// full once and corrected H-1 times; C and V are re-formed from (M, ring) at every call entry.  Same numbers to
// rounding (float64 prototype: 1e-15; held to 1e-4 against the float64 oracle like the direct form).
template <int D, int M, int H, int TPB, int MINB, int WSRC, bool INCR>
__global__ void __launch_bounds__(TPB, MINB) lds_gpc_quad_kernel(const LdsArgs a) {
  using Lay = QuadLayout<D, M, H, TPB>;
  constexpr int DC = Lay::DC, HT = Lay::HT, HS = Lay::HS, HP = H / 2, P = Lay::P, EPW = Lay::EPW, S1 = Lay::S1;
  constexpr int NPI = HT / 2, TI = HT & 1, NPIA = NPI > 0 ? NPI : 1;   // tap pairs of a tap half
  constexpr int R = H + HT;                                            // ring slots one lane touches per pass
  constexpr int KMAX = H - 1;                                          // Abar^KMAX closes the sliding sum q
  static_assert(D % 2 == 0 && H % 2 == 0 && H >= 4, "the quad split needs even D and even H >= 4");
  extern __shared__ __align__(16) unsigned char smraw[];
  const DlcLdsParams& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, qc = q & 1, qi = q >> 1, e = lane >> 2;
  const int partner = (lane & ~3) | (qc << 1) | qi;                    // the lane with the transposed role (qi, qc)
  const bool diag = qc == qi;
  const int64_t n_raw = ((int64_t)blockIdx.x * TPB + tid) >> 2;
  const bool live = n_raw < p.N;
  const int64_t n = live ? n_raw : p.N - 1;
  const int i0 = qi * HT, k0 = qc * DC, j0 = qi * DC;
  float* const ring = reinterpret_cast<float*>(smraw) + (size_t)warp * P * S1 + e;    // w[slot][k] = ring[slot*S1 + k*EPW]
  float* const rown = ring + k0 * EPW;                                                // own coordinates (S_qc)
  float* const rcol = ring + j0 * EPW;                                                // coordinates S_qi
  float4* const sPar = reinterpret_cast<float4*>(smraw + Lay::kRingBytes) + tid;      // sPar[idx4 * TPB]
  // INCR: the warp's column ring of window Gram sums, entry (row, env) at colw[row * EPW]
  float* const colw = reinterpret_cast<float*>(smraw + Lay::kSmemBytes) + (size_t)warp * Lay::CROWS * EPW + e;
  static_assert(!INCR || ((HT & 1) == 1), "INCR is written for an odd tap half (H = 10)");
  // INCR: V_(t-1)[h][c] at vsh[(h * M + c) * EPW] (kept in shared memory: 15 registers less across the pass)
  float* const vsh = reinterpret_cast<float*>(smraw + Lay::kSmemBytes + Lay::kColBytes) + (size_t)warp * Lay::VROWS * EPW + e;
  constexpr unsigned FULL = 0xffffffffu;

  // ================= set-up: operators of the recursion-free form (once per rollout)
  float Br[DC][M], Kc[M][DC];             // B rows S_qc, K columns S_qc: registers for the whole rollout
  float qrow[DC];                         // q split by rows (S_qc)
  {
    const int64_t np = p.shared_dynamics ? 0 : n;
    const float* gA = a.A + np * (D * D);
    const float* gB = a.B + np * (D * M);
    const float* gK = a.K + np * (M * D);
    const float* gh = a.hist_io + n * (P * D);
    float Bq[DC][M];                      // B rows S_qi (for G = Abar^k B summed over the column halves)
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        Br[kk][c] = __ldg(gB + (k0 + kk) * M + c);
        Kc[c][kk] = __ldg(gK + c * D + k0 + kk);
        Bq[kk][c] = __ldg(gB + (j0 + kk) * M + c);
      }
    float Aown[DC][DC], Aoth[DC][DC];     // Abar[S_qc][S_qi], Abar[S_qc][S_(1-qi)]
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) {
        float v = __ldg(gA + (k0 + kk) * D + j0 + jj);
#pragma unroll
        for (int c = 0; c < M; ++c) v = fmaf(-Br[kk][c], __ldg(gK + c * D + j0 + jj), v);
        Aown[kk][jj] = v;
      }
    {
      float flat[Lay::nA4 * 4];           // Abar block, row-major [kk][jj], padded to float4
#pragma unroll
      for (int i = 0; i < Lay::nA4 * 4; ++i) flat[i] = i < DC * DC ? Aown[i / DC][i % DC] : 0.f;
#pragma unroll
      for (int i = 0; i < Lay::nA4; ++i) sPar[i * TPB] = make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]);
    }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) Aoth[kk][jj] = __shfl_xor_sync(FULL, Aown[kk][jj], 2);
    // Abar[S_qc][S_qc] and Abar[S_qc][S_(1-qc)]: what multiplies the own / the other row block of a power
    float Asame[DC][DC], Adiff[DC][DC];
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) {
        Asame[kk][jj] = diag ? Aown[kk][jj] : Aoth[kk][jj];
        Adiff[kk][jj] = diag ? Aoth[kk][jj] : Aown[kk][jj];
      }
    float Pk[DC][DC];                     // block (qc, qi) of Abar^k
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) Pk[kk][jj] = (diag && kk == jj) ? 1.f : 0.f;
    float qacc[DC];
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) qacc[kk] = 0.f;
    float* const sG = reinterpret_cast<float*>(sPar + Lay::oG * TPB);          // G rows as scalars: float i at sG[(i/4)*4*TPB + i%4]
    // the lane's window subset is h in [qi*HS, qi*HS + HS); window h uses G_(H-2-h); h = H-1 has no recursion term
    for (int i = 0; i < Lay::nG4 * 4; ++i) sG[(i >> 2) * 4 * TPB + (i & 3)] = 0.f;
#pragma unroll 1
    for (int k = 0; k <= KMAX; ++k) {
      if (k <= H - 2) {
        // G_k[S_qc][:] = sum over the column halves of Abar^k(qc, .) B[S_.][:]
        float g[DC][M];
#pragma unroll
        for (int kk = 0; kk < DC; ++kk)
#pragma unroll
          for (int c = 0; c < M; ++c) {
            float s = 0.f;
#pragma unroll
            for (int jj = 0; jj < DC; ++jj) s = fmaf(Pk[kk][jj], Bq[jj][c], s);
            g[kk][c] = s + __shfl_xor_sync(FULL, s, 2);
          }
        const int hh = (H - 2 - k) - qi * HS;                          // position of window h = H-2-k in the lane's subset
        if (hh >= 0 && hh < HS) {
#pragma unroll
          for (int kk = 0; kk < DC; ++kk)
#pragma unroll
            for (int c = 0; c < M; ++c) {
              const int i = (hh * DC + kk) * M + c;
              sG[(i >> 2) * 4 * TPB + (i & 3)] = g[kk][c];
            }
        }
        // q of the entry ring: sum_k Abar^k w[2H-2-k]  (slots H .. 2H-2), partial over the column halves
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) {
          float s = qacc[kk];
#pragma unroll
          for (int jj = 0; jj < DC; ++jj) s = fmaf(Pk[kk][jj], gh[(2 * H - 2 - k) * D + j0 + jj], s);
          qacc[kk] = s;
        }
      }
      if (k == KMAX) {
        float flat[Lay::nA4 * 4];
#pragma unroll
        for (int i = 0; i < Lay::nA4 * 4; ++i) flat[i] = i < DC * DC ? Pk[i / DC][i % DC] : 0.f;
#pragma unroll
        for (int i = 0; i < Lay::nA4; ++i)
          sPar[(Lay::oA9 + i) * TPB] = make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]);
      } else {
        // Abar^(k+1)(qc, qi) = Abar(qc, qc) P(qc, qi) + Abar(qc, 1-qc) P(1-qc, qi)
        float Po[DC][DC], Pnew[DC][DC];
#pragma unroll
        for (int kk = 0; kk < DC; ++kk)
#pragma unroll
          for (int jj = 0; jj < DC; ++jj) Po[kk][jj] = __shfl_xor_sync(FULL, Pk[kk][jj], 1);   // block (1-qc, qi)
#pragma unroll
        for (int kk = 0; kk < DC; ++kk)
#pragma unroll
          for (int jj = 0; jj < DC; ++jj) {
            float s = 0.f;
#pragma unroll
            for (int l = 0; l < DC; ++l) s = fmaf(Asame[kk][l], Pk[l][jj], s);
#pragma unroll
            for (int l = 0; l < DC; ++l) s = fmaf(Adiff[kk][l], Po[l][jj], s);
            Pnew[kk][jj] = s;
          }
#pragma unroll
        for (int kk = 0; kk < DC; ++kk)
#pragma unroll
          for (int jj = 0; jj < DC; ++jj) Pk[kk][jj] = Pnew[kk][jj];
      }
    }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) qrow[kk] = qacc[kk] + __shfl_xor_sync(FULL, qacc[kk], 2);
  }
  float qcol[DC];                         // q split by columns (S_qi): the input layout of the next block product
#pragma unroll
  for (int kk = 0; kk < DC; ++kk) qcol[kk] = __shfl_sync(FULL, qrow[kk], partner);

  // M: tap pairs (2p, 2p+1) as packed accumulators, the odd tap as a scalar
  f2 Mp[NPIA][M][DC];
  float Ms[M][DC];
  {
    const float* gM = a.M_io + n * (H * M * D);
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
#pragma unroll
        for (int pp = 0; pp < NPI; ++pp)
          Mp[pp][c][kk] = pk(gM[((i0 + 2 * pp) * M + c) * D + k0 + kk], gM[((i0 + 2 * pp + 1) * M + c) * D + k0 + kk]);
        if (TI) Ms[c][kk] = gM[((i0 + HT - 1) * M + c) * D + k0 + kk];
      }
  }
  auto Mtap = [&](int ii, int c, int kk) -> float {
    if (TI && ii == HT - 1) return Ms[c][kk];
    return (ii & 1) ? hi(Mp[ii / 2][c][kk]) : lo(Mp[ii / 2][c][kk]);
  };
  // noise ring: the four lanes of an env fill it together
  {
    const float* gh = a.hist_io + n * (P * D);
    for (int idx = q; idx < P * D; idx += 4) {
      const int s = idx / D, k = idx - s * D;
      ring[s * S1 + k * EPW] = gh[s * D + k];
    }
  }
  __syncwarp();

  // INCR set-up: C_0 and the previous step's window products, both formed in full from (M, ring) once per call
  int cb = 0;                             // t mod H
  if constexpr (INCR) {
    float Vp[HS][M];                      // V_(-1) on the lane's window subset h in [qi*HS, qi*HS + HS)
    int cnt = 0;
#pragma unroll 1
    for (int mn = 0; mn < H; ++mn)
#pragma unroll 1
      for (int mx = mn; mx < H; ++mx) {
        if ((cnt++ & 3) != q) continue;
        float sacc = 0.f;
#pragma unroll 1
        for (int i = 0; i < H; ++i)
#pragma unroll
          for (int k = 0; k < D; ++k)
            sacc = fmaf(ring[(mn + i) * S1 + k * EPW], ring[(mx + i) * S1 + k * EPW], sacc);
        colw[(mn * H + (mx - mn)) * EPW] = sacc;                       // (head = 0 here: logical slot = physical slot)
      }
    if (q == 0) colw[H * H * EPW] = 0.f;                               // pad row (read by a discarded term only)
#pragma unroll
    for (int hh = 0; hh < HS; ++hh)
#pragma unroll
      for (int c = 0; c < M; ++c) Vp[hh][c] = 0.f;
    const float* gM = a.M_io + n * (H * M * D);
#pragma unroll 1
    for (int i = 0; i < H; ++i)
#pragma unroll 1
      for (int k = 0; k < D; ++k) {
        float mv[M];
#pragma unroll
        for (int c = 0; c < M; ++c) mv[c] = gM[(i * M + c) * D + k];
#pragma unroll
        for (int hh = 0; hh < HS; ++hh) {                              // V_(-1)[h] = sum_i M[i] w[h + i]
          const float wv = ring[(qi * HS + hh + i) * S1 + k * EPW];
#pragma unroll
          for (int c = 0; c < M; ++c) Vp[hh][c] = fmaf(mv[c], wv, Vp[hh][c]);
        }
      }
    if (qc == 0) {
#pragma unroll
      for (int hh = 0; hh < HS; ++hh)
#pragma unroll
        for (int c = 0; c < M; ++c) vsh[((qi * HS + hh) * M + c) * EPW] = Vp[hh][c];
    }
    if (q == 0) {
#pragma unroll
      for (int c = 0; c < M; ++c) vsh[(H * M + c) * EPW] = 0.f;        // pad row (read by a discarded term only)
    }
    __syncwarp();
  }

  // per-lane operators back from shared memory (float4 columns), at the point of use
  auto load_block = [&](int o4, float blk[DC][DC]) {
    float flat[Lay::nA4 * 4];
#pragma unroll
    for (int i = 0; i < Lay::nA4; ++i) {
      const float4 v = sPar[(o4 + i) * TPB];
      flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < DC * DC; ++i) blk[i / DC][i % DC] = flat[i];
  };
  auto load_G = [&](float G[HS][DC][M]) {
    float flat[Lay::nG4 * 4];
#pragma unroll
    for (int i = 0; i < Lay::nG4; ++i) {
      const float4 v = sPar[(Lay::oG + i) * TPB];
      flat[4 * i] = v.x; flat[4 * i + 1] = v.y; flat[4 * i + 2] = v.z; flat[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < HS * DC * M; ++i) G[i / (DC * M)][(i / M) % DC][i % M] = flat[i];
  };
  auto K_dot = [&](const float vr[DC], float out[M]) {                // K v (v split by rows), summed over qc
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = Kc[c][0] * vr[0];
#pragma unroll
      for (int kk = 1; kk < DC; ++kk) s = fmaf(Kc[c][kk], vr[kk], s);
      out[c] = s;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) out[c] += __shfl_xor_sync(FULL, out[c], 1);
  };
  auto B_acc = [&](const float v[M], float acc[DC]) {                 // acc[S_qc] += B[S_qc][:] v
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int c = 0; c < M; ++c) acc[kk] = fmaf(Br[kk][c], v[c], acc[kk]);
  };

  // ---- state
  float xr[DC], ax[DC], uprev[M];
#pragma unroll
  for (int kk = 0; kk < DC; ++kk) xr[kk] = a.x_io[n * D + k0 + kk];
  {
    float Ab[DC][DC], xpc[DC], xpr[DC], kxp[M];
    load_block(0, Ab);
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) { xpr[kk] = a.xprev_io[n * D + k0 + kk]; xpc[kk] = a.xprev_io[n * D + j0 + kk]; }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) {
      float s = 0.f;
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) s = fmaf(Ab[kk][jj], xpc[jj], s);
      ax[kk] = s + __shfl_xor_sync(FULL, s, 2);                        // Abar x_prev ...
    }
    K_dot(xpr, kxp);
    B_acc(kxp, ax);                                                   // ... + B (K x_prev) = A x_prev
#pragma unroll
    for (int c = 0; c < M; ++c) uprev[c] = a.uprev_io ? a.uprev_io[n * M + c] : 0.f;
  }
  const bool sampled = live && (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;

  auto load_w = [&](int t, float w[DC]) {
    if (WSRC == 1) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * D + k0;
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) w[kk] = __ldg(gw + kk);
    } else {
      // lane q draws the normals of coordinates 4q .. 4q+3 (the stream of dlc_gaussian_disturbance_f32); each
      // coordinate is then fetched from the lane that drew it
      float z[4];
      normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), (uint32_t)q, 0u, z);
      float all[D];
#pragma unroll
      for (int c = 0; c < D; ++c) all[c] = __shfl_sync(FULL, z[c & 3], (lane & ~3) | (c >> 2));
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) w[kk] = __fmul_rn(p.w_std, qc ? all[DC + kk] : all[kk]);
    }
  };
  float wnext[DC];
  if (p.T > 0) load_w(0, wnext);
  float sprev[H][M];                      // s of the previous step (all windows); zero: no update in the first pass
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int c = 0; c < M; ++c) sprev[h][c] = 0.f;

  int head = 0;
  // T + 1 trips: the pass at the top of trip t applies step t-1's update, so trip T only runs the pass.  (Written as a
  // counted loop with a guarded body so that the compiler does not rotate the loop and duplicate the pass.)
#pragma unroll 1
  for (int t = 0; t <= p.T; ++t) {
    // ================= work that does not depend on this step's window products, issued ahead of the pass so that its
    // shuffle / shared-memory latencies run under the pass's FFMA2 stream (ring indices here: before the roll)
    float kx[M], z[DC], nzp[DC], wt[DC];
    {
      float Ab[DC][DC], A9[DC][DC], wold[DC], xc[DC];
      load_block(0, Ab);
      load_block(Lay::oA9, A9);
      // q_t = Abar q_(t-1) - Abar^(H-1) w[H] + w[2H-1]          (the ring only shifts: see the header)
      {
        int so = head + H, sn = head + 2 * H - 1;
        so -= so >= P ? P : 0;
        sn -= sn >= P ? P : 0;
#pragma unroll
        for (int jj = 0; jj < DC; ++jj) wold[jj] = rcol[so * S1 + jj * EPW];
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) xc[kk] = __shfl_sync(FULL, xr[kk], partner);   // x split by columns
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) {
          float s0 = Ab[kk][0] * qcol[0], s1 = -A9[kk][0] * wold[0], s2 = Ab[kk][0] * xc[0];
#pragma unroll
          for (int jj = 1; jj < DC; ++jj) {
            s0 = fmaf(Ab[kk][jj], qcol[jj], s0);
            s1 = fmaf(-A9[kk][jj], wold[jj], s1);
            s2 = fmaf(Ab[kk][jj], xc[jj], s2);
          }
          const float s = s0 + s1;
          qrow[kk] = (s + __shfl_xor_sync(FULL, s, 2)) + rown[sn * S1 + kk * EPW];
          z[kk] = s2 + __shfl_xor_sync(FULL, s2, 2);                   // Abar x
        }
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) qcol[kk] = __shfl_sync(FULL, qrow[kk], partner);
      }
      K_dot(xr, kx);                                                  // K x                     _lqr.py:72
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
        nzp[kk] = xr[kk] - ax[kk];                                    // x - A x_prev (the noise estimate minus B u)
        ax[kk] = z[kk];
        wt[kk] = wnext[kk];
      }
      B_acc(kx, ax);                                                  // A x = Abar x + B (K x): next step's A x_prev
      const int tn = t + 1 < p.T ? t + 1 : (p.T > 0 ? p.T - 1 : 0);
      if (p.T > 0) load_w(tn, wnext);
    }
    // ================= one pass over the ring (old indexing): M += sum_h s[h] (x) w[h+i];  V[h] = sum_i M[i] w[h+i+1]
    constexpr int NACC = INCR ? 1 : HP;
    f2 acc[NACC][M];
    float acct[M];                        // INCR: the odd tap's share of the newest window
    f2 g2[H / 2];                         // INCR: pairs of ring inner products (in-terms on qi = 1, out-terms on qi = 0)
    {
#pragma unroll
      for (int b = 0; b < NACC; ++b)
#pragma unroll
        for (int c = 0; c < M; ++c) acc[b][c] = pk(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < M; ++c) acct[c] = 0.f;
#pragma unroll
      for (int j = 0; j < H / 2; ++j) g2[j] = pk(0.f, 0.f);
      // logical slot i0 + r lives at byte address rlo + r*S1*4, or, past the wrap, at rhi + r*S1*4
      const unsigned rlo = (unsigned)__cvta_generic_to_shared(rown + (head + i0) * S1);
      const unsigned rhi = rlo - P * S1 * 4;
      const int rw = P - head - i0;
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
        // every slot pair (w[i0 + r], w[i0 + r + 1]) is loaded straight into its own aligned register pair (two
        // LDS.32 per pair; a slot is therefore loaded twice, which is cheaper than the register moves that sharing
        // one load between an even and an odd pair costs)
        f2 wp[R - 1];
#pragma unroll
        for (int r = 0; r < R - 1; ++r) {
          const unsigned a0 = ((r < rw) ? rlo : rhi) + (r * S1 + kk * EPW) * 4;
          const unsigned a1 = ((r + 1 < rw) ? rlo : rhi) + ((r + 1) * S1 + kk * EPW) * 4;
          asm volatile("{ .reg .f32 lo_, hi_; ld.volatile.shared.f32 lo_, [%1]; ld.volatile.shared.f32 hi_, [%2]; mov.b64 %0, {lo_, hi_}; }"
                       : "=l"(wp[r]) : "r"(a0), "r"(a1));
        }
        auto pair = [&](int r) -> f2 { return wp[r]; };
        // backward (step t-1): M[ii][c][k] += s[h][c] * w[h + i0 + ii][k]                       _gpc.py:163
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int c = 0; c < M; ++c) {
#pragma unroll
            for (int pp = 0; pp < NPI; ++pp) Mp[pp][c][kk] = fma2(dup(sprev[h][c]), pair(h + 2 * pp), Mp[pp][c][kk]);
            if (TI) Ms[c][kk] = fmaf(sprev[h][c], lo(wp[h + HT - 1]), Ms[c][kk]);
          }
        if constexpr (INCR) {
          // forward, newest window only: V[H-1][c] += M[ii][c][k] * w[H + i0 + ii][k]; tap pairs against slot pairs
#pragma unroll
          for (int c = 0; c < M; ++c) {
#pragma unroll
            for (int pp = 0; pp < NPI; ++pp) acc[0][c] = fma2(Mp[pp][c][kk], wp[H + 2 * pp], acc[0][c]);
            if (TI) acct[c] = fmaf(Ms[c][kk], hi(wp[R - 2]), acct[c]);
          }
          // ring inner products for the last column of C_(t+1): qi = 1 holds slots H .. 2H-1 (in-terms against the
          // newest slot), qi = 0 holds slots 0 .. H-1 (out-terms against slot H-1)
          const f2 pv = dup(qi ? hi(wp[R - 2]) : lo(wp[H - 1]));
#pragma unroll
          for (int j = 0; j < H / 2; ++j) {
            const f2 src = qi ? wp[HT + 2 * j] : wp[2 * j];
            g2[j] = fma2(src, pv, g2[j]);
          }
        } else {
        // forward (step t): (V[2b], V[2b+1])[c] += M[ii][c][k] * (w'[2b + i], w'[2b + i + 1]),  w'[j] = w[j+1]
#pragma unroll
        for (int b = 0; b < NACC; ++b)
#pragma unroll
          for (int ii = 0; ii < HT; ++ii)
#pragma unroll
            for (int c = 0; c < M; ++c) acc[b][c] = fma2(dup(Mtap(ii, c, kk)), pair(2 * b + ii + 1), acc[b][c]);
        }
      }
    }
    if (t < p.T) {                                                    // (trip T: the pass applied the last update)
    head = (head + 1 == P) ? 0 : head + 1;                            // roll(-1)               _gpc.py:156-157
    float G[HS][DC][M];
    if constexpr (!INCR) load_G(G);

    // ---- window products: the tap halves swap the partial sums of each other's window subset, then the coordinate
    // halves are summed; lane (qc, qi) ends up with the complete V[h] of its subset h in [qi*HS, qi*HS + HS)
    float Vown[HS][M], mw[M];
    if constexpr (INCR) {
      // newest window, summed over the four lanes (every lane ends up with it: it is the action's M-part)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float v = (lo(acc[0][c]) + hi(acc[0][c])) + acct[c];
        v += __shfl_xor_sync(FULL, v, 1);
        v += __shfl_xor_sync(FULL, v, 2);
        mw[c] = v;
      }
      // ---- rows of the column ring this step touches.  The lane sums h' = a over [A0, A0 + HT) for the windows
      // b = h + 1 in [B0, B0 + HS); entry (mn, mx) of C_t is at row ((t + mn) mod H) * H + (mx - mn).
      constexpr int RS = H * EPW;                                      // floats per slot
      const int A0 = qc * HT, B0 = qi * HS + 1, D0 = B0 - A0;
      int offA[HT + 1], offB[HS];
#pragma unroll
      for (int x = 0; x <= HT; ++x) {
        int sl = cb + A0 + x;
        sl -= sl >= H ? H : 0;
        offA[x] = sl * RS;
      }
#pragma unroll
      for (int hh = 0; hh < HS; ++hh) {
        int sl = cb + B0 + hh;
        sl -= sl >= H ? H : 0;
        offB[hh] = sl * RS;
      }
      // ---- last column of C_(t+1) first (it frees the pass's accumulators): reduce the inner products over the
      // coordinate halves (each keeps its a half), exchange in- and out-terms between the tap halves, add to the old
      // column.  (No entry of C_t that this step reads lives where the new column goes: see the header.)
      {
        float gl[H], keep[HT];
#pragma unroll
        for (int j = 0; j < H / 2; ++j) { gl[2 * j] = lo(g2[j]); gl[2 * j + 1] = hi(g2[j]); }
#pragma unroll
        for (int x = 0; x < HT; ++x) {
          const float send = qc ? gl[x] : gl[HT + x];
          const float mine = qc ? gl[HT + x] : gl[x];
          keep[x] = mine + __shfl_xor_sync(FULL, send, 1);
        }
#pragma unroll
        for (int x = 0; x < HT; ++x) {
          const float other = __shfl_xor_sync(FULL, keep[x], 2);
          const float delta = qi ? keep[x] - other : other - keep[x];  // in - out
          const int pos = (H - 1 - A0 - x) * EPW;                      // entry (a, H-1), a = A0 + x; both tap halves
          colw[offA[x + 1] + pos] = colw[offA[x] + pos] + delta;       // store the same value (no divergent branch)
        }
      }
      // ---- V_t[h] = V_(t-1)[h+1] + sum_h' s_(t-1)[h'] C_t[h'][h+1] on the lane's windows, h' split over qc.
      // a <= b: row offA[x] + (b - a) = ua[x] + hh ; a >= b: row offB[hh] + (a - b) = ub[hh] + x.  Which order holds is
      // a compile-time property of (x, hh) on the diagonal lanes (x <= hh + 1 <=> a <= b) and fixed on the other two.
      const bool is10 = qc == 1 && qi == 0, is01 = qc == 0 && qi == 1;
      int QA[HS], QB[HS];
#pragma unroll
      for (int hh = 0; hh < HS; ++hh) {
        const int ub = offB[hh] - (D0 + hh) * EPW;
        QA[hh] = is10 ? ub : hh * EPW;
        QB[hh] = is01 ? hh * EPW : ub;
      }
      // the accumulators start from the shifted previous windows on qc = 0 and from zero on qc = 1, so that the sum
      // over the coordinate-half pair adds V_(t-1)[h+1] once; the subset's last window continues in the other tap half
      {
        const float* vnext = vsh + (qi * HS + 1) * M * EPW;            // V_(t-1)[h + 1], h = qi*HS + hh
#pragma unroll
        for (int hh = 0; hh < HS; ++hh)
#pragma unroll
          for (int c = 0; c < M; ++c) {
            const float nxt = vnext[(hh * M + c) * EPW];
            Vown[hh][c] = qc ? 0.f : nxt;
          }
      }
#pragma unroll
      for (int x = 0; x < HT; ++x) {
        const int ua = offA[x] + (D0 - x) * EPW;
        const int pa = is10 ? x * EPW : ua, pb = is01 ? ua : x * EPW;
        float shx[M];                                                  // s_(t-1)[A0 + x]
#pragma unroll
        for (int c = 0; c < M; ++c) shx[c] = qc ? sprev[HT + x][c] : sprev[x][c];
#pragma unroll
        for (int hh = 0; hh < HS; ++hh) {
          const float cv = colw[(x <= hh + 1) ? pa + QA[hh] : pb + QB[hh]];
#pragma unroll
          for (int c = 0; c < M; ++c) Vown[hh][c] = fmaf(shx[c], cv, Vown[hh][c]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < HS; ++hh)
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const float v = Vown[hh][c] + __shfl_xor_sync(FULL, Vown[hh][c], 1);
          Vown[hh][c] = (hh == HS - 1 && qi) ? mw[c] : v;              // h = H-1: the direct one
        }
      __syncwarp();                                                    // (all of V_(t-1) has been read)
      if (qc == 0) {
#pragma unroll
        for (int hh = 0; hh < HS; ++hh)
#pragma unroll
          for (int c = 0; c < M; ++c) vsh[((qi * HS + hh) * M + c) * EPW] = Vown[hh][c];
      }
    } else {
    {
      float v[H][M];
#pragma unroll
      for (int b = 0; b < HP; ++b)
#pragma unroll
        for (int c = 0; c < M; ++c) { v[2 * b][c] = lo(acc[b][c]); v[2 * b + 1][c] = hi(acc[b][c]); }
#pragma unroll
      for (int hh = 0; hh < HS; ++hh)
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const float mine = qi ? v[HS + hh][c] : v[hh][c];
          const float theirs = qi ? v[hh][c] : v[HS + hh][c];
          const float s = mine + __shfl_xor_sync(FULL, theirs, 2);
          Vown[hh][c] = s + __shfl_xor_sync(FULL, s, 1);
        }
#pragma unroll
      for (int c = 0; c < M; ++c) mw[c] = __shfl_sync(FULL, Vown[HS - 1][c], lane | 2);   // V[H-1] lives in the upper tap half
    }
    }
    if constexpr (INCR) load_G(G);                                    // (after the correction block: 75 registers)
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = mw[c] - kx[c];                 // get_action              _gpc.py:171-183

    // ---- noise = x - A x_prev - B u ; append as the newest slot                              _gpc.py:155-157
    {
      float un[M], nz[DC];
#pragma unroll
      for (int c = 0; c < M; ++c) un[c] = p.noise_uses_prev_action ? uprev[c] : u[c];
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) nz[kk] = 0.f;
      B_acc(un, nz);
      int sn = head + P - 1;
      sn -= sn >= P ? P : 0;
      float* newest = rown + sn * S1;                                 // (both tap halves store the same values)
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) newest[kk * EPW] = nzp[kk] - nz[kk];
    }
    __syncwarp();

    // ---- y_f = q + sum_h G_(H-2-h) v_h ;  u_f = -K y_f + mw ; du = 2 u_f ; lam = 2 y_f - K' du   _gpc.py:109-125
    const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
    float sown[HS][M], du[M];
    {
      float yf[DC];
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
        float sc[M];
#pragma unroll
        for (int c = 0; c < M; ++c) sc[c] = G[0][kk][c] * Vown[0][c];
#pragma unroll
        for (int hh = 1; hh < HS; ++hh)
#pragma unroll
          for (int c = 0; c < M; ++c) sc[c] = fmaf(G[hh][kk][c], Vown[hh][c], sc[c]);
        float s = sc[0];
#pragma unroll
        for (int c = 1; c < M; ++c) s += sc[c];
        yf[kk] = (s + __shfl_xor_sync(FULL, s, 2)) + qrow[kk];
      }
      float ky[M], lam[DC];
      K_dot(yf, ky);
#pragma unroll
      for (int c = 0; c < M; ++c) du[c] = 2.0f * (mw[c] - ky[c]);
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
        float s = 2.0f * yf[kk];
#pragma unroll
        for (int c = 0; c < M; ++c) s = fmaf(-Kc[c][kk], du[c], s);
        lam[kk] = s;
      }
      // s[h] = -lr G_(H-2-h)' lam on the lane's window subset, summed over the coordinate halves       SURVEY A-3.5
#pragma unroll
      for (int hh = 0; hh < HS; ++hh)
#pragma unroll
        for (int c = 0; c < M; ++c) {
          float s = G[hh][0][c] * lam[0];
#pragma unroll
          for (int kk = 1; kk < DC; ++kk) s = fmaf(G[hh][kk][c], lam[kk], s);
          sown[hh][c] = s;
        }
    }
#pragma unroll
    for (int hh = 0; hh < HS; ++hh)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const float s = sown[hh][c] + __shfl_xor_sync(FULL, sown[hh][c], 1);
        sown[hh][c] = -lr * ((hh == HS - 1 && qi) ? du[c] : s);       // window H-1: s = -lr du (the final action)
      }
    // every lane needs every window's s for the next pass: the tap halves exchange their subsets
#pragma unroll
    for (int hh = 0; hh < HS; ++hh)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const float other = __shfl_xor_sync(FULL, sown[hh][c], 2);
        sprev[hh][c] = qi ? other : sown[hh][c];
        sprev[HS + hh][c] = qi ? sown[hh][c] : other;
      }
#pragma unroll
    for (int c = 0; c < M; ++c) uprev[c] = u[c];

    // ---- realised cost x'x + u'u                                                             _gpc.py:29-40
    float cost;
    {
      float s = xr[0] * xr[0];
#pragma unroll
      for (int kk = 1; kk < DC; ++kk) s = fmaf(xr[kk], xr[kk], s);
      s += __shfl_xor_sync(FULL, s, 1);
      float cu = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
      cost = s + cu;
    }
    csum += (double)cost;
    if (live && q == 0 && a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (live && qi == 0 && t == p.T - 1) {
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) a.xprev_io[n * D + k0 + kk] = xr[kk];
    }
    if (sampled) {
      if (a.x_out && qi == 0) {
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) a.x_out[((int64_t)t * Ns + ns) * D + k0 + kk] = xr[kk];
      }
      if (a.u_out && q == 0) {
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
      }
    }

    // ---- env: x <- (A x + B u) + w_t = (Abar x + B mw) + w_t                                   _lds.py:102-111
    B_acc(mw, z);
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) xr[kk] = z[kk] + wt[kk];
    cb = (cb + 1 == H) ? 0 : cb + 1;
    }  // t < T
  }

  // ---- write back
  if (!live) return;
  if (qi == 0) {
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) a.x_io[n * D + k0 + kk] = xr[kk];
  }
  if (q == 0) {
    if (a.cost_sum_out) a.cost_sum_out[n] = csum;
    if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
    if (a.uprev_io) {
#pragma unroll
      for (int c = 0; c < M; ++c) a.uprev_io[n * M + c] = uprev[c];
    }
  }
  {
    float* gM = a.M_io + n * (H * M * D);
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
#pragma unroll
        for (int pp = 0; pp < NPI; ++pp) {
          gM[((i0 + 2 * pp) * M + c) * D + k0 + kk] = lo(Mp[pp][c][kk]);
          gM[((i0 + 2 * pp + 1) * M + c) * D + k0 + kk] = hi(Mp[pp][c][kk]);
        }
        if (TI) gM[((i0 + HT - 1) * M + c) * D + k0 + kk] = Ms[c][kk];
      }
  }
  __syncwarp();
  {
    float* gh = a.hist_io + n * (P * D);
    for (int idx = q; idx < P * D; idx += 4) {
      const int s = idx / D, k = idx - s * D;
      int ph = head + s;
      ph = ph >= P ? ph - P : ph;
      gh[s * D + k] = ring[ph * S1 + k * EPW];
    }
  }
}

template <int D, int M, int H, int TPB, int MINB, int WSRC, bool INCR>
static int32_t launch_quad_w(const LdsArgs& a, cudaStream_t st, bool one_cta_per_sm = false) {
  using Lay = QuadLayout<D, M, H, TPB>;
  constexpr size_t kBytes = INCR ? Lay::kSmemBytesIncr : Lay::kSmemBytes;
  static_assert((kBytes + 1024) * MINB <= 228 * 1024, "ring + operators do not fit shared memory");
  auto kern = lds_gpc_quad_kernel<D, M, H, TPB, MINB, WSRC, INCR>;
  // (measurement only: asking for more than half of the SM's shared memory keeps a second CTA off the SM)
  const size_t smem = one_cta_per_sm && kBytes < 120 * 1024 ? 120 * 1024 : kBytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_gpc_quad)");
  const int64_t threads = a.p.N * 4;
  const unsigned grid = (unsigned)((threads + TPB - 1) / TPB);
  kern<<<grid, TPB, smem, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_gpc_quad_kernel launch");
}

template <int D, int M, int H, int TPB, int MINB, bool INCR = false>
static int32_t launch_quad(const LdsArgs& a, cudaStream_t st, bool one_cta_per_sm = false) {
  return a.W ? launch_quad_w<D, M, H, TPB, MINB, 1, INCR>(a, st, one_cta_per_sm)
             : launch_quad_w<D, M, H, TPB, MINB, 2, INCR>(a, st, one_cta_per_sm);
}

// DLC_QUAD_INCR=1 makes kernel_variant 0 the incremental-window kernel (variant 17) instead of the direct one (13).
// Measured (B200, 262,144 envs x 1000 steps): the incremental form executes 27 % fewer FMA-pipe cycles but MORE
// instructions (ring addressing, lane-role selects, extra shared-memory traffic) and is slower, so it is not the default.
static bool quad_incr_default() {
  static const bool on = [] { const char* e = getenv("DLC_QUAD_INCR"); return e && atoi(e) != 0; }();
  return on;
}

// kernel_variant 0 (automatic) or 13..16 (direct window products; launch shapes 128 x 2, 64 x 4, 256 x 1, 128 x 1)
// or 17 (incremental window products, 128 x 2)
bool launch_gpc_quad(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  if (p.policy != DLC_POLICY_GPC || p.mode != DLC_MODE_CLOSED_LOOP) return false;
  if (!(p.kernel_variant == 0 || (p.kernel_variant >= 13 && p.kernel_variant <= 17))) return false;
  if (p.d == 10 && p.m == 3 && p.H == 10 && p.HH == 10) {
    switch (p.kernel_variant) {
      case 14: *rc = launch_quad<10, 3, 10, 64, 4>(a, st); break;
      case 15: *rc = launch_quad<10, 3, 10, 256, 1>(a, st); break;
      case 16: *rc = launch_quad<10, 3, 10, 128, 2>(a, st, true); break;   // one warp per scheduler (measurement only)
      case 13: *rc = launch_quad<10, 3, 10, 128, 2>(a, st); break;
      case 17: *rc = launch_quad<10, 3, 10, 128, 2, true>(a, st); break;
      default:
        *rc = quad_incr_default() ? launch_quad<10, 3, 10, 128, 2, true>(a, st) : launch_quad<10, 3, 10, 128, 2>(a, st);
        break;
    }
    return true;
  }
  return false;
}

}  // namespace dlc
