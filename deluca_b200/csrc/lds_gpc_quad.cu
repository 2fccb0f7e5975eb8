// LDS + GPC rollout for long even histories (H = HH, the north-star target d = 10, m = 3, H = HH = 10):
// four lanes own an env and split it TWO ways -- by state coordinate (halves S_0, S_1 of D/2) and by GPC tap
// (halves I_0, I_1 of H/2) -- so that every lane holds exactly a quarter of everything and nothing is padded:
//
//     lane (qc, qi):   M[i in I_qi][c][k in S_qc]   (H/2 * m * D/2 floats: 75)
//                      (A - B K)[rows S_qc][cols S_qi]   (25),  B[rows S_qc][:] (15),  K[:][cols S_qc] (15)
//
// (lds_f2_kernel's split of 10 coordinates over 4 lanes as 3 + 3 + 3 + 1 pads every contraction 12-for-10.)
//
// The M-side work of a step -- the rank-H update of step t-1 and the H window products of step t,
//     M[i] += sum_h s[h] (x) w[h + i]        V[h] = sum_i M[i] w'[h + i],   w'[j] = w[j + 1] (the ring rolled by one)
// -- is ONE pass over the noise ring: for each own coordinate the 14 slot pairs a lane needs are loaded once
// (LDS.64 of (w[j], w[j+1]), the ring stores every element with its successor), M[.][.][k] is updated in registers
// and immediately used for the window products of the next step.  Both sides are packed FFMA2 (tap pairs as
// accumulators backward, window pairs forward); the partial window products are summed over the four lanes.
//
// The two recursions of policy_loss are latency chains (2 (H - 2) dependent mat-vecs per step), so they are kept short:
// lane (qc, qi) multiplies with the block Abar[S_qc][S_qi] when the vector arrives split by columns (S_qi) and with the
// block Abar[S_qi][S_qc] -- its partner lane's block -- when it arrives split by rows (S_qc); the result of either
// product is a 1-bit xor-reduce away from the input layout of the other, so consecutive steps alternate between the
// two and no vector is ever exchanged.  The term B v_h + w[h+H], which does not depend on the recursion state, is the
// initial value of the dot-product chains on the two diagonal lanes.  The blocks live in shared memory and are loaded
// into registers only for the recursion phase, which keeps the M pass free of spills.
//
// Reference arithmetic: deluca/agents/_gpc.py:109-183 (policy_loss, update, get_action), deluca/envs/_lds.py:102-111.
#include "lds_args.cuh"
#include "f2.cuh"

#pragma nv_diag_suppress 550  // unused half of a mov.b64 unpack

namespace dlc {

template <int D, int M, int H, int TPB>
struct QuadLayout {
  static constexpr int DC = D / 2, HT = H / 2, P = 2 * H, EPW = 8, WARPS = TPB / 32;
  // float2 per ring slot of one warp.  The two tap halves of an env read slots H/2 apart in the same instruction; the
  // pad puts those reads on different banks (H/2 * S * 2 words = 8 mod 32 for D = 10, H = 10)
  static constexpr int S = D * EPW + 4;
  static constexpr size_t kRingBytes = (size_t)WARPS * P * S * sizeof(float2);
  static constexpr size_t kScratchBytes = (size_t)H * M * TPB * sizeof(float);
  static constexpr int NPAR = DC * DC + 2 * DC * M;                   // Abar block, B rows, K columns of one lane
  static constexpr size_t kParamBytes = (size_t)NPAR * TPB * sizeof(float);
  static constexpr size_t kSmemBytes = kRingBytes + kScratchBytes + kParamBytes;
};

// dot product of two DC-vectors as two interleaved chains (shorter dependency chain than one), on top of `init`
template <int DC>
__device__ __forceinline__ float dot_chain(const float* a, const float* v, float init) {
  float p0 = fmaf(a[0], v[0], init);
  if (DC == 1) return p0;
  float p1 = a[1] * v[1];
#pragma unroll
  for (int i = 2; i < DC; ++i) {
    if (i & 1) p1 = fmaf(a[i], v[i], p1);
    else p0 = fmaf(a[i], v[i], p0);
  }
  return p0 + p1;
}

// arrivals per SM: every second CTA that lands on an SM starts half a step late (see `phase_delay` in the kernel)
__device__ unsigned int g_quad_arrivals[1024];

// WSRC: 1 = disturbances from a.W, 2 = in-kernel Philox stream
template <int D, int M, int H, int TPB, int MINB, int WSRC>
__global__ void __launch_bounds__(TPB, MINB) lds_gpc_quad_kernel(const LdsArgs a) {
  using Lay = QuadLayout<D, M, H, TPB>;
  constexpr int DC = Lay::DC, HT = Lay::HT, HP = H / 2, P = Lay::P, EPW = Lay::EPW, S = Lay::S;
  constexpr int NPI = HT / 2, TI = HT & 1, NPIA = NPI > 0 ? NPI : 1;   // tap pairs of a tap half
  constexpr int R = H + HT - 1;                                        // ring slots one lane touches per pass
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
  float2* const ring = reinterpret_cast<float2*>(smraw) + (size_t)warp * P * S + e;   // entry(slot, k) = ring[slot*S + k*EPW]
  float2* const rown = ring + k0 * EPW;                                               // own coordinates
  float* const sV = reinterpret_cast<float*>(smraw + Lay::kRingBytes) + tid;          // scratch: V[h][c] then s[h][c]
  float* const sPar = reinterpret_cast<float*>(smraw + Lay::kRingBytes + Lay::kScratchBytes) + tid;
  const float* const sParT = sPar - lane + partner;                                   // the partner lane's blocks
  constexpr int oB = DC * DC, oK = DC * DC + DC * M;       // sPar[idx * TPB]: Abar block [kk][jj] | B rows [kk][c] | K cols [c][kk]
  constexpr unsigned FULL = 0xffffffffu;

  // ---- parameters -> shared memory.  Abar = A - B K (the counterfactual recursion is y <- Abar y + B v + w and its
  // adjoint Abar' lam; the env step recovers A x = Abar x + B (K x)).
  {
    const int64_t np = p.shared_dynamics ? 0 : n;
    const float* gA = a.A + np * (D * D);
    const float* gB = a.B + np * (D * M);
    const float* gK = a.K + np * (M * D);
    float Br0[DC][M];
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        Br0[kk][c] = __ldg(gB + (k0 + kk) * M + c);
        sPar[(oB + kk * M + c) * TPB] = Br0[kk][c];
        sPar[(oK + c * DC + kk) * TPB] = __ldg(gK + c * D + k0 + kk);
      }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) {
        float v = __ldg(gA + (k0 + kk) * D + j0 + jj);
#pragma unroll
        for (int c = 0; c < M; ++c) v = fmaf(-Br0[kk][c], __ldg(gK + c * D + j0 + jj), v);
        sPar[(kk * DC + jj) * TPB] = v;
      }
  }
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
  // noise ring: every element with its successor, (w[s][k], w[s+1][k]); the four lanes of an env fill it together
  {
    const float* gh = a.hist_io + n * (P * D);
    for (int idx = q; idx < P * D; idx += 4) {
      const int s = idx / D, k = idx - s * D;
      ring[s * S + k * EPW] = make_float2(gh[s * D + k], s + 1 < P ? gh[(s + 1) * D + k] : 0.f);
    }
  }
  __syncwarp();

  // the blocks of the recursion phase, loaded from shared memory where they are needed
  struct Blocks { float Ab[DC][DC], AbT[DC][DC], Br[DC][M], Kc[M][DC]; };
  auto load_blocks = [&](Blocks& b) {
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) {
        b.Ab[kk][jj] = sPar[(kk * DC + jj) * TPB];                     // Abar[k0 + kk][j0 + jj]
        b.AbT[kk][jj] = sParT[(kk * DC + jj) * TPB];                   // Abar[j0 + kk][k0 + jj]
      }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        b.Br[kk][c] = sPar[(oB + kk * M + c) * TPB];
        b.Kc[c][kk] = sPar[(oK + c * DC + kk) * TPB];
      }
  };
  // y[S_qc] = init + Abar[S_qc][S_qi] vc   (vc split by columns)  -> summed over qi: split by rows
  auto mul_cols_to_rows = [&](const Blocks& b, const float vc[DC], const float init[DC], float out[DC]) {
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) out[kk] = dot_chain<DC>(b.Ab[kk], vc, init[kk]);
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) out[kk] += __shfl_xor_sync(FULL, out[kk], 2);
  };
  // y[S_qi] = init + Abar[S_qi][S_qc] vr   (vr split by rows)     -> summed over qc: split by columns
  auto mul_rows_to_cols = [&](const Blocks& b, const float vr[DC], const float init[DC], float out[DC]) {
#pragma unroll
    for (int jj = 0; jj < DC; ++jj) out[jj] = dot_chain<DC>(b.AbT[jj], vr, init[jj]);
#pragma unroll
    for (int jj = 0; jj < DC; ++jj) out[jj] += __shfl_xor_sync(FULL, out[jj], 1);
  };
  // adjoint products: lam'[S_qi] = Abar[S_qc][S_qi]' lam[S_qc] (rows -> columns), lam'[S_qc] = Abar[S_qi][S_qc]' lam[S_qi]
  auto mulT_rows_to_cols = [&](const Blocks& b, const float vr[DC], float out[DC]) {
#pragma unroll
    for (int jj = 0; jj < DC; ++jj) {
      float col[DC];
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) col[kk] = b.Ab[kk][jj];
      out[jj] = dot_chain<DC>(col, vr, 0.f);
    }
#pragma unroll
    for (int jj = 0; jj < DC; ++jj) out[jj] += __shfl_xor_sync(FULL, out[jj], 1);
  };
  auto mulT_cols_to_rows = [&](const Blocks& b, const float vc[DC], float out[DC]) {
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) {
      float col[DC];
#pragma unroll
      for (int jj = 0; jj < DC; ++jj) col[jj] = b.AbT[jj][kk];
      out[kk] = dot_chain<DC>(col, vc, 0.f);
    }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) out[kk] += __shfl_xor_sync(FULL, out[kk], 2);
  };
  auto K_dot = [&](const Blocks& b, const float vr[DC], float out[M]) {   // K v (v split by rows), summed over qc
#pragma unroll
    for (int c = 0; c < M; ++c) out[c] = dot_chain<DC>(b.Kc[c], vr, 0.f);
#pragma unroll
    for (int c = 0; c < M; ++c) out[c] += __shfl_xor_sync(FULL, out[c], 1);
  };
  auto B_acc = [&](const Blocks& b, const float v[M], float acc[DC]) {     // acc[S_qc] += B[S_qc][:] v
#pragma unroll
    for (int kk = 0; kk < DC; ++kk)
#pragma unroll
      for (int c = 0; c < M; ++c) acc[kk] = fmaf(b.Br[kk][c], v[c], acc[kk]);
  };
  float zero[DC];
#pragma unroll
  for (int kk = 0; kk < DC; ++kk) zero[kk] = 0.f;

  // ---- state
  float xr[DC], ax[DC], uprev[M];
#pragma unroll
  for (int kk = 0; kk < DC; ++kk) xr[kk] = a.x_io[n * D + k0 + kk];
  {
    Blocks b;
    load_blocks(b);
    float xpr[DC], xpc[DC], kxp[M];
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) { xpr[kk] = a.xprev_io[n * D + k0 + kk]; xpc[kk] = a.xprev_io[n * D + j0 + kk]; }
    mul_cols_to_rows(b, xpc, zero, ax);                               // Abar x_prev ...
    K_dot(b, xpr, kxp);
    B_acc(b, kxp, ax);                                                // ... + B (K x_prev) = A x_prev
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
#pragma unroll
  for (int r = 0; r < H * M; ++r) sV[r * TPB] = 0.f;                   // s of "step -1": no update in the first pass

  // A step alternates between a phase that saturates the FMA pipe (the pass) and a phase of dependent mat-vecs that
  // mostly waits (the recursions).  All warps run identical steps, so the two warps that share a scheduler stay in
  // whatever relative phase they started with -- and they start together.  Delaying every second arrival on an SM by
  // about half a step puts one warp's pass under the other's recursions for the whole launch.
  if (a.phase_delay > 0) {
    __shared__ unsigned int s_arrival;
    if (tid == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_arrival = atomicAdd(&g_quad_arrivals[smid & 1023], 1u);
    }
    __syncthreads();
    const bool late = TPB >= 256 ? (((warp >> 2) ^ s_arrival) & 1) : (s_arrival & 1);
    if (late) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)a.phase_delay) { }
    }
  }
  int head = 0;
  // T + 1 trips: the pass at the top of trip t applies step t-1's update, so trip T only runs the pass.  (Written as a
  // counted loop with a guarded body so that the compiler does not rotate the loop and duplicate the 11 KB pass.)
#pragma unroll 1
  for (int t = 0; t <= p.T; ++t) {
    // ================= one pass over the ring: M += sum_h s[h] (x) w[h+i];  V[h] = sum_i M[i] w[h+i+1]
    f2 acc[HP][M];
    {
      float sprev[H * M];
#pragma unroll
      for (int r = 0; r < H * M; ++r) sprev[r] = sV[r * TPB];
#pragma unroll
      for (int b = 0; b < HP; ++b)
#pragma unroll
        for (int c = 0; c < M; ++c) acc[b][c] = pk(0.f, 0.f);
      const float2* rlo = rown + (head + i0) * S;                     // logical slot i0 + r lives at rlo + r*S ...
      const float2* rhi = rlo - P * S;                                // ... or, past the wrap, at rhi + r*S
      const int rw = P - head - i0;
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) {
        f2 w2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 v = ((r < rw) ? rlo : rhi)[r * S + kk * EPW];
          w2[r] = pk(v.x, v.y);
        }
        // backward (step t-1): M[ii][c][k] += s[h][c] * w[h + i0 + ii][k]                       _gpc.py:163
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int c = 0; c < M; ++c) {
#pragma unroll
            for (int pp = 0; pp < NPI; ++pp) Mp[pp][c][kk] = fma2(dup(sprev[h * M + c]), w2[h + 2 * pp], Mp[pp][c][kk]);
            if (TI) Ms[c][kk] = fmaf(sprev[h * M + c], lo(w2[h + HT - 1]), Ms[c][kk]);
          }
        // forward (step t): (V[2b], V[2b+1])[c] += M[ii][c][k] * (w'[2b + i], w'[2b + i + 1]),  w'[j] = w[j+1]
#pragma unroll
        for (int b = 0; b < HP; ++b)
#pragma unroll
          for (int ii = 0; ii < HT; ++ii)
#pragma unroll
            for (int c = 0; c < M; ++c) acc[b][c] = fma2(dup(Mtap(ii, c, kk)), w2[2 * b + ii + 1], acc[b][c]);
      }
    }
    if (t < p.T) {                                                    // (trip T: the pass applied the last update)
    head = (head + 1 == P) ? 0 : head + 1;                            // roll(-1)               _gpc.py:156-157
    const float2* const nlo = rown + head * S;                        // new logical slot j at nlo + j*S (j < nw) ...
    const float2* const nhi = nlo - P * S;                            // ... else at nhi + j*S
    const int nw = P - head;
    auto w_at = [&](int j, int kk) -> float {                         // w[j][own coordinate kk], new indexing
      return reinterpret_cast<const float*>(((j < nw) ? nlo : nhi) + j * S + kk * EPW)[0];
    };
    float wt[DC];
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) wt[kk] = wnext[kk];
    if (t + 1 < p.T) load_w(t + 1, wnext);
    Blocks b;
    load_blocks(b);

    // ---- sum the partial window products over the four lanes
    float V0[M], mw[M];
#pragma unroll
    for (int bb = 0; bb < HP; ++bb)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float v0 = lo(acc[bb][c]), v1 = hi(acc[bb][c]);
        v0 += __shfl_xor_sync(FULL, v0, 1);
        v1 += __shfl_xor_sync(FULL, v1, 1);
        v0 += __shfl_xor_sync(FULL, v0, 2);
        v1 += __shfl_xor_sync(FULL, v1, 2);
        if (bb == 0) V0[c] = v0; else sV[((2 * bb) * M + c) * TPB] = v0;
        if (bb == HP - 1) mw[c] = v1; else sV[((2 * bb + 1) * M + c) * TPB] = v1;
      }
    float kx[M], u[M];
    K_dot(b, xr, kx);                                                 // K x                     _lqr.py:72
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = mw[c] - kx[c];                 // get_action              _gpc.py:171-183

    // ---- noise = x - A x_prev - B u ; append as the newest slot                              _gpc.py:155-157
    {
      float un[M], nz[DC];
#pragma unroll
      for (int c = 0; c < M; ++c) un[c] = p.noise_uses_prev_action ? uprev[c] : u[c];
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) nz[kk] = 0.f;
      B_acc(b, un, nz);
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) nz[kk] = (xr[kk] - ax[kk]) - nz[kk];
      if (qi == 0) {
        float* newest = reinterpret_cast<float*>(const_cast<float2*>(((P - 1 < nw) ? nlo : nhi) + (P - 1) * S));      // .x
        float* second = reinterpret_cast<float*>(const_cast<float2*>(((P - 2 < nw) ? nlo : nhi) + (P - 2) * S)) + 1;  // .y
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) {
          newest[2 * kk * EPW] = nz[kk];
          second[2 * kk * EPW] = nz[kk];
        }
      }
    }
    __syncwarp();

    // ---- policy_loss forward: y <- Abar y + (B v_h + w[h+H])                                 _gpc.py:118-124
    // zb(h) = B v_h + w[h+H] on the own rows; only the two diagonal lanes feed it into the reduced sums
    auto zb_of = [&](int h, const float v[M], float zb[DC]) {
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) zb[kk] = w_at(h + H, kk);
      B_acc(b, v, zb);
    };
    float yr[DC];
    zb_of(0, V0, yr);                                                 // y_1 = B v_0 + w[H]  (y_0 = 0), split by rows
#pragma unroll 1
    for (int h = 1; h < H - 1; h += 2) {
      float v1[M], v2[M], z1[DC], z2[DC], yc[DC];
#pragma unroll
      for (int c = 0; c < M; ++c) { v1[c] = sV[(h * M + c) * TPB]; v2[c] = sV[((h + 1) * M + c) * TPB]; }
      zb_of(h, v1, z1);
      zb_of(h + 1, v2, z2);
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) { z1[kk] = diag ? z1[kk] : 0.f; z2[kk] = diag ? z2[kk] : 0.f; }
      mul_rows_to_cols(b, yr, z1, yc);                                // y_{h+1}, split by columns
      mul_cols_to_rows(b, yc, z2, yr);                                // y_{h+2}, split by rows
    }
    // ---- u_f = -K y + mw ; du = 2 u_f ; lam = 2 y - K' du                                    SURVEY A-3.5
    const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
    float du[M], lam[DC];
    {
      float ky[M];
      K_dot(b, yr, ky);
#pragma unroll
      for (int c = 0; c < M; ++c) du[c] = 2.0f * (mw[c] - ky[c]);
    }
#pragma unroll
    for (int kk = 0; kk < DC; ++kk) {
      float s = 2.0f * yr[kk];
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(-b.Kc[c][kk], du[c], s);
      lam[kk] = s;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) sV[((H - 1) * M + c) * TPB] = -lr * du[c];
    // ---- adjoint recursion: s[h] = -lr B' lam_h ; lam_{h-1} = Abar' lam_h, alternating layouts
    auto s_from_rows = [&](int h, const float lr_[DC]) {              // lam split by rows: every lane has its B rows
      float lu[M];
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float col[DC];
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) col[kk] = b.Br[kk][c];
        lu[c] = dot_chain<DC>(col, lr_, 0.f);
      }
#pragma unroll
      for (int c = 0; c < M; ++c) lu[c] += __shfl_xor_sync(FULL, lu[c], 1);
#pragma unroll
      for (int c = 0; c < M; ++c) sV[(h * M + c) * TPB] = -lr * lu[c];
    };
    auto s_from_cols = [&](int h, const float lc[DC]) {               // lam split by columns: the diagonal lanes' B rows
      float lu[M];
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float col[DC];
#pragma unroll
        for (int kk = 0; kk < DC; ++kk) col[kk] = b.Br[kk][c];
        const float v = dot_chain<DC>(col, lc, 0.f);
        lu[c] = diag ? v : 0.f;
      }
#pragma unroll
      for (int c = 0; c < M; ++c) lu[c] += __shfl_xor_sync(FULL, lu[c], 1);
#pragma unroll
      for (int c = 0; c < M; ++c) lu[c] += __shfl_xor_sync(FULL, lu[c], 2);
#pragma unroll
      for (int c = 0; c < M; ++c) sV[(h * M + c) * TPB] = -lr * lu[c];
    };
#pragma unroll 1
    for (int h = H - 2; h >= 2; h -= 2) {
      float lc[DC];
      s_from_rows(h, lam);
      mulT_rows_to_cols(b, lam, lc);                                  // lam_{h-1}, split by columns
      s_from_cols(h - 1, lc);
      mulT_cols_to_rows(b, lc, lam);                                  // lam_{h-2}, split by rows
    }
    s_from_rows(0, lam);
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

    // ---- env: x <- (A x + B u) + w_t = (Abar x + B mw) + w_t;  A x = Abar x + B (K x)          _lds.py:102-111
    {
      float xc[DC], z[DC];
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) xc[kk] = __shfl_sync(FULL, xr[kk], partner);   // split by columns
      mul_cols_to_rows(b, xc, zero, z);
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) ax[kk] = z[kk];
      B_acc(b, kx, ax);
      B_acc(b, mw, z);
#pragma unroll
      for (int kk = 0; kk < DC; ++kk) xr[kk] = z[kk] + wt[kk];
    }
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
      gh[s * D + k] = ring[ph * S + k * EPW].x;
    }
  }
}

template <int D, int M, int H, int TPB, int MINB, int WSRC>
static int32_t launch_quad_w(const LdsArgs& a, cudaStream_t st, bool one_cta_per_sm = false) {
  using Lay = QuadLayout<D, M, H, TPB>;
  static_assert((Lay::kSmemBytes + 1024) * MINB <= 228 * 1024, "ring + scratch + blocks do not fit shared memory");
  auto kern = lds_gpc_quad_kernel<D, M, H, TPB, MINB, WSRC>;
  // (measurement only: asking for more than half of the SM's shared memory keeps a second CTA off the SM)
  const size_t smem = one_cta_per_sm && Lay::kSmemBytes < 120 * 1024 ? 120 * 1024 : Lay::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_gpc_quad)");
  const int64_t threads = a.p.N * 4;
  const unsigned grid = (unsigned)((threads + TPB - 1) / TPB);
  kern<<<grid, TPB, smem, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_gpc_quad_kernel launch");
}

template <int D, int M, int H, int TPB, int MINB>
static int32_t launch_quad(const LdsArgs& a, cudaStream_t st, bool one_cta_per_sm = false) {
  return a.W ? launch_quad_w<D, M, H, TPB, MINB, 1>(a, st, one_cta_per_sm) : launch_quad_w<D, M, H, TPB, MINB, 2>(a, st, one_cta_per_sm);
}

// kernel_variant 0 (automatic) or 13..15 (launch shapes 128 x 2, 64 x 4, 256 x 1)
bool launch_gpc_quad(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  if (p.policy != DLC_POLICY_GPC || p.mode != DLC_MODE_CLOSED_LOOP) return false;
  if (!(p.kernel_variant == 0 || (p.kernel_variant >= 13 && p.kernel_variant <= 16))) return false;
  if (p.d == 10 && p.m == 3 && p.H == 10 && p.HH == 10) {
    switch (p.kernel_variant) {
      case 14: *rc = launch_quad<10, 3, 10, 64, 4>(a, st); break;
      case 15: *rc = launch_quad<10, 3, 10, 256, 1>(a, st); break;
      case 16: *rc = launch_quad<10, 3, 10, 128, 2>(a, st, true); break;   // one warp per scheduler (measurement only)
      default: *rc = launch_quad<10, 3, 10, 128, 2>(a, st); break;
    }
    return true;
  }
  return false;
}

}  // namespace dlc
