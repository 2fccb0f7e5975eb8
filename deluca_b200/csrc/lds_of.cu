// lds_of.cu -- partially observed LDS (y = C x) driven by deluca's output-feedback controllers, fused over T steps
// (SURVEY 8(f) rank 3).
//
//   LQG   deluca/agents/_lqg.py:83-98     u = -K x_hat;  x_hat <- A x_hat + B u + L (y - C x_hat)
//   DRC   deluca/agents/_drc.py:131-203   y_nat <- y - sum_i G[i] us[i];  u = -K y + sum_j M[j] y_nat[j];
//                                         M <- clip_RM(M - lr * d policy_loss / dM)
//   GRC   deluca/agents/_grc.py:90-169    }  one family: the action is linear in the parameters and in a fixed bank
//   SFC   deluca/agents/_sfc.py:130-214   }  of linear filters of the y_nat history,
//   DSC   deluca/agents/_dsc.py:150-292   }      a(k) = sum_f Theta[f] phi_f(k),  phi_f(k) = sum_o Phi[f,o] ynat[k+o],
//                                            with (F, L, Phi) chosen by the host mirror (oracle/of.py filter_bank)
//   env   deluca/envs/_lds.py:102-111     x <- A x + B u + w,  y = C x;   cost = quad_loss(y, u)  (_grc.py:27-39)
//
// Mapping: a group of 8, 16 or 32 LANES owns an env (chosen from the contraction widths; see of_kernel).  Everything the env needs between steps (A, B, C, the Markov operators, the
// parameters, the history rings) stays in the warp's slice of shared memory for all T steps; lanes split the
// elements of each small contraction and meet through warp shuffles / __syncwarp.  Runtime dims.
//
// Algebra that is exact, not approximate:
//   (i)  policy_loss rolls delta <- A delta + B a(k) for k < m and reads only C delta_m (_grc.py:97-105).  That is
//        sum_k (C A^(m-1-k) B) a(k): the m Markov operators are formed once per call, so the m-step recursion and
//        its adjoint become 2 m p n FMAs per step instead of 2 m (d^2 + d n)  (DRC keeps G the same way, _drc.py:97-102).
//   (ii) the filtered windows phi_f(k), k = 0..m, of step t are those of step t-1 shifted by one: only the newest
//        window is filtered each step, the others are kept in a ring.
//   (iii) jit(grad(policy_loss)) is a hand adjoint: dTheta[f] = sum_k g(k) phi_f(k)^T with g(m) = 2 a(m) and
//        g(k) = G[m-1-k]^T (2 final_y) for k < m.
#include <stdlib.h>

#include "common.cuh"
#include <type_traits>

#include "f2.cuh"
#include "row_ring.cuh"

namespace dlc {
namespace {

struct OfArgs {
  DlcOfParams p;
  const float *A, *B, *C, *K, *Lk, *Phi, *W, *y_in;
  float *x_io, *theta_io, *z_io, *ynat_io, *us_io;
  float* cost_out;
  double* cost_sum_out;
  float* u_out;
  int32_t* status_out;
  int wpc;   // envs (lane groups) per CTA
  int S;     // floats of shared memory per env
  int J;     // Markov operators kept (m for the filter family, h for DRC)
  int FP;    // length of one feature window: F * p (filter family), m * p (DRC)
  int nPhi;  // CTA-wide floats at the start of dynamic shared memory (the filter bank)
  int oA, oB, oC, oG, oK, oL, oTh, oFeat, oYn, oUs, oX, oZ, oY, oU, oAct, oGact, oFy, oTmp;
};

// sum over the GL lanes that own an env (GL = 32: the whole warp)
template <int GL>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = GL / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// e / den for 0 <= e < 2^20 (inv = 1 / den): one multiply and a conversion instead of an integer division
__device__ __forceinline__ int fast_div(int e, float inv) { return (int)(((float)e + 0.5f) * inv); }

// r[k] (+)= sum_c Mat[row][c] v[c] for row = gl + k GL < rows (rows <= 2 GL); the caller decides where the result goes.
// ACC continues the accumulation in r (A z + B u + L e in one register per row).
template <int GL, bool ACC = false>
__device__ __forceinline__ void matvec(const float* Mat, const float* v, int rows, int cols, int gl, float (&r)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (!ACC) r[k] = 0.f;
    if (k * GL < rows) {   // warp-uniform
      const int row = gl + k * GL;
      if (row < rows) {
        const float* mr = Mat + row * cols;
        float acc = r[k];
        for (int c = 0; c < cols; ++c) acc = fmaf(mr[c], v[c], acc);
        r[k] = acc;
      }
    }
  }
}

// Reduce-scatter over the 8 lanes of a group: v[0 .. 8 Q) are per-lane partial sums of 8 Q quantities; afterwards lane gl holds
// the group totals of quantities base .. base + Q - 1 in v[0 .. Q), base = Q * (4 b2 + 2 b1 + b0) for gl = (b2 b1 b0).
// 7 Q shuffles (an all-reduce of every quantity would take 24 Q).
template <int Q>
__device__ __forceinline__ void group8_reduce_scatter(float (&v)[8 * Q], int gl) {
#pragma unroll
  for (int half = 4 * Q, off = 4; off >= 1; half >>= 1, off >>= 1) {
    const bool up = (gl & off) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// The same over a whole warp: v[0 .. 32 Q) per-lane partials; afterwards lane l holds the warp totals of quantities
// Q l .. Q l + Q - 1 in v[0 .. Q).  31 Q shuffles.
template <int Q>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[32 * Q], int lane) {
#pragma unroll
  for (int half = 16 * Q, off = 16; off >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

__device__ __forceinline__ float pick4(const float z[4], int k) { return k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3])); }

// buf[i] <- (i < step) ? fresh[i] : buf[i - step]  for i < tot: a newest-first history takes one entry.  GL-wide blocks,
// back to front, so no element is overwritten before it has moved.
template <int GL>
__device__ __forceinline__ void shift_in(float* buf, int tot, int step, const float* fresh, int gl) {
  for (int base = ((tot - 1) / GL) * GL; base >= 0; base -= GL) {
    const int i = base + gl;
    const float v = (i < tot) ? (i < step ? fresh[i] : buf[i - step]) : 0.f;
    __syncwarp();
    if (i < tot) buf[i] = v;
    __syncwarp();
  }
}

// GL lanes own an env (32 / GL envs per warp): the contractions of the colab-sized controllers are 10 to 60 elements
// wide, so a whole warp per env leaves most lanes idle and the kernel issue-bound on their index arithmetic.  All
// groups of a warp run the same trip counts (the dims are shared), so warp-wide __syncwarp / full-mask shuffles stay
// legal; a group past the last env recomputes env N-1 and writes nothing.
//
// SD..SH: compile-time (d, n, p, m, F, L, h) of the shapes the reference's agents default to on the colab's system
// (d_hidden = 10, d_action = 3, d_obs = 2); SD = 0 keeps every dim a runtime value.  With the dims known the
// contractions unroll into LDS + FFMA with immediate offsets; the runtime-dims loops spend three integer
// instructions per FMA on addresses and trip counts (ncu: 1660 warp-instructions per LQG step, 300 of them FFMA).
template <int AGENT, int GL, int SD = 0, int SN = 0, int SP = 0, int SM = 0, int SF = 0, int SL = 0, int SH = 0>
// (the window-blocked SFC instance -- 8 lanes, m = 30 -- holds Theta / dTheta in registers and is limited to 16 envs per CTA by
// its shared memory anyway: it may use the whole register file)
__global__ void __launch_bounds__(256, GL == 32 ? ((AGENT == DLC_OF_FGRC && SD != 0 && SF * SP > 24) ? 1 : 4)
                                                : (AGENT != DLC_OF_FGRC ? 3 : ((GL == 8 && SM >= 20) ? 1 : 2)))
of_kernel(const OfArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NR = 2;   // rows per lane of a mat-vec: the launcher keeps max(d, n, p) <= 2 GL (code size: the step loop must stay
                          // inside the instruction cache -- a 64 / GL-way unroll made it 167 KB and no_instruction the top stall)
  const int gl = threadIdx.x & (GL - 1), grp = threadIdx.x / GL;
  const DlcOfParams& P = a.p;
  const int d = SD ? SD : P.d, n = SD ? SN : P.n, p = SD ? SP : P.p, m = SD ? SM : P.m, T = P.T;
  const int F = SD ? SF : P.F, L = SD ? SL : P.L, h = SD ? SH : P.h;
  const int FP = SD ? (AGENT == DLC_OF_DRC ? SM * SP : SF * SP) : a.FP;
  const int J = SD ? (AGENT == DLC_OF_DRC ? SH : SM) : a.J;
  const bool closed = P.mode == DLC_MODE_CLOSED_LOOP;
  // Window-blocked form of the filter family's three contractions (compile-time dims, 8 lanes, n F p <= 72: the default GRC
  // and SFC).  The element-per-lane loops below read two shared-memory operands per FMA and are LSU-bound (ncu,
  // profiles/r01_of_grc_reg_full.txt: LSU pipe 61 %); here a lane owns the windows in ring slots gl, gl + 8, ... and
  // holds Theta (then the partial dTheta) in registers, so a window's features are read once per pass and each feeds
  // n FMAs; the per-lane partial sums meet in a 7/8-shuffle-per-value reduce-scatter.  Same products, other summation order.
  constexpr bool BLK = AGENT == DLC_OF_FGRC && SD != 0 && GL == 8 && SN * SF * SP <= 72 && SF * SP <= 24;
  constexpr int CFP = BLK ? SF * SP : 1, CN = BLK ? SN : 1, CP = BLK ? SP : 1, CM1 = BLK ? SM + 1 : 1;
  constexpr int WPL = (CM1 + 7) / 8;   // windows per lane
  // Feature-blocked form for wide filter banks on a whole warp (compile-time dims, F p > 24: the default DSC, 242 features
  // x 31 windows).  Lane l owns features 8 l .. 8 l + 7 of EVERY window and keeps its 3 x 8 block of Theta in registers for
  // the whole rollout: a(k) needs one reduce-scatter over the warp (lane k ends up with a(k)), dTheta needs no cross-lane
  // sum at all, and the filter bank is staged transposed ([tap][filter]) so a lane reads its 4 filters of a tap as one
  // 16-byte word.  Window rows are padded to 8 * 31 = 248 floats.
  constexpr bool FBL = AGENT == DLC_OF_FGRC && SD != 0 && GL == 32 && SF * SP > 24 && SF * SP <= 248 && SM + 1 <= 32 &&
                       SN == 3 && SP == 2;
  constexpr int FST = 248;                       // padded feature-row stride (FBL)
  constexpr int FPD = FBL ? ((SF + 3) & ~3) : 1; // padded filter count of the transposed bank
  // FBL feature ownership: lane l < 31 owns features 4 l .. 4 l + 3 and 124 + 4 l .. 124 + 4 l + 3, so that each of its
  // two 16-byte reads of a window row is part of one contiguous 496-byte sweep of the warp (conflict-free); in filter
  // terms (p = 2) that is filters 2 l, 2 l + 1 and 62 + 2 l, 62 + 2 l + 1.  Lane 31 idles through the passes (its block of
  // Theta is zero and it reads lane 0's addresses).
  const int fl = gl < 31 ? gl : 0;
  auto fbl_feature = [&](int j) { return (j < 4 ? 0 : 124) + 4 * gl + (j & 3); };
  if (AGENT == DLC_OF_FGRC) {
    if constexpr (FBL) {
      for (int i = threadIdx.x; i < a.nPhi; i += blockDim.x) {
        const int o = i / FPD, f = i - o * FPD;
        smem[i] = (o < SL && f < SF) ? a.Phi[f * SL + o] : 0.f;
      }
    } else {
      for (int i = threadIdx.x; i < a.nPhi; i += blockDim.x) smem[i] = a.Phi[i];
    }
    __syncthreads();  // the only CTA-wide barrier: warps are independent from here on
  }
  int64_t env = (int64_t)blockIdx.x * a.wpc + grp;
  if (GL == 32 && env >= P.N) return;   // a whole warp: nobody left to shuffle with
  const bool live = env < P.N;
  if (!live) env = P.N - 1;
  const float* sPhi = smem;
  float* s = smem + a.nPhi + (size_t)grp * a.S;
  float *sA = s + a.oA, *sB = s + a.oB, *sC = s + a.oC, *sG = s + a.oG, *sK = s + a.oK, *sL = s + a.oL;
  float *sTh = s + a.oTh, *sFeat = s + a.oFeat, *sYn = s + a.oYn, *sUs = s + a.oUs;
  float *sx = s + a.oX, *sz = s + a.oZ, *sy = s + a.oY, *su = s + a.oU;
  float *sAct = s + a.oAct, *sGact = s + a.oGact, *sFy = s + a.oFy, *sTmp = s + a.oTmp;

  // ---------------------------------------------------------------- stage the env
  const size_t eD = P.shared_dynamics ? 0 : (size_t)env;
  for (int i = gl; i < d * d; i += GL) sA[i] = a.A[eD * d * d + i];
  for (int i = gl; i < d * n; i += GL) sB[i] = a.B[eD * d * n + i];
  for (int i = gl; i < p * d; i += GL) sC[i] = a.C[eD * p * d + i];
  if (closed) for (int i = gl; i < d; i += GL) sx[i] = a.x_io[(size_t)env * d + i];
  if (AGENT == DLC_OF_LQG) {
    for (int i = gl; i < n * d; i += GL) sK[i] = a.K[eD * n * d + i];
    for (int i = gl; i < d * p; i += GL) sL[i] = a.Lk[eD * d * p + i];
    for (int i = gl; i < d; i += GL) sz[i] = a.z_io[(size_t)env * d + i];
  }
  if (AGENT == DLC_OF_DRC) {
    for (int i = gl; i < n * p; i += GL) sK[i] = a.K ? a.K[eD * n * p + i] : 0.f;  // _drc.py:104: K defaults to 0
    // M[j][a][q] -> sTh[a][j*p + q]
    for (int e = gl; e < m * n * p; e += GL) {
      const int j = e / (n * p), rem = e - j * n * p, aa = rem / p, q = rem - aa * p;
      sTh[aa * FP + j * p + q] = a.theta_io[(size_t)env * m * n * p + e];
    }
    for (int i = gl; i < m * p; i += GL) sYn[i] = a.ynat_io[(size_t)env * m * p + i];
    for (int i = gl; i < h * n; i += GL) sUs[i] = a.us_io[(size_t)env * h * n + i];
  }
  // FBL: this lane's 3 x 8 block of Theta, resident for the whole rollout as packed pairs (feature j = 2 i, 2 i + 1): the
  // passes run on fma.rn.f32x2 (two FMAs per issued instruction), the window rows arriving as ready-made pairs
  f2 thr2[3][4];
#pragma unroll
  for (int aa = 0; aa < 3; ++aa)
#pragma unroll
    for (int i = 0; i < 4; ++i) thr2[aa][i] = 0ull;
  if (AGENT == DLC_OF_FGRC) {
    if constexpr (FBL) {
#pragma unroll
      for (int aa = 0; aa < 3; ++aa)
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          float v[2] = {0.f, 0.f};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int i = fbl_feature(2 * i2 + h2), f = i >> 1, q = i & 1;      // (p = 2)
            if (gl < 31 && i < SF * SP) v[h2] = a.theta_io[(size_t)env * SF * 6 + (f * 3 + aa) * 2 + q];
          }
          thr2[aa][i2] = pk(v[0], v[1]);
        }
    } else {
      for (int e = gl; e < F * n * p; e += GL) {
        const int f = e / (n * p), rem = e - f * n * p, aa = rem / p, q = rem - aa * p;
        sTh[aa * FP + f * p + q] = a.theta_io[(size_t)env * F * n * p + e];
      }
    }
    for (int i = gl; i < (L + m) * p; i += GL) {
      const float v = a.ynat_io[(size_t)env * (L + m) * p + i];
      sYn[i] = v;
      if constexpr (FBL) sYn[(L + m) * p + i] = v;   // the mirror copy (see fbl_filter)
    }
    for (int i = gl; i < d; i += GL) sz[i] = a.z_io[(size_t)env * d + i];
  }
  __syncwarp();

  // ---------------------------------------------------------------- Markov operators G[j] = C A^j B, j < J
  if (AGENT != DLC_OF_LQG) {
    float* ab0 = sTmp;
    float* ab1 = sTmp + d * n;
    for (int i = gl; i < d * n; i += GL) ab0[i] = sB[i];
    __syncwarp();
    for (int j = 0; j < J; ++j) {
      for (int e = gl; e < p * n; e += GL) {
        const int q = e / n, aa = e - q * n;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(sC[q * d + c], ab0[c * n + aa], acc);
        sG[j * p * n + e] = acc;
      }
      for (int e = gl; e < d * n; e += GL) {
        const int r = e / n, aa = e - r * n;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(sA[r * d + c], ab0[c * n + aa], acc);
        ab1[e] = acc;
      }
      __syncwarp();
      float* tmp = ab0; ab0 = ab1; ab1 = tmp;
    }
  }
  // ---------------------------------------------------------------- filtered windows of the initial history
  const int M1 = m + 1, R = L + m;
  // FBL: one window's 8 outputs of this lane: sum_o PhiT[o][own filters] (x) y_nat[s0 + o][0..1].  The y_nat ring is kept
  // twice (slot and slot + R), so the L taps starting anywhere are contiguous and the loop unrolls without wrap tests.
  auto fbl_filter = [&](int s0, float* fw) {
    f2 acc[4] = {0ull, 0ull, 0ull, 0ull};        // (filter, q = 0 / 1) pairs: own filters 2 l, 2 l + 1, 62 + 2 l, 62 + 2 l + 1
    const float* yb = sYn + s0 * 2;
    const float* pb = sPhi + 2 * fl;
#pragma unroll 4
    for (int o = 0; o < SL; ++o) {
      const f2 y2 = *reinterpret_cast<const f2*>(yb + 2 * o);
      const float2 pa = *reinterpret_cast<const float2*>(pb + o * FPD);
      const float2 pc = *reinterpret_cast<const float2*>(pb + o * FPD + 62);
      acc[0] = fma2(dup(pa.x), y2, acc[0]);
      acc[1] = fma2(dup(pa.y), y2, acc[1]);
      acc[2] = fma2(dup(pc.x), y2, acc[2]);
      acc[3] = fma2(dup(pc.y), y2, acc[3]);
    }
    if (gl < 31) {
      *reinterpret_cast<ulonglong2*>(fw + 4 * gl) = make_ulonglong2(acc[0], acc[1]);
      *reinterpret_cast<ulonglong2*>(fw + 124 + 4 * gl) = make_ulonglong2(acc[2], acc[3]);
    }
  };
  if constexpr (FBL) {
    for (int k = 0; k < M1; ++k) fbl_filter(k, sFeat + k * FST);
    __syncwarp();
  } else
  if (AGENT == DLC_OF_FGRC) {
    for (int k = 0; k < M1; ++k)
      for (int i = gl; i < FP; i += GL) {
        const int f = i / p, q = i - f * p;
        float acc = 0.f;
        for (int o = 0; o < L; ++o) acc = fmaf(sPhi[f * L + o], sYn[(k + o) * p + q], acc);
        sFeat[k * FP + i] = acc;
      }
    __syncwarp();
  }
  int ybase = 0;  // ring slot of the oldest y_nat (filter family)
  int fbase = 0;  // ring slot of feature window k = 0
  const float inv_n = 1.f / (float)n, inv_FP = 1.f / (float)FP, inv_p = 1.f / (float)p;
  const uint32_t genv = (uint32_t)(P.env_offset + env);
  double csum = 0.0;
  int status = 0;

  for (int t = 0; t < T; ++t) {
    // disturbance of this step, fetched early
    float w[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      w[k] = 0.f;
      const int row = gl + k * GL;
      if (closed && k * GL < d && row < d) {
        if (a.W) {
          w[k] = a.W[((size_t)t * P.N + env) * d + row];
        } else {
          float z4[4];
          normal4(P.seed, genv, (uint32_t)(P.env_t0 + t), (uint32_t)(row >> 2), 0u, z4);
          w[k] = __fmul_rn(P.w_std, pick4(z4, row & 3));
        }
      }
    }
    // ---- observation y = C x  (_lds.py:110), or the caller's observation (Agent.__call__ on its own)
    if (closed) {
      float r[NR];
      matvec<GL>(sC, sx, p, d, gl, r);
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (gl + k * GL < p) sy[gl + k * GL] = r[k];
    } else {
      for (int i = gl; i < p; i += GL) sy[i] = a.y_in[((size_t)t * P.N + env) * p + i];
    }
    __syncwarp();

    if (AGENT == DLC_OF_LQG) {
      // u = -K x_hat; residual = y - C x_hat; x_hat <- A x_hat + B u + L residual   (_lqg.py:94-97)
      float r[NR];
      matvec<GL>(sK, sz, n, d, gl, r);
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (gl + k * GL < n) su[gl + k * GL] = -r[k];
      matvec<GL>(sC, sz, p, d, gl, r);
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (gl + k * GL < p) sFy[gl + k * GL] = sy[gl + k * GL] - r[k];
      __syncwarp();
      matvec<GL>(sA, sz, d, d, gl, r);
      matvec<GL, true>(sB, su, d, n, gl, r);
      matvec<GL, true>(sL, sFy, d, p, gl, r);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (gl + k * GL < d) sz[gl + k * GL] = r[k];
      __syncwarp();
    }

    if (AGENT == DLC_OF_DRC) {
      const float lr = P.decay ? P.lr / (float)(P.t0 + t + 1) : 1.0f;  // _drc.py:184: without decay the rate is 1.0 (sic)
      // gu[q] = sum_i sum_a G[i][q][a] us[i][a]     (_drc.py:169)
      for (int q = 0; q < p; ++q) {
        float part = 0.f;
        for (int i = gl; i < h; i += GL)
          for (int aa = 0; aa < n; ++aa) part = fmaf(sG[(i * p + q) * n + aa], sUs[i * n + aa], part);
        part = group_sum<GL>(part);
        if (gl == 0) sFy[q] = part;
      }
      __syncwarp();
      // y_nat: shift by one position (newest first), y_nat[0] = y - gu    (_drc.py:169-171)
      for (int i = gl; i < p; i += GL) sTmp[i] = sy[i] - sFy[i];
      __syncwarp();
      shift_in<GL>(sYn, m * p, p, sTmp, gl);
      // mdot[a] = sum_{j,q} M[j][a][q] y_nat[j][q];  u = -K y + mdot;  final = y_nat[0] + gu;  act = -K final + mdot
      for (int aa = 0; aa < n; ++aa) {
        float part = 0.f;
        for (int i = gl; i < FP; i += GL) part = fmaf(sTh[aa * FP + i], sYn[i], part);
        part = group_sum<GL>(part);
        if (gl == 0) {
          float ky = 0.f, kf = 0.f;
          for (int q = 0; q < p; ++q) {
            ky = fmaf(sK[aa * p + q], sy[q], ky);
            kf = fmaf(sK[aa * p + q], sYn[q] + sFy[q], kf);
          }
          su[aa] = part - ky;
          sAct[aa] = 2.f * (part - kf);   // 2 * action(final_state): the adjoint seed of quad_loss w.r.t. the action
        }
      }
      __syncwarp();
      // M <- M - lr * 2 act (x) y_nat, then the norm clip   (_drc.py:181-193)
      float nrm = 0.f;
      for (int e = gl; e < n * FP; e += GL) {
        const int aa = fast_div(e, inv_FP), i = e - aa * FP;
        const float th = fmaf(-lr * sAct[aa], sYn[i], sTh[e]);
        sTh[e] = th;
        nrm = fmaf(th, th, nrm);
      }
      nrm = sqrtf(group_sum<GL>(nrm));
      if (nrm > P.R_M) {
        const float sc = P.R_M / nrm;
        for (int e = gl; e < n * FP; e += GL) sTh[e] *= sc;
      }
      // us: shift, us[0] = u    (_drc.py:196-197)
      shift_in<GL>(sUs, h * n, n, su, gl);
    }

    if (AGENT == DLC_OF_FGRC) {
      const float lr = P.decay ? P.lr / (float)(P.t0 + t + 1) : P.lr;  // _grc.py:145
      // ---- get_action: newest window of the history as it stands (_grc.py:157-169)
      if constexpr (FBL) {
        int snew = fbase + m; if (snew >= M1) snew -= M1;
        const float* fw = sFeat + snew * FST + 4 * fl;
        const ulonglong2 p0 = *reinterpret_cast<const ulonglong2*>(fw), p1 = *reinterpret_cast<const ulonglong2*>(fw + 124);
#pragma unroll
        for (int aa = 0; aa < 3; ++aa) {
          const f2 acc = fma2(thr2[aa][3], p1.y, fma2(thr2[aa][2], p1.x, fma2(thr2[aa][1], p0.y, mul2(thr2[aa][0], p0.x))));
          float part = lo(acc) + hi(acc);                               // (lane 31's block of Theta is zero)
          part = group_sum<32>(part);
          if (gl == 0) su[aa] = part;
        }
      } else {
        int snew = fbase + m; if (snew >= M1) snew -= M1;
        const float* fw = sFeat + snew * FP;
        for (int aa = 0; aa < n; ++aa) {
          float part = 0.f;
          for (int i = gl; i < FP; i += GL) part = fmaf(sTh[aa * FP + i], fw[i], part);
          part = group_sum<GL>(part);
          if (gl == 0) su[aa] = part;
        }
      }
      __syncwarp();
      // ---- update: z <- A z + B u; y_nat = y - C z; push    (_grc.py:139-142)
      {
        float rc[NR];
        matvec<GL>(sA, sz, d, d, gl, rc);
        matvec<GL, true>(sB, su, d, n, gl, rc);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NR; ++k)
          if (gl + k * GL < d) sz[gl + k * GL] = rc[k];
        __syncwarp();
        matvec<GL>(sC, sz, p, d, gl, rc);
        float* slot = sYn + ybase * p;   // the oldest entry is replaced by the newest
#pragma unroll
        for (int k = 0; k < NR; ++k)
          if (gl + k * GL < p) {
            slot[gl + k * GL] = sy[gl + k * GL] - rc[k];
            if constexpr (FBL) slot[R * p + gl + k * GL] = sy[gl + k * GL] - rc[k];
          }
        ybase = (ybase + 1 == R) ? 0 : ybase + 1;
      }
      __syncwarp();
      // ---- filter the newest window (history positions m .. m+L-1) into the ring slot of the oldest one
      if constexpr (FBL) {
        int s0 = ybase + m; if (s0 >= R) s0 -= R;
        fbl_filter(s0, sFeat + fbase * FST);
        fbase = (fbase + 1 == M1) ? 0 : fbase + 1;
      } else if constexpr (BLK) {
        // taps dealt to lanes: a lane reads y_nat[o] once and feeds all F filters with it
        float* fw = sFeat + fbase * FP;
        int s0 = ybase + m; if (s0 >= R) s0 -= R;
        float part[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) part[i] = 0.f;
        for (int o = gl; o < L; o += 8) {
          int sl = s0 + o; if (sl >= R) sl -= R;
          float yq[CP];
#pragma unroll
          for (int q = 0; q < CP; ++q) yq[q] = sYn[sl * CP + q];
#pragma unroll
          for (int f = 0; f < CFP / CP; ++f) {
            const float ph = sPhi[f * L + o];
#pragma unroll
            for (int q = 0; q < CP; ++q) part[f * CP + q] = fmaf(ph, yq[q], part[f * CP + q]);
          }
        }
        group8_reduce_scatter<3>(part, gl);
        const int base = 3 * gl;   // (4 b2 + 2 b1 + b0 = gl)
#pragma unroll
        for (int r = 0; r < 3; ++r)
          if (base + r < CFP) fw[base + r] = part[r];
        fbase = (fbase + 1 == M1) ? 0 : fbase + 1;
      } else {
        float* fw = sFeat + fbase * FP;
        int s0 = ybase + m; if (s0 >= R) s0 -= R;
        for (int i = gl; i < FP; i += GL) {
          const int f = fast_div(i, inv_p), q = i - f * p;
          const float* ph = sPhi + f * L;
          int sl = s0;
          float acc = 0.f;
          for (int o = 0; o < L; ++o) {
            acc = fmaf(ph[o], sYn[sl * p + q], acc);
            sl = (sl + 1 == R) ? 0 : sl + 1;
          }
          fw[i] = acc;
        }
        fbase = (fbase + 1 == M1) ? 0 : fbase + 1;
      }
      __syncwarp();
      if constexpr (FBL) {
        // ---- pass 1: this lane's share of a(k) = Theta phi(k) for every window; reduce-scatter: lane k gets a(k)
        constexpr int CM1f = SM + 1;
        float ap[96];
#pragma unroll
        for (int e = 0; e < 96; ++e) ap[e] = 0.f;
        const int fo = 4 * fl;
        {
          int sl = fbase;
#pragma unroll
          for (int k = 0; k < CM1f; ++k) {
            const float* fw = sFeat + sl * FST + fo;
            const ulonglong2 p0 = *reinterpret_cast<const ulonglong2*>(fw), p1 = *reinterpret_cast<const ulonglong2*>(fw + 124);
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
              const f2 acc = fma2(thr2[aa][3], p1.y, fma2(thr2[aa][2], p1.x, fma2(thr2[aa][1], p0.y, mul2(thr2[aa][0], p0.x))));
              ap[3 * k + aa] = lo(acc) + hi(acc);
            }
            sl = (sl + 1 == CM1f) ? 0 : sl + 1;
          }
        }
        warp_reduce_scatter<3>(ap, gl);                    // lane k: a(k)[0..2] in ap[0..2]  (lane 31: zeros)
        // ---- final_y and the adjoint seeds; lane k owns window k, so its Markov operator index is fixed: m - 1 - gl
        const int kk = gl < SM ? SM - 1 - gl : 0;
        const float* g = sG + kk * 6;
        float fy0 = 0.f, fy1 = 0.f;
        if (gl < SM) {
          fy0 = fmaf(g[2], ap[2], fmaf(g[1], ap[1], g[0] * ap[0]));
          fy1 = fmaf(g[5], ap[2], fmaf(g[4], ap[1], g[3] * ap[0]));
        }
        int snew = ybase + R - 1; if (snew >= R) snew -= R;
        fy0 = 2.f * (sYn[snew * 2] + group_sum<32>(fy0));
        fy1 = 2.f * (sYn[snew * 2 + 1] + group_sum<32>(fy1));
        float gk[3];
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
          gk[aa] = gl == SM ? 2.f * ap[aa] : (gl < SM ? fmaf(g[3 + aa], fy1, g[aa] * fy0) : 0.f);
        __syncwarp();
        if (gl < CM1f) *reinterpret_cast<float4*>(sAct + 4 * gl) = make_float4(gk[0], gk[1], gk[2], 0.f);
        __syncwarp();
        // ---- pass 2: dTheta block of this lane = sum_k g(k) (x) phi(k)[own features]; no cross-lane sum
        f2 dth[3][4];
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
          for (int i = 0; i < 4; ++i) dth[aa][i] = 0ull;
        {
          int sl = fbase;
          // (rolled 4 at a time: with both passes fully unrolled the step loop is ~45 KB of SASS, past the 32 KB
          // instruction cache -- ncu: no_instruction 0.42 stall cycles per issue -- and nothing here needs a
          // compile-time window index)
#pragma unroll 4
          for (int k = 0; k < CM1f; ++k) {
            const float* fw = sFeat + sl * FST + fo;
            const ulonglong2 p0 = *reinterpret_cast<const ulonglong2*>(fw), p1 = *reinterpret_cast<const ulonglong2*>(fw + 124);
            const float4 g4 = *reinterpret_cast<const float4*>(sAct + 4 * k);
            const float gg[3] = {g4.x, g4.y, g4.z};
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
              const f2 g2 = dup(gg[aa]);
              dth[aa][0] = fma2(g2, p0.x, dth[aa][0]);
              dth[aa][1] = fma2(g2, p0.y, dth[aa][1]);
              dth[aa][2] = fma2(g2, p1.x, dth[aa][2]);
              dth[aa][3] = fma2(g2, p1.y, dth[aa][3]);
            }
            sl = (sl + 1 == CM1f) ? 0 : sl + 1;
          }
        }
        // ---- Theta <- proj_RM(Theta - lr dTheta)     (_grc.py:144-151).  (Padding features carry zeros: their filters are
        // zero, so their dTheta is zero and their Theta stays zero -- no masks.)
        f2 nrm2 = 0ull;
        const f2 nlr2 = dup(gl < 31 ? -lr : 0.f);   // (lane 31 reads lane 0's rows: its block of Theta must stay zero)
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            thr2[aa][i] = fma2(nlr2, dth[aa][i], thr2[aa][i]);
            nrm2 = fma2(thr2[aa][i], thr2[aa][i], nrm2);
          }
        float nrm = gl < 31 ? lo(nrm2) + hi(nrm2) : 0.f;
        nrm = sqrtf(group_sum<32>(nrm));
        const float sc = fminf(1.0f, P.R_M / (nrm + 1e-8f));
        if (sc < 1.0f) {
          const f2 sc2 = dup(sc);
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int i = 0; i < 4; ++i) thr2[aa][i] = mul2(thr2[aa][i], sc2);
        }
        __syncwarp();
      } else if constexpr (BLK) {
        // ---- pass 1: a(k) = Theta phi(k) for this lane's windows; their share of final_y    (_grc.py:97-104)
        float am[CN], fyp[CP];
#pragma unroll
        for (int aa = 0; aa < CN; ++aa) am[aa] = 0.f;
#pragma unroll
        for (int q = 0; q < CP; ++q) fyp[q] = 0.f;
        {
          static_assert(!BLK || CFP % 2 == 0, "feature rows are read as float2");
          float th[CN][CFP];   // (every per-env offset is a multiple of 4 floats and CFP is even: 8-byte aligned rows)
#pragma unroll
          for (int aa = 0; aa < CN; ++aa)
#pragma unroll
            for (int i = 0; i < CFP; i += 2) {
              const float2 t2 = *reinterpret_cast<const float2*>(sTh + aa * CFP + i);
              th[aa][i] = t2.x; th[aa][i + 1] = t2.y;
            }
          // (branch-free: an out-of-range slot is clamped and weighted 0, so the WPL windows' chains interleave)
#pragma unroll
          for (int j = 0; j < WPL; ++j) {
            const int sl0 = gl + 8 * j;
            const bool valid = sl0 < CM1;
            const int sl = valid ? sl0 : CM1 - 1;
            int k = sl - fbase; if (k < 0) k += CM1;
            const bool newest = k == m;
            const float* fw = sFeat + sl * CFP;
            float ak[CN];
#pragma unroll
            for (int aa = 0; aa < CN; ++aa) ak[aa] = 0.f;
#pragma unroll
            for (int i = 0; i < CFP; i += 2) {
              const float2 ph = *reinterpret_cast<const float2*>(fw + i);
#pragma unroll
              for (int aa = 0; aa < CN; ++aa) ak[aa] = fmaf(th[aa][i + 1], ph.y, fmaf(th[aa][i], ph.x, ak[aa]));
            }
            const float* g = sG + (newest ? 0 : m - 1 - k) * CP * CN;
            const float wgt = (valid && !newest) ? 1.f : 0.f;
#pragma unroll
            for (int aa = 0; aa < CN; ++aa) {
              am[aa] = (valid && newest) ? ak[aa] : am[aa];
              ak[aa] *= wgt;
            }
#pragma unroll
            for (int q = 0; q < CP; ++q)
#pragma unroll
              for (int aa = 0; aa < CN; ++aa) fyp[q] = fmaf(g[q * CN + aa], ak[aa], fyp[q]);
          }
        }
        int snew = ybase + R - 1; if (snew >= R) snew -= R;
        float fy[CP];
#pragma unroll
        for (int q = 0; q < CP; ++q) fy[q] = 2.f * (sYn[snew * CP + q] + group_sum<8>(fyp[q]));
        // ---- pass 2: dTheta = sum_k g(k) phi(k)^T, g(m) = 2 a(m), g(k) = G[m-1-k]^T fy     (_grc.py:144-151)
        float dth[72];
#pragma unroll
        for (int e = 0; e < 72; ++e) dth[e] = 0.f;
#pragma unroll
        for (int j = 0; j < WPL; ++j) {
          const int sl0 = gl + 8 * j;
          const bool valid = sl0 < CM1;
          const int sl = valid ? sl0 : CM1 - 1;
          int k = sl - fbase; if (k < 0) k += CM1;
          const bool newest = k == m;
          const float* fw = sFeat + sl * CFP;
          const float* g = sG + (newest ? 0 : m - 1 - k) * CP * CN;
          float gk[CN];
#pragma unroll
          for (int aa = 0; aa < CN; ++aa) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < CP; ++q) acc = fmaf(g[q * CN + aa], fy[q], acc);
            gk[aa] = valid ? (newest ? 2.f * am[aa] : acc) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < CFP; i += 2) {
            const float2 ph = *reinterpret_cast<const float2*>(fw + i);
#pragma unroll
            for (int aa = 0; aa < CN; ++aa) {
              dth[aa * CFP + i] = fmaf(gk[aa], ph.x, dth[aa * CFP + i]);
              dth[aa * CFP + i + 1] = fmaf(gk[aa], ph.y, dth[aa * CFP + i + 1]);
            }
          }
        }
        group8_reduce_scatter<9>(dth, gl);
        // ---- Theta <- proj_RM(Theta - lr dTheta): this lane applies entries 9 gl .. 9 gl + 8
        __syncwarp();                                  // every lane has read the old Theta
        float nrm = 0.f;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const int e = 9 * gl + r;
          if (e < CN * CFP) {
            const float thn = fmaf(-lr, dth[r], sTh[e]);
            sTh[e] = thn;
            nrm = fmaf(thn, thn, nrm);
          }
        }
        nrm = sqrtf(group_sum<8>(nrm));
        const float sc = fminf(1.0f, P.R_M / (nrm + 1e-8f));
        if (sc < 1.0f) {
#pragma unroll
          for (int r = 0; r < 9; ++r)
            if (9 * gl + r < CN * CFP) sTh[9 * gl + r] *= sc;
        }
        __syncwarp();
      } else {
      // ---- a(k), k = 0..m, with the parameters before the update
      for (int e = gl; e < M1 * n; e += GL) {
        const int k = fast_div(e, inv_n), aa = e - k * n;
        int sl = fbase + k; if (sl >= M1) sl -= M1;
        const float* fw = sFeat + sl * FP;
        const float* th = sTh + aa * FP;
        float acc = 0.f;
        for (int i = 0; i < FP; ++i) acc = fmaf(th[i], fw[i], acc);
        sAct[e] = acc;
      }
      __syncwarp();
      // ---- final_y = sum_{k<m} G[m-1-k] a(k) + y_nat(newest)     (_grc.py:100-104)
      {
        int snew = ybase + R - 1; if (snew >= R) snew -= R;
        for (int q = gl; q < p; q += GL) {
          float acc = sYn[snew * p + q];
          for (int k = 0; k < m; ++k) {
            const float* g = sG + ((m - 1 - k) * p + q) * n;
            for (int aa = 0; aa < n; ++aa) acc = fmaf(g[aa], sAct[k * n + aa], acc);
          }
          sFy[q] = 2.f * acc;
        }
      }
      __syncwarp();
      // ---- adjoint of the actions: g(m) = 2 a(m), g(k) = G[m-1-k]^T (2 final_y)
      for (int e = gl; e < M1 * n; e += GL) {
        const int k = fast_div(e, inv_n), aa = e - k * n;
        float acc;
        if (k == m) {
          acc = 2.f * sAct[e];
        } else {
          acc = 0.f;
          const float* g = sG + (m - 1 - k) * p * n + aa;
          for (int q = 0; q < p; ++q) acc = fmaf(g[q * n], sFy[q], acc);
        }
        sGact[e] = acc;
      }
      __syncwarp();
      // ---- Theta <- proj_RM(Theta - lr * sum_k g(k) phi(k)^T)     (_grc.py:144-151)
      float nrm = 0.f;
      for (int e = gl; e < n * FP; e += GL) {
        const int aa = fast_div(e, inv_FP), i = e - aa * FP;
        float acc = 0.f;
        int sl = fbase;
        for (int k = 0; k < M1; ++k) {
          acc = fmaf(sGact[k * n + aa], sFeat[sl * FP + i], acc);
          sl = (sl + 1 == M1) ? 0 : sl + 1;
        }
        const float th = fmaf(-lr, acc, sTh[e]);
        sTh[e] = th;
        nrm = fmaf(th, th, nrm);
      }
      nrm = sqrtf(group_sum<GL>(nrm));
      const float sc = fminf(1.0f, P.R_M / (nrm + 1e-8f));
      if (sc < 1.0f)
        for (int e = gl; e < n * FP; e += GL) sTh[e] *= sc;
      __syncwarp();
      }
    }

    // ---- realised cost quad_loss(y, u), outputs
    {
      float c = 0.f;
      for (int i = gl; i < p; i += GL) c = fmaf(sy[i], sy[i], c);
      for (int i = gl; i < n; i += GL) c = fmaf(su[i], su[i], c);
      c = group_sum<GL>(c);
      if (gl == 0) {
        if (a.cost_out && live) a.cost_out[(size_t)t * P.N + env] = c;
        csum += (double)c;
        if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
      }
      if (a.u_out && live) for (int i = gl; i < n; i += GL) a.u_out[((size_t)t * P.N + env) * n + i] = su[i];
    }
    // ---- env step x <- A x + B u + w    (_lds.py:107-109)
    if (closed) {
      float r[NR];
      matvec<GL>(sA, sx, d, d, gl, r);
      matvec<GL, true>(sB, su, d, n, gl, r);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (gl + k * GL < d) sx[gl + k * GL] = r[k] + w[k];
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- write the state back
  if (!live) return;
  if (closed) for (int i = gl; i < d; i += GL) a.x_io[(size_t)env * d + i] = sx[i];
  if (AGENT == DLC_OF_LQG) for (int i = gl; i < d; i += GL) a.z_io[(size_t)env * d + i] = sz[i];
  if (AGENT == DLC_OF_DRC) {
    for (int e = gl; e < m * n * p; e += GL) {
      const int j = e / (n * p), rem = e - j * n * p, aa = rem / p, q = rem - aa * p;
      a.theta_io[(size_t)env * m * n * p + e] = sTh[aa * FP + j * p + q];
    }
    for (int i = gl; i < m * p; i += GL) a.ynat_io[(size_t)env * m * p + i] = sYn[i];
    for (int i = gl; i < h * n; i += GL) a.us_io[(size_t)env * h * n + i] = sUs[i];
  }
  if (AGENT == DLC_OF_FGRC) {
    if constexpr (FBL) {
#pragma unroll
      for (int aa = 0; aa < 3; ++aa)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = fbl_feature(j), f = i >> 1, q = i & 1;
          const float v = (j & 1) ? hi(thr2[aa][j >> 1]) : lo(thr2[aa][j >> 1]);
          if (gl < 31 && i < SF * SP) a.theta_io[(size_t)env * SF * 6 + (f * 3 + aa) * 2 + q] = v;
        }
    } else {
      for (int e = gl; e < F * n * p; e += GL) {
        const int f = e / (n * p), rem = e - f * n * p, aa = rem / p, q = rem - aa * p;
        a.theta_io[(size_t)env * F * n * p + e] = sTh[aa * FP + f * p + q];
      }
    }
    for (int i = gl; i < R * p; i += GL) {   // oldest -> newest
      const int pos = i / p, q = i - pos * p;
      int sl = ybase + pos; if (sl >= R) sl -= R;
      a.ynat_io[(size_t)env * R * p + i] = sYn[sl * p + q];
    }
    for (int i = gl; i < d; i += GL) a.z_io[(size_t)env * d + i] = sz[i];
  }
  if (gl == 0) {
    if (a.cost_sum_out) a.cost_sum_out[env] = csum;
    if (a.status_out) a.status_out[env] = status;
  }
}

// LQG on a small compile-time system, one THREAD per env with A, B, C, K, L and both states in registers
// (200 + 20 floats at d = 10, n = 3, p = 2): a step is 355 FFMA with no shared memory and no warp synchronisation.
// Same arithmetic per element as of_kernel<LQG> (row sums accumulated in the same order).
// RS > 0 (closed loop, W from HBM, 16-byte tileable rows): the disturbance rows arrive through the warp's cp.async ring
// (row_ring.cuh) RS steps ahead; outputs go through running pointers either way (the kernel is issue-bound at 8 warps
// per SM: every per-step index computation is paid in FFMA slots).
template <int D, int NA, int P, int RS>
__global__ void __launch_bounds__(128) of_lqg_reg_kernel(const OfArgs a) {
  const DlcOfParams& Pm = a.p;
  const int64_t nt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t nw0 = nt - lane;                       // first env of this warp
  if (nw0 >= Pm.N) return;                             // (warp-uniform)
  const bool live = nt < Pm.N;                         // idle lanes of the last warp shadow its last env, store nothing
  const int64_t env = live ? nt : Pm.N - 1;
  const bool closed = RS > 0 || Pm.mode == DLC_MODE_CLOSED_LOOP;
  const size_t eD = Pm.shared_dynamics ? 0 : (size_t)env;
  float A[D][D], B[D][NA], C[P][D], K[NA][D], L[D][P], x[D], z[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + eD * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < NA; ++c) B[i][c] = __ldg(a.B + eD * D * NA + i * NA + c);
#pragma unroll
    for (int q = 0; q < P; ++q) L[i][q] = __ldg(a.Lk + eD * D * P + i * P + q);
    x[i] = closed ? a.x_io[(size_t)env * D + i] : 0.f;
    z[i] = a.z_io[(size_t)env * D + i];
  }
#pragma unroll
  for (int q = 0; q < P; ++q)
#pragma unroll
    for (int j = 0; j < D; ++j) C[q][j] = __ldg(a.C + eD * P * D + q * D + j);
#pragma unroll
  for (int c = 0; c < NA; ++c)
#pragma unroll
    for (int j = 0; j < D; ++j) K[c][j] = __ldg(a.K + eD * NA * D + c * D + j);
  const uint32_t genv = (uint32_t)(Pm.env_offset + env);
  double csum = 0.0;
  int status = 0;
  const bool store_cost = a.cost_out != nullptr && live, store_u = a.u_out != nullptr && live;
  float* cost_p = a.cost_out + env;                    // running pointer (not dereferenced when absent)
  // disturbance of step t from HBM (register path) or the generator
  auto fetch = [&](int t, float (&w)[D]) {
    if (a.W) {
      const float* wr = a.W + ((size_t)t * Pm.N + env) * D;
      if (D % 2 == 0) {   // rows of an even d start on 8-byte boundaries: half the load instructions
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(wr) + i);
          w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = __ldg(wr + i);
      }
    } else {
#pragma unroll
      for (int g = 0; g < (D + 3) / 4; ++g) {
        float z4[4];
        normal4(Pm.seed, genv, (uint32_t)(Pm.env_t0 + t), (uint32_t)g, 0u, z4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < D) w[4 * g + k] = __fmul_rn(Pm.w_std, z4[k]);
      }
    }
  };
  auto advance = [&](int t, const float (&w)[D]) {
    // y = C x (_lds.py:110), or the caller's observation
    float y[P], u[NA], res[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (closed) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(C[q][j], x[j], s);
        y[q] = s;
      } else {
        y[q] = a.y_in[((size_t)t * Pm.N + env) * P + q];
      }
    }
    // u = -K x_hat; residual = y - C x_hat; x_hat <- A x_hat + B u + L residual   (_lqg.py:94-97)
#pragma unroll
    for (int c = 0; c < NA; ++c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(K[c][j], z[j], s);
      u[c] = -s;
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(C[q][j], z[j], s);
      res[q] = y[q] - s;
    }
    float zn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(A[i][j], z[j], s);
#pragma unroll
      for (int c = 0; c < NA; ++c) s = fmaf(B[i][c], u[c], s);
#pragma unroll
      for (int q = 0; q < P; ++q) s = fmaf(L[i][q], res[q], s);
      zn[i] = s;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = zn[i];
    // realised cost quad_loss(y, u), outputs
    float c = 0.f;
#pragma unroll
    for (int q = 0; q < P; ++q) c = fmaf(y[q], y[q], c);
#pragma unroll
    for (int k = 0; k < NA; ++k) c = fmaf(u[k], u[k], c);
    if constexpr (RS > 0) {
      if (store_cost) *cost_p = c;
      cost_p += Pm.N;
    } else {   // (the register path has no room for the pointer: it would spill)
      if (store_cost) a.cost_out[(size_t)t * Pm.N + env] = c;
    }
    csum += (double)c;
    if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
    if (store_u)
#pragma unroll
      for (int k = 0; k < NA; ++k) a.u_out[((size_t)t * Pm.N + env) * NA + k] = u[k];
    // env step x <- A x + B u + w    (_lds.py:107-109)
    if (closed) {
      float xn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(A[i][j], x[j], s);
#pragma unroll
        for (int k = 0; k < NA; ++k) s = fmaf(B[i][k], u[k], s);
        xn[i] = s + w[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = xn[i];
    }
  };
  if constexpr (RS > 0) {
    static_assert(RS % 4 == 0, "the step loop is unrolled by 4");
    using Ring = WarpRowRing<D, RS>;
    extern __shared__ __align__(16) unsigned char ring_smem[];
    Ring ring;
    const int64_t left = Pm.N - nw0;
    ring.init(ring_smem + (size_t)(threadIdx.x >> 5) * Ring::kBytesPerWarp, a.W + nw0 * D, Pm.N * D,
              (int)(left < 32 ? left : 32), lane, live ? lane : 0);
    ring.prime(Pm.T);
    for (int tb = 0; tb < Pm.T; tb += 4) {
      const int sb = tb & (RS - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tb + k < Pm.T) {
          const uint32_t row = ring.acquire(sb + k);
          float w[D];
          if (D % 2 == 0) {
#pragma unroll
            for (int i = 0; i < D / 2; ++i) {
              const float2 v = lds_f2(row + 8u * i);
              w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < D; ++i) w[i] = lds_f1(row + 4u * i);
          }
          ring.release(sb + k, tb + k + RS < Pm.T);
          advance(tb + k, w);
        }
      }
    }
  } else {
    for (int t = 0; t < Pm.T; ++t) {
      float w[D];
      if (closed) fetch(t, w);   // fetched early
      advance(t, w);
    }
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (closed) a.x_io[(size_t)env * D + i] = x[i];
    a.z_io[(size_t)env * D + i] = z[i];
  }
  if (a.cost_sum_out) a.cost_sum_out[env] = csum;
  if (a.status_out) a.status_out[env] = status;
}

// GRC (identity filter bank: F = L = m) on a small compile-time system, one THREAD per env.  Theta, A and both states
// live in registers; B, C, the Markov operators, the y_nat ring and the m+1 counterfactual actions live in shared
// memory as [element][thread] (bank = thread).  With Phi = I the filtered window k IS the y_nat window starting at
// slot k, so there is no feature ring; the y_nat ring is a mirrored linear buffer (every entry stored at slots s and
// s + R), so a window is L*p consecutive floats at a register base + immediate offsets.  A step is ~1,900 FFMA and
// ~600 LDS with no warp synchronisation.  The gradient is accumulated straight into Theta, candidate by candidate
// (Theta is only read by the actions, which are all formed before), instead of into a separate sum.
template <int D, int NA, int P, int MH>
struct GrcRegLayout {
  static constexpr int R = 2 * MH, FP = MH * P, M1 = MH + 1;
  static constexpr int oRing = 0, oG = oRing + 2 * R * P, oB = oG + MH * P * NA,
                       oC = oB + D * NA, kFloats = oC + P * D;
};

template <int D, int NA, int P, int MH, int TPB>
__global__ void __launch_bounds__(TPB) of_grc_reg_kernel(const OfArgs a) {
  using Lay = GrcRegLayout<D, NA, P, MH>;
  constexpr int R = Lay::R, FP = Lay::FP, M1 = Lay::M1;
  extern __shared__ __align__(16) float smem[];
  const DlcOfParams& Pm = a.p;
  const int64_t env_raw = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (env_raw >= Pm.N) return;   // no warp-level communication: idle threads may leave
  const int64_t env = env_raw;
  float* const sm = smem + threadIdx.x;
  auto S = [&](int off) -> float& { return sm[off * TPB]; };
  // the y_nat ring as 8-byte slots [slot][thread] (p = 2: a slot IS one (q = 0, q = 1) pair, a window is MH consecutive
  // slots), so every window element pair is one LDS.64 feeding one fma.rn.f32x2 per action row
  static_assert(P == 2 && Lay::oRing == 0, "the packed ring assumes two observations per slot at the start of the slice");
  f2* const ring2 = reinterpret_cast<f2*>(smem) + threadIdx.x;
  const bool closed = Pm.mode == DLC_MODE_CLOSED_LOOP;
  const size_t eD = Pm.shared_dynamics ? 0 : (size_t)env;
  float A[D][D], x[D], z[D];
  f2 Th[NA][MH];   // Theta[f][aa][.] as the pair (q = 0, q = 1)
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + eD * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < NA; ++c) S(Lay::oB + i * NA + c) = __ldg(a.B + eD * D * NA + i * NA + c);
    x[i] = closed ? a.x_io[(size_t)env * D + i] : 0.f;
    z[i] = a.z_io[(size_t)env * D + i];
  }
#pragma unroll
  for (int q = 0; q < P; ++q)
#pragma unroll
    for (int j = 0; j < D; ++j) S(Lay::oC + q * D + j) = __ldg(a.C + eD * P * D + q * D + j);
  // Theta[f][aa][q] -> Th[aa][f*P + q]
#pragma unroll
  for (int f = 0; f < MH; ++f)
#pragma unroll
    for (int aa = 0; aa < NA; ++aa)
      Th[aa][f] = pk(a.theta_io[(size_t)env * MH * NA * P + (f * NA + aa) * P],
                     a.theta_io[(size_t)env * MH * NA * P + (f * NA + aa) * P + 1]);
  // y_nat ring, oldest -> newest, mirrored
  for (int i = 0; i < R; ++i) {
    const f2 v = pk(a.ynat_io[(size_t)env * R * P + 2 * i], a.ynat_io[(size_t)env * R * P + 2 * i + 1]);
    ring2[i * TPB] = v;
    ring2[(i + R) * TPB] = v;
  }
  // Markov operators G[j][q][aa] = (C A^j B)[q][aa], j < m            (_grc.py:97-105 collapsed, see the file header)
  {
    float ab[D][NA];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int c = 0; c < NA; ++c) ab[i][c] = S(Lay::oB + i * NA + c);
    for (int j = 0; j < MH; ++j) {
#pragma unroll
      for (int q = 0; q < P; ++q)
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < D; ++k) acc = fmaf(S(Lay::oC + q * D + k), ab[k][c], acc);
          S(Lay::oG + (j * P + q) * NA + c) = acc;
        }
      float nb[D][NA];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < D; ++k) acc = fmaf(A[i][k], ab[k][c], acc);
          nb[i][c] = acc;
        }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int c = 0; c < NA; ++c) ab[i][c] = nb[i][c];
    }
  }
  int ybase = 0;   // physical slot of the oldest entry
  const uint32_t genv = (uint32_t)(Pm.env_offset + env);
  double csum = 0.0;
  int status = 0;
  {   // the caller promised Phi == I (DLC_OF_FLAG_PHI_IDENTITY); a broken promise must not pass silently
    bool ident = true;
    for (int i = 0; i < MH * MH; ++i) ident = ident && (__ldg(a.Phi + i) == ((i / MH == i % MH) ? 1.0f : 0.0f));
    if (!ident) {
      if (a.cost_sum_out) a.cost_sum_out[env] = (double)nanf("");
      if (a.status_out) a.status_out[env] = DLC_STATUS_BAD_HINT;
      return;
    }
  }

  for (int t = 0; t < Pm.T; ++t) {
    const float lr = Pm.decay ? Pm.lr / (float)(Pm.t0 + t + 1) : Pm.lr;  // _grc.py:145
    float y[P], u[NA];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (closed) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(S(Lay::oC + q * D + j), x[j], s);
        y[q] = s;
      } else {
        y[q] = a.y_in[((size_t)t * Pm.N + env) * P + q];
      }
    }
    // ---- get_action: the newest window of the history as it stands (_grc.py:157-169)
    {
      const f2* win = ring2 + (ybase + MH) * TPB;
      f2 acc[NA];
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) acc[aa] = 0ull;
#pragma unroll
      for (int i = 0; i < MH; ++i) {
        const f2 v = win[i * TPB];
#pragma unroll
        for (int aa = 0; aa < NA; ++aa) acc[aa] = fma2(Th[aa][i], v, acc[aa]);
      }
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) u[aa] = lo(acc[aa]) + hi(acc[aa]);
    }
    // ---- update: z <- A z + B u; y_nat = y - C z; push (the oldest slot and its mirror take it)   (_grc.py:139-142)
    {
      float zn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(A[i][j], z[j], s);
#pragma unroll
        for (int c = 0; c < NA; ++c) s = fmaf(S(Lay::oB + i * NA + c), u[c], s);
        zn[i] = s;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = zn[i];
      float yn[P];
#pragma unroll
      for (int q = 0; q < P; ++q) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(S(Lay::oC + q * D + j), z[j], s);
        yn[q] = y[q] - s;
      }
      ring2[ybase * TPB] = pk(yn[0], yn[1]);
      ring2[(ybase + R) * TPB] = pk(yn[0], yn[1]);
      ybase = (ybase + 1 == R) ? 0 : ybase + 1;
    }
    const f2* ring = ring2 + ybase * TPB;   // logical slot s at ring[s * TPB]
    // ---- a(k), k = 0..m, with the parameters before the update
    // (windows k .. k + NW - 1 overlap in all but NW - 1 slots: a group per trip reads MH + NW - 1 slots once and runs
    // NW x NA independent chains -- the kernel is latency-bound at six warps per SM, so the chains are what it needs)
    float fy[P], aM[NA];
    {
      const f2 newest = ring[(R - 1) * TPB];
      fy[0] = lo(newest); fy[1] = hi(newest);
    }
    auto act_windows = [&](int k0, auto nwc) {
      constexpr int NW = decltype(nwc)::value;
      const f2* win = ring + k0 * TPB;
      f2 acc[NW][NA];
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int aa = 0; aa < NA; ++aa) acc[w][aa] = 0ull;
#pragma unroll
      for (int i = 0; i < MH + NW - 1; ++i) {
        const f2 v = win[i * TPB];
#pragma unroll
        for (int w = 0; w < NW; ++w)
          if (i - w >= 0 && i - w < MH) {
#pragma unroll
            for (int aa = 0; aa < NA; ++aa) acc[w][aa] = fma2(Th[aa][i - w], v, acc[w][aa]);
          }
      }
      // final_y = sum_{k<m} G[m-1-k] a(k) + y_nat(newest) (_grc.py:100-104) rides along: a(k) is still in registers,
      // and only a(m) is needed again (the seed of the newest window)
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        float ak[NA];
#pragma unroll
        for (int aa = 0; aa < NA; ++aa) ak[aa] = lo(acc[w][aa]) + hi(acc[w][aa]);
        if (k0 + w == MH) {
#pragma unroll
          for (int aa = 0; aa < NA; ++aa) aM[aa] = ak[aa];
        } else {
#pragma unroll
          for (int q = 0; q < P; ++q)
#pragma unroll
            for (int aa = 0; aa < NA; ++aa)
              fy[q] = fmaf(S(Lay::oG + ((MH - 1 - (k0 + w)) * P + q) * NA + aa), ak[aa], fy[q]);
        }
      }
    };
    constexpr int NWG = 4, NFULL = M1 / NWG, NREST = M1 % NWG;
#pragma unroll 1
    for (int k = 0; k < NFULL * NWG; k += NWG) act_windows(k, std::integral_constant<int, NWG>{});
    if constexpr (NREST > 0) act_windows(NFULL * NWG, std::integral_constant<int, NREST>{});
    // ---- this step's disturbance row: asked for here (not at the top of the step) so that it is not live during the
    // window pass, the phase with the most registers in use; the update below covers its latency
    float w[D];
    if (closed) {
      if (a.W) {
        const float* wr = a.W + ((size_t)t * Pm.N + env) * D;
        if (D % 2 == 0) {   // rows of an even d start on 8-byte boundaries: half the load instructions
#pragma unroll
          for (int i = 0; i < D / 2; ++i) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(wr) + i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < D; ++i) w[i] = __ldg(wr + i);
        }
      } else {
#pragma unroll
        for (int g = 0; g < (D + 3) / 4; ++g) {
          float z4[4];
          normal4(Pm.seed, genv, (uint32_t)(Pm.env_t0 + t), (uint32_t)g, 0u, z4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (4 * g + k < D) w[4 * g + k] = __fmul_rn(Pm.w_std, z4[k]);
        }
      }
    }
    // ---- fy = 2 final_y
#pragma unroll
    for (int q = 0; q < P; ++q) fy[q] *= 2.f;
    // ---- Theta <- Theta - lr * sum_k g(k) phi(k)^T, g(m) = 2 a(m), g(k) = G[m-1-k]^T fy      (_grc.py:144-151)
    auto seed_of = [&](int k, float (&g)[NA]) {   // -lr g(k)
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) {
        float s;
        if (k == MH) {
          s = 2.f * aM[aa];
        } else {
          s = 0.f;
#pragma unroll
          for (int q = 0; q < P; ++q) s = fmaf(S(Lay::oG + ((MH - 1 - k) * P + q) * NA + aa), fy[q], s);
        }
        g[aa] = -lr * s;
      }
    };
    auto grad_windows = [&](int k0, auto nwc) {
      constexpr int NW = decltype(nwc)::value;
      float g[NW][NA];
#pragma unroll
      for (int w = 0; w < NW; ++w) seed_of(k0 + w, g[w]);
      const f2* win = ring + k0 * TPB;
#pragma unroll
      for (int i = 0; i < MH + NW - 1; ++i) {
        const f2 v = win[i * TPB];
#pragma unroll
        for (int w = 0; w < NW; ++w)
          if (i - w >= 0 && i - w < MH) {
#pragma unroll
            for (int aa = 0; aa < NA; ++aa) Th[aa][i - w] = fma2(dup(g[w][aa]), v, Th[aa][i - w]);
          }
      }
    };
#pragma unroll 1
    for (int k = 0; k < NFULL * NWG; k += NWG) grad_windows(k, std::integral_constant<int, NWG>{});
    if constexpr (NREST > 0) grad_windows(NFULL * NWG, std::integral_constant<int, NREST>{});
    {
      f2 nrm2[NA];   // (a chain per action row)
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) {
        nrm2[aa] = 0ull;
#pragma unroll
        for (int i = 0; i < MH; ++i) nrm2[aa] = fma2(Th[aa][i], Th[aa][i], nrm2[aa]);
      }
      float nrm = 0.f;
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) nrm += lo(nrm2[aa]) + hi(nrm2[aa]);
      const float sc = fminf(1.0f, Pm.R_M / (sqrtf(nrm) + 1e-8f));
      if (sc < 1.0f) {
#pragma unroll
        for (int aa = 0; aa < NA; ++aa)
#pragma unroll
          for (int i = 0; i < MH; ++i) Th[aa][i] = mul2(Th[aa][i], dup(sc));
      }
    }
    // ---- realised cost quad_loss(y, u), outputs
    float c = 0.f;
#pragma unroll
    for (int q = 0; q < P; ++q) c = fmaf(y[q], y[q], c);
#pragma unroll
    for (int k = 0; k < NA; ++k) c = fmaf(u[k], u[k], c);
    if (a.cost_out) a.cost_out[(size_t)t * Pm.N + env] = c;
    csum += (double)c;
    if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
    if (a.u_out)
#pragma unroll
      for (int k = 0; k < NA; ++k) a.u_out[((size_t)t * Pm.N + env) * NA + k] = u[k];
    // ---- env step x <- A x + B u + w    (_lds.py:107-109)
    if (closed) {
      float xn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(A[i][j], x[j], s);
#pragma unroll
        for (int k = 0; k < NA; ++k) s = fmaf(S(Lay::oB + i * NA + k), u[k], s);
        xn[i] = s + w[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) x[i] = xn[i];
    }
  }
  // ---- write the state back
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (closed) a.x_io[(size_t)env * D + i] = x[i];
    a.z_io[(size_t)env * D + i] = z[i];
  }
#pragma unroll
  for (int f = 0; f < MH; ++f)
#pragma unroll
    for (int aa = 0; aa < NA; ++aa) {
      a.theta_io[(size_t)env * MH * NA * P + (f * NA + aa) * P] = lo(Th[aa][f]);
      a.theta_io[(size_t)env * MH * NA * P + (f * NA + aa) * P + 1] = hi(Th[aa][f]);
    }
  for (int i = 0; i < R; ++i) {
    const f2 v = ring2[(ybase + i) * TPB];
    a.ynat_io[(size_t)env * R * P + 2 * i] = lo(v);
    a.ynat_io[(size_t)env * R * P + 2 * i + 1] = hi(v);
  }
  if (a.cost_sum_out) a.cost_sum_out[env] = csum;
  if (a.status_out) a.status_out[env] = status;
}

// DRC on a small compile-time system, one THREAD per env (the caller's DLC_OF_FLAG_STABLE_A promise).  The as-written
// agent contracts the truncated Markov operator with the action history twice a step, gu = sum_{i<h} (C A^i B) us[i]
// (_drc.py:97-102,169,113): h p n = 300 multiply-adds on 450 operands that fit no register file.  Because
// G[i] = C A^i B, gu = C s with s = sum_{i<h} A^i B us[i], and pushing u into the history moves s by
//   s <- A s + B u - (A^h B) us[h-1]                  (the action that drops out of the window)
// -- one d x d mat-vec, exact algebra, and a contraction of rounding errors exactly when the truncated response decays
// (|A^h B| <= |B|, which is what the flag asserts and what the kernel verifies per env on entry: a broken promise
// returns DLC_STATUS_BAD_HINT and a NaN cost sum, never wrong numbers).  s is re-formed from the stored history at
// every call entry (Horner), so chunked calls do not carry the recursion's rounding along.  A, M (as Th[a][j p + q]),
// the y_nat window, x and s live in registers; B, C, K, A^h B and the action ring in shared memory as [element][thread].
template <int D, int NA, int P, int MH, int HH>
struct DrcRegLayout {
  static constexpr int oUs = 0, oB = oUs + HH * NA, oC = oB + D * NA, oAhB = oC + P * D, oK = oAhB + D * NA,
                       kFloats = oK + NA * P;
};

template <int D, int NA, int P, int MH, int HH, int TPB, int RS>
__global__ void __launch_bounds__(TPB) of_drc_reg_kernel(const OfArgs a) {
  using Lay = DrcRegLayout<D, NA, P, MH, HH>;
  constexpr int FP = MH * P;
  extern __shared__ __align__(16) float smem[];
  const DlcOfParams& Pm = a.p;
  const int64_t nt = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t nw0 = nt - lane;                       // first env of this warp
  if (nw0 >= Pm.N) return;                             // (warp-uniform)
  const bool live = nt < Pm.N;                         // idle lanes of the last warp shadow its last env, store nothing
  const int64_t env = live ? nt : Pm.N - 1;
  float* const sm = smem + threadIdx.x;
  auto S = [&](int off) -> float& { return sm[off * TPB]; };
  const bool closed = RS > 0 || Pm.mode == DLC_MODE_CLOSED_LOOP;
  const size_t eD = Pm.shared_dynamics ? 0 : (size_t)env;
  static_assert(P == 2, "M and the y_nat window are held as (q = 0, q = 1) pairs");
  float A[D][D], x[D], s[D];
  f2 Th[NA][MH], yn[MH];   // M[j][aa][.] and y_nat[j][.] as packed pairs: every FMA on them is a fma.rn.f32x2
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + eD * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < NA; ++c) S(Lay::oB + i * NA + c) = __ldg(a.B + eD * D * NA + i * NA + c);
    x[i] = closed ? a.x_io[(size_t)env * D + i] : 0.f;
    s[i] = 0.f;
  }
#pragma unroll
  for (int q = 0; q < P; ++q)
#pragma unroll
    for (int j = 0; j < D; ++j) S(Lay::oC + q * D + j) = __ldg(a.C + eD * P * D + q * D + j);
#pragma unroll
  for (int i = 0; i < NA * P; ++i) S(Lay::oK + i) = a.K ? __ldg(a.K + eD * NA * P + i) : 0.f;   // _drc.py:104
  // the action ring: logical us[i] (newest first) sits in slot (head + i) mod h
  for (int i = 0; i < HH * NA; ++i) S(Lay::oUs + i) = a.us_io[(size_t)env * HH * NA + i];
  int head = 0;
  int status = 0;
  bool bad = false;
  // A^h B column by column (h applications of A each), and s = sum_i A^i B us[i] by Horner from the oldest action
  // (before M and the y_nat window are loaded: the setup then fits the register file)
  {
    float nb2 = 0.f, na2 = 0.f;
#pragma unroll
    for (int c = 0; c < NA; ++c) {
      float v[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        v[i] = S(Lay::oB + i * NA + c);
        nb2 = fmaf(v[i], v[i], nb2);
      }
#pragma unroll 1
      for (int k = 0; k < HH; ++k) {
        float nv[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) acc = fmaf(A[i][j], v[j], acc);
          nv[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) v[i] = nv[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        S(Lay::oAhB + i * NA + c) = v[i];
        na2 = fmaf(v[i], v[i], na2);
      }
    }
    // the truncated response does not decay (or is not finite): the recursion would not contract.  The thread stays with
    // its warp (the row ring synchronises it) and reports nothing but the flag.
    bad = !(na2 <= nb2);
#pragma unroll 1
    for (int k = HH - 1; k >= 0; --k) {
      float ns[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fmaf(A[i][j], s[j], acc);
#pragma unroll
        for (int c = 0; c < NA; ++c) acc = fmaf(S(Lay::oB + i * NA + c), S(Lay::oUs + k * NA + c), acc);
        ns[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) s[i] = ns[i];
    }
  }
  // M[j][aa][q] -> Th[aa][j*P + q]; y_nat newest first
#pragma unroll
  for (int j = 0; j < MH; ++j)
#pragma unroll
    for (int aa = 0; aa < NA; ++aa)
      Th[aa][j] = pk(a.theta_io[(size_t)env * MH * NA * P + (j * NA + aa) * P],
                     a.theta_io[(size_t)env * MH * NA * P + (j * NA + aa) * P + 1]);
#pragma unroll
  for (int i = 0; i < MH; ++i) yn[i] = pk(a.ynat_io[(size_t)env * FP + 2 * i], a.ynat_io[(size_t)env * FP + 2 * i + 1]);
  const uint32_t genv = (uint32_t)(Pm.env_offset + env);
  double csum = 0.0;
  // RS > 0 (closed loop, W from HBM, 16-byte tileable rows): the disturbance rows arrive through the warp's cp.async
  // ring (row_ring.cuh) RS steps ahead and are read where the env step needs them -- no registers held across the step
  using Ring = WarpRowRing<D, RS ? RS : 4>;
  Ring ring;
  if constexpr (RS > 0) {
    const int64_t left = Pm.N - nw0;
    ring.init(reinterpret_cast<unsigned char*>(smem + Lay::kFloats * TPB) + (size_t)(threadIdx.x >> 5) * Ring::kBytesPerWarp,
              a.W + nw0 * D, Pm.N * D, (int)(left < 32 ? left : 32), lane, live ? lane : 0);
    ring.prime(Pm.T);
  }
  const bool store_cost = a.cost_out != nullptr && live, store_u = a.u_out != nullptr && live;

#pragma unroll 1
  for (int t = 0; t < Pm.T; ++t) {
    float w[D] = {};
    if (RS == 0 && closed) {
      if (a.W) {
        const float* wr = a.W + ((size_t)t * Pm.N + env) * D;
        if (D % 2 == 0) {
#pragma unroll
          for (int i = 0; i < D / 2; ++i) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(wr) + i);
            w[2 * i] = v.x; w[2 * i + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < D; ++i) w[i] = __ldg(wr + i);
        }
      } else {
#pragma unroll
        for (int g = 0; g < (D + 3) / 4; ++g) {
          float z4[4];
          normal4(Pm.seed, genv, (uint32_t)(Pm.env_t0 + t), (uint32_t)g, 0u, z4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (4 * g + k < D) w[4 * g + k] = __fmul_rn(Pm.w_std, z4[k]);
        }
      }
    }
    const float lr = Pm.decay ? Pm.lr / (float)(Pm.t0 + t + 1) : 1.0f;  // _drc.py:184: without decay the rate is 1.0 (sic)
    float y[P], gu[P], u[NA], act[NA];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (closed) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fmaf(S(Lay::oC + q * D + j), x[j], acc);
        y[q] = acc;
      } else {
        y[q] = a.y_in[((size_t)t * Pm.N + env) * P + q];
      }
      // gu = sum_i G[i] us[i] = C s     (_drc.py:169)
      float g = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) g = fmaf(S(Lay::oC + q * D + j), s[j], g);
      gu[q] = g;
    }
    // y_nat: shift by one position (newest first), y_nat[0] = y - gu    (_drc.py:169-171)
#pragma unroll
    for (int i = MH - 1; i >= 1; --i) yn[i] = yn[i - 1];
    const float yn0[P] = {y[0] - gu[0], y[1] - gu[1]};
    yn[0] = pk(yn0[0], yn0[1]);
    // mdot = sum_{j,q} M[j][.][q] y_nat[j][q];  u = -K y + mdot;  final = y_nat[0] + gu;  act = 2 (-K final + mdot)
#pragma unroll
    for (int aa = 0; aa < NA; ++aa) {
      f2 part2 = 0ull;
#pragma unroll
      for (int i = 0; i < MH; ++i) part2 = fma2(Th[aa][i], yn[i], part2);
      const float part = lo(part2) + hi(part2);
      float ky = 0.f, kf = 0.f;
#pragma unroll
      for (int q = 0; q < P; ++q) {
        const float kq = S(Lay::oK + aa * P + q);
        ky = fmaf(kq, y[q], ky);
        kf = fmaf(kq, yn0[q] + gu[q], kf);
      }
      u[aa] = part - ky;
      act[aa] = 2.f * (part - kf);
    }
    // M <- M - lr * 2 act (x) y_nat, then the norm clip   (_drc.py:181-193)
    {
      f2 nrm2[NA];   // (a chain per action row)
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) {
        const f2 g = dup(-lr * act[aa]);
        nrm2[aa] = 0ull;
#pragma unroll
        for (int i = 0; i < MH; ++i) {
          Th[aa][i] = fma2(g, yn[i], Th[aa][i]);
          nrm2[aa] = fma2(Th[aa][i], Th[aa][i], nrm2[aa]);
        }
      }
      float nrm = 0.f;
#pragma unroll
      for (int aa = 0; aa < NA; ++aa) nrm += lo(nrm2[aa]) + hi(nrm2[aa]);
      nrm = sqrtf(nrm);
      if (nrm > Pm.R_M) {
        const f2 sc = dup(Pm.R_M / nrm);
#pragma unroll
        for (int aa = 0; aa < NA; ++aa)
#pragma unroll
          for (int i = 0; i < MH; ++i) Th[aa][i] = mul2(Th[aa][i], sc);
      }
    }
    // us: shift, us[0] = u (_drc.py:196-197); the oldest action leaves the window and s follows
    {
      head = head == 0 ? HH - 1 : head - 1;
      float uo[NA];
#pragma unroll
      for (int c = 0; c < NA; ++c) {
        uo[c] = S(Lay::oUs + head * NA + c);
        S(Lay::oUs + head * NA + c) = u[c];
      }
      float ns[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fmaf(A[i][j], s[j], acc);
#pragma unroll
        for (int c = 0; c < NA; ++c) acc = fmaf(-S(Lay::oAhB + i * NA + c), uo[c], acc);
        ns[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) s[i] = ns[i];   // (+ B u below: the env step forms the same product)
    }
    // ---- realised cost quad_loss(y, u), outputs
    float c = 0.f;
#pragma unroll
    for (int q = 0; q < P; ++q) c = fmaf(y[q], y[q], c);
#pragma unroll
    for (int k = 0; k < NA; ++k) c = fmaf(u[k], u[k], c);
    if (store_cost) a.cost_out[(size_t)t * Pm.N + env] = c;
    csum += (double)c;
    if (!isfinite(c)) status |= DLC_STATUS_NONFINITE;
    if (store_u)
#pragma unroll
      for (int k = 0; k < NA; ++k) a.u_out[((size_t)t * Pm.N + env) * NA + k] = u[k];
    // ---- env step x <- A x + B u + w    (_lds.py:107-109)
    {
      if constexpr (RS > 0) {
        const int st = t & (RS - 1);
        const uint32_t row = ring.acquire(st);
        static_assert(RS == 0 || D % 2 == 0, "rows are read as 8-byte pairs");
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
          const float2 v = lds_f2(row + 8u * i);
          w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
        ring.release(st, t + RS < Pm.T);
      }
      float xn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float bu = 0.f;
#pragma unroll
        for (int k = 0; k < NA; ++k) bu = fmaf(S(Lay::oB + i * NA + k), u[k], bu);
        s[i] += bu;
        float acc = bu;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fmaf(A[i][j], x[j], acc);
        xn[i] = acc + w[i];
      }
      if (closed)
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = xn[i];
    }
  }
  // ---- write the state back
  if (!live) return;
  if (bad) {
    if (a.cost_sum_out) a.cost_sum_out[env] = (double)nanf("");
    if (a.status_out) a.status_out[env] = DLC_STATUS_BAD_HINT;
    return;
  }
  if (closed)
#pragma unroll
    for (int i = 0; i < D; ++i) a.x_io[(size_t)env * D + i] = x[i];
#pragma unroll
  for (int j = 0; j < MH; ++j)
#pragma unroll
    for (int aa = 0; aa < NA; ++aa) {
      a.theta_io[(size_t)env * MH * NA * P + (j * NA + aa) * P] = lo(Th[aa][j]);
      a.theta_io[(size_t)env * MH * NA * P + (j * NA + aa) * P + 1] = hi(Th[aa][j]);
    }
#pragma unroll
  for (int i = 0; i < MH; ++i) {
    a.ynat_io[(size_t)env * FP + 2 * i] = lo(yn[i]);
    a.ynat_io[(size_t)env * FP + 2 * i + 1] = hi(yn[i]);
  }
  for (int i = 0; i < HH; ++i) {
    int sl = head + i; if (sl >= HH) sl -= HH;
#pragma unroll
    for (int c = 0; c < NA; ++c) a.us_io[(size_t)env * HH * NA + i * NA + c] = S(Lay::oUs + sl * NA + c);
  }
  if (a.cost_sum_out) a.cost_sum_out[env] = csum;
  if (a.status_out) a.status_out[env] = status;
}

// the shape served by the feature-blocked instance of_kernel<FGRC, 32, 10, 3, 2, 30, 121, 61> (DSC(m = 30, h = 10) on the
// colab's system) unless the caller forces 8 or 16 lanes
bool of_feature_blocked(const DlcOfParams& P) {
  return P.agent == DLC_OF_FGRC && P.d == 10 && P.n == 3 && P.p == 2 && P.m == 30 && P.F == 121 && P.L == 61 &&
         (P.lanes_per_env == 0 || P.lanes_per_env == 32);
}

// shared-memory layout of one env's slice; returns floats per env
int layout(OfArgs& a) {
  const DlcOfParams& P = a.p;
  const int d = P.d, n = P.n, p = P.p, m = P.m;
  int o = 0;
  auto take = [&](int cnt) { const int at = o; o += (cnt + 3) & ~3; return at; };
  a.J = P.agent == DLC_OF_DRC ? P.h : (P.agent == DLC_OF_FGRC ? m : 0);
  a.FP = P.agent == DLC_OF_DRC ? m * p : (P.agent == DLC_OF_FGRC ? P.F * p : 0);
  a.nPhi = P.agent == DLC_OF_FGRC ? ((P.F * P.L + 3) & ~3) : 0;
  const bool fbl = of_feature_blocked(P);   // the default DSC: transposed filter bank, padded feature rows, Theta in registers
  if (fbl) a.nPhi = P.L * ((P.F + 3) & ~3);
  a.oA = take(d * d); a.oB = take(d * n); a.oC = take(p * d);
  a.oG = take(a.J * p * n);
  a.oK = take(P.agent == DLC_OF_LQG ? n * d : (P.agent == DLC_OF_DRC ? n * p : 0));
  a.oL = take(P.agent == DLC_OF_LQG ? d * p : 0);
  a.oTh = take(fbl ? 0 : n * a.FP);
  a.oFeat = take(P.agent == DLC_OF_FGRC ? (m + 1) * (fbl ? 248 : a.FP) : 0);
  a.oYn = take(P.agent == DLC_OF_FGRC ? (fbl ? 2 : 1) * (P.L + m) * p : (P.agent == DLC_OF_DRC ? m * p : 0));
  a.oUs = take(P.agent == DLC_OF_DRC ? P.h * n : 0);
  a.oX = take(d); a.oZ = take(d); a.oY = take(p); a.oU = take(n);
  a.oAct = take((m + 1) * n); a.oGact = take((m + 1) * n); a.oFy = take(p);
  a.oTmp = take(2 * d * n > p ? 2 * d * n : p);
  return o;
}

}  // namespace
}  // namespace dlc

using namespace dlc;

static int32_t of_check(const DlcOfParams* p) {
  if (!p) return fail(DLC_ERR_ARG, "dlc_lds_of: params is NULL");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_lds_of: negative N or T");
  if (p->agent < DLC_OF_LQG || p->agent > DLC_OF_FGRC) return fail(DLC_ERR_ARG, "dlc_lds_of: unknown agent");
  if (p->mode != DLC_MODE_CLOSED_LOOP && p->mode != DLC_MODE_AGENT_ONLY) return fail(DLC_ERR_ARG, "dlc_lds_of: unknown mode");
  if (p->d < 1 || p->n < 1 || p->p < 1 || p->d > 64 || p->n > 64 || p->p > 64)
    return fail(DLC_ERR_SHAPE, "dlc_lds_of: d, n, p must be in 1..64");
  if (p->agent == DLC_OF_DRC && (p->m < 1 || p->h < 1))
    return fail(DLC_ERR_SHAPE, "dlc_lds_of: DRC needs m >= 1 and h >= 1");
  if (p->agent == DLC_OF_FGRC && (p->m < 1 || p->F < 1 || p->L < 1))
    return fail(DLC_ERR_SHAPE, "dlc_lds_of: the filter family needs m, F, L >= 1");
  if (p->agent == DLC_OF_LQG && p->m != 0) return fail(DLC_ERR_ARG, "dlc_lds_of: LQG takes m = 0");
  return DLC_OK;
}

extern "C" size_t dlc_lds_of_smem_bytes(const DlcOfParams* p) {
  if (of_check(p) != DLC_OK) return 0;
  OfArgs a{};
  a.p = *p;
  const int S = layout(a);
  return ((size_t)a.nPhi + S) * 4;
}

extern "C" int32_t dlc_lds_of_rollout_f32(const DlcOfParams* p, const float* A, const float* B, const float* C,
                                          const float* K, const float* Lk, const float* Phi, const float* W,
                                          const float* y_in, float* x_io, float* theta_io, float* z_io, float* ynat_io,
                                          float* us_io, float* cost_out, double* cost_sum_out, float* u_out,
                                          int32_t* status_out, void* stream) {
  int32_t rc = of_check(p);
  if (rc != DLC_OK) return rc;
  if (p->N == 0 || p->T == 0) return DLC_OK;
  const bool closed = p->mode == DLC_MODE_CLOSED_LOOP;
  if (!A || !B || !C) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: A, B and C are required");
  if (closed && !x_io) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: x_io is required in closed-loop mode");
  if (!closed && !y_in) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: y_in is required in agent-only mode");
  if (p->agent == DLC_OF_LQG && (!K || !Lk || !z_io)) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: LQG needs K, Lk and z_io (x_hat)");
  if (p->agent == DLC_OF_DRC && (!theta_io || !ynat_io || !us_io)) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: DRC needs theta_io (M), ynat_io and us_io");
  if (p->agent == DLC_OF_FGRC && (!Phi || !theta_io || !z_io || !ynat_io)) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: the filter family needs Phi, theta_io, z_io and ynat_io");
  OfArgs a{};
  a.p = *p;
  a.A = A; a.B = B; a.C = C; a.K = K; a.Lk = Lk; a.Phi = Phi; a.W = W; a.y_in = y_in;
  a.x_io = x_io; a.theta_io = theta_io; a.z_io = z_io; a.ynat_io = ynat_io; a.us_io = us_io;
  a.cost_out = cost_out; a.cost_sum_out = cost_sum_out; a.u_out = u_out; a.status_out = status_out;
  a.S = layout(a);
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_status(e, "cudaGetDevice");
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (p->agent == DLC_OF_LQG && p->d == 10 && p->n == 3 && p->p == 2 && p->lanes_per_env == 0) {
    // the colab's system: register-resident thread-per-env kernel (lanes_per_env = 8 / 16 / 32 still forces lane groups;
    // the identity-filter flag means nothing to LQG and is ignored)
    const unsigned grid = (unsigned)((p->N + 127) / 128);
    const char* renv = getenv("DLC_ROW_RING");   // profiling switch: 0 = disturbance rows through registers
    if ((!renv || atoi(renv) > 0) && p->mode == DLC_MODE_CLOSED_LOOP && p->T > 2 && WarpRowRing<10, 4>::usable(W, p->N)) {
      of_lqg_reg_kernel<10, 3, 2, 4><<<grid, 128, 4 * WarpRowRing<10, 4>::kBytesPerWarp, (cudaStream_t)stream>>>(a);
      return cuda_status(cudaGetLastError(), "of_lqg_reg_kernel (row ring) launch");
    }
    of_lqg_reg_kernel<10, 3, 2, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "of_lqg_reg_kernel launch");
  }
  if (p->agent == DLC_OF_FGRC && p->d == 10 && p->n == 3 && p->p == 2 && p->m == 10 && p->F == 10 && p->L == 10 &&
      (p->flags & DLC_OF_FLAG_PHI_IDENTITY) && p->lanes_per_env == 0) {
    // GRC(m = 10) on the colab's system: register-resident thread-per-env kernel
    constexpr int TPB = 64;
    const size_t smem = (size_t)GrcRegLayout<10, 3, 2, 10>::kFloats * TPB * sizeof(float);
    auto kern = of_grc_reg_kernel<10, 3, 2, 10, TPB>;
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return cuda_status(e2, "cudaFuncSetAttribute(of_grc_reg_kernel)");
    kern<<<(unsigned)((p->N + TPB - 1) / TPB), TPB, smem, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "of_grc_reg_kernel launch");
  }
  if (p->agent == DLC_OF_DRC && p->d == 10 && p->n == 3 && p->p == 2 && p->m == 10 && p->h == 50 &&
      (p->flags & DLC_OF_FLAG_STABLE_A) && p->lanes_per_env == 0) {
    // DRC(m = 10, h = 50) on the colab's system with a decaying truncated response: register-resident thread-per-env kernel
    constexpr int TPB = 64;
    size_t smem = (size_t)DrcRegLayout<10, 3, 2, 10, 50>::kFloats * TPB * sizeof(float);
    auto kern = of_drc_reg_kernel<10, 3, 2, 10, 50, TPB, 0>;
    const char* renv = getenv("DLC_ROW_RING");   // profiling switch: 0 = disturbance rows through registers
    if ((!renv || atoi(renv) > 0) && p->mode == DLC_MODE_CLOSED_LOOP && p->T > 2 && WarpRowRing<10, 4>::usable(W, p->N)) {
      kern = of_drc_reg_kernel<10, 3, 2, 10, 50, TPB, 4>;
      smem += (TPB / 32) * WarpRowRing<10, 4>::kBytesPerWarp;
    }
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return cuda_status(e2, "cudaFuncSetAttribute(of_drc_reg_kernel)");
    kern<<<(unsigned)((p->N + TPB - 1) / TPB), TPB, smem, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "of_drc_reg_kernel launch");
  }
  // lanes per env from the widest per-step contraction (lanes_per_env = 8 / 16 / 32 overrides)
  const int width = p->agent == DLC_OF_LQG ? p->d : (p->d > a.FP ? p->d : a.FP);
  int GL = width <= 20 ? 8 : (width <= 48 ? 16 : 32);   // measured: GRC / DRC / LQG 8, SFC (22 wide) 16, DSC 32
  // the default SFC (22 features wide) with compile-time dims is 12 % faster on 8 lanes than on 16 (137.9 vs 154.7 ms)
  if (p->agent == DLC_OF_FGRC && p->d == 10 && p->n == 3 && p->p == 2 && p->m == 30 && p->F == 11 && p->L == 31) GL = 8;
  int widest = p->d > p->n ? p->d : p->n;
  if (p->p > widest) widest = p->p;
  while (2 * GL < widest) GL <<= 1;   // a lane owns at most two rows of a mat-vec
  if (p->lanes_per_env == 8 || p->lanes_per_env == 16 || p->lanes_per_env == 32) {
    if (2 * p->lanes_per_env < widest) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: lanes_per_env must be >= max(d, n, p) / 2");
    GL = p->lanes_per_env;
  }
  else if (p->lanes_per_env != 0) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: lanes_per_env must be 0 (automatic), 8, 16 or 32");
  if (p->flags & ~(DLC_OF_FLAG_PHI_IDENTITY | DLC_OF_FLAG_STABLE_A)) return fail(DLC_ERR_ARG, "dlc_lds_of_rollout_f32: unknown flag bits");
  // groups of one warp hit the same offsets of their slices at once: a slice stride of GL (mod 32) banks keeps
  // their unit-stride accesses apart
  // (for the window-blocked SFC instance a stride of 16 mod 32 would spread its 8-byte feature-row reads over all even
  // banks, but it measured slower than GL mod 32 -- 61.1 vs 58.3 ms -- because every other access then collides)
  if (GL < 32) a.S += ((GL - a.S % 32) + 32) % 32;
  const size_t per_env = (size_t)a.S * 4, fixed = (size_t)a.nPhi * 4;
  if (fixed + per_env > (size_t)max_smem)
    return fail(DLC_ERR_SHAPE, "dlc_lds_of_rollout_f32: one env's controller state does not fit in shared memory");
  // envs per CTA: whole warps, as many as half the SM's shared memory holds (at most 256 threads)
  int epc = 256 / GL;
  while (epc > 32 / GL && fixed + per_env * epc > (size_t)max_smem / 2) epc >>= 1;
  // the feature-blocked DSC instance is a warp per env with 34 KB of windows each and a 30 KB filter bank per CTA: one CTA
  // per SM with as many envs as fit (5) instead of two CTAs of two
  if (GL == 32 && of_feature_blocked(*p)) {
    epc = 8;
    while (epc > 1 && fixed + per_env * epc > (size_t)max_smem) --epc;
  }
  if (fixed + per_env * epc > (size_t)max_smem) {   // not even one warp of groups: fall back to a warp per env
    GL = 32;
    epc = 1;
  }
  a.wpc = epc;
  const size_t smem = fixed + per_env * epc;
  void (*kern)(const OfArgs) = nullptr;
#define OF_PICK(AG)                                                                                          \
  if (p->agent == AG) kern = GL == 8 ? of_kernel<AG, 8> : (GL == 16 ? of_kernel<AG, 16> : of_kernel<AG, 32>);
  OF_PICK(DLC_OF_LQG) OF_PICK(DLC_OF_DRC) OF_PICK(DLC_OF_FGRC)
#undef OF_PICK
  // compile-time dims for the reference's default controllers on the colab's system
  if (p->d == 10 && p->n == 3 && p->p == 2) {
    if (p->agent == DLC_OF_LQG && GL == 8) kern = of_kernel<DLC_OF_LQG, 8, 10, 3, 2>;
    if (p->agent == DLC_OF_DRC && GL == 8 && p->m == 10 && p->h == 50) kern = of_kernel<DLC_OF_DRC, 8, 10, 3, 2, 10, 0, 0, 50>;
    if (p->agent == DLC_OF_FGRC && GL == 8 && p->m == 10 && p->F == 10 && p->L == 10)       // GRC(m = 10)
      kern = of_kernel<DLC_OF_FGRC, 8, 10, 3, 2, 10, 10, 10>;
    if (p->agent == DLC_OF_FGRC && GL == 16 && p->m == 30 && p->F == 11 && p->L == 31)      // SFC(m = 30, h = 10)
      kern = of_kernel<DLC_OF_FGRC, 16, 10, 3, 2, 30, 11, 31>;
    if (p->agent == DLC_OF_FGRC && GL == 8 && p->m == 30 && p->F == 11 && p->L == 31)       // (forced 8 lanes)
      kern = of_kernel<DLC_OF_FGRC, 8, 10, 3, 2, 30, 11, 31>;
    if (p->agent == DLC_OF_FGRC && GL == 32 && p->m == 30 && p->F == 11 && p->L == 31)      // (forced 32 lanes)
      kern = of_kernel<DLC_OF_FGRC, 32, 10, 3, 2, 30, 11, 31>;
    if (p->agent == DLC_OF_FGRC && GL == 32 && p->m == 30 && p->F == 121 && p->L == 61)     // DSC(m = 30, h = 10)
      kern = of_kernel<DLC_OF_FGRC, 32, 10, 3, 2, 30, 121, 61>;
  }
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(of_kernel)");
  const unsigned grid = (unsigned)((p->N + epc - 1) / epc);
  kern<<<grid, GL * epc, smem, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "of_kernel launch");
}
