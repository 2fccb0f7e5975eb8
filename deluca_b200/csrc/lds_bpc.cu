// lds_bpc.cu -- LDS + BPC (bandit perturbation controller) rollout, fused over T steps, for compile-time shapes.
//
//   BPC.__call__   deluca/agents/_bpc.py:97-158      u = -K x + sum_i M[i] hist[i]
//                  :124-126                           noise = x - A x_prev - B u ; hist <- roll(hist, -1), append
//                  :128-132                           M <- M - lr * cost * sum_q eps[q]      (cost of the previous step)
//                  :134-139                           eps <- roll(eps, -1), append a unit-norm Gaussian [H,m,d]; M += delta eps[-1]
//   LDS.__call__   deluca/envs/_lds.py:102-111        x <- A x + B u + w
//
// Mapping (the lane-split layout of lds_split_kernel): L adjacent lanes own an env; lane q keeps, in registers, the rows
// S_q of A and B and the columns S_q of K and of every M[i], S_q = [q DL, q DL + DL).  What is indexed by time -- the
// noise history and the H-deep ring of perturbations (H * H m d floats per env: 3 KB at the reference's H = 5) -- lives
// in shared memory as [slot][element][thread] (bank = thread).  The new perturbation is the expensive part of a step
// (H m d normals: 38 Philox blocks at H = 5): its blocks are dealt round-robin to the L lanes, parked in a per-env
// staging strip, normalised by the group-reduced norm and picked up by the lane that owns each column.
// The runtime-dims lds_generic_kernel stays the path for every other shape (and kernel_variant = 1 forces it).
#include "common.cuh"
#include "lds_args.cuh"

namespace dlc {
namespace {

template <int L>
__device__ __forceinline__ float bpc_allreduce(float v) {
#pragma unroll
  for (int o = 1; o < L; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int D, int M, int H, int L, int TPB>
struct BpcLayout {
  static constexpr int DL = (D + L - 1) / L;
  static constexpr int HMD = H * M * D;
  static constexpr int NG = (HMD + 3) / 4;           // Philox blocks of one perturbation
  static constexpr int EPC = TPB / L;                // envs per CTA
  static constexpr int EL = H * M * DL;              // perturbation elements a lane owns
  static constexpr int kPerThread = H * DL + H * EL; // noise ring + perturbation ring
  static constexpr int kStage = NG * 4 * (EPC + 1);  // staging strip, [k][env] with a one-float skew
  static constexpr size_t kSmemBytes = ((size_t)kPerThread * TPB + kStage) * sizeof(float);
};

template <int D, int M, int H, int L, int TPB>
__global__ void __launch_bounds__(TPB) lds_bpc_kernel(const LdsArgs a, float* __restrict__ cost_io) {
  using Lay = BpcLayout<D, M, H, L, TPB>;
  constexpr int DL = Lay::DL, DF = DL * L, HMD = Lay::HMD, NG = Lay::NG, EPC = Lay::EPC, EL = Lay::EL;
  extern __shared__ float sm[];
  const DlcLdsParams& p = a.p;
  const int tid = threadIdx.x, q = tid % L, e = tid / L;
  const int64_t n_raw = (int64_t)blockIdx.x * EPC + e;
  const bool live = n_raw < p.N;
  const int64_t n = live ? n_raw : p.N - 1;   // idle groups shadow the last env (keeps shuffles whole)
  float* const sHist = sm + tid;                                   // (slot, k)      -> sHist[(slot*DL + k) * TPB]
  float* const sEps = sm + (size_t)H * DL * TPB + tid;             // (slot, el)     -> sEps[(slot*EL + el) * TPB]
  float* const sStage = sm + (size_t)Lay::kPerThread * TPB + e;    // flat element k -> sStage[k * (EPC + 1)]

  // ---- parameters -> registers (A's columns in lane-relative order: chunk c holds the coordinates of lane q^c)
  const int64_t np = p.shared_dynamics ? 0 : n;
  float Ar[DL][DF], Br[DL][M], Kc[M][DL], Mr[H][M][DL];
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
#pragma unroll
  for (int i = 0; i < H; ++i)
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k)
        Mr[i][c][k] = (q * DL + k < D) ? a.M_io[n * HMD + (i * M + c) * D + q * DL + k] : 0.f;
  // rings: logical index s (oldest -> newest) lives at physical slot (head + s) % H
#pragma unroll
  for (int s = 0; s < H; ++s) {
#pragma unroll
    for (int k = 0; k < DL; ++k)
      sHist[(s * DL + k) * TPB] = (q * DL + k < D) ? a.hist_io[n * (H * D) + s * D + q * DL + k] : 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          sEps[(s * EL + (i * M + c) * DL + k) * TPB] =
              (q * DL + k < D) ? a.eps_io[(n * H + s) * HMD + (i * M + c) * D + q * DL + k] : 0.f;
  }
  int hh = 0;  // hist head
  int he = 0;  // eps head
  float x[DL], xp[DL], ax[DL], uprev[M];
#pragma unroll
  for (int k = 0; k < DL; ++k) {
    x[k] = (q * DL + k < D) ? a.x_io[n * D + q * DL + k] : 0.f;
    xp[k] = (q * DL + k < D) ? a.xprev_io[n * D + q * DL + k] : 0.f;
  }
#pragma unroll
  for (int c = 0; c < M; ++c) uprev[c] = a.uprev_io ? a.uprev_io[n * M + c] : 0.f;
  // gather a vector split over the group into lane-relative order
  auto gather = [&](const float v[DL], float f[DF]) {
#pragma unroll
    for (int k = 0; k < DL; ++k) f[k] = v[k];
#pragma unroll
    for (int c = 1; c < L; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k) f[c * DL + k] = __shfl_xor_sync(0xffffffffu, v[k], c);
  };
  {
    float xf[DF];
    gather(xp, xf);
#pragma unroll
    for (int i = 0; i < DL; ++i) {   // A x_prev (_bpc.py:124); afterwards the env step hands A x to the next step
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < DF; ++r) s = fmaf(Ar[i][r], xf[r], s);
      ax[i] = s;
    }
  }
  const bool sampled = live && (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;
  float prev_cost = cost_io ? cost_io[n] : 0.f;
  __syncwarp();

  for (int t = 0; t < p.T; ++t) {
    // ---- disturbance of this step (own coordinates), fetched early
    float w[DL];
    if (a.W) {
      const float* gw = a.W + ((int64_t)t * p.N + n) * D + q * DL;
#pragma unroll
      for (int k = 0; k < DL; ++k) w[k] = (q * DL + k < D) ? __ldg(gw + k) : 0.f;
    } else {
      constexpr int NGW = L == 1 ? (DL + 3) / 4 : (DL + 6) / 4;   // own coordinates span at most this many blocks
      const int g0 = (q * DL) / 4;
      float z[NGW * 4];
#pragma unroll
      for (int g = 0; g < NGW; ++g) normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), g0 + g, 0u, z + 4 * g);
      const int off = q * DL - g0 * 4;
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        float zz = z[k];
#pragma unroll
        for (int o = 1; o < 4; ++o)
          if (k + o < NGW * 4) zz = (off == o) ? z[k + o] : zz;
        w[k] = (q * DL + k < D) ? __fmul_rn(p.w_std, zz) : 0.f;
      }
    }
    // ---- the new perturbation: blocks dealt round-robin to the lanes, parked unnormalised
    float ss = 0.f;
    if (a.eps_new) {
      const float* src = a.eps_new + ((int64_t)t * p.N + n) * HMD;
      for (int k = q; k < HMD; k += L) sStage[k * (EPC + 1)] = src[k];
    } else {
#pragma unroll 2
      for (int g = q; g < NG; g += L) {
        float z[4];
        normal4(p.seed, env_id, (uint32_t)(p.t0 + t), (uint32_t)g, 1u, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sStage[(4 * g + k) * (EPC + 1)] = z[k];
          if (4 * g + k < HMD) ss = fmaf(z[k], z[k], ss);
        }
      }
    }
    // ---- u = -K x + sum_i M[i] hist[i]                              _bpc.py:156-158
    float red[2 * M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DL; ++k) s = fmaf(Kc[c][k], x[k], s);
      red[c] = s;
      red[M + c] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      int sl = hh + i; if (sl >= H) sl -= H;
#pragma unroll
      for (int k = 0; k < DL; ++k) {
        const float hk = sHist[(sl * DL + k) * TPB];
#pragma unroll
        for (int c = 0; c < M; ++c) red[M + c] = fmaf(Mr[i][c][k], hk, red[M + c]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2 * M; ++r) red[r] = bpc_allreduce<L>(red[r]);
    ss = bpc_allreduce<L>(ss);   // also orders the staging writes before the reads below (shuffles converge the group)
    __syncwarp();
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) u[c] = red[M + c] - red[c];
    // ---- noise = x - A x_prev - B u ; the oldest history slot takes it      _bpc.py:124-126
#pragma unroll
    for (int i = 0; i < DL; ++i) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], p.noise_uses_prev_action ? uprev[c] : u[c], s);
      sHist[(hh * DL + i) * TPB] = (x[i] - ax[i]) - s;
    }
    hh = (hh + 1 == H) ? 0 : hh + 1;
    // ---- M <- M - lr * cost * sum_q eps[q]; ring takes the new perturbation; M += delta eps[-1]   _bpc.py:128-139
    {
      float lr = p.lr_scale;
      if (p.decay) lr = p.lr_scale * (1.0f / (powf((float)(p.t0 + t), 0.75f) + 1.0f));
      const float inv = a.eps_new ? 1.0f : 1.0f / sqrtf(ss);
      float* enew = sEps + (size_t)he * EL * TPB;   // the oldest slot becomes the newest
#pragma unroll
      for (int i = 0; i < H; ++i)
#pragma unroll
        for (int c = 0; c < M; ++c)
#pragma unroll
          for (int k = 0; k < DL; ++k) {
            const int el = (i * M + c) * DL + k;
            float s = 0.f;
            int sl = he;
#pragma unroll
            for (int qq = 0; qq < H; ++qq) {   // oldest -> newest, like jnp.sum(eps, axis=0)
              s += sEps[(sl * EL + el) * TPB];
              sl = (sl + 1 == H) ? 0 : sl + 1;
            }
            float mv = Mr[i][c][k] - lr * (prev_cost * s);
            const bool own = q * DL + k < D;
            const float en = own ? sStage[((i * M + c) * D + q * DL + k) * (EPC + 1)] * inv : 0.f;
            enew[el * TPB] = en;
            Mr[i][c][k] = own ? fmaf(p.bpc_delta, en, mv) : 0.f;
          }
      he = (he + 1 == H) ? 0 : he + 1;
    }
    // ---- realised cost quad_loss(x, u)
    float cx = 0.f, cu = 0.f;
#pragma unroll
    for (int k = 0; k < DL; ++k) cx = fmaf(x[k], x[k], cx);
    cx = bpc_allreduce<L>(cx);
#pragma unroll
    for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
    const float cost = cx + cu;
    prev_cost = cost;
    csum += (double)cost;
    if (live && q == 0 && a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (sampled) {
      if (a.x_out)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (q * DL + k < D) a.x_out[((int64_t)t * Ns + ns) * D + q * DL + k] = x[k];
      if (a.u_out && q == 0)
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
    }
    // ---- env step x <- A x + B u + w; A x is kept for the next step's noise        _lds.py:107-109
    {
      float xf[DF];
      gather(x, xf);
#pragma unroll
      for (int k = 0; k < DL; ++k) xp[k] = x[k];
#pragma unroll
      for (int i = 0; i < DL; ++i) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < DF; ++r) s = fmaf(Ar[i][r], xf[r], s);
        ax[i] = s;
#pragma unroll
        for (int c = 0; c < M; ++c) s = fmaf(Br[i][c], u[c], s);
        x[i] = s + w[i];
      }
#pragma unroll
      for (int c = 0; c < M; ++c) uprev[c] = u[c];
    }
    __syncwarp();   // the staging strip is rewritten at the top of the next step
  }

  // ---- write the state back (logical order)
  if (!live) return;
#pragma unroll
  for (int k = 0; k < DL; ++k)
    if (q * DL + k < D) {
      a.x_io[n * D + q * DL + k] = x[k];
      a.xprev_io[n * D + q * DL + k] = xp[k];
    }
#pragma unroll
  for (int i = 0; i < H; ++i)
#pragma unroll
    for (int c = 0; c < M; ++c)
#pragma unroll
      for (int k = 0; k < DL; ++k)
        if (q * DL + k < D) a.M_io[n * HMD + (i * M + c) * D + q * DL + k] = Mr[i][c][k];
  for (int s = 0; s < H; ++s) {
    int sh = hh + s; if (sh >= H) sh -= H;
    int se = he + s; if (se >= H) se -= H;
#pragma unroll
    for (int k = 0; k < DL; ++k)
      if (q * DL + k < D) a.hist_io[n * (H * D) + s * D + q * DL + k] = sHist[(sh * DL + k) * TPB];
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int c = 0; c < M; ++c)
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (q * DL + k < D)
            a.eps_io[(n * H + s) * HMD + (i * M + c) * D + q * DL + k] = sEps[(se * EL + (i * M + c) * DL + k) * TPB];
  }
  if (q == 0) {
    if (a.uprev_io)
#pragma unroll
      for (int c = 0; c < M; ++c) a.uprev_io[n * M + c] = uprev[c];
    if (cost_io) cost_io[n] = prev_cost;
    if (a.cost_sum_out) a.cost_sum_out[n] = csum;
    if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
  }
}

template <int D, int M, int H, int L, int TPB>
int32_t launch_bpc_shape(const LdsArgs& a, float* cost_io, cudaStream_t st) {
  using Lay = BpcLayout<D, M, H, L, TPB>;
  auto kern = lds_bpc_kernel<D, M, H, L, TPB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::kSmemBytes);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_bpc_kernel)");
  const unsigned grid = (unsigned)((a.p.N + Lay::EPC - 1) / Lay::EPC);
  kern<<<grid, TPB, Lay::kSmemBytes, st>>>(a, cost_io);
  return cuda_status(cudaGetLastError(), "lds_bpc_kernel launch");
}

}  // namespace

// BPC on the compile-time shapes: (d, m, H) = (10, 3, 5) the reference default, (10, 3, 3), and a small one for tests.
#define DLC_BPC_SHAPES(X) X(10, 3, 5, 4, 64) X(10, 3, 3, 4, 128) X(4, 2, 2, 2, 64)

bool bpc_supported(const DlcLdsParams& p) {
  if (p.policy != DLC_POLICY_BPC || p.mode != DLC_MODE_CLOSED_LOOP || p.kernel_variant == 1) return false;
#define X(D_, M_, H_, L_, TPB_) \
  if (p.d == D_ && p.m == M_ && p.H == H_) return true;
  DLC_BPC_SHAPES(X)
#undef X
  return false;
}

bool launch_bpc(const LdsArgs& a, float* cost_io, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  if (!bpc_supported(p)) return false;
#define X(D_, M_, H_, L_, TPB_)                                      \
  if (p.d == D_ && p.m == M_ && p.H == H_) {                         \
    *rc = launch_bpc_shape<D_, M_, H_, L_, TPB_>(a, cost_io, st);    \
    return true;                                                     \
  }
  DLC_BPC_SHAPES(X)
#undef X
  return false;
}

}  // namespace dlc
