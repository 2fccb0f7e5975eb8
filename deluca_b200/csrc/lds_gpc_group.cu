// lds_gpc_group.cu -- LDS + GPC for ANY (d, m, H, HH) with d, m <= 64: the runtime-dims path that is not slow.
//
//   GPC.__call__ / get_action   deluca/agents/_gpc.py:130-183     u = -K x + sum_i M[i] hist[HH + i]
//   GPC.update                  deluca/agents/_gpc.py:145-169     noise = x - A x_prev - B u; roll, append; M -= lr dM
//   policy_loss                 deluca/agents/_gpc.py:109-125     H-step counterfactual with lax.dynamic_slice clamping
//   LDS.__call__                deluca/envs/_lds.py:102-111
//
// The compile-time-shape kernels (lds_split_kernel, lds_f2_kernel) cover the benchmark shapes; every other shape used
// to run lds_generic_kernel, one thread per env with its state in global memory (~9e7 env-steps/s).  Here a group of
// GL = 8 / 16 / 32 lanes owns an env and A, B, K, M, the noise ring and every work vector stay in the group's slice of
// shared memory for all T steps (the mapping of lds_of.cu): lanes split the rows of each mat-vec, the columns of each
// transposed mat-vec and the d-axis of the M-window products and rank-1 updates, and meet through group shuffles.
// The arithmetic follows lds_generic_kernel statement by statement (same clamping when HH < H - 1, same hand adjoint).
#include "common.cuh"
#include "lds_args.cuh"

namespace dlc {
namespace {

struct GroupArgs {
  LdsArgs a;
  int epc;  // envs per CTA
  int S;    // floats of shared memory per env
};

template <int GL>
__device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int off = GL / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ float pick4g(const float z[4], int k) { return k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3])); }

// r[k] (+)= sum_c Mat[row][c] v[c], row = gl + k GL < rows (rows <= 2 GL)
template <int GL, bool ACC>
__device__ __forceinline__ void mv(const float* Mat, const float* v, int rows, int cols, int gl, float (&r)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (!ACC) r[k] = 0.f;
    if (k * GL < rows) {
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
// r[k] (+)= sum_i Mat[i][col] v[i], col = gl + k GL < cols: Mat^T v (consecutive lanes read consecutive floats)
template <int GL, bool ACC>
__device__ __forceinline__ void mvt(const float* Mat, const float* v, int rows, int cols, int gl, float (&r)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (!ACC) r[k] = 0.f;
    if (k * GL < cols) {
      const int col = gl + k * GL;
      if (col < cols) {
        float acc = r[k];
        for (int i = 0; i < rows; ++i) acc = fmaf(Mat[i * cols + col], v[i], acc);
        r[k] = acc;
      }
    }
  }
}
template <int GL>
__device__ __forceinline__ void put(float* dst, int len, int gl, const float (&r)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (gl + k * GL < len) dst[gl + k * GL] = r[k];
}

// SD..SHH: compile-time (d, m, H, HH), 0 = runtime (the contractions then unroll into LDS + FFMA with immediate offsets)
template <int GL, int SD = 0, int SM = 0, int SH = 0, int SHH = 0>
__global__ void __launch_bounds__(256, GL == 32 ? 4 : 2) lds_gpc_group_kernel(const GroupArgs g) {
  extern __shared__ __align__(16) float smem[];
  const LdsArgs& a = g.a;
  const DlcLdsParams& p = a.p;
  const int gl = threadIdx.x & (GL - 1), grp = threadIdx.x / GL;
  const int d = SD ? SD : p.d, m = SD ? SM : p.m, H = SD ? SH : p.H, HH = SD ? SHH : p.HH, P = H + HH, HMD = H * m * d;
  int64_t n = (int64_t)blockIdx.x * g.epc + grp;
  if (GL == 32 && n >= p.N) return;
  const bool live = n < p.N;
  if (!live) n = p.N - 1;   // idle groups shadow the last env: warp-wide __syncwarp / shuffles stay whole
  float* s = smem + (size_t)grp * g.S;
  float *sA = s, *sB = sA + d * d, *sK = sB + d * m, *sM = sK + m * d, *sH = sM + HMD;
  float *sx = sH + P * d, *sxp = sx + d, *sy = sxp + d, *slam = sy + d;
  float *su = slam + d, *sup = su + m, *sv = sup + m, *sdu = sv + m, *smw = sdu + m;
  const int64_t np = p.shared_dynamics ? 0 : n;
  for (int i = gl; i < d * d; i += GL) sA[i] = a.A[np * d * d + i];
  for (int i = gl; i < d * m; i += GL) sB[i] = a.B[np * d * m + i];
  for (int i = gl; i < m * d; i += GL) sK[i] = a.K[np * m * d + i];
  for (int i = gl; i < HMD; i += GL) sM[i] = a.M_io[n * (int64_t)HMD + i];
  for (int i = gl; i < P * d; i += GL) sH[i] = a.hist_io[n * (int64_t)P * d + i];
  for (int i = gl; i < d; i += GL) { sx[i] = a.x_io[n * d + i]; sxp[i] = a.xprev_io[n * d + i]; }
  for (int i = gl; i < m; i += GL) sup[i] = a.uprev_io ? a.uprev_io[n * m + i] : 0.f;
  __syncwarp();
  int hh = 0;   // physical slot of the oldest noise: logical slot k lives at (hh + k) % P
  auto slot = [&](int k) -> const float* { int q = hh + k; if (q >= P) q -= P; return sH + q * d; };
  auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
  // out[c] = sum_i sum_j M[i][c][j] hist[start + i][j], c < m: lanes split j, one group reduction per c
  auto win_dot = [&](int start, float* out) {
    for (int c = 0; c < m; ++c) {
      float part = 0.f;
      for (int i = 0; i < H; ++i) {
        const float* hrow = slot(start + i);
        const float* mrow = sM + (i * m + c) * d;
        for (int j = gl; j < d; j += GL) part = fmaf(mrow[j], hrow[j], part);
      }
      part = gsum<GL>(part);
      if (gl == 0) out[c] = part;
    }
    __syncwarp();
  };
  const bool sampled = live && (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;

  for (int t = 0; t < p.T; ++t) {
    // disturbance of this step (own rows), fetched early
    float w[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      w[k] = 0.f;
      const int row = gl + k * GL;
      if (k * GL < d && row < d) {
        if (a.W) {
          w[k] = a.W[((int64_t)t * p.N + n) * d + row];
        } else {
          float z4[4];
          normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), (uint32_t)(row >> 2), 0u, z4);
          w[k] = __fmul_rn(p.w_std, pick4g(z4, row & 3));
        }
      }
    }
    // ---- action u = -K x + sum_i M[i] hist[HH + i]                      _gpc.py:171-183
    win_dot(P - H, smw);
    {
      float r[2];
      mv<GL, false>(sK, sx, m, d, gl, r);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < m) su[gl + k * GL] = smw[gl + k * GL] - r[k];
    }
    __syncwarp();
    // ---- noise = x - A x_prev - B u (the action just computed, as written); roll(-1), append      _gpc.py:155-157
    {
      float ra[2], rb[2];
      mv<GL, false>(sA, sxp, d, d, gl, ra);
      mv<GL, false>(sB, p.noise_uses_prev_action ? sup : su, d, m, gl, rb);
      float* oldest = sH + hh * d;   // becomes the newest logical slot P - 1
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < d) oldest[gl + k * GL] = (sx[gl + k * GL] - ra[k]) - rb[k];
      hh = (hh + 1 == P) ? 0 : hh + 1;
    }
    __syncwarp();
    // ---- policy_loss forward (windows clamp like lax.dynamic_slice)       _gpc.py:109-125
    for (int i = gl; i < d; i += GL) sy[i] = 0.f;
    __syncwarp();
    for (int h = 0; h < H - 1; ++h) {
      win_dot(clampi(h, 0, P - H), sv);
      float r[2];
      mv<GL, false>(sK, sy, m, d, gl, r);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < m) sv[gl + k * GL] -= r[k];
      __syncwarp();
      const float* wh = slot((h + H < P - 1) ? h + H : P - 1);
      mv<GL, false>(sA, sy, d, d, gl, r);
      mv<GL, true>(sB, sv, d, m, gl, r);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < d) sy[gl + k * GL] = r[k] + wh[gl + k * GL];
      __syncwarp();
    }
    const int fstart = clampi(HH - 1, 0, P - H);
    win_dot(fstart, sv);
    const float lr = p.decay ? p.lr_scale * (1.0f / (float)(p.t0 + t + 1)) : p.lr_scale;
    {
      float r[2];
      mv<GL, false>(sK, sy, m, d, gl, r);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < m) {                                                    // dL/du_f = 2 r (.) u_f
          const float wr = a.cost_r ? 2.0f * __ldg(a.cost_r + gl + k * GL) : 2.0f;
          sdu[gl + k * GL] = wr * (sv[gl + k * GL] - r[k]);
        }
    }
    __syncwarp();
    {
      float r[2];
      mvt<GL, false>(sK, sdu, m, d, gl, r);   // K' du
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < d) {                                                    // lam = 2 q (.) y_f - K' du
          const float wq = a.cost_q ? 2.0f * __ldg(a.cost_q + gl + k * GL) : 2.0f;
          slam[gl + k * GL] = wq * sy[gl + k * GL] - r[k];
        }
    }
    // adjoint, M updated in place (the backward pass never reads M): M[i][c][:] -= lr du[c] hist[fstart + i][:]
    auto rank1 = [&](int start, const float* coef) {
      for (int i = 0; i < H; ++i) {
        const float* hrow = slot(start + i);
        for (int c = 0; c < m; ++c) {
          const float f = -lr * coef[c];
          float* mrow = sM + (i * m + c) * d;
          for (int j = gl; j < d; j += GL) mrow[j] = fmaf(f, hrow[j], mrow[j]);
        }
      }
    };
    rank1(fstart, sdu);
    __syncwarp();
    for (int h = H - 2; h >= 0; --h) {
      float r[2];
      mvt<GL, false>(sB, slam, d, m, gl, r);   // lu = B' lam
      put<GL>(sv, m, gl, r);
      __syncwarp();
      rank1(clampi(h, 0, P - H), sv);
      mvt<GL, false>(sA, slam, d, d, gl, r);   // lam <- A' lam - K' lu
      float rk[2];
      mvt<GL, false>(sK, sv, m, d, gl, rk);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < d) slam[gl + k * GL] = r[k] - rk[k];
      __syncwarp();
    }
    // ---- x_prev = x, u_prev = u; realised cost quad_loss(x, u); outputs
    float cx = 0.f, cu = 0.f;
    for (int i = gl; i < d; i += GL) { cx = fmaf(sx[i], sx[i], cx); sxp[i] = sx[i]; }
    for (int i = gl; i < m; i += GL) { cu = fmaf(su[i], su[i], cu); sup[i] = su[i]; }
    const float cost = gsum<GL>(cx) + gsum<GL>(cu);
    csum += (double)cost;
    if (gl == 0 && live && a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (sampled) {
      if (a.x_out) for (int i = gl; i < d; i += GL) a.x_out[((int64_t)t * Ns + ns) * d + i] = sx[i];
      if (a.u_out) for (int i = gl; i < m; i += GL) a.u_out[((int64_t)t * Ns + ns) * m + i] = su[i];
    }
    // ---- env step x <- A x + B u + w                                      _lds.py:107-109
    {
      float r[2];
      mv<GL, false>(sA, sx, d, d, gl, r);
      mv<GL, true>(sB, su, d, m, gl, r);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (gl + k * GL < d) sx[gl + k * GL] = r[k] + w[k];
    }
    __syncwarp();
  }
  if (!live) return;
  for (int i = gl; i < d; i += GL) { a.x_io[n * d + i] = sx[i]; a.xprev_io[n * d + i] = sxp[i]; }
  for (int i = gl; i < HMD; i += GL) a.M_io[n * (int64_t)HMD + i] = sM[i];
  for (int e = gl; e < P * d; e += GL) {   // oldest -> newest
    const int k = e / d, j = e - k * d;
    a.hist_io[n * (int64_t)P * d + e] = slot(k)[j];
  }
  if (a.uprev_io) for (int i = gl; i < m; i += GL) a.uprev_io[n * m + i] = sup[i];
  if (gl == 0) {
    if (a.cost_sum_out) a.cost_sum_out[n] = csum;
    if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
  }
}

}  // namespace

bool gpc_group_supported(const DlcLdsParams& p) {
  return p.policy == DLC_POLICY_GPC && p.mode == DLC_MODE_CLOSED_LOOP && p.kernel_variant != 1 && p.d <= 64 &&
         p.m <= 64 && p.H >= 1 && p.HH >= 1;
}

// Returns false (and launches nothing) if the shape is not for this kernel or one env does not fit shared memory.
bool launch_gpc_group(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  const DlcLdsParams& p = a.p;
  if (!gpc_group_supported(p)) return false;
  const int d = p.d, m = p.m, P = p.H + p.HH;
  int S = d * d + 2 * d * m + p.H * m * d + P * d + 4 * d + 5 * m;
  S = (S + 3) & ~3;
  const int widest = d > m ? d : m;
  const int GL = widest <= 16 ? 8 : (widest <= 32 ? 16 : 32);
  if (GL < 32) S += ((GL - S % 32) + 32) % 32;
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int epc = 256 / GL;
  while (epc > 32 / GL && (size_t)S * 4 * epc > (size_t)max_smem / 2) epc >>= 1;
  if ((size_t)S * 4 * epc > (size_t)max_smem) return false;   // falls through to the thread-per-env kernel
  GroupArgs g{a, epc, S};
  const size_t smem = (size_t)S * 4 * epc;
  void (*kern)(const GroupArgs) = GL == 8 ? lds_gpc_group_kernel<8> : (GL == 16 ? lds_gpc_group_kernel<16>
                                                                                : lds_gpc_group_kernel<32>);
  // the reference's own defaults: LDS(d_hidden = 25, d_in = 1) + GPC(H = 3, HH = 2)
  if (d == 25 && m == 1 && p.H == 3 && p.HH == 2 && GL == 16) kern = lds_gpc_group_kernel<16, 25, 1, 3, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *rc = cuda_status(e, "cudaFuncSetAttribute(lds_gpc_group_kernel)"); return true; }
  const unsigned grid = (unsigned)((p.N + epc - 1) / epc);
  kern<<<grid, GL * epc, smem, st>>>(g);
  *rc = cuda_status(cudaGetLastError(), "lds_gpc_group_kernel launch");
  return true;
}

}  // namespace dlc
