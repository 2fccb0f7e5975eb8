// lds_lqr_reg.cu -- LDS + LQR (u = -K x, deluca/agents/_lqr.py:62-72; env deluca/envs/_lds.py:102-111) for small
// compile-time systems, one THREAD per env with A, B, K and the state in registers (170 floats at d = 10, m = 3).
// A step is 160 FFMA; the kernel is bound by the disturbance stream it reads (44 B per env-step), so W is fetched two
// steps ahead (two register buffers) to keep enough bytes in flight at 8 warps per SM.
#include "common.cuh"
#include "lds_args.cuh"

namespace dlc {
namespace {

template <int D, int M>
__global__ void __launch_bounds__(128) lds_lqr_reg_kernel(const LdsArgs a) {
  const DlcLdsParams& p = a.p;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int64_t np = p.shared_dynamics ? 0 : n;
  float A[D][D], B[D][M], K[M][D], x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + np * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      B[i][c] = __ldg(a.B + np * D * M + i * M + c);
      K[c][i] = __ldg(a.K + np * M * D + c * D + i);
    }
    x[i] = a.x_io[n * D + i];
  }
  const bool sampled = (a.x_out || a.u_out) && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;
  auto fetch = [&](int t, float (&w)[D]) {
    if (a.W) {
      const float* wr = a.W + ((int64_t)t * p.N + n) * D;
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
        float z[4];
        normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), (uint32_t)g, 0u, z);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < D) w[4 * g + k] = __fmul_rn(p.w_std, z[k]);
      }
    }
  };
  auto advance = [&](int t, const float (&w)[D]) {
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(K[c][j], x[j], s);
      u[c] = -s;
    }
    float cost = 0.f, cu = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) cost = fmaf(x[i], x[i], cost);
#pragma unroll
    for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
    cost += cu;
    csum += (double)cost;
    if (a.cost_out) a.cost_out[(int64_t)t * p.N + n] = cost;
    if (sampled) {
      if (a.x_out)
#pragma unroll
        for (int i = 0; i < D; ++i) a.x_out[((int64_t)t * Ns + ns) * D + i] = x[i];
      if (a.u_out)
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
    }
    float xn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(A[i][j], x[j], s);
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(B[i][c], u[c], s);
      xn[i] = s + w[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
  };
  float w0[D], w1[D];
  if (p.T > 0) fetch(0, w0);
  if (p.T > 1) fetch(1, w1);
  for (int t = 0; t < p.T; t += 2) {
    {
      float w[D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = w0[i];
      if (t + 2 < p.T) fetch(t + 2, w0);
      advance(t, w);
    }
    if (t + 1 < p.T) {
      float w[D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = w1[i];
      if (t + 3 < p.T) fetch(t + 3, w1);
      advance(t + 1, w);
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) a.x_io[n * D + i] = x[i];
  if (a.cost_sum_out) a.cost_sum_out[n] = csum;
  if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
}

}  // namespace

bool lqr_reg_supported(const DlcLdsParams& p) {
  return p.policy == DLC_POLICY_LQR && p.mode == DLC_MODE_CLOSED_LOOP && p.kernel_variant == 0 && p.d == 10 && p.m == 3;
}

bool launch_lqr_reg(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  if (!lqr_reg_supported(a.p)) return false;
  lds_lqr_reg_kernel<10, 3><<<(unsigned)((a.p.N + 127) / 128), 128, 0, st>>>(a);
  *rc = cuda_status(cudaGetLastError(), "lds_lqr_reg_kernel launch");
  return true;
}

}  // namespace dlc
