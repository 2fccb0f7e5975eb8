// Reverse mode through the LDS + linear-policy rollout: loss and d loss / d K per env.
// One thread per env; A, B, K in shared memory laid out [element][thread] (conflict-free), the state
// checkpoints x_t in an HBM workspace laid out [t][k][env] (coalesced), adjoint and dK in registers.
#include "common.cuh"

namespace dlc {

struct LdsGradArgs {
  int64_t N, env_offset;
  int T, shared, t0;
  const float *A, *B, *K, *x0, *W;
  uint64_t seed;
  float w_std;
  float *loss_out, *gradK_out, *gradx0_out;
  double* sums_out;
  float* ws;
};

template <int D, int M, int TPB>
__global__ void __launch_bounds__(TPB) lds_lqr_grad_kernel(const LdsGradArgs a) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  const int64_t n_raw = (int64_t)blockIdx.x * TPB + tid;
  const bool live = n_raw < a.N;
  const int64_t n = live ? n_raw : a.N - 1;
  float* sA = sm + tid;                       // A[i][j] at sA[(i*D+j)*TPB]
  float* sB = sm + (D * D) * TPB + tid;       // B[i][c]
  float* sK = sm + (D * D + D * M) * TPB + tid;   // K[c][j]
  const int64_t np = a.shared ? 0 : n;
  for (int i = 0; i < D * D; ++i) sA[i * TPB] = a.A[np * D * D + i];
  for (int i = 0; i < D * M; ++i) sB[i * TPB] = a.B[np * D * M + i];
  for (int i = 0; i < M * D; ++i) sK[i * TPB] = a.K[np * M * D + i];
  float x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = a.x0[n * D + i];
  const uint32_t env_id = (uint32_t)(a.env_offset + n);
  float* ws = a.ws + n;
  float total = 0.f;
  // ---- forward (same arithmetic as the rollout kernels), checkpointing x_t
  for (int t = 0; t < a.T; ++t) {
#pragma unroll
    for (int i = 0; i < D; ++i) ws[((int64_t)t * D + i) * a.N] = x[i];
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(sK[(c * D + j) * TPB], x[j], s);
      u[c] = -s;
    }
    float cost = 0.f, cu = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) cost = fmaf(x[i], x[i], cost);
#pragma unroll
    for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
    total += cost + cu;
    float w[(D + 3) / 4 * 4];
    if (a.W) {
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = a.W[((int64_t)t * a.N + n) * D + i];
    } else {
#pragma unroll
      for (int g = 0; g < (D + 3) / 4; ++g) {
        float z[4];
        normal4(a.seed, env_id, (uint32_t)(a.t0 + t), g, 0u, z);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[g * 4 + k] = __fmul_rn(a.w_std, z[k]);
      }
    }
    float xn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(sA[(i * D + j) * TPB], x[j], s);
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(sB[(i * M + c) * TPB], u[c], s);
      xn[i] = s + w[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
  }
  // ---- backward
  float lam[D], gK[M * D];
#pragma unroll
  for (int i = 0; i < D; ++i) lam[i] = 0.f;
#pragma unroll
  for (int i = 0; i < M * D; ++i) gK[i] = 0.f;
  for (int t = a.T - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = ws[((int64_t)t * D + i) * a.N];
    float gu[M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f, bl = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(sK[(c * D + j) * TPB], x[j], s);       // -u_c
#pragma unroll
      for (int i = 0; i < D; ++i) bl = fmaf(sB[(i * M + c) * TPB], lam[i], bl);   // (B' lam)_c
      gu[c] = bl - 2.0f * s;                                                      // 2 u + B' lam
#pragma unroll
      for (int j = 0; j < D; ++j) gK[c * D + j] = fmaf(-gu[c], x[j], gK[c * D + j]);
    }
    float ln[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float s = 2.0f * x[j];
#pragma unroll
      for (int i = 0; i < D; ++i) s = fmaf(sA[(i * D + j) * TPB], lam[i], s);
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(-sK[(c * D + j) * TPB], gu[c], s);
      ln[j] = s;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) lam[j] = ln[j];
  }
  if (live) {
    a.loss_out[n] = total;
#pragma unroll
    for (int i = 0; i < M * D; ++i) a.gradK_out[n * M * D + i] = gK[i];
    if (a.gradx0_out) {
#pragma unroll
      for (int i = 0; i < D; ++i) a.gradx0_out[n * D + i] = lam[i];
    }
  }
  if (a.sums_out) {   // warp reduction then one atomic per warp and component
    for (int i = -1; i < M * D; ++i) {
      double v = live ? (double)(i < 0 ? total : gK[i < 0 ? 0 : i]) : 0.0;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) atomicAdd(&a.sums_out[i + 1], v);
    }
  }
}

template <int D, int M>
static int32_t launch_grad(const LdsGradArgs& a, cudaStream_t st) {
  constexpr int TPB = 64;
  const size_t smem = (size_t)(D * D + 2 * D * M) * TPB * sizeof(float);
  auto kern = lds_lqr_grad_kernel<D, M, TPB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_lqr_grad)");
  kern<<<(unsigned)((a.N + TPB - 1) / TPB), TPB, smem, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_lqr_grad_kernel launch");
}

}  // namespace dlc

using namespace dlc;

extern "C" size_t dlc_lds_lqr_grad_workspace_bytes(int64_t N, int32_t T, int32_t d) {
  if (N <= 0 || T <= 0 || d <= 0) return 0;
  return (size_t)N * (size_t)T * (size_t)d * sizeof(float);
}

extern "C" int32_t dlc_lds_lqr_grad_f32(int64_t N, int32_t T, int32_t d, int32_t m, int32_t shared_dynamics,
                                        const float* A, const float* B, const float* K, const float* x0,
                                        const float* W, uint64_t seed, int64_t env_offset, int32_t t0, float w_std,
                                        float* loss_out, float* gradK_out, float* gradx0_out, double* sums_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || T < 0) return fail(DLC_ERR_ARG, "dlc_lds_lqr_grad_f32: negative N or T");
  if (N == 0) return DLC_OK;
  if (!A || !B || !K || !x0 || !loss_out || !gradK_out) return fail(DLC_ERR_ARG, "dlc_lds_lqr_grad_f32: NULL buffer");
  if (T > 0 && (!workspace || workspace_bytes < dlc_lds_lqr_grad_workspace_bytes(N, T, d)))
    return fail(DLC_ERR_WORKSPACE, "dlc_lds_lqr_grad_f32: workspace smaller than dlc_lds_lqr_grad_workspace_bytes()");
  LdsGradArgs a{};
  a.N = N; a.env_offset = env_offset; a.T = T; a.shared = shared_dynamics; a.t0 = t0;
  a.A = A; a.B = B; a.K = K; a.x0 = x0; a.W = W; a.seed = seed; a.w_std = w_std;
  a.loss_out = loss_out; a.gradK_out = gradK_out; a.gradx0_out = gradx0_out; a.sums_out = sums_out;
  a.ws = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
#define G(D_, M_) if (d == D_ && m == M_) return launch_grad<D_, M_>(a, st);
  G(10, 3) G(4, 1) G(4, 2) G(6, 2) G(8, 2) G(2, 1) G(3, 1) G(16, 4) G(16, 8) G(12, 4)
#undef G
  return fail(DLC_ERR_SHAPE, "dlc_lds_lqr_grad_f32: (d, m) not instantiated: (2,1) (3,1) (4,1) (4,2) (6,2) (8,2) (10,3) "
                             "(12,4) (16,4) (16,8)");
}
