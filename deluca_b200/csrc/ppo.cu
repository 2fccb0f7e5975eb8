// Experience post-processing of the actor-critic rollout:
//   gae_advantages   deluca/training/ppo.py:44-84     reverse-time first-order recurrence per agent
//   ac_rollout tail  deluca/training/ppo.py:131-134   rewards = -losses, returns = advantages + t[4]
// One pass over five [N,T] series: 12 B read + 8 B written per env-step, nothing else -- HBM-bound.
//
// Batch-major [N,T] (the layout vmap(ac_rollout) produces): a CTA owns 256 agents and walks the time
// axis backwards in chunks of 32 steps.  Load phase: one warp per agent row, lanes along t (128 B
// segments), delta_t and the carry coefficient are formed in registers and parked in shared memory
// (row stride 33: conflict-free for both the row-wise and the column-wise phase).  Scan phase: one
// thread per agent runs gae = delta + coef * gae down its 32 parked steps.  Store phase: row-wise
// again, returns formed on the way out.  Time-major [T,N] needs no staging (thread per agent).
//
// Arithmetic is float32 with every operation rounded separately (no FMA contraction), in the
// reference's order, so oracle/ppo.py (NumPy float32) agrees bit for bit.
#include "common.cuh"

namespace dlc {

struct GaeArgs {
  int64_t N;
  int T;
  const float *losses, *values, *log_probs, *masks;
  float discount, dl;  // dl = discount * gae_param (float32 product, ppo.py:68)
  int returns_use_values;
  float *adv, *ret;
};

constexpr int GAE_ROWS = 256, GAE_TC = 32, GAE_LD = GAE_TC + 1;

__global__ void __launch_bounds__(GAE_ROWS) gae_batch_major_kernel(const GaeArgs a) {
  extern __shared__ float sm[];
  float* s_delta = sm;                      // [GAE_ROWS][GAE_LD]; overwritten by the advantages
  float* s_coef = sm + GAE_ROWS * GAE_LD;   // [GAE_ROWS][GAE_LD]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * GAE_ROWS;
  const int T = a.T;
  float gae = 0.f;  // gae_{T} = 0: the last step is rewards[-1] - values[-1]  (ppo.py:64-65)
  for (int t1 = T; t1 > 0; t1 -= GAE_TC) {
    const int t0 = max(t1 - GAE_TC, 0);
    const int t = t0 + lane;
    // ---- load: rows warp, warp+8, ... ; lanes along t
#pragma unroll 4
    for (int r = warp; r < GAE_ROWS; r += 8) {
      const int64_t n = n0 + r;
      if (n < a.N && t < t1) {
        const int64_t i = n * T + t;
        const float rew = -a.losses[i];  // ppo.py:132 (-1.0 * losses is exact)
        const float v = a.values[i];
        float delta, coef;
        if (t == T - 1) {
          delta = __fsub_rn(rew, v);
          coef = 0.f;
        } else {
          const float mk = a.masks ? a.masks[i] : 1.0f;
          // rewards[t] + discount * values[t+1] * terminal_masks[t] - values[t]     ppo.py:67
          delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(a.discount, a.values[i + 1]), mk)), v);
          coef = __fmul_rn(a.dl, mk);  // discount * gae_param * terminal_masks[t]   ppo.py:68
        }
        s_delta[r * GAE_LD + lane] = delta;
        s_coef[r * GAE_LD + lane] = coef;
      }
    }
    __syncthreads();
    // ---- scan: one thread per agent, newest to oldest
    {
      float* dl = s_delta + threadIdx.x * GAE_LD;
      const float* cf = s_coef + threadIdx.x * GAE_LD;
      for (int j = t1 - t0 - 1; j >= 0; --j) {
        gae = __fadd_rn(dl[j], __fmul_rn(cf[j], gae));
        dl[j] = gae;
      }
    }
    __syncthreads();
    // ---- store
#pragma unroll 4
    for (int r = warp; r < GAE_ROWS; r += 8) {
      const int64_t n = n0 + r;
      if (n < a.N && t < t1) {
        const int64_t i = n * T + t;
        const float ad = s_delta[r * GAE_LD + lane];
        a.adv[i] = ad;
        if (a.ret) a.ret[i] = __fadd_rn(ad, a.returns_use_values ? a.values[i] : a.log_probs[i]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) gae_time_major_kernel(const GaeArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int T = a.T;
  const int64_t N = a.N;
  float gae = 0.f, vnext = 0.f;
#pragma unroll 4
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * N + n;
    const float rew = -a.losses[i];
    const float v = a.values[i];
    if (t == T - 1) {
      gae = __fsub_rn(rew, v);
    } else {
      const float mk = a.masks ? a.masks[i] : 1.0f;
      const float delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(a.discount, vnext), mk)), v);
      gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(a.dl, mk), gae));
    }
    vnext = v;
    a.adv[i] = gae;
    if (a.ret) a.ret[i] = __fadd_rn(gae, a.returns_use_values ? v : a.log_probs[i]);
  }
}

}  // namespace dlc

using namespace dlc;

extern "C" int32_t dlc_gae_f32(int64_t N, int32_t T, int32_t time_major, const float* losses, const float* values,
                               const float* log_probs, const float* done_masks, float discount, float gae_param,
                               int32_t returns_use_values, float* advantages_out, float* returns_out,
                               void* stream) {
  if (N < 0 || T < 1) return fail(DLC_ERR_ARG, "dlc_gae_f32: bad N / T");
  if (N == 0) return DLC_OK;
  if (!losses || !values || !advantages_out) return fail(DLC_ERR_ARG, "dlc_gae_f32: NULL buffer");
  if (returns_out && !returns_use_values && !log_probs)
    return fail(DLC_ERR_ARG, "dlc_gae_f32: returns_out needs log_probs (or returns_use_values = 1)");
  GaeArgs a{};
  a.N = N; a.T = T; a.losses = losses; a.values = values; a.log_probs = log_probs; a.masks = done_masks;
  a.discount = discount; a.dl = discount * gae_param; a.returns_use_values = returns_use_values;
  a.adv = advantages_out; a.ret = returns_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (time_major) {
    gae_time_major_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
  } else {
    static_assert(2 * GAE_ROWS * GAE_LD * sizeof(float) > 48 * 1024, "opt-in below is needed");
    const size_t smem = 2 * GAE_ROWS * GAE_LD * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(gae_batch_major_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gae_batch_major_kernel smem opt-in");
    gae_batch_major_kernel<<<(unsigned)((N + GAE_ROWS - 1) / GAE_ROWS), GAE_ROWS, smem, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "gae kernel launch");
}
