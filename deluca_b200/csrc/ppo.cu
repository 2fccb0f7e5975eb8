// Experience post-processing of the actor-critic rollout:
//   gae_advantages   deluca/training/ppo.py:44-84     reverse-time first-order recurrence per agent
//   ac_rollout tail  deluca/training/ppo.py:131-134   rewards = -losses, returns = advantages + t[4]
// One pass over five [N,T] series: 12 B read + 8 B written per env-step, nothing else -- HBM-bound.
//
// Batch-major [N,T] (the layout vmap(ac_rollout) produces): a WARP owns 32 agents and walks the time
// axis backwards in chunks of 32 steps; warps never meet (no CTA barrier), so their load, scan and
// store phases overlap freely.  Load: for T <= 32 the warp's 32 rows are one contiguous run of 32*T
// floats and are copied flat (perfectly coalesced, whatever T is -- the PPO lung horizon is 29);
// for longer series row-wise 128-byte segments on 32-step boundaries.  Both are 4-byte cp.async
// copies, all of a chunk in flight at once, and park the series in the
// warp's shared-memory tile (row stride 33: conflict-free row-wise and column-wise).  Scan: one lane
// per agent runs the recurrence down its parked steps, overwriting the losses with the advantages.
// Store: the same mapping as the load, returns formed on the way out.
// Time-major [T,N] needs no staging (thread per agent, loads hoisted eight steps deep).
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

constexpr int GAE_WARPS = 4, GAE_TC = 32, GAE_LD = GAE_TC + 1;

// e / den for 0 <= e < 2^20 (inv = 1 / den)
__device__ __forceinline__ int gae_div(int e, float inv) { return (int)(((float)e + 0.5f) * inv); }

// 4-byte asynchronous global -> shared copy (LDGSTS): no register staging, so one lane keeps ~100 of
// them in flight; alignment-free, which the unaligned rows of an odd T need.
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <bool MASKS>
__global__ void __launch_bounds__(GAE_WARPS * 32) gae_batch_major_kernel(const GaeArgs a) {
  extern __shared__ float sm[];
  constexpr int NARR = MASKS ? 4 : 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sL = sm + warp * (NARR * 32 * GAE_LD);  // [32][GAE_LD] losses, overwritten by the advantages
  float* sV = sL + 32 * GAE_LD;                  // [32][GAE_LD] values
  float* sT = sV + 32 * GAE_LD;                  // [32][GAE_LD] the series added for the returns
  float* sM = sT + 32 * GAE_LD;                  // [32][GAE_LD] masks (MASKS only)
  const float* __restrict__ losses = a.losses;
  const float* __restrict__ values = a.values;
  const float* __restrict__ masks = a.masks;
  const float* __restrict__ tail = a.ret ? (a.returns_use_values ? a.values : a.log_probs) : nullptr;
  float* __restrict__ adv = a.adv;
  float* __restrict__ ret = a.ret;
  const int64_t nw = ((int64_t)blockIdx.x * GAE_WARPS + warp) * 32;  // first agent of this warp
  if (nw >= a.N) return;
  const int rows = (int)min((int64_t)32, a.N - nw);
  const int T = a.T;
  // T <= 32: the warp's rows are one contiguous run of rows*T floats -> flat, fully coalesced copies;
  // otherwise row-wise 128-byte segments on 32-step boundaries of each row.
  const bool flat = T <= GAE_TC;
  const float invT = 1.0f / (float)T;
  // gae_T = 0 and values[T] := 0 make the general step equal rewards[-1] - values[-1] at t = T-1  (ppo.py:64-65)
  float gae = 0.f, vnext = 0.f;
  for (int t0 = ((T - 1) / GAE_TC) * GAE_TC; t0 >= 0; t0 -= GAE_TC) {
    const int tc = min(GAE_TC, T - t0);
    // ---- load (asynchronous; every copy of the chunk is in flight before the first is awaited)
    if (flat) {
      const int64_t base = nw * T;
      const int tot = rows * T;
      for (int i = lane; i < tot; i += 32) {
        const int r = gae_div(i, invT), o = r * GAE_LD + (i - r * T);
        cp_async4(sL + o, losses + base + i);
        cp_async4(sV + o, values + base + i);
        if (tail) cp_async4(sT + o, tail + base + i);
        if (MASKS) cp_async4(sM + o, masks + base + i);
      }
    } else if (lane < tc) {
      for (int r = 0; r < rows; ++r) {
        const int64_t i = (nw + r) * T + t0 + lane;
        const int o = r * GAE_LD + lane;
        cp_async4(sL + o, losses + i);
        cp_async4(sV + o, values + i);
        if (tail) cp_async4(sT + o, tail + i);
        if (MASKS) cp_async4(sM + o, masks + i);
      }
    }
    cp_async_wait_all();
    __syncwarp();
    // ---- scan: one lane per agent, newest to oldest
    if (lane < rows) {
      float* l = sL + lane * GAE_LD;
      const float* v = sV + lane * GAE_LD;
      const float* mkp = sM + lane * GAE_LD;
#pragma unroll 4
      for (int j = tc - 1; j >= 0; --j) {
        const float rew = -l[j];  // ppo.py:132 (-1.0 * losses is exact)
        const float vj = v[j];
        const float mk = MASKS ? mkp[j] : 1.0f;
        // rewards[t] + discount * values[t+1] * terminal_masks[t] - values[t]     ppo.py:67
        const float delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(a.discount, vnext), mk)), vj);
        // delta + discount * gae_param * terminal_masks[t] * gae                   ppo.py:68
        gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(a.dl, mk), gae));
        vnext = vj;
        l[j] = gae;
      }
    }
    __syncwarp();
    // ---- store (returns formed on the way out)
    if (flat) {
      const int64_t base = nw * T;
      const int tot = rows * T;
#pragma unroll 4
      for (int i = lane; i < tot; i += 32) {
        const int r = gae_div(i, invT), o = r * GAE_LD + (i - r * T);
        const float ad = sL[o];
        adv[base + i] = ad;
        if (ret) ret[base + i] = __fadd_rn(ad, sT[o]);
      }
    } else if (lane < tc) {
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        const int64_t i = (nw + r) * T + t0 + lane;
        const float ad = sL[r * GAE_LD + lane];
        adv[i] = ad;
        if (ret) ret[i] = __fadd_rn(ad, sT[r * GAE_LD + lane]);
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) gae_time_major_kernel(const GaeArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int T = a.T;
  const int64_t N = a.N;
  const float* __restrict__ losses = a.losses;
  const float* __restrict__ values = a.values;
  const float* __restrict__ tail = a.returns_use_values ? nullptr : a.log_probs;
  float* __restrict__ adv = a.adv;
  float* __restrict__ ret = a.ret;
  float gae = 0.f, vnext = 0.f;
#pragma unroll 8
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * N + n;
    const float rew = -losses[i];
    const float v = values[i];
    // gae_T = 0, values[T] := 0: the general step is rewards[-1] - values[-1] at t = T-1
    const float mk = a.masks ? a.masks[i] : 1.0f;
    const float delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(a.discount, vnext), mk)), v);
    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(a.dl, mk), gae));
    vnext = v;
    adv[i] = gae;
    if (ret) ret[i] = __fadd_rn(gae, tail ? tail[i] : v);
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
    const size_t smem = (size_t)GAE_WARPS * (done_masks ? 4 : 3) * 32 * GAE_LD * sizeof(float);
    const unsigned grid = (unsigned)((N + GAE_WARPS * 32 - 1) / (GAE_WARPS * 32));
    auto kern = done_masks ? gae_batch_major_kernel<true> : gae_batch_major_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gae_batch_major_kernel smem opt-in");
    kern<<<grid, GAE_WARPS * 32, smem, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "gae kernel launch");
}
