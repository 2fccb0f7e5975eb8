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

constexpr int GAE_WARPS = 4, GAE_TC = 32, GAE_LD = GAE_TC + 1;   // flat tile: 32 agents x 32 steps
constexpr int GAE_LR = 16, GAE_LC = 64, GAE_LLD = GAE_LC + 4;      // long-series tile: 16 agents x 64 steps, 16-byte aligned rows

// e / den for 0 <= e < 2^20 (inv = 1 / den)
__device__ __forceinline__ int gae_div(int e, float inv) { return (int)(((float)e + 0.5f) * inv); }

// 4-byte asynchronous global -> shared copy (LDGSTS): no register staging, so one lane keeps ~100 of
// them in flight; alignment-free, which the unaligned rows of an odd T need.
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

constexpr int GAE_TILE = GAE_LR * GAE_LLD > 32 * GAE_LD ? GAE_LR * GAE_LLD : 32 * GAE_LD;
static_assert(GAE_TILE % 4 == 0, "tiles must keep 16-byte alignment");

template <bool MASKS>
__global__ void __launch_bounds__(GAE_WARPS * 32) gae_batch_major_kernel(const GaeArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NARR = MASKS ? 4 : 3;
  constexpr int TILE = GAE_TILE;   // floats of one series tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = a.T;
  // T <= 32: a warp owns 32 agents; their rows are one contiguous run of 32*T floats -> one flat, fully coalesced
  // chunk.  Longer series: a warp owns 16 agents and walks their rows backwards in 64-step chunks (256 contiguous
  // bytes per row and series), double-buffered: chunk k-1 is in flight while chunk k is scanned and stored.
  const bool flat = T <= GAE_TC;
  const int R = flat ? 32 : GAE_LR, C = flat ? GAE_TC : GAE_LC, LD = flat ? GAE_LD : GAE_LLD;
  // rows that start on 16-byte boundaries move as float4 (a quarter of the copy / store instructions: the kernel
  // is issue-bound, ncu profiles/r01_gae_batch_major_T1000_v3_full.txt)
  const bool vec = !flat && (T & 3) == 0;
  const int nbuf = flat ? 1 : 2;
  float* const base = sm + warp * (nbuf * NARR * TILE);
  const float* __restrict__ losses = a.losses;
  const float* __restrict__ values = a.values;
  const float* __restrict__ masks = a.masks;
  const float* __restrict__ tail = a.ret ? (a.returns_use_values ? a.values : a.log_probs) : nullptr;
  float* __restrict__ adv = a.adv;
  float* __restrict__ ret = a.ret;
  const int64_t nw = ((int64_t)blockIdx.x * GAE_WARPS + warp) * R;  // first agent of this warp
  if (nw >= a.N) return;
  const int rows = (int)min((int64_t)R, a.N - nw);
  const float invT = 1.0f / (float)T;
  // flat chunk of a full warp with odd T: the 32*T floats are parked as they lie (float4 copies) -- row stride T is odd,
  // so the per-agent scan is still bank-conflict-free
  const bool flatv = flat && (T & 1) && rows == 32;
  const int LDs = flatv ? T : LD;   // row stride of the parked tile

  auto issue = [&](int t0, float* buf) {   // asynchronous copies of chunk [t0, t0 + C) into buf: L | V | tail | masks
    float *sL = buf, *sV = buf + TILE, *sT = buf + 2 * TILE, *sM = buf + 3 * TILE;
    if (flatv) {
      const int64_t gb = nw * T;
      for (int i = 4 * lane; i < 32 * T; i += 128) {
        cp_async16(sL + i, losses + gb + i);
        cp_async16(sV + i, values + gb + i);
        if (tail) cp_async16(sT + i, tail + gb + i);
        if (MASKS) cp_async16(sM + i, masks + gb + i);
      }
    } else if (flat) {
      const int64_t gb = nw * T;
      const int tot = rows * T;
      for (int i = lane; i < tot; i += 32) {
        const int r = gae_div(i, invT), o = r * LD + (i - r * T);
        cp_async4(sL + o, losses + gb + i);
        cp_async4(sV + o, values + gb + i);
        if (tail) cp_async4(sT + o, tail + gb + i);
        if (MASKS) cp_async4(sM + o, masks + gb + i);
      }
    } else if (vec) {
      const int tc4 = min(C, T - t0) >> 2;                 // float4 per row (C / 4 = 16: two rows per warp instruction)
      const int rr = lane >> 4, j4 = lane & 15;
      for (int r = rr; r < rows; r += 2) {
        if (j4 < tc4) {
          const int64_t gi = (nw + r) * T + t0 + 4 * j4;
          const int o = r * LD + 4 * j4;
          cp_async16(sL + o, losses + gi);
          cp_async16(sV + o, values + gi);
          if (tail) cp_async16(sT + o, tail + gi);
          if (MASKS) cp_async16(sM + o, masks + gi);
        }
      }
    } else {
      const int tc = min(C, T - t0);
      for (int r = 0; r < rows; ++r) {
        const int64_t gb = (nw + r) * T + t0;
        for (int j = lane; j < tc; j += 32) {
          const int o = r * LD + j;
          cp_async4(sL + o, losses + gb + j);
          cp_async4(sV + o, values + gb + j);
          if (tail) cp_async4(sT + o, tail + gb + j);
          if (MASKS) cp_async4(sM + o, masks + gb + j);
        }
      }
    }
    cp_async_commit();
  };

  // gae_T = 0 and values[T] := 0 make the general step equal rewards[-1] - values[-1] at t = T-1  (ppo.py:64-65)
  float gae = 0.f, vnext = 0.f;
  int cur = 0;
  const int t_last = ((T - 1) / C) * C;
  issue(t_last, base);
  for (int t0 = t_last; t0 >= 0; t0 -= C) {
    const int tc = min(C, T - t0);
    float* buf = base + cur * (NARR * TILE);
    if (t0 > 0) {
      issue(t0 - C, base + (cur ^ 1) * (NARR * TILE));
      cp_async_wait<1>();   // everything but the chunk just issued has landed
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    float *sL = buf, *sT = buf + 2 * TILE;
    // ---- scan: one lane per agent, newest to oldest
    if (lane < rows) {
      float* l = sL + lane * LDs;
      const float* v = buf + TILE + lane * LDs;
      const float* mkp = buf + 3 * TILE + lane * LDs;
#pragma unroll 4
      for (int j = tc - 1; j >= 0; --j) {
        const float rew = -l[j];  // ppo.py:132 (-1.0 * losses is exact)
        const float vj = v[j];
        // (multiplying by the constant mask 1.0 is exact, so the mask-free instance omits it)
        // rewards[t] + discount * values[t+1] * terminal_masks[t] - values[t]     ppo.py:67
        float dv = __fmul_rn(a.discount, vnext), cg = a.dl;
        if (MASKS) { const float mk = mkp[j]; dv = __fmul_rn(dv, mk); cg = __fmul_rn(cg, mk); }
        const float delta = __fsub_rn(__fadd_rn(rew, dv), vj);
        // delta + discount * gae_param * terminal_masks[t] * gae                   ppo.py:68
        gae = __fadd_rn(delta, __fmul_rn(cg, gae));
        vnext = vj;
        l[j] = gae;
      }
    }
    __syncwarp();
    // ---- store (returns formed on the way out)
    if (flatv) {
      const int64_t gb = nw * T;
#pragma unroll 2
      for (int i = 4 * lane; i < 32 * T; i += 128) {
        const float4 ad = *reinterpret_cast<const float4*>(sL + i);
        *reinterpret_cast<float4*>(adv + gb + i) = ad;
        if (ret) {
          const float4 tl = *reinterpret_cast<const float4*>(sT + i);
          *reinterpret_cast<float4*>(ret + gb + i) = make_float4(__fadd_rn(ad.x, tl.x), __fadd_rn(ad.y, tl.y),
                                                                __fadd_rn(ad.z, tl.z), __fadd_rn(ad.w, tl.w));
        }
      }
    } else if (flat) {
      const int64_t gb = nw * T;
      const int tot = rows * T;
#pragma unroll 4
      for (int i = lane; i < tot; i += 32) {
        const int r = gae_div(i, invT), o = r * LD + (i - r * T);
        const float ad = sL[o];
        adv[gb + i] = ad;
        if (ret) ret[gb + i] = __fadd_rn(ad, sT[o]);
      }
    } else if (vec) {
      const int tc4 = tc >> 2;
      const int rr = lane >> 4, j4 = lane & 15;
#pragma unroll 4
      for (int r = rr; r < rows; r += 2) {
        if (j4 < tc4) {
          const int64_t gi = (nw + r) * T + t0 + 4 * j4;
          const float4 ad = *reinterpret_cast<const float4*>(sL + r * LD + 4 * j4);
          *reinterpret_cast<float4*>(adv + gi) = ad;
          if (ret) {
            const float4 tl = *reinterpret_cast<const float4*>(sT + r * LD + 4 * j4);
            *reinterpret_cast<float4*>(ret + gi) = make_float4(__fadd_rn(ad.x, tl.x), __fadd_rn(ad.y, tl.y),
                                                               __fadd_rn(ad.z, tl.z), __fadd_rn(ad.w, tl.w));
          }
        }
      }
    } else {
      for (int r = 0; r < rows; ++r) {
        const int64_t gb = (nw + r) * T + t0;
#pragma unroll 4
        for (int j = lane; j < tc; j += 32) {
          const float ad = sL[r * LD + j];
          adv[gb + j] = ad;
          if (ret) ret[gb + j] = __fadd_rn(ad, sT[r * LD + j]);
        }
      }
    }
    __syncwarp();   // the tile is rewritten by the copies issued in the next iteration
    cur ^= nbuf - 1;
  }
}

// Time-major [T,N]: V adjacent agents per thread (float2 when N is even: twice the bytes in flight per thread), loads
// hoisted eight steps deep.
template <int V>
__global__ void __launch_bounds__(128) gae_time_major_kernel(const GaeArgs a) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (n >= a.N) return;
  const int T = a.T;
  const int64_t N = a.N;
  const float* __restrict__ losses = a.losses;
  const float* __restrict__ values = a.values;
  const float* __restrict__ masks = a.masks;
  const float* __restrict__ tail = a.ret ? (a.returns_use_values ? a.values : a.log_probs) : nullptr;
  float* __restrict__ adv = a.adv;
  float* __restrict__ ret = a.ret;
  float gae[V], vnext[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { gae[k] = 0.f; vnext[k] = 0.f; }
  // gae_T = 0, values[T] := 0: the general step is rewards[-1] - values[-1] at t = T-1
#pragma unroll 8
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * N + n;
    float l[V], v[V], mk[V], tl[V], o_adv[V], o_ret[V];
    if (V == 2) {
      const float2 l2 = *reinterpret_cast<const float2*>(losses + i), v2 = *reinterpret_cast<const float2*>(values + i);
      l[0] = l2.x; l[V - 1] = l2.y; v[0] = v2.x; v[V - 1] = v2.y;
      if (masks) { const float2 m2 = *reinterpret_cast<const float2*>(masks + i); mk[0] = m2.x; mk[V - 1] = m2.y; }
      if (tail) { const float2 t2 = *reinterpret_cast<const float2*>(tail + i); tl[0] = t2.x; tl[V - 1] = t2.y; }
    } else {
      l[0] = losses[i]; v[0] = values[i];
      if (masks) mk[0] = masks[i];
      if (tail) tl[0] = tail[i];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float rew = -l[k];
      float dv = __fmul_rn(a.discount, vnext[k]), cg = a.dl;
      if (masks) { dv = __fmul_rn(dv, mk[k]); cg = __fmul_rn(cg, mk[k]); }
      const float delta = __fsub_rn(__fadd_rn(rew, dv), v[k]);
      gae[k] = __fadd_rn(delta, __fmul_rn(cg, gae[k]));
      vnext[k] = v[k];
      o_adv[k] = gae[k];
      o_ret[k] = tail ? __fadd_rn(gae[k], tl[k]) : 0.f;
    }
    if (V == 2) {
      *reinterpret_cast<float2*>(adv + i) = make_float2(o_adv[0], o_adv[V - 1]);
      if (ret) *reinterpret_cast<float2*>(ret + i) = make_float2(o_ret[0], o_ret[V - 1]);
    } else {
      adv[i] = o_adv[0];
      if (ret) ret[i] = o_ret[0];
    }
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
    if ((N & 1) == 0) gae_time_major_kernel<2><<<(unsigned)((N / 2 + 127) / 128), 128, 0, st>>>(a);
    else gae_time_major_kernel<1><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
  } else {
    const size_t smem = (size_t)GAE_WARPS * (T <= GAE_TC ? 1 : 2) * (done_masks ? 4 : 3) * GAE_TILE * sizeof(float);
    const int per_cta = GAE_WARPS * (T <= GAE_TC ? 32 : GAE_LR);   // agents per CTA
    const unsigned grid = (unsigned)((N + per_cta - 1) / per_cta);
    auto kern = done_masks ? gae_batch_major_kernel<true> : gae_batch_major_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gae_batch_major_kernel smem opt-in");
    kern<<<grid, GAE_WARPS * 32, smem, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "gae kernel launch");
}
