// Fused ventilator episode: PID + Expiratory + BalloonLung / DelayLung.
// One thread per breath; all state in registers (DelayLung's two valve histories in a local
// ring); per step five/six coalesced 4-byte stores per thread ([T,N] time-major series), which is
// what bounds the kernel (HBM writes).
#include <mutex>

#include "common.cuh"

namespace dlc {

// single-instruction special functions (MUFU.EX2 / LG2 / RCP); the wrappers pin the PTX so no slow path is emitted
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct LungArgs {
  DlcLungParams p;
  DlcLungBuffers b;
};

// t mod period for t >= 0, bit-identical to fmodf (IEEE remainders are exactly representable):
// q = floor(t / period) with a correctly rounded divide can only be one too large (when t / period
// rounds up to an integer); the FMA residual t - q*period is then exact (a small negative multiple of
// the operands' common granularity) and adding the period back is exact as well.
__device__ __forceinline__ float fmod_pos(float t, float period) {
  if (!(t >= 0.f) || !(t < 1e30f)) return fmodf(t, period);
  const float q = floorf(__fdiv_rn(t, period));
  float r = fmaf(-q, period, t);
  if (r < 0.f) r += period;
  return r;
}

// BreathWaveform.at: jnp.interp(t, xp, fp, period) on the host-built table (lung/core.py:54-60).
// `x` is t mod period (fmodf: exact, and t >= 0 on this path, so it equals jnp.mod).  NK > 0 is the
// compile-time knot count (7 for the default 5-knot waveform): the search is unrolled and the
// per-env ordinates stay in registers; NK == 0 reads the count from the parameter block.
template <int NK>
__device__ __forceinline__ float interp_knots(const DlcLungParams& p, const float* kf, float x) {
  if (NK > 0) {
    float x0 = p.kx[0], x1 = p.kx[1], f0 = kf[0], f1 = kf[1];
#pragma unroll
    for (int k = 1; k < NK - 1; ++k) {  // searchsorted(side='right'), clipped to [1, NK-1]
      const bool adv = p.kx[k] <= x;
      x0 = adv ? p.kx[k] : x0;
      x1 = adv ? p.kx[k + 1] : x1;
      f0 = adv ? kf[k] : f0;
      f1 = adv ? kf[k + 1] : f1;
    }
    const float dx = x1 - x0;
    return dx == 0.f ? f0 : f0 + __fdiv_rn(x - x0, dx) * (f1 - f0);
  }
  int i = 1;
#pragma unroll 1
  for (int k = 1; k < p.nk - 1; ++k) i += (p.kx[k] <= x) ? 1 : 0;
  const float x0 = p.kx[i - 1], dx = p.kx[i] - x0;
  const float f0 = kf[i - 1], df = kf[i] - f0;
  return dx == 0.f ? f0 : f0 + __fdiv_rn(x - x0, dx) * df;
}

// t mod period for 0 <= t < 2^20 * period, bit-identical to fmodf without a divide: q = floor(t * (1/period)) is
// within one of the true quotient; the FMA residual t - q*period is then the correctly rounded value of a number
// that is exactly representable whenever it lies in [-period, period), so its sign / size tells which way q was
// off, and the residual recomputed with the corrected quotient is exact.
__device__ __forceinline__ float fmod_fast(float t, float period, float inv_period, float t_max) {
  // one unsigned compare: negative t (sign bit), NaN and t >= t_max all take the library path
  if (__builtin_expect(__float_as_uint(t) >= __float_as_uint(t_max), 0)) return fmodf(t, period);
  const float q = floorf(t * inv_period);
  float r = fmaf(-q, period, t);
  if (__builtin_expect(r < 0.f || r >= period, 0)) r = fmaf(-(r < 0.f ? q - 1.0f : q + 1.0f), period, t);
  return r;
}

// One segment of the waveform table: the interpolant on [x0, x1) is f0 + (x - x0) * sl.  Episodes dwell ~30 steps
// in a segment, so the knot search (searchsorted(side='right') clipped to [1, NK-1], as jnp.interp) only runs when
// x leaves the cached segment; the values are the same either way.
struct Seg {
  float x0, x1, f0, sl;
};

template <int NK>
__device__ __forceinline__ void seg_search(const DlcLungParams& p, const float* kf, float x, Seg& s) {
  float x0, x1, f0, f1;
  if (NK > 0) {
    x0 = p.kx[0]; x1 = p.kx[1]; f0 = kf[0]; f1 = kf[1];
#pragma unroll
    for (int k = 1; k < NK - 1; ++k) {
      const bool adv = p.kx[k] <= x;
      x0 = adv ? p.kx[k] : x0;
      x1 = adv ? p.kx[k + 1] : x1;
      f0 = adv ? kf[k] : f0;
      f1 = adv ? kf[k + 1] : f1;
    }
  } else {
    int i = 1;
#pragma unroll 1
    for (int k = 1; k < p.nk - 1; ++k) i += (p.kx[k] <= x) ? 1 : 0;
    x0 = p.kx[i - 1]; x1 = p.kx[i]; f0 = kf[i - 1]; f1 = kf[i];
  }
  const float dx = x1 - x0;
  s.x0 = x0; s.x1 = x1; s.f0 = f0;
  s.sl = dx == 0.f ? 0.f : __fdiv_rn(f1 - f0, dx);
}

template <int NK>
__device__ __forceinline__ float interp_seg(const DlcLungParams& p, const float* kf, float x, Seg& s) {
  if (__builtin_expect(!(s.x0 <= x && x < s.x1), 0)) seg_search<NK>(p, kf, x, s);
  return fmaf(x - s.x0, s.sl, s.f0);
}

// tanh(z) + 1 = 2 e / (1 + e), e = exp(2z): relative accuracy ~3e-7 everywhere (tanhf(z) + 1 loses it for z << 0);
// z is capped at 10, where the float32 value is already 2.
__device__ __forceinline__ float tanh_plus_one(float z) {
  const float e = exp2f_approx(fminf(z, 10.f) * 2.8853900817779268f);
  return __fdividef(e + e, 1.0f + e);
}

// x^(1/3) for x > 0 (NaN for x < 0, like the reference's pow): ex2(lg2(x)/3) refined by one Newton step (<= 1 ulp).
__device__ __forceinline__ float cbrt_pos(float x) {
  float r = exp2f_approx(lg2f_approx(x) * 0.3333333333333333f);
  const float r2 = r * r;
  const float res = fmaf(r2, r, -x);
  const float rc = rcpf_approx(3.0f * r2);
  if (x > 0.f && x < 3.0e38f) r = fmaf(-res, rc, r);
  return r;
}

// MODE >= 0 fixes p.mode at compile time (the closed-loop episode is the hot instance), MODE < 0 reads it; ALLOUT
// promises that all six series pointers are non-NULL.
template <int SIM, int NK, int MODE, bool ALLOUT>
__global__ void __launch_bounds__(128) lung_pid_kernel(const LungArgs a) {
  extern __shared__ float ring[];                                  // DelayLung valve histories [2][delay][blockDim]
  const DlcLungParams& p = a.p;
  const DlcLungBuffers& b = a.b;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  float kf[DLC_MAX_KNOTS];
  {
    const float* g = b.kf + (p.kf_per_env ? n * p.nk : 0);
#pragma unroll
    for (int k = 0; k < DLC_MAX_KNOTS; ++k) kf[k] = k < p.nk ? g[k] : 0.f;
  }
  const float k0 = b.K[n * 3 + 0], k1 = b.K[n * 3 + 1], k2 = b.K[n * 3 + 2];
  float time = b.time_io[n], volume = b.volume_io[n], pressure = b.pressure_io[n];
  float steps = b.steps_io[n], target = b.target_io[n];
  float pipe = SIM == DLC_SIM_DELAY ? b.pipe_pressure_io[n] : 0.f;
  float obs_p = b.obs_pressure_io[n], obs_t = b.obs_time_io[n];
  float cp = b.pid_p_io[n], ci = b.pid_i_io[n], cd = b.pid_d_io[n];
  // controller bookkeeping (dt = max(default_dt, t - proper_time(time)); time = t; steps += 1) only depends on the
  // last two observation times: carried as (previous, current) and formed once after the loop
  float ctime = b.pid_time_io[n], etime = b.exp_time_io[n], prev_ctime = ctime, prev_etime = etime;
  // ring layout [slot][thread] with a row stride of blockDim + 1 words: the per-step access (slot fixed, thread
  // varies) and the transposing copies below (thread fixed, slot varies) are both bank-conflict-free
  const int rs = blockDim.x + 1;
  float* hin = ring + threadIdx.x;
  float* hout = ring + (size_t)p.delay * rs + threadIdx.x;
  int pos = 0;  // ring slot of logical index 0 (the newest entry)
  // the caller's histories are [N, delay]: a full CTA owns one contiguous chunk and copies it with coalesced
  // accesses (the per-thread rows are 4 * delay bytes apart); the last, partial CTA copies row by row
  const bool full_cta = (int64_t)(blockIdx.x + 1) * blockDim.x <= p.N;
  if (SIM == DLC_SIM_DELAY) {
    if (full_cta) {
      const int64_t base = (int64_t)blockIdx.x * blockDim.x * p.delay;
      const int count = blockDim.x * p.delay;
      for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const int e = i / p.delay, k = i - e * p.delay;
        ring[k * rs + e] = b.in_history_io[base + i];
        ring[(p.delay + k) * rs + e] = b.out_history_io[base + i];
      }
      __syncthreads();
    } else {
      for (int k = 0; k < p.delay; ++k) {
        hin[k * rs] = b.in_history_io[n * p.delay + k];
        hout[k * rs] = b.out_history_io[n * p.delay + k];
      }
    }
  }
  bool bad = false, ran_env = false;
  const int mode = MODE >= 0 ? MODE : p.mode;
  const bool run_ctrl = mode != DLC_LUNG_MODE_ENV_ONLY, run_env = mode != DLC_LUNG_MODE_CTRL_ONLY;
  // running output pointers: one add per series and step instead of re-forming t * N + n
  const int64_t N = p.N;
  float* o_ts = b.timestamp_out ? b.timestamp_out + n : nullptr;
  float* o_uin = b.u_in_out ? b.u_in_out + n : nullptr;
  float* o_uout = b.u_out_out ? b.u_out_out + n : nullptr;
  float* o_pr = b.pressure_out ? b.pressure_out + n : nullptr;
  float* o_flow = b.flow_out ? b.flow_out + n : nullptr;
  float* o_tgt = b.target_out ? b.target_out + n : nullptr;
  const float* i_uin = b.u_in_in ? b.u_in_in + n : nullptr;
  const float* i_uout = b.u_out_in ? b.u_out_in + n : nullptr;
  const float period = p.period, xp_in = p.xp_in, dec = p.pid_decay, dt = p.dt, run_dt = p.run_dt;
  const float inv_period = __fdiv_rn(1.0f, p.period), t_max = 1048576.0f * p.period;
  const float pc_r0sq = __fdiv_rn(SIM == DLC_SIM_BALLOON ? p.PC : p.d_C, p.r0_sq), inv_dR = __fdiv_rn(1.0f, p.d_R);
  const float early_end = (float)p.delay;
  Seg seg_pid = {1.f, 0.f, 0.f, 0.f}, seg_out = {1.f, 0.f, 0.f, 0.f};   // empty segments: first use searches

  for (int t = 0; t < p.T; ++t) {
    const float pr_emit = obs_p;                                   // run_controller.py:118
    float u_in, u_out;
    if (run_ctrl) {
      // ---- PID                                                   controllers/_pid.py:77-102
      const float xo = fmod_fast(obs_t, period, inv_period, t_max);   // shared by PID (at) and Expiratory (is_in)
      const float err = interp_seg<NK>(p, kf, xo, seg_pid) - obs_p;
      const float ni = ci + dec * (err - ci);
      const float nd = cd + dec * ((err - cp) - cd);
      cp = err; ci = ni; cd = nd;
      u_in = (err * k0 + ni * k1) + nd * k2;                        // Dense(1, use_bias=False)
      u_in = fminf(fmaxf(u_in, 0.f), 100.f);                        // lax.clamp(0, u, 100)
      // ---- Expiratory                                            controllers/_expiratory.py:35-45
      u_out = (xo <= xp_in) ? 0.f : 1.f;
      prev_ctime = ctime; prev_etime = etime;
      ctime = obs_t; etime = obs_t;
    } else {
      u_in = *i_uin; i_uin += N;
      u_out = *i_uout; i_uout += N;
    }
    // ---- simulator
    if (!run_env) {
    } else if (SIM == DLC_SIM_BALLOON) {                            // envs/_balloon_lung.py:100-146
      float valve = tanh_plus_one(0.03f * (3.0f * u_in - 130.f));
      valve = fminf(fmaxf(valve, 0.f), 1.72f);
      float flow = fminf(fmaxf(valve * p.R, 0.f), 2.0f);
      if (pressure > p.peep_valve) flow -= (u_out > 0.f ? 1.0f : 0.0f) * 0.05f * pressure;
      volume += flow * dt;
      if (p.leak) volume += p.leak_coef * (p.min_volume - volume);
      const float r = cbrt_pos(volume * 0.238732414637843f);       // (3V / 4pi)^(1/3)
      const float ir = rcpf_approx(r);
      const float q = p.r0 * ir, q2 = q * q;
      pressure = fmaf(pc_r0sq * fmaf(-q2 * q2, q2, 1.0f), ir, p.P0);   // P0 + PC (1 - (r0/r)^6) / (r0^2 r)
      steps = time + 1.0f;                                          // sic (:135)
      time = time + dt;                                             // target = waveform.at(time): only the
      ran_env = true;                                               // last one survives, formed after the loop
      obs_p = pressure;
      obs_t = time;
    } else {                                                        // envs/_delay_lung.py:75-133
      const float uin_pos = u_in > 0.f ? u_in : 0.f;
      pos = pos == 0 ? p.delay - 1 : pos - 1;                       // roll(+1); [0] <- u
      hin[pos * rs] = uin_pos;
      hout[pos * rs] = u_out;
      obs_p = pressure;                                             // observation of the pre-dynamics state
      obs_t = time;
      const float flow = pressure * inv_dR;
      volume = fmaxf(volume + flow * dt, p.min_volume);
      const float r = cbrt_pos(volume * 0.238732414637843f);
      const float ir = rcpf_approx(r);
      const float q = p.r0 * ir, q2 = q * q;
      const float lung_p = pc_r0sq * fmaf(-q2 * q2, q2, 1.0f) * ir;
      const bool early = time < early_end;                          // seconds vs steps, sic (:88-96)
      const float impulse = early ? 0.f : p.control_gain * uin_pos;
      const float peep = early ? 0.f : u_out;
      pipe = p.inertia * pipe + impulse;
      pressure = fmaxf(0.f, pipe - lung_p);
      if (peep != 0.f) pipe *= 0.995f;
      steps = time + 1.0f;
      time = time + dt;
      ran_env = true;
    }
    // ---- emitted series                                          run_controller.py:128-141
    const float ts = dt * steps - run_dt;
    if (ALLOUT || o_ts) { *o_ts = ts; o_ts += N; }
    if (ALLOUT || o_uin) { *o_uin = u_in; o_uin += N; }
    if (ALLOUT || o_uout) { *o_uout = u_out; o_uout += N; }
    if (ALLOUT || o_pr) { *o_pr = pr_emit; o_pr += N; }
    if (ALLOUT || o_flow) { *o_flow = p.env_flow; o_flow += N; }
    if (ALLOUT || o_tgt) {                                          // waveform.at(timestamps), :207
      *o_tgt = interp_seg<NK>(p, kf, fmod_fast(ts, period, inv_period, t_max), seg_out);
      o_tgt += N;
    }
    bad |= !isfinite(pressure);
  }

  if (ran_env) target = interp_knots<NK>(p, kf, fmod_pos(time, p.period));   // state.target (:140)
  b.steps_io[n] = steps; b.time_io[n] = time; b.volume_io[n] = volume; b.pressure_io[n] = pressure;
  b.target_io[n] = target;
  b.obs_pressure_io[n] = obs_p; b.obs_time_io[n] = obs_t;
  b.pid_p_io[n] = cp; b.pid_i_io[n] = ci; b.pid_d_io[n] = cd;
  if (run_ctrl && p.T > 0) {
    b.pid_time_io[n] = ctime;
    b.pid_dt_io[n] = fmaxf(p.default_dt, ctime - (isinf(prev_ctime) ? 0.f : prev_ctime));
    b.pid_steps_io[n] += p.T;
    b.exp_time_io[n] = etime;
    b.exp_dt_io[n] = fmaxf(p.default_dt, etime - (isinf(prev_etime) ? 0.f : prev_etime));
    b.exp_steps_io[n] += p.T;
  }
  if (SIM == DLC_SIM_DELAY) {
    b.pipe_pressure_io[n] = pipe;
    if (full_cta) {                                                 // pos is the same in every thread (T steps each)
      __syncthreads();
      const int64_t base = (int64_t)blockIdx.x * blockDim.x * p.delay;
      const int count = blockDim.x * p.delay;
      for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const int e = i / p.delay, k = i - e * p.delay;
        int s = pos + k;
        s -= s >= p.delay ? p.delay : 0;
        b.in_history_io[base + i] = ring[s * rs + e];
        b.out_history_io[base + i] = ring[(p.delay + s) * rs + e];
      }
    } else {
      for (int k = 0; k < p.delay; ++k) {
        int s = pos + k;
        s -= s >= p.delay ? p.delay : 0;
        b.in_history_io[n * p.delay + k] = hin[s * rs];
        b.out_history_io[n * p.delay + k] = hout[s * rs];
      }
    }
  }
  if (b.status_out) b.status_out[n] = bad ? DLC_STATUS_NONFINITE : 0;
}

}  // namespace dlc

extern "C" int32_t dlc_lung_pid_rollout_f32(const DlcLungParams* p, const DlcLungBuffers* b, void* stream) {
  using namespace dlc;
  if (!p || !b) return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: NULL params/buffers");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: negative N or T");
  if (p->sim != DLC_SIM_BALLOON && p->sim != DLC_SIM_DELAY)
    return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: unknown simulator");
  if (p->nk < 2 || p->nk > DLC_MAX_KNOTS) return fail(DLC_ERR_SHAPE, "dlc_lung_pid_rollout_f32: 2 <= nk <= 16 knots");
  if (p->sim == DLC_SIM_DELAY && (p->delay < 1 || p->delay > DLC_MAX_DELAY))
    return fail(DLC_ERR_SHAPE, "dlc_lung_pid_rollout_f32: 1 <= delay <= 64");
  if (p->N == 0) return DLC_OK;
  const bool need = b->K && b->kf && b->steps_io && b->time_io && b->volume_io && b->pressure_io && b->target_io &&
                    b->obs_pressure_io && b->obs_time_io && b->pid_p_io && b->pid_i_io && b->pid_d_io &&
                    b->pid_time_io && b->pid_dt_io && b->pid_steps_io && b->exp_time_io && b->exp_dt_io &&
                    b->exp_steps_io;
  if (!need) return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: a required state buffer is NULL");
  if (p->mode < 0 || p->mode > DLC_LUNG_MODE_ENV_ONLY) return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: unknown mode");
  if (p->mode == DLC_LUNG_MODE_ENV_ONLY && !(b->u_in_in && b->u_out_in))
    return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: env-only mode needs u_in_in and u_out_in");
  if (p->sim == DLC_SIM_DELAY && !(b->pipe_pressure_io && b->in_history_io && b->out_history_io))
    return fail(DLC_ERR_ARG, "dlc_lung_pid_rollout_f32: DelayLung needs pipe_pressure/in_history/out_history");
  LungArgs a;
  a.p = *p;
  a.b = *b;
  const unsigned grid = (unsigned)((p->N + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  // hot instance: default 7-knot waveform table, closed loop, all six series requested
  const bool hot = p->nk == 7 && p->mode == DLC_LUNG_MODE_CLOSED_LOOP && b->timestamp_out && b->u_in_out &&
                   b->u_out_out && b->pressure_out && b->flow_out && b->target_out;
  if (p->sim == DLC_SIM_BALLOON) {
    if (hot) lung_pid_kernel<DLC_SIM_BALLOON, 7, DLC_LUNG_MODE_CLOSED_LOOP, true><<<grid, 128, 0, st>>>(a);
    else if (p->nk == 7) lung_pid_kernel<DLC_SIM_BALLOON, 7, -1, false><<<grid, 128, 0, st>>>(a);
    else lung_pid_kernel<DLC_SIM_BALLOON, 0, -1, false><<<grid, 128, 0, st>>>(a);
  } else {
    const size_t smem = 2 * (size_t)p->delay * 129 * sizeof(float);   // 64.5 KB at the maximum delay of 64
    if (smem > 48 * 1024) {
      static std::once_flag once;
      static cudaError_t attr_rc[3];
      std::call_once(once, [] {
        const int cap = 2 * DLC_MAX_DELAY * 129 * 4;
        attr_rc[0] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_rc[1] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 7, -1, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_rc[2] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 0, -1, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
      });
      for (int i = 0; i < 3; ++i)
        if (attr_rc[i] != cudaSuccess) return cuda_status(attr_rc[i], "lung_pid_kernel shared-memory attribute");
    }
    if (hot) lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, true><<<grid, 128, smem, st>>>(a);
    else if (p->nk == 7) lung_pid_kernel<DLC_SIM_DELAY, 7, -1, false><<<grid, 128, smem, st>>>(a);
    else lung_pid_kernel<DLC_SIM_DELAY, 0, -1, false><<<grid, 128, smem, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "lung_pid_kernel launch");
}

namespace dlc {
__global__ void __launch_bounds__(256) pid_kernel(int64_t N, int T, float decay, const float* KP, const float* KI,
                                                  const float* KD, const float* err, float* P_io, float* I_io,
                                                  float* D_io, float* u_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float kp = KP[n], ki = KI[n], kd = KD[n];
  float P = P_io[n], I = I_io[n], D = D_io[n];
  for (int t = 0; t < T; ++t) {
    const float e = err[(int64_t)t * N + n];
    I = I + decay * (e - I);
    D = D + decay * ((e - P) - D);
    P = e;
    u_out[(int64_t)t * N + n] = (kp * P + ki * I) + kd * D;
  }
  P_io[n] = P; I_io[n] = I; D_io[n] = D;
}
}  // namespace dlc

extern "C" int32_t dlc_pid_f32(int64_t N, int32_t T, float decay, const float* K_P, const float* K_I,
                               const float* K_D, const float* err, float* P_io, float* I_io, float* D_io,
                               float* u_out, void* stream) {
  using namespace dlc;
  if (N < 0 || T < 0) return fail(DLC_ERR_ARG, "dlc_pid_f32: negative N or T");
  if (N == 0 || T == 0) return DLC_OK;
  if (!K_P || !K_I || !K_D || !err || !P_io || !I_io || !D_io || !u_out)
    return fail(DLC_ERR_ARG, "dlc_pid_f32: NULL buffer");
  pid_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(N, T, decay, K_P, K_I, K_D, err, P_io,
                                                                            I_io, D_io, u_out);
  return cuda_status(cudaGetLastError(), "pid_kernel launch");
}

// ------------------------------------------------------------------------------------------
// value + gradient of train_controller.rollout w.r.t. the PID gains (BalloonLung)
// ------------------------------------------------------------------------------------------
namespace dlc {

struct LungGradArgs {
  DlcLungParams p;
  const float *K, *kf, *tt;
  int loss_kind;
  float noise, reset_pressure;
  float *loss_out, *grad_out;
  double* sums_out;
  float* ws;   // [T][6][N]: P (after noise), V, p, i, d, time at the start of each step
};

template <int NK>
__global__ void __launch_bounds__(128) lung_pid_grad_kernel(const LungGradArgs a) {
  const DlcLungParams& p = a.p;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = n < p.N;
  const int64_t nn = live ? n : p.N - 1;
  float kf[DLC_MAX_KNOTS];
  {
    const float* g = a.kf + (p.kf_per_env ? nn * p.nk : 0);
#pragma unroll
    for (int k = 0; k < DLC_MAX_KNOTS; ++k) kf[k] = k < p.nk ? g[k] : 0.f;
  }
  const float k0 = a.K[nn * 3 + 0], k1 = a.K[nn * 3 + 1], k2 = a.K[nn * 3 + 2];
  const float dec = p.pid_decay;
  const float c4pi = 12.566370614359172f;
  float* ws = a.ws + nn;
  const int64_t N = p.N;
  // ---- forward (train_controller.py:44-67), checkpointing the state at the start of every step
  float P = a.reset_pressure, V = p.min_volume, cp = 0.f, ci = 0.f, cd = 0.f, time = 0.f, total = 0.f;
  for (int t = 0; t < p.T; ++t) {
    P += a.noise;                                                  // :51-53 (the noise also enters the state)
    float* w = ws + (int64_t)t * 6 * N;
    w[0] = P; w[N] = V; w[2 * N] = cp; w[3 * N] = ci; w[4 * N] = cd; w[5 * N] = time;
    const float xo = fmod_pos(time, p.period);
    const float err = interp_knots<NK>(p, kf, xo) - P;
    const float np_ = err, ni = ci + dec * (err - ci), nd = cd + dec * ((err - cp) - cd);
    cp = np_; ci = ni; cd = nd;
    const float u_raw = (np_ * k0 + ni * k1) + nd * k2;
    const float u = fminf(fmaxf(u_raw, 0.f), 100.f);
    const bool inhale = xo <= p.xp_in;                             // u_out == 0
    const float vraw = tanhf(0.03f * (3.0f * u - 130.f)) + 1.0f;
    const float valve = fminf(fmaxf(vraw, 0.f), 1.72f);
    float flow = fminf(fmaxf(valve * p.R, 0.f), 2.0f);
    if (P > p.peep_valve) flow -= (inhale ? 0.0f : 1.0f) * 0.05f * P;
    const float diff = interp_knots<NK>(p, kf, fmod_pos(a.tt[t], p.period)) - P;
    if (inhale) total += a.loss_kind == DLC_LOSS_L1 ? fabsf(diff) : diff * diff;
    V += flow * p.dt;
    if (p.leak) V += p.leak_coef * (p.min_volume - V);
    const float r = cbrtf(__fdiv_rn(3.0f * V, c4pi));
    const float q = __fdiv_rn(p.r0, r), q2 = q * q;
    P = p.P0 + __fdiv_rn(p.PC * (1.0f - q2 * q2 * q2), p.r0_sq * r);
    time += p.dt;
  }
  // ---- backward: adjoints of (P, V, p, i, d) after step t, walked down to t = 0
  float lP = 0.f, lV = 0.f, lp = 0.f, li = 0.f, ld = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int t = p.T - 1; t >= 0; --t) {
    const float* w = ws + (int64_t)t * 6 * N;
    const float Pi = w[0], Vi = w[N], pi_ = w[2 * N], ii = w[3 * N], di = w[4 * N], ti = w[5 * N];
    // recompute the step's intermediates
    const float xo = fmod_pos(ti, p.period);
    const float err = interp_knots<NK>(p, kf, xo) - Pi;
    const float np_ = err, ni = ii + dec * (err - ii), nd = di + dec * ((err - pi_) - di);
    const float u_raw = (np_ * k0 + ni * k1) + nd * k2;
    const float u = fminf(fmaxf(u_raw, 0.f), 100.f);
    const bool inhale = xo <= p.xp_in;
    const float th = tanhf(0.03f * (3.0f * u - 130.f));
    const float vraw = th + 1.0f;
    const float vr = fminf(fmaxf(vraw, 0.f), 1.72f) * p.R;
    float flow = fminf(fmaxf(vr, 0.f), 2.0f);
    const bool valve_open = Pi > p.peep_valve;
    if (valve_open) flow -= (inhale ? 0.0f : 1.0f) * 0.05f * Pi;
    float Vn = Vi + flow * p.dt;
    if (p.leak) Vn += p.leak_coef * (p.min_volume - Vn);
    const float r = cbrtf(__fdiv_rn(3.0f * Vn, c4pi));
    // P' = P0 + (PC / r0^2) (1/r - r0^6 / r^7)
    const float ir = __fdiv_rn(1.0f, r), ir2 = ir * ir, q = p.r0 * ir, q2 = q * q, q6 = q2 * q2 * q2;
    const float dPdr = __fdiv_rn(p.PC, p.r0_sq) * (-ir2 + 7.0f * q6 * ir2);
    const float drdV = __fdiv_rn(r, 3.0f * Vn);
    float gV = lV + lP * dPdr * drdV;
    if (p.leak) gV *= (1.0f - p.leak_coef);
    const float gflow = p.dt * gV;
    float nlP = 0.f;
    if (valve_open) nlP -= (inhale ? 0.0f : 1.0f) * 0.05f * gflow;
    const float gvr = (vr >= 0.f && vr <= 2.0f) ? gflow : 0.f;        // clamp: gradient passes inside the range
    const float gvraw = (vraw >= 0.f && vraw <= 1.72f) ? gvr * p.R : 0.f;
    const float gu = 0.09f * (gvraw * (1.0f - th * th));               // d tanh(0.03 (3u - 130)) / du
    const float guraw = (u_raw >= 0.f && u_raw <= 100.f) ? gu : 0.f;
    g0 = fmaf(guraw, np_, g0); g1 = fmaf(guraw, ni, g1); g2 = fmaf(guraw, nd, g2);
    const float gp = lp + k0 * guraw, gi = li + k1 * guraw, gd = ld + k2 * guraw;
    const float gerr = gp + dec * gi + dec * gd;
    nlP -= gerr;                                                       // err = target - P
    if (inhale) {
      const float diff = interp_knots<NK>(p, kf, fmod_pos(a.tt[t], p.period)) - Pi;
      nlP -= a.loss_kind == DLC_LOSS_L1 ? (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.f)) : 2.0f * diff;
    }
    lP = nlP; lV = gV; lp = -dec * gd; li = (1.0f - dec) * gi; ld = (1.0f - dec) * gd;
  }
  if (live) {
    a.loss_out[n] = total;
    a.grad_out[n * 3 + 0] = g0; a.grad_out[n * 3 + 1] = g1; a.grad_out[n * 3 + 2] = g2;
  }
  if (a.sums_out) {   // block reduction, one atomic per block and component
    __shared__ double red[4][4];
    double v[4] = {live ? (double)total : 0.0, live ? (double)g0 : 0.0, live ? (double)g1 : 0.0,
                   live ? (double)g2 : 0.0};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((threadIdx.x & 31) == 0)
      for (int c = 0; c < 4; ++c) red[c][threadIdx.x >> 5] = v[c];
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&a.sums_out[threadIdx.x], (red[threadIdx.x][0] + red[threadIdx.x][1]) +
                                                                  (red[threadIdx.x][2] + red[threadIdx.x][3]));
  }
}

}  // namespace dlc

extern "C" size_t dlc_lung_pid_grad_workspace_bytes(const DlcLungParams* p) {
  if (!p || p->N <= 0 || p->T <= 0) return 0;
  return (size_t)p->T * 6 * (size_t)p->N * sizeof(float);
}

extern "C" int32_t dlc_lung_pid_grad_f32(const DlcLungParams* p, const float* K, const float* kf, const float* tt,
                                         int32_t loss_kind, float noise, float reset_pressure, float* loss_out,
                                         float* gradK_out, double* sums_out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  using namespace dlc;
  if (!p) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: NULL params");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: negative N or T");
  if (p->sim != DLC_SIM_BALLOON) return fail(DLC_ERR_SHAPE, "dlc_lung_pid_grad_f32: BalloonLung only");
  if (p->nk < 2 || p->nk > DLC_MAX_KNOTS) return fail(DLC_ERR_SHAPE, "dlc_lung_pid_grad_f32: 2 <= nk <= 16 knots");
  if (loss_kind != DLC_LOSS_L1 && loss_kind != DLC_LOSS_L2) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: bad loss");
  if (p->N == 0) return DLC_OK;
  if (!K || !kf || (!tt && p->T > 0) || !loss_out || !gradK_out)
    return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: NULL buffer");
  if (p->T > 0 && (!workspace || workspace_bytes < dlc_lung_pid_grad_workspace_bytes(p)))
    return fail(DLC_ERR_WORKSPACE, "dlc_lung_pid_grad_f32: workspace smaller than dlc_lung_pid_grad_workspace_bytes()");
  LungGradArgs a;
  a.p = *p; a.K = K; a.kf = kf; a.tt = tt; a.loss_kind = loss_kind; a.noise = noise; a.reset_pressure = reset_pressure;
  a.loss_out = loss_out; a.grad_out = gradK_out; a.sums_out = sums_out; a.ws = (float*)workspace;
  const unsigned grid = (unsigned)((p->N + 127) / 128);
  if (p->nk == 7) lung_pid_grad_kernel<7><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  else lung_pid_grad_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "lung_pid_grad_kernel launch");
}
