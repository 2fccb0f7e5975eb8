// Fused ventilator episode: PID + Expiratory + BalloonLung / DelayLung.
// One thread per breath; all state in registers (DelayLung's two valve histories in a local
// ring); per step five/six coalesced 4-byte stores per thread ([T,N] time-major series), which is
// what bounds the kernel (HBM writes).
#include <stdlib.h>

#include <mutex>
#include <type_traits>

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
  float x0, w, f0, sl;   // w = x1 - x0 (as computed in float32: x in [x0, x1) <=> 0 <= x - x0 < w up to that rounding,
};                       // and an x within rounding of x1 interpolates to the same value from either side)

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
  s.x0 = x0; s.w = dx; s.f0 = f0;
  s.sl = dx == 0.f ? 0.f : __fdiv_rn(f1 - f0, dx);
}

// The hit test is ONE unsigned compare on x - x0 (which the interpolant needs anyway): a negative or NaN difference
// has its sign / exponent bits above any finite positive width.  The miss is voted warp-wide so that the common
// path carries no divergence bookkeeping (BSSY / BSYNC).
template <int NK>
__device__ __forceinline__ float interp_seg(const DlcLungParams& p, const float* kf, float x, Seg& s, unsigned mask) {
  float d = x - s.x0;
  const bool miss = !(__float_as_uint(d) < __float_as_uint(s.w)) || !(s.x0 <= x);
  if (__builtin_expect(__any_sync(mask, miss), 0)) {
    if (miss) {
      seg_search<NK>(p, kf, x, s);
      d = x - s.x0;
    }
  }
  return fmaf(d, s.sl, s.f0);
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

// MODE >= 0 fixes p.mode at compile time (the closed-loop episode is the hot instance), MODE < 0 reads it; OUTS = 1
// promises that all six series pointers are non-NULL, OUTS = 2 that the four per-breath ones are and that timestamp /
// flow are not wanted (see lung_pid_pair_kernel's LEAN), OUTS = 0 tests every pointer.
template <int SIM, int NK, int MODE, int OUTS>
__global__ void __launch_bounds__(128) lung_pid_kernel(const LungArgs a) {
  extern __shared__ float ring[];                                  // DelayLung valve histories [2][delay][blockDim]
  const DlcLungParams& p = a.p;
  const DlcLungBuffers& b = a.b;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const unsigned lanes = __activemask();   // the warp's live lanes: constant from here on (no early exits below)
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
  // The episode only WRITES the histories (the simulator reads the current action, _delay_lung.py:88-96), so after
  // T >= delay steps nothing of the caller's contents survives: the load is skipped then.
  const bool keep_old = p.T < p.delay || p.mode == DLC_LUNG_MODE_CTRL_ONLY;   // (ctrl-only: the ring passes through)
  const int di = blockDim.x / p.delay, dr = blockDim.x - di * p.delay;     // i += blockDim without a divide per element
  if (SIM == DLC_SIM_DELAY && keep_old) {
    if (full_cta) {
      const int64_t base = (int64_t)blockIdx.x * blockDim.x * p.delay;
      const int count = blockDim.x * p.delay;
      int e = threadIdx.x / p.delay, k = threadIdx.x - e * p.delay;
      for (int i = threadIdx.x; i < count; i += blockDim.x) {
        ring[k * rs + e] = b.in_history_io[base + i];
        ring[(p.delay + k) * rs + e] = b.out_history_io[base + i];
        k += dr; e += di;
        if (k >= p.delay) { k -= p.delay; ++e; }
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
  Seg seg_pid = {0.f, 0.f, 0.f, 0.f}, seg_out = {0.f, 0.f, 0.f, 0.f};   // empty segments: first use searches
  float chk = 0.f;                                                 // 0 * pressure accumulated: NaN iff any non-finite

  // The two clocks the waveform is read on -- the observation time and the emitted timestamp -- advance by dt and
  // dt * dt per step.  When every thread of the warp can promise nondecreasing arguments with steps below one period
  // (the only case in practice), t mod period is carried as a cached quotient: one FMA, and a predicated bump when the
  // phase wraps -- exact, like fmodf (see fmod_fast).  Otherwise the warp runs the loop with the general reduction.
  float q_obs = 0.f, q_ts = 0.f;
  bool fast_clock = false;
  if (MODE == DLC_LUNG_MODE_CLOSED_LOOP) {
    const float nxt = SIM == DLC_SIM_BALLOON ? time + dt : time;   // the observation time after the first step
    const float ts0 = dt * (time + 1.0f) - run_dt;
    const bool valid = dt > 0.f && dt < period && dt * dt < period && obs_t >= 0.f && nxt >= obs_t &&
                       nxt - obs_t < period && ts0 >= 0.f && (time + 2.0f) + (float)p.T * dt < t_max &&
                       dt * ((time + 2.0f) + (float)p.T * dt) < t_max;
    fast_clock = __all_sync(lanes, valid);
    if (fast_clock) {
      q_obs = floorf(obs_t * inv_period);
      float r = fmaf(-q_obs, period, obs_t);
      q_obs += r < 0.f ? -1.0f : (r >= period ? 1.0f : 0.f);
      q_ts = floorf(ts0 * inv_period);
      r = fmaf(-q_ts, period, ts0);
      q_ts += r < 0.f ? -1.0f : (r >= period ? 1.0f : 0.f);
    }
  }

  auto episode = [&](auto fast_tag) {
  constexpr bool FAST = decltype(fast_tag)::value;
  auto phase = [&](float x, float& q) -> float {
    if (!FAST) return fmod_fast(x, period, inv_period, t_max);
    float r = fmaf(-q, period, x);
    if (r >= period) {
      q += 1.0f;
      r = fmaf(-q, period, x);
    }
    return r;
  };
  for (int t = 0; t < p.T; ++t) {
    const float pr_emit = obs_p;                                   // run_controller.py:118
    float u_in, u_out;
    if (run_ctrl) {
      // ---- PID                                                   controllers/_pid.py:77-102
      const float xo = phase(obs_t, q_obs);                        // shared by PID (at) and Expiratory (is_in)
      const float err = interp_seg<NK>(p, kf, xo, seg_pid, lanes) - obs_p;
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
    if (OUTS == 1 || (OUTS == 0 && o_ts)) { *o_ts = ts; o_ts += N; }
    if (OUTS >= 1 || o_uin) { *o_uin = u_in; o_uin += N; }
    if (OUTS >= 1 || o_uout) { *o_uout = u_out; o_uout += N; }
    if (OUTS >= 1 || o_pr) { *o_pr = pr_emit; o_pr += N; }
    if (OUTS == 1 || (OUTS == 0 && o_flow)) { *o_flow = p.env_flow; o_flow += N; }
    if (OUTS >= 1 || o_tgt) {                                          // waveform.at(timestamps), :207
      *o_tgt = interp_seg<NK>(p, kf, phase(ts, q_ts), seg_out, lanes);
      o_tgt += N;
    }
    chk = fmaf(0.f, pressure, chk);
  }
  };
  if (fast_clock) episode(std::true_type{});
  else episode(std::false_type{});
  bad = chk != chk;

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
      int e = threadIdx.x / p.delay, k = threadIdx.x - e * p.delay;
      for (int i = threadIdx.x; i < count; i += blockDim.x) {
        int s = pos + k;
        s -= s >= p.delay ? p.delay : 0;
        b.in_history_io[base + i] = ring[s * rs + e];
        b.out_history_io[base + i] = ring[(p.delay + s) * rs + e];
        k += dr; e += di;
        if (k >= p.delay) { k -= p.delay; ++e; }
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


// ------------------------------------------------------------------------------------------------------------------
// The hot instance (default 7-knot waveform, closed loop, all six series) with TWO breaths per thread.  The kernel is
// bound by its stores (24 B per env-step, nothing read): measured on B200 (tools/bench_series_store.py), six [T,N]
// series written with one 4-byte store per thread top out at ~5.0 TB/s, with 8-byte stores (two adjacent envs per
// thread, 256 contiguous bytes per warp and row) at ~6.3 TB/s -- so each thread owns envs (2j, 2j+1), runs their
// recurrences interleaved (two independent dependency chains) and emits float2 rows.  Arithmetic per env is the
// same sequence of operations as lung_pid_kernel (bit-identical results, tests/test_lung_gpu.py).
struct LungEnv {
  float time, volume, pressure, steps, pipe, obs_p, obs_t, cp, ci, cd, ctime, etime, prev_ctime, prev_etime;
  float q_obs, q_ts, chk;
  Seg seg_pid, seg_out;
};

// LEAN: the timestamp and flow series are not stored.  Both are the same for every breath of a batch that starts at a
// common time -- flow is the env's static field, the timestamp depends on the step only -- so a caller can keep ONE
// column of each (16 B per env-step instead of 24; the Python mirror does, lung/utils/scripts/run_controller.py).
template <int SIM, bool LEAN>
__global__ void __launch_bounds__(128) lung_pid_pair_kernel(const LungArgs a) {
  constexpr int NK = 7;
  extern __shared__ float ring[];                 // DelayLung: [2][delay][2 * blockDim + 1], column = e * blockDim + tid
  const DlcLungParams& p = a.p;
  const DlcLungBuffers& b = a.b;
  const int64_t n0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);      // N is even (launcher)
  const bool live = n0 < p.N;
  const int64_t cta_lo = 2 * (int64_t)blockIdx.x * blockDim.x;
  const int cta_envs = (int)min((int64_t)2 * blockDim.x, p.N - cta_lo);
  const int rs = 2 * blockDim.x + 1;
  const int di = blockDim.x / p.delay, dr = blockDim.x - di * p.delay;       // i += blockDim without a divide per element
  if (SIM == DLC_SIM_DELAY && p.T < p.delay) {    // (after T >= delay steps nothing of the old contents survives)
    const int64_t base = cta_lo * p.delay;        // coalesced transposing copy of the CTA's [envs, delay] chunk
    const int count = cta_envs * p.delay;
    int el = threadIdx.x / p.delay, k = threadIdx.x - el * p.delay;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
      const int col = (el & 1) * blockDim.x + (el >> 1);
      ring[k * rs + col] = b.in_history_io[base + i];
      ring[(p.delay + k) * rs + col] = b.out_history_io[base + i];
      k += dr; el += di;
      if (k >= p.delay) { k -= p.delay; ++el; }
    }
    __syncthreads();
  }
  int pos = 0;
  const unsigned lanes = __ballot_sync(0xffffffffu, live);
  if (live) {
    float kf[2][NK], kg[2][3];
    LungEnv E[2];
    float target[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t n = n0 + e;
      const float* g = b.kf + (p.kf_per_env ? n * NK : 0);
#pragma unroll
      for (int k = 0; k < NK; ++k) kf[e][k] = g[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) kg[e][k] = b.K[n * 3 + k];
    }
    {
      auto ld2 = [&](const float* src, float& x0, float& x1) {
        const float2 v = *reinterpret_cast<const float2*>(src + n0);
        x0 = v.x; x1 = v.y;
      };
      ld2(b.time_io, E[0].time, E[1].time); ld2(b.volume_io, E[0].volume, E[1].volume);
      ld2(b.pressure_io, E[0].pressure, E[1].pressure); ld2(b.steps_io, E[0].steps, E[1].steps);
      ld2(b.target_io, target[0], target[1]);
      if (SIM == DLC_SIM_DELAY) ld2(b.pipe_pressure_io, E[0].pipe, E[1].pipe);
      else E[0].pipe = E[1].pipe = 0.f;
      ld2(b.obs_pressure_io, E[0].obs_p, E[1].obs_p); ld2(b.obs_time_io, E[0].obs_t, E[1].obs_t);
      ld2(b.pid_p_io, E[0].cp, E[1].cp); ld2(b.pid_i_io, E[0].ci, E[1].ci); ld2(b.pid_d_io, E[0].cd, E[1].cd);
      ld2(b.pid_time_io, E[0].ctime, E[1].ctime); ld2(b.exp_time_io, E[0].etime, E[1].etime);
    }
    const float period = p.period, xp_in = p.xp_in, dec = p.pid_decay, dt = p.dt, run_dt = p.run_dt;
    const float inv_period = __fdiv_rn(1.0f, p.period), t_max = 1048576.0f * p.period;
    const float pc_r0sq = __fdiv_rn(SIM == DLC_SIM_BALLOON ? p.PC : p.d_C, p.r0_sq), inv_dR = __fdiv_rn(1.0f, p.d_R);
    const float early_end = (float)p.delay;
    bool valid = true;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      LungEnv& s = E[e];
      s.prev_ctime = s.ctime; s.prev_etime = s.etime;
      s.seg_pid = Seg{0.f, 0.f, 0.f, 0.f}; s.seg_out = Seg{0.f, 0.f, 0.f, 0.f};
      s.chk = 0.f; s.q_obs = 0.f; s.q_ts = 0.f;
      const float nxt = SIM == DLC_SIM_BALLOON ? s.time + dt : s.time;
      const float ts0 = dt * (s.time + 1.0f) - run_dt;
      valid = valid && dt > 0.f && dt < period && dt * dt < period && s.obs_t >= 0.f && nxt >= s.obs_t &&
              nxt - s.obs_t < period && ts0 >= 0.f && (s.time + 2.0f) + (float)p.T * dt < t_max &&
              dt * ((s.time + 2.0f) + (float)p.T * dt) < t_max;
    }
    const bool fast_clock = __all_sync(lanes, valid);   // see lung_pid_kernel: cached-quotient t mod period
    if (fast_clock) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        LungEnv& s = E[e];
        s.q_obs = floorf(s.obs_t * inv_period);
        float r = fmaf(-s.q_obs, period, s.obs_t);
        s.q_obs += r < 0.f ? -1.0f : (r >= period ? 1.0f : 0.f);
        const float ts0 = dt * (s.time + 1.0f) - run_dt;
        s.q_ts = floorf(ts0 * inv_period);
        r = fmaf(-s.q_ts, period, ts0);
        s.q_ts += r < 0.f ? -1.0f : (r >= period ? 1.0f : 0.f);
      }
    }
    const int64_t N2 = p.N >> 1;
    float2* o_ts = LEAN ? nullptr : reinterpret_cast<float2*>(b.timestamp_out + n0);
    float2* o_uin = reinterpret_cast<float2*>(b.u_in_out + n0);
    float2* o_uout = reinterpret_cast<float2*>(b.u_out_out + n0);
    float2* o_pr = reinterpret_cast<float2*>(b.pressure_out + n0);
    float2* o_flow = LEAN ? nullptr : reinterpret_cast<float2*>(b.flow_out + n0);
    float2* o_tgt = reinterpret_cast<float2*>(b.target_out + n0);

    auto episode = [&](auto fast_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      auto phase = [&](float x, float& q) -> float {
        if (!FAST) return fmod_fast(x, period, inv_period, t_max);
        float r = fmaf(-q, period, x);
        if (r >= period) {
          q += 1.0f;
          r = fmaf(-q, period, x);
        }
        return r;
      };
      for (int t = 0; t < p.T; ++t) {
        float v_ts[2], v_uin[2], v_uout[2], v_pr[2], v_tgt[2];
        if (SIM == DLC_SIM_DELAY) pos = pos == 0 ? p.delay - 1 : pos - 1;          // roll(+1); [0] <- u
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          LungEnv& s = E[e];
          v_pr[e] = s.obs_p;                                           // run_controller.py:118
          // ---- PID                                                  controllers/_pid.py:77-102
          const float xo = phase(s.obs_t, s.q_obs);
          const float err = interp_seg<NK>(p, kf[e], xo, s.seg_pid, lanes) - s.obs_p;
          const float ni = s.ci + dec * (err - s.ci);
          const float nd = s.cd + dec * ((err - s.cp) - s.cd);
          s.cp = err; s.ci = ni; s.cd = nd;
          float u_in = (err * kg[e][0] + ni * kg[e][1]) + nd * kg[e][2];
          u_in = fminf(fmaxf(u_in, 0.f), 100.f);
          const float u_out = (xo <= xp_in) ? 0.f : 1.f;               // controllers/_expiratory.py:35-45
          s.prev_ctime = s.ctime; s.prev_etime = s.etime;
          s.ctime = s.obs_t; s.etime = s.obs_t;
          // ---- simulator
          if (SIM == DLC_SIM_BALLOON) {                                // envs/_balloon_lung.py:100-146
            float valve = tanh_plus_one(0.03f * (3.0f * u_in - 130.f));
            valve = fminf(fmaxf(valve, 0.f), 1.72f);
            float flow = fminf(fmaxf(valve * p.R, 0.f), 2.0f);
            if (s.pressure > p.peep_valve) flow -= (u_out > 0.f ? 1.0f : 0.0f) * 0.05f * s.pressure;
            s.volume += flow * dt;
            if (p.leak) s.volume += p.leak_coef * (p.min_volume - s.volume);
            const float r = cbrt_pos(s.volume * 0.238732414637843f);
            const float ir = rcpf_approx(r);
            const float q = p.r0 * ir, q2 = q * q;
            s.pressure = fmaf(pc_r0sq * fmaf(-q2 * q2, q2, 1.0f), ir, p.P0);
            s.steps = s.time + 1.0f;                                    // sic (:135)
            s.time = s.time + dt;
            s.obs_p = s.pressure;
            s.obs_t = s.time;
          } else {                                                     // envs/_delay_lung.py:75-133
            const float uin_pos = u_in > 0.f ? u_in : 0.f;
            const int col = e * blockDim.x + threadIdx.x;
            ring[pos * rs + col] = uin_pos;
            ring[(p.delay + pos) * rs + col] = u_out;
            s.obs_p = s.pressure;
            s.obs_t = s.time;
            const float flow = s.pressure * inv_dR;
            s.volume = fmaxf(s.volume + flow * dt, p.min_volume);
            const float r = cbrt_pos(s.volume * 0.238732414637843f);
            const float ir = rcpf_approx(r);
            const float q = p.r0 * ir, q2 = q * q;
            const float lung_p = pc_r0sq * fmaf(-q2 * q2, q2, 1.0f) * ir;
            const bool early = s.time < early_end;                      // seconds vs steps, sic (:88-96)
            const float impulse = early ? 0.f : p.control_gain * uin_pos;
            const float peep = early ? 0.f : u_out;
            s.pipe = p.inertia * s.pipe + impulse;
            s.pressure = fmaxf(0.f, s.pipe - lung_p);
            if (peep != 0.f) s.pipe *= 0.995f;
            s.steps = s.time + 1.0f;
            s.time = s.time + dt;
          }
          const float ts = dt * s.steps - run_dt;                       // run_controller.py:128-141
          v_ts[e] = ts; v_uin[e] = u_in; v_uout[e] = u_out;
          v_tgt[e] = interp_seg<NK>(p, kf[e], phase(ts, s.q_ts), s.seg_out, lanes);
          s.chk = fmaf(0.f, s.pressure, s.chk);
        }
        if (!LEAN) { *o_ts = make_float2(v_ts[0], v_ts[1]); o_ts += N2; }
        *o_uin = make_float2(v_uin[0], v_uin[1]); o_uin += N2;
        *o_uout = make_float2(v_uout[0], v_uout[1]); o_uout += N2;
        *o_pr = make_float2(v_pr[0], v_pr[1]); o_pr += N2;
        if (!LEAN) { *o_flow = make_float2(p.env_flow, p.env_flow); o_flow += N2; }
        *o_tgt = make_float2(v_tgt[0], v_tgt[1]); o_tgt += N2;
      }
    };
    if (fast_clock) episode(std::true_type{});
    else episode(std::false_type{});

    if (p.T > 0) {
#pragma unroll
      for (int e = 0; e < 2; ++e) target[e] = interp_knots<NK>(p, kf[e], fmod_pos(E[e].time, p.period));   // :140
    }
    auto st2 = [&](float* dst, float x0, float x1) { *reinterpret_cast<float2*>(dst + n0) = make_float2(x0, x1); };
    st2(b.steps_io, E[0].steps, E[1].steps); st2(b.time_io, E[0].time, E[1].time);
    st2(b.volume_io, E[0].volume, E[1].volume); st2(b.pressure_io, E[0].pressure, E[1].pressure);
    st2(b.target_io, target[0], target[1]);
    st2(b.obs_pressure_io, E[0].obs_p, E[1].obs_p); st2(b.obs_time_io, E[0].obs_t, E[1].obs_t);
    st2(b.pid_p_io, E[0].cp, E[1].cp); st2(b.pid_i_io, E[0].ci, E[1].ci); st2(b.pid_d_io, E[0].cd, E[1].cd);
    if (p.T > 0) {
      auto bk = [&](float cur, float prev) { return fmaxf(p.default_dt, cur - (isinf(prev) ? 0.f : prev)); };
      st2(b.pid_time_io, E[0].ctime, E[1].ctime);
      st2(b.pid_dt_io, bk(E[0].ctime, E[0].prev_ctime), bk(E[1].ctime, E[1].prev_ctime));
      st2(b.exp_time_io, E[0].etime, E[1].etime);
      st2(b.exp_dt_io, bk(E[0].etime, E[0].prev_etime), bk(E[1].etime, E[1].prev_etime));
      int2* ps = reinterpret_cast<int2*>(b.pid_steps_io + n0);
      int2 v = *ps; v.x += p.T; v.y += p.T; *ps = v;
      int2* xs = reinterpret_cast<int2*>(b.exp_steps_io + n0);
      v = *xs; v.x += p.T; v.y += p.T; *xs = v;
    }
    if (SIM == DLC_SIM_DELAY) st2(b.pipe_pressure_io, E[0].pipe, E[1].pipe);
    if (b.status_out)
      *reinterpret_cast<int2*>(b.status_out + n0) = make_int2(E[0].chk != E[0].chk ? DLC_STATUS_NONFINITE : 0,
                                                              E[1].chk != E[1].chk ? DLC_STATUS_NONFINITE : 0);
  }
  if (SIM == DLC_SIM_DELAY) {                      // pos is the same in every live thread (T steps each)
    pos = p.T % p.delay == 0 ? 0 : p.delay - p.T % p.delay;
    __syncthreads();
    const int64_t base = cta_lo * p.delay;
    const int count = cta_envs * p.delay;
    int el = threadIdx.x / p.delay, k = threadIdx.x - el * p.delay;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
      const int col = (el & 1) * blockDim.x + (el >> 1);
      int s = pos + k;
      s -= s >= p.delay ? p.delay : 0;
      b.in_history_io[base + i] = ring[s * rs + col];
      b.out_history_io[base + i] = ring[(p.delay + s) * rs + col];
      k += dr; el += di;
      if (k >= p.delay) { k -= p.delay; ++el; }
    }
  }
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
  // hot instances: default 7-knot waveform table, closed loop, and either all six series or the four that differ between
  // breaths (timestamp_out and flow_out both NULL: "lean")
  const bool four = p->nk == 7 && p->mode == DLC_LUNG_MODE_CLOSED_LOOP && b->u_in_out && b->u_out_out &&
                    b->pressure_out && b->target_out;
  const bool lean = four && !b->timestamp_out && !b->flow_out;
  const bool hot = four && b->timestamp_out && b->flow_out;
  // two breaths per thread (float2 rows) when the batch is even and every [N] array is 8-byte aligned
  auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
  // (measured, tools/bench_lung_n.py: the wider stores win from ~48 steps on BalloonLung, ~96 on DelayLung; below that
  // the one-per-thread kernel's higher occupancy hides the state load / store better)
  const bool long_enough = p->kernel_variant == 2 || p->T >= (p->sim == DLC_SIM_BALLOON ? 48 : 96);
  bool pair = (hot || lean) && p->kernel_variant != 1 && long_enough && (p->N % 2 == 0) && al8(b->timestamp_out) &&
              al8(b->u_in_out) && al8(b->u_out_out) && al8(b->pressure_out) && al8(b->flow_out) && al8(b->target_out) &&
              al8(b->steps_io) && al8(b->time_io) && al8(b->volume_io) && al8(b->pressure_io) && al8(b->target_io) &&
              al8(b->obs_pressure_io) && al8(b->obs_time_io) && al8(b->pid_p_io) && al8(b->pid_i_io) &&
              al8(b->pid_d_io) && al8(b->pid_time_io) && al8(b->pid_dt_io) && al8(b->pid_steps_io) &&
              al8(b->exp_time_io) && al8(b->exp_dt_io) && al8(b->exp_steps_io) &&
              (!b->status_out || al8(b->status_out)) && (p->sim != DLC_SIM_DELAY || al8(b->pipe_pressure_io));
  const unsigned grid2 = (unsigned)((p->N / 2 + 127) / 128);
  if (pair && p->sim == DLC_SIM_BALLOON) {
    if (lean) lung_pid_pair_kernel<DLC_SIM_BALLOON, true><<<grid2, 128, 0, st>>>(a);
    else lung_pid_pair_kernel<DLC_SIM_BALLOON, false><<<grid2, 128, 0, st>>>(a);
    return cuda_status(cudaGetLastError(), "lung_pid_pair_kernel launch");
  }
  if (pair) {
    const size_t smem2 = 2 * (size_t)p->delay * 257 * sizeof(float);   // 128.5 KB at the maximum delay of 64
    static std::once_flag once2;
    static cudaError_t rc2;
    std::call_once(once2, [] {
      rc2 = cudaFuncSetAttribute(lung_pid_pair_kernel<DLC_SIM_DELAY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 2 * DLC_MAX_DELAY * 257 * 4);
      if (rc2 == cudaSuccess)
        rc2 = cudaFuncSetAttribute(lung_pid_pair_kernel<DLC_SIM_DELAY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   2 * DLC_MAX_DELAY * 257 * 4);
    });
    if (rc2 != cudaSuccess) return cuda_status(rc2, "lung_pid_pair_kernel shared-memory attribute");
    if (lean) lung_pid_pair_kernel<DLC_SIM_DELAY, true><<<grid2, 128, smem2, st>>>(a);
    else lung_pid_pair_kernel<DLC_SIM_DELAY, false><<<grid2, 128, smem2, st>>>(a);
    return cuda_status(cudaGetLastError(), "lung_pid_pair_kernel launch");
  }
  if (p->sim == DLC_SIM_BALLOON) {
    if (hot) lung_pid_kernel<DLC_SIM_BALLOON, 7, DLC_LUNG_MODE_CLOSED_LOOP, 1><<<grid, 128, 0, st>>>(a);
    else if (lean) lung_pid_kernel<DLC_SIM_BALLOON, 7, DLC_LUNG_MODE_CLOSED_LOOP, 2><<<grid, 128, 0, st>>>(a);
    else if (p->nk == 7) lung_pid_kernel<DLC_SIM_BALLOON, 7, -1, 0><<<grid, 128, 0, st>>>(a);
    else lung_pid_kernel<DLC_SIM_BALLOON, 0, -1, 0><<<grid, 128, 0, st>>>(a);
  } else {
    const size_t smem = 2 * (size_t)p->delay * 129 * sizeof(float);   // 64.5 KB at the maximum delay of 64
    if (smem > 48 * 1024) {
      static std::once_flag once;
      static cudaError_t attr_rc[4];
      std::call_once(once, [] {
        const int cap = 2 * DLC_MAX_DELAY * 129 * 4;
        attr_rc[3] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, 2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_rc[0] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, 1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_rc[1] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 7, -1, 0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_rc[2] = cudaFuncSetAttribute(lung_pid_kernel<DLC_SIM_DELAY, 0, -1, 0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
      });
      for (int i = 0; i < 4; ++i)
        if (attr_rc[i] != cudaSuccess) return cuda_status(attr_rc[i], "lung_pid_kernel shared-memory attribute");
    }
    if (hot) lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, 1><<<grid, 128, smem, st>>>(a);
    else if (lean) lung_pid_kernel<DLC_SIM_DELAY, 7, DLC_LUNG_MODE_CLOSED_LOOP, 2><<<grid, 128, smem, st>>>(a);
    else if (p->nk == 7) lung_pid_kernel<DLC_SIM_DELAY, 7, -1, 0><<<grid, 128, smem, st>>>(a);
    else lung_pid_kernel<DLC_SIM_DELAY, 0, -1, 0><<<grid, 128, smem, st>>>(a);
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


// The same value and gradient in FORWARD mode: the leaves are three gains, so their tangents ride along the episode --
// (dP, dV, dp, di, dd, dloss) x 3 = 18 registers -- and nothing is checkpointed: no workspace, no second sweep, no HBM
// traffic beyond the gains in and loss + gradient out (the reverse-mode kernel above writes and re-reads 48 B per
// env-step and recomputes every step).  Same step as lung_pid_grad_kernel's forward sweep, operation for operation.
template <int NK, int SIM>
__global__ void __launch_bounds__(128) lung_pid_grad_fwd_kernel(const LungGradArgs a) {
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
  const float kg[3] = {a.K[nn * 3 + 0], a.K[nn * 3 + 1], a.K[nn * 3 + 2]};
  const float dec = p.pid_decay;
  const float pc_r0sq = __fdiv_rn(SIM == DLC_SIM_BALLOON ? p.PC : p.d_C, p.r0_sq);
  float P = a.reset_pressure, V = p.min_volume, cp = 0.f, ci = 0.f, cd = 0.f, time = 0.f, total = 0.f;
  float dP[3] = {0.f, 0.f, 0.f}, dV[3] = {0.f, 0.f, 0.f}, dcp[3] = {0.f, 0.f, 0.f}, dci[3] = {0.f, 0.f, 0.f},
        dcd[3] = {0.f, 0.f, 0.f}, dtot[3] = {0.f, 0.f, 0.f};
  if (SIM == DLC_SIM_DELAY) {
    // DelayLung (envs/_delay_lung.py:75-133) under train_controller.rollout: the observation is the PRE-dynamics
    // state, so the controllers' clock lags the simulator's by a step; the action enters the pipe through
    // in_history[0] -- the entry just written (:120-123) -- once `time < delay` (seconds against a step count, sic
    // :88-96) stops holding, which is why the gradient is exactly zero over the first 25 s with the default delay.
    // Forward arithmetic as lung_pid_kernel's DelayLung branch, operation for operation.
    const float inv_dR = __fdiv_rn(1.0f, p.d_R), early_end = (float)p.delay;
    float pipe = 0.f, obs_t = 0.f, dpipe[3] = {0.f, 0.f, 0.f};
    for (int t = 0; t < p.T; ++t) {
      P += a.noise;                                                // train_controller.py:51-53
      const float xo = fmod_pos(obs_t, p.period);
      const float err = interp_knots<NK>(p, kf, xo) - P;
      const float ni = ci + dec * (err - ci), nd = cd + dec * ((err - cp) - cd);
      const float pid[3] = {err, ni, nd};
      const float u_raw = (err * kg[0] + ni * kg[1]) + nd * kg[2];
      const float u = fminf(fmaxf(u_raw, 0.f), 100.f);
      const bool inhale = xo <= p.xp_in;                           // u_out == 0
      const bool early = time < early_end;
      // d impulse / d u_raw: the clamp passes inside its range, `u_in > 0` (:113-115) on the positive side
      const float chain = (!early && u_raw > 0.f && u_raw <= 100.f) ? p.control_gain : 0.f;
      const float diff = interp_knots<NK>(p, kf, fmod_pos(a.tt[t], p.period)) - P;
      float dl = 0.f;
      if (inhale) {
        total += a.loss_kind == DLC_LOSS_L1 ? fabsf(diff) : diff * diff;
        dl = a.loss_kind == DLC_LOSS_L1 ? (diff > 0.f ? -1.0f : (diff < 0.f ? 1.0f : 0.f)) : -2.0f * diff;
      }
      const float uin_pos = u > 0.f ? u : 0.f;
      obs_t = time;
      const float Vraw = V + (P * inv_dR) * p.dt;
      const bool v_free = Vraw > p.min_volume;                     // jnp.maximum(volume, min_volume)
      const float Vn = fmaxf(Vraw, p.min_volume);
      const float r = cbrt_pos(Vn * 0.238732414637843f);
      const float ir = rcpf_approx(r), ir2 = ir * ir, q = p.r0 * ir, q2 = q * q, q6 = q2 * q2 * q2;
      const float lung_p = pc_r0sq * fmaf(-q2 * q2, q2, 1.0f) * ir;
      const float dLdV = pc_r0sq * (-ir2 + 7.0f * q6 * ir2) * (r * rcpf_approx(3.0f * Vn));
      const float pipe_n = p.inertia * pipe + (early ? 0.f : p.control_gain * uin_pos);
      const float Praw = pipe_n - lung_p;
      const bool p_free = Praw > 0.f;                              // jnp.maximum(0, pipe_pressure - lung_pressure)
      const float bleed = (!early && !inhale) ? 0.995f : 1.0f;     // peep = out_history[0] = u_out (:93-102)
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        dtot[g] = fmaf(dl, dP[g], dtot[g]);
        const float derr = -dP[g];
        const float dni = dci[g] + dec * (derr - dci[g]);
        const float dnd = dcd[g] + dec * ((derr - dcp[g]) - dcd[g]);
        const float du_raw = ((derr * kg[0] + dni * kg[1]) + dnd * kg[2]) + pid[g];
        dcp[g] = derr; dci[g] = dni; dcd[g] = dnd;
        dV[g] = v_free ? dV[g] + p.dt * inv_dR * dP[g] : 0.f;
        const float dpn = p.inertia * dpipe[g] + chain * du_raw;
        dP[g] = p_free ? dpn - dLdV * dV[g] : 0.f;
        dpipe[g] = bleed * dpn;
      }
      cp = err; ci = ni; cd = nd;
      V = Vn;
      P = fmaxf(0.f, Praw);
      pipe = bleed == 1.0f ? pipe_n : pipe_n * 0.995f;
      time += p.dt;
    }
  }
  for (int t = 0; SIM == DLC_SIM_BALLOON && t < p.T; ++t) {
    P += a.noise;                                                  // train_controller.py:51-53
    const float xo = fmod_pos(time, p.period);
    const float err = interp_knots<NK>(p, kf, xo) - P;
    const float ni = ci + dec * (err - ci), nd = cd + dec * ((err - cp) - cd);
    const float pid[3] = {err, ni, nd};
    const float u_raw = (err * kg[0] + ni * kg[1]) + nd * kg[2];
    const float u = fminf(fmaxf(u_raw, 0.f), 100.f);
    const bool u_in = u_raw >= 0.f && u_raw <= 100.f;             // clamp: the tangent passes inside the range
    const bool inhale = xo <= p.xp_in;                             // u_out == 0
    const float vraw = tanh_plus_one(0.03f * (3.0f * u - 130.f));   // (the rollout kernel's MUFU forms throughout)
    const float th = vraw - 1.0f;
    const float vr = fminf(fmaxf(vraw, 0.f), 1.72f) * p.R;
    float flow = fminf(fmaxf(vr, 0.f), 2.0f);
    const bool valve_open = P > p.peep_valve;
    const float bleed = (valve_open && !inhale) ? 0.05f : 0.f;
    flow -= bleed * P;
    // d flow / d u_raw through tanh and the three clamps
    const float chain = (u_in && vraw >= 0.f && vraw <= 1.72f && vr >= 0.f && vr <= 2.0f)
                            ? 0.09f * (1.0f - th * th) * p.R : 0.f;
    const float diff = interp_knots<NK>(p, kf, fmod_pos(a.tt[t], p.period)) - P;
    float dl = 0.f;                                                // d loss_t / d P
    if (inhale) {
      total += a.loss_kind == DLC_LOSS_L1 ? fabsf(diff) : diff * diff;
      dl = a.loss_kind == DLC_LOSS_L1 ? (diff > 0.f ? -1.0f : (diff < 0.f ? 1.0f : 0.f)) : -2.0f * diff;
    }
    float Vn = V + flow * p.dt;
    if (p.leak) Vn += p.leak_coef * (p.min_volume - Vn);
    const float r = cbrt_pos(Vn * 0.238732414637843f);              // (3V / 4pi)^(1/3)
    const float ir = rcpf_approx(r), ir2 = ir * ir, q = p.r0 * ir, q2 = q * q, q6 = q2 * q2 * q2;
    const float dPdV = pc_r0sq * (-ir2 + 7.0f * q6 * ir2) * (r * rcpf_approx(3.0f * Vn));
    const float keep = p.leak ? 1.0f - p.leak_coef : 1.0f;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      dtot[g] = fmaf(dl, dP[g], dtot[g]);
      const float derr = -dP[g];
      const float dni = dci[g] + dec * (derr - dci[g]);
      const float dnd = dcd[g] + dec * ((derr - dcp[g]) - dcd[g]);
      const float du_raw = ((derr * kg[0] + dni * kg[1]) + dnd * kg[2]) + pid[g];
      const float dflow = chain * du_raw - bleed * dP[g];
      dcp[g] = derr; dci[g] = dni; dcd[g] = dnd;
      dV[g] = keep * (dV[g] + p.dt * dflow);
      dP[g] = dPdV * dV[g];
    }
    cp = err; ci = ni; cd = nd;
    V = Vn;
    P = fmaf(pc_r0sq * fmaf(-q2 * q2, q2, 1.0f), ir, p.P0);         // P0 + PC (1 - (r0/r)^6) / (r0^2 r)
    time += p.dt;
  }
  if (live) {
    a.loss_out[n] = total;
    a.grad_out[n * 3 + 0] = dtot[0]; a.grad_out[n * 3 + 1] = dtot[1]; a.grad_out[n * 3 + 2] = dtot[2];
  }
  if (a.sums_out) {   // block reduction, one atomic per block and component
    __shared__ double red[4][4];
    double v[4] = {live ? (double)total : 0.0, live ? (double)dtot[0] : 0.0, live ? (double)dtot[1] : 0.0,
                   live ? (double)dtot[2] : 0.0};
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

// DLC_LUNG_GRAD_REVERSE=1 selects the checkpointed reverse-mode kernel (kept for cross-checks and measurements); the
// default forward-mode kernel needs no workspace.
static bool lung_grad_reverse() {
  static const bool rev = [] { const char* e = getenv("DLC_LUNG_GRAD_REVERSE"); return e && atoi(e) != 0; }();
  return rev;
}

extern "C" size_t dlc_lung_pid_grad_workspace_bytes(const DlcLungParams* p) {
  if (!p || p->N <= 0 || p->T <= 0 || !lung_grad_reverse()) return 0;
  return (size_t)p->T * 6 * (size_t)p->N * sizeof(float);
}

extern "C" int32_t dlc_lung_pid_grad_f32(const DlcLungParams* p, const float* K, const float* kf, const float* tt,
                                         int32_t loss_kind, float noise, float reset_pressure, float* loss_out,
                                         float* gradK_out, double* sums_out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  using namespace dlc;
  if (!p) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: NULL params");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: negative N or T");
  if (p->sim != DLC_SIM_BALLOON && p->sim != DLC_SIM_DELAY)
    return fail(DLC_ERR_SHAPE, "dlc_lung_pid_grad_f32: sim must be DLC_SIM_BALLOON or DLC_SIM_DELAY");
  if (p->sim == DLC_SIM_DELAY && lung_grad_reverse())
    return fail(DLC_ERR_SHAPE, "dlc_lung_pid_grad_f32: the reverse-mode cross-check kernel covers BalloonLung only");
  if (p->nk < 2 || p->nk > DLC_MAX_KNOTS) return fail(DLC_ERR_SHAPE, "dlc_lung_pid_grad_f32: 2 <= nk <= 16 knots");
  if (loss_kind != DLC_LOSS_L1 && loss_kind != DLC_LOSS_L2) return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: bad loss");
  if (p->N == 0) return DLC_OK;
  if (!K || !kf || (!tt && p->T > 0) || !loss_out || !gradK_out)
    return fail(DLC_ERR_ARG, "dlc_lung_pid_grad_f32: NULL buffer");
  const bool reverse = lung_grad_reverse();
  if (reverse && p->T > 0 && (!workspace || workspace_bytes < dlc_lung_pid_grad_workspace_bytes(p)))
    return fail(DLC_ERR_WORKSPACE, "dlc_lung_pid_grad_f32: workspace smaller than dlc_lung_pid_grad_workspace_bytes()");
  LungGradArgs a;
  a.p = *p; a.K = K; a.kf = kf; a.tt = tt; a.loss_kind = loss_kind; a.noise = noise; a.reset_pressure = reset_pressure;
  a.loss_out = loss_out; a.grad_out = gradK_out; a.sums_out = sums_out; a.ws = (float*)workspace;
  const unsigned grid = (unsigned)((p->N + 127) / 128);
  if (reverse) {
    if (p->nk == 7) lung_pid_grad_kernel<7><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    else lung_pid_grad_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  } else {
    const cudaStream_t st = (cudaStream_t)stream;
    if (p->sim == DLC_SIM_DELAY) {
      if (p->nk == 7) lung_pid_grad_fwd_kernel<7, DLC_SIM_DELAY><<<grid, 128, 0, st>>>(a);
      else lung_pid_grad_fwd_kernel<0, DLC_SIM_DELAY><<<grid, 128, 0, st>>>(a);
    } else {
      if (p->nk == 7) lung_pid_grad_fwd_kernel<7, DLC_SIM_BALLOON><<<grid, 128, 0, st>>>(a);
      else lung_pid_grad_fwd_kernel<0, DLC_SIM_BALLOON><<<grid, 128, 0, st>>>(a);
    }
  }
  return cuda_status(cudaGetLastError(), "lung_pid_grad_kernel launch");
}
