// Fused ventilator episode: PID + Expiratory + BalloonLung / DelayLung.
// One thread per breath; all state in registers (DelayLung's two valve histories in a local
// ring); per step five/six coalesced 4-byte stores per thread ([T,N] time-major series), which is
// what bounds the kernel (HBM writes).
#include "common.cuh"

namespace dlc {

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

template <int SIM, int NK>
__global__ void __launch_bounds__(128) lung_pid_kernel(const LungArgs a) {
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
  float cp = b.pid_p_io[n], ci = b.pid_i_io[n], cd = b.pid_d_io[n], ctime = b.pid_time_io[n], cdt = b.pid_dt_io[n];
  int csteps = b.pid_steps_io[n];
  float etime = b.exp_time_io[n], edt = b.exp_dt_io[n];
  int esteps = b.exp_steps_io[n];
  float hin[SIM == DLC_SIM_DELAY ? DLC_MAX_DELAY : 1], hout[SIM == DLC_SIM_DELAY ? DLC_MAX_DELAY : 1];
  int pos = 0;  // ring slot of logical index 0 (the newest entry)
  if (SIM == DLC_SIM_DELAY) {
    for (int k = 0; k < p.delay; ++k) {
      hin[k] = b.in_history_io[n * p.delay + k];
      hout[k] = b.out_history_io[n * p.delay + k];
    }
  }
  bool bad = false, ran_env = false;
  const bool run_ctrl = p.mode != DLC_LUNG_MODE_ENV_ONLY, run_env = p.mode != DLC_LUNG_MODE_CTRL_ONLY;

  for (int t = 0; t < p.T; ++t) {
    const float pr_emit = obs_p;                                   // run_controller.py:118
    float u_in, u_out;
    if (run_ctrl) {
    // ---- PID                                                     controllers/_pid.py:77-102
    const float xo = fmod_pos(obs_t, p.period);                      // shared by PID (at) and Expiratory (is_in)
    const float tgt = interp_knots<NK>(p, kf, xo);
    const float err = tgt - obs_p;
    const float np_ = err;
    const float ni = ci + p.pid_decay * (err - ci);
    const float nd = cd + p.pid_decay * ((err - cp) - cd);
    cp = np_; ci = ni; cd = nd;
    u_in = (np_ * k0 + ni * k1) + nd * k2;                          // Dense(1, use_bias=False)
    u_in = fminf(fmaxf(u_in, 0.f), 100.f);                          // lax.clamp(0, u, 100)
    cdt = fmaxf(p.default_dt, obs_t - (isinf(ctime) ? 0.f : ctime));
    ctime = obs_t;
    ++csteps;
    // ---- Expiratory                                              controllers/_expiratory.py:35-45
    u_out = (xo <= p.xp_in) ? 0.f : 1.f;
    edt = fmaxf(p.default_dt, obs_t - (isinf(etime) ? 0.f : etime));
    etime = obs_t;
    ++esteps;
    } else {
      u_in = b.u_in_in[(int64_t)t * p.N + n];
      u_out = b.u_out_in[(int64_t)t * p.N + n];
    }
    // ---- simulator
    if (!run_env) {
    } else if (SIM == DLC_SIM_BALLOON) {                                   // envs/_balloon_lung.py:100-146
      float valve = 1.0f * (tanhf(0.03f * (3.0f * u_in - 130.f)) + 1.0f);
      valve = fminf(fmaxf(valve, 0.f), 1.72f);
      float flow = fminf(fmaxf(valve * p.R, 0.f), 2.0f);
      if (pressure > p.peep_valve) flow -= (u_out > 0.f ? 1.0f : 0.0f) * 0.05f * pressure;
      volume += flow * p.dt;
      if (p.leak) volume += p.leak_coef * (p.min_volume - volume);
      // (3V/4pi)^(1/3): cbrtf differs from the reference's pow(x, float32(1/3)) by < 1 ulp
      const float r = cbrtf(__fdiv_rn(3.0f * volume, 12.566370614359172f));
      const float q = __fdiv_rn(p.r0, r);
      const float q2 = q * q;
      pressure = p.P0 + __fdiv_rn(p.PC * (1.0f - q2 * q2 * q2), p.r0_sq * r);
      steps = time + 1.0f;                                          // sic (:135)
      time = time + p.dt;                                           // target = waveform.at(time): only the
      ran_env = true;                                               // last one survives, formed after the loop
      obs_p = pressure;
      obs_t = time;
    } else {                                                        // envs/_delay_lung.py:75-133
      const float uin_pos = u_in > 0.f ? u_in : 0.f;
      pos = pos == 0 ? p.delay - 1 : pos - 1;                       // roll(+1); [0] <- u
      hin[pos] = uin_pos;
      hout[pos] = u_out;
      obs_p = pressure;                                             // observation of the pre-dynamics state
      obs_t = time;
      const float flow = __fdiv_rn(pressure, p.d_R);
      volume = fmaxf(volume + flow * p.dt, p.min_volume);
      const float r = cbrtf(__fdiv_rn(3.0f * volume, 12.566370614359172f));
      const float q = __fdiv_rn(p.r0, r);
      const float q2 = q * q;
      const float lung_p = __fdiv_rn(p.d_C * (1.0f - q2 * q2 * q2), p.r0_sq * r);
      const bool early = time < (float)p.delay;                     // seconds vs steps, sic (:88-96)
      const float impulse = early ? 0.f : p.control_gain * uin_pos;
      const float peep = early ? 0.f : u_out;
      pipe = p.inertia * pipe + impulse;
      pressure = fmaxf(0.f, pipe - lung_p);
      if (peep != 0.f) pipe *= 0.995f;
      steps = time + 1.0f;
      time = time + p.dt;
      ran_env = true;
    }
    // ---- emitted series                                          run_controller.py:128-141
    const float ts = p.dt * steps - p.run_dt;
    const int64_t o = (int64_t)t * p.N + n;
    if (b.timestamp_out) b.timestamp_out[o] = ts;
    if (b.u_in_out) b.u_in_out[o] = u_in;
    if (b.u_out_out) b.u_out_out[o] = u_out;
    if (b.pressure_out) b.pressure_out[o] = pr_emit;
    if (b.flow_out) b.flow_out[o] = p.env_flow;
    if (b.target_out) b.target_out[o] = interp_knots<NK>(p, kf, fmod_pos(ts, p.period));   // waveform.at(timestamps), :207
    bad |= !isfinite(pressure);
  }

  if (ran_env) target = interp_knots<NK>(p, kf, fmod_pos(time, p.period));   // state.target (:140)
  b.steps_io[n] = steps; b.time_io[n] = time; b.volume_io[n] = volume; b.pressure_io[n] = pressure;
  b.target_io[n] = target;
  b.obs_pressure_io[n] = obs_p; b.obs_time_io[n] = obs_t;
  b.pid_p_io[n] = cp; b.pid_i_io[n] = ci; b.pid_d_io[n] = cd; b.pid_time_io[n] = ctime; b.pid_dt_io[n] = cdt;
  b.pid_steps_io[n] = csteps;
  b.exp_time_io[n] = etime; b.exp_dt_io[n] = edt; b.exp_steps_io[n] = esteps;
  if (SIM == DLC_SIM_DELAY) {
    b.pipe_pressure_io[n] = pipe;
    for (int k = 0; k < p.delay; ++k) {
      int s = pos + k;
      s -= s >= p.delay ? p.delay : 0;
      b.in_history_io[n * p.delay + k] = hin[s];
      b.out_history_io[n * p.delay + k] = hout[s];
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
  if (p->sim == DLC_SIM_BALLOON) {
    if (p->nk == 7) lung_pid_kernel<DLC_SIM_BALLOON, 7><<<grid, 128, 0, st>>>(a);
    else lung_pid_kernel<DLC_SIM_BALLOON, 0><<<grid, 128, 0, st>>>(a);
  } else {
    if (p->nk == 7) lung_pid_kernel<DLC_SIM_DELAY, 7><<<grid, 128, 0, st>>>(a);
    else lung_pid_kernel<DLC_SIM_DELAY, 0><<<grid, 128, 0, st>>>(a);
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
