// XLA FFI custom-call handlers over the C ABI of libdeluca_b200.so (include/deluca_b200.h).
//
// This is the binding BASELINE.json's north_star names: JAX arrays reach the kernels zero-copy as device pointers and
// the launch is enqueued on XLA's stream.  Each handler only unwraps buffers and attributes, copies donated state
// inputs into the result buffers (the C ABI updates state in place; XLA results are fresh buffers) and forwards.
//
// jaxlib ships the only header this needs, `xla/ffi/api/ffi.h` (header-only C++ API).  It is absent from this image
// (SURVEY F2), so the body is guarded: without the header the translation unit compiles to nothing and
// `make -C deluca_b200/csrc xla_ffi` reports that.  tests/test_host_logic.py compiles it against a small API-shaped
// stand-in of the header (tests/xla_ffi_stub) so that signatures and the binding chains are at least type-checked.
//
//   build:  g++ -std=c++17 -shared -fPIC -I$(python -c 'import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), "include"))')
//               -I/usr/local/cuda/include -Iinclude deluca_b200/csrc/xla_ffi/deluca_xla_ffi.cc -Ldeluca_b200 -ldeluca_b200 -lcudart
//               -o libdeluca_xla_ffi.so
//   use:    INTEGRATION.md section 2 (jax.ffi.register_ffi_target + jax.ffi.ffi_call)
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define DLC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef DLC_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstring>

#include "xla/ffi/api/ffi.h"
#include "deluca_b200.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status(int32_t rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, dlc_last_error());
}

template <typename Src, typename Dst>
void donate(cudaStream_t stream, const Src& src, Dst& dst) {   // state leaf in -> result leaf (updated in place below)
  cudaMemcpyAsync(dst->untyped_data(), src.untyped_data(), src.size_bytes(), cudaMemcpyDeviceToDevice, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// dlc_lds_rollout_f32: LDS + LQR / GPC for T steps (deluca/training/apg.py:39-100 on envs/_lds.py, agents/_gpc.py,
// agents/_lqr.py).  `W` may be an empty buffer: the env then draws its own disturbances (in-kernel Philox stream).
ffi::Error LdsRollout(cudaStream_t stream, ffi::Buffer<ffi::F32> A, ffi::Buffer<ffi::F32> B, ffi::Buffer<ffi::F32> K,
                      ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> M0, ffi::Buffer<ffi::F32> hist0,
                      ffi::Buffer<ffi::F32> xprev0, ffi::Buffer<ffi::F32> W, int32_t policy, int32_t T, int32_t H,
                      int32_t HH, int32_t t0, int32_t env_t0, float lr_scale, bool decay, float w_std, uint64_t seed,
                      int64_t env_offset, ffi::ResultBuffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> M,
                      ffi::ResultBuffer<ffi::F32> hist, ffi::ResultBuffer<ffi::F32> xprev,
                      ffi::ResultBuffer<ffi::F32> costs, ffi::ResultBuffer<ffi::F64> cost_sum,
                      ffi::ResultBuffer<ffi::S32> status_out) {
  DlcLdsParams p;
  std::memset(&p, 0, sizeof p);
  p.N = x0.dimensions()[0];
  p.d = static_cast<int32_t>(x0.dimensions()[1]);
  p.m = static_cast<int32_t>(B.dimensions().back());
  p.T = T; p.H = H; p.HH = HH; p.policy = policy; p.t0 = t0; p.env_t0 = env_t0;
  p.decay = decay; p.lr_scale = lr_scale; p.w_std = w_std; p.seed = seed; p.env_offset = env_offset;
  p.shared_dynamics = A.dimensions().size() == 2;
  p.traj_stride = 1;
  if (dlc_lds_workspace_bytes(&p) != 0)
    return ffi::Error(ffi::ErrorCode::kUnimplemented, "this (d, m, H, HH) runs the runtime-dims kernel, which needs a workspace operand");
  donate(stream, x0, x);
  const bool learn = policy != DLC_POLICY_LQR;
  if (learn) { donate(stream, M0, M); donate(stream, hist0, hist); donate(stream, xprev0, xprev); }
  const float* w = W.element_count() ? W.typed_data() : nullptr;
  return status(dlc_lds_rollout_f32(&p, A.typed_data(), B.typed_data(), K.typed_data(), w, /*u_in=*/nullptr,
                                    x->typed_data(), learn ? M->typed_data() : nullptr,
                                    learn ? hist->typed_data() : nullptr, learn ? xprev->typed_data() : nullptr,
                                    /*uprev_io=*/nullptr, nullptr, nullptr, nullptr, costs->typed_data(),
                                    cost_sum->typed_data(), nullptr, nullptr, status_out->typed_data(), nullptr, 0, stream));
}

// ---------------------------------------------------------------------------------------------------------------------
// dlc_lung_pid_rollout_f32: run_controller_scan (deluca/lung/utils/scripts/run_controller.py:144-220) for N breaths:
// lung PID + Expiratory on BalloonLung (sim = 0) or DelayLung (sim = 1).  The state operands are the batched leaves of
// env.reset(), controller.init(waveform) and expiratory.init(); `knots` = [kx(nk) | period | xp_in] and `sim_params`
// are the static fields, packed by the Python side exactly as deluca_b200/lung/_launch.py does.
ffi::Error LungPidRollout(cudaStream_t stream, ffi::Buffer<ffi::F32> K, ffi::Buffer<ffi::F32> kf,
                          ffi::Buffer<ffi::F32> env0 /*[5,N]: steps,time,volume,pressure,target*/,
                          ffi::Buffer<ffi::F32> obs0 /*[2,N]*/, ffi::Buffer<ffi::F32> pid0 /*[5,N]: p,i,d,time,dt*/,
                          ffi::Buffer<ffi::F32> exp0 /*[2,N]: time,dt*/, ffi::Buffer<ffi::F32> delay0 /*[1+2*delay,N] or empty*/,
                          ffi::Span<const float> knots, ffi::Span<const float> sim_params, int32_t T, int32_t sim,
                          int32_t leak, int32_t delay, ffi::ResultBuffer<ffi::F32> env, ffi::ResultBuffer<ffi::F32> obs,
                          ffi::ResultBuffer<ffi::F32> pid, ffi::ResultBuffer<ffi::F32> expi,
                          ffi::ResultBuffer<ffi::F32> delay_state, ffi::ResultBuffer<ffi::S32> steps,
                          ffi::ResultBuffer<ffi::F32> series /*[6,T,N]*/, ffi::ResultBuffer<ffi::S32> status_out) {
  const int64_t N = K.dimensions()[0];
  const size_t nk = knots.size() - 2;
  if (nk > DLC_MAX_KNOTS || sim_params.size() < 16) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad knots / sim_params");
  DlcLungParams p;
  std::memset(&p, 0, sizeof p);
  p.N = N; p.T = T; p.sim = sim; p.nk = static_cast<int32_t>(nk); p.kf_per_env = kf.dimensions().size() == 2;
  for (size_t i = 0; i < nk; ++i) p.kx[i] = knots[i];
  p.period = knots[nk]; p.xp_in = knots[nk + 1];
  const float* s = sim_params.begin();
  p.pid_decay = s[0]; p.default_dt = s[1]; p.run_dt = s[2]; p.env_flow = s[3];
  p.leak = leak; p.peep_valve = s[4]; p.PC = s[5]; p.P0 = s[6]; p.R = s[7]; p.dt = s[8]; p.min_volume = s[9];
  p.r0 = s[10]; p.r0_sq = s[11]; p.leak_coef = s[12];
  p.delay = delay; p.d_R = s[13]; p.d_C = s[14]; p.inertia = s[15]; p.control_gain = sim_params.size() > 16 ? s[16] : 0.f;
  p.mode = DLC_LUNG_MODE_CLOSED_LOOP;
  donate(stream, env0, env); donate(stream, obs0, obs); donate(stream, pid0, pid); donate(stream, exp0, expi);
  if (sim == DLC_SIM_DELAY) donate(stream, delay0, delay_state);
  cudaMemsetAsync(steps->untyped_data(), 0, steps->size_bytes(), stream);      // controller / expiratory step counters
  DlcLungBuffers b;
  std::memset(&b, 0, sizeof b);
  float *e = env->typed_data(), *o = obs->typed_data(), *c = pid->typed_data(), *x = expi->typed_data();
  b.K = K.typed_data(); b.kf = kf.typed_data();
  b.steps_io = e; b.time_io = e + N; b.volume_io = e + 2 * N; b.pressure_io = e + 3 * N; b.target_io = e + 4 * N;
  b.obs_pressure_io = o; b.obs_time_io = o + N;
  b.pid_p_io = c; b.pid_i_io = c + N; b.pid_d_io = c + 2 * N; b.pid_time_io = c + 3 * N; b.pid_dt_io = c + 4 * N;
  b.exp_time_io = x; b.exp_dt_io = x + N;
  b.pid_steps_io = steps->typed_data(); b.exp_steps_io = steps->typed_data() + N;
  if (sim == DLC_SIM_DELAY) {
    float* d = delay_state->typed_data();
    b.pipe_pressure_io = d; b.in_history_io = d + N; b.out_history_io = d + N + (int64_t)delay * N;
  }
  float* ts = series->typed_data();
  const int64_t TN = (int64_t)T * N;
  b.timestamp_out = ts; b.u_in_out = ts + TN; b.u_out_out = ts + 2 * TN; b.pressure_out = ts + 3 * TN;
  b.flow_out = ts + 4 * TN; b.target_out = ts + 5 * TN;
  b.status_out = status_out->typed_data();
  return status(dlc_lung_pid_rollout_f32(&p, &b, stream));
}

// ---------------------------------------------------------------------------------------------------------------------
// dlc_plan_f32: iLQR / iGPC_closed / iLC_closed (deluca/igpc/ilqr.py:27-93, igpc.py:129-179, ilc.py:23-69), the whole
// outer loop on the device.  `env_true` / `env_sim` = [kind | 8 static fields]; `cost` = [q(d) | goal(d) | r(m) | goal_u(m)].
ffi::Error Plan(cudaStream_t stream, ffi::Buffer<ffi::F32> U0, ffi::Buffer<ffi::F32> x0,
                ffi::Span<const float> env_true, ffi::Span<const float> env_sim, ffi::Span<const float> cost,
                int32_t T, int32_t algo, bool backtracking, bool fix_backtracking, int32_t n_alpha, int32_t w_is,
                float alpha0, float lr, ffi::ResultBuffer<ffi::F32> U, ffi::ResultBuffer<ffi::F32> X,
                ffi::ResultBuffer<ffi::F32> k, ffi::ResultBuffer<ffi::F32> Kg, ffi::ResultBuffer<ffi::F32> c,
                ffi::ResultBuffer<ffi::F32> alpha, ffi::ResultBuffer<ffi::S32> status_out,
                ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t N = U0.dimensions()[0];
  const int d = static_cast<int>(x0.dimensions()[1]), m = static_cast<int>(U0.dimensions()[2]);
  if (env_true.size() != 9 || env_sim.size() != 9 || cost.size() != size_t(2 * d + 2 * m) || d > DLC_PLAN_MAXD ||
      m > DLC_PLAN_MAXM)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bad env / cost attribute lengths");
  DlcPlanParams p;
  std::memset(&p, 0, sizeof p);
  p.N = N; p.H = static_cast<int32_t>(U0.dimensions()[1]); p.T = T;
  p.env_true.kind = static_cast<int32_t>(env_true[0]);
  p.env_sim.kind = static_cast<int32_t>(env_sim[0]);
  for (int i = 0; i < 8; ++i) { p.env_true.p[i] = env_true[1 + i]; p.env_sim.p[i] = env_sim[1 + i]; }
  for (int i = 0; i < d; ++i) { p.q[i] = cost[i]; p.goal[i] = cost[d + i]; }
  for (int i = 0; i < m; ++i) { p.r[i] = cost[2 * d + i]; p.goal_u[i] = cost[2 * d + m + i]; }
  p.algo = algo; p.backtracking = backtracking; p.fix_backtracking = fix_backtracking; p.n_alpha = n_alpha;
  p.w_is = w_is; p.alpha0 = alpha0; p.lr = lr;
  const size_t need = dlc_plan_workspace_bytes(&p);
  if (workspace->size_bytes() < need) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "workspace result is too small");
  donate(stream, U0, U);
  return status(dlc_plan_f32(&p, U->typed_data(), x0.typed_data(), X->typed_data(), k->typed_data(), Kg->typed_data(),
                             c->typed_data(), alpha->typed_data(), status_out->typed_data(), /*decision_out=*/nullptr,
                             /*cost_trace_out=*/nullptr, workspace->untyped_data(), workspace->size_bytes(), stream));
}

}  // namespace

using F32 = ffi::Buffer<ffi::F32>;

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    dlc_xla_lds_rollout, LdsRollout,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Attr<int32_t>("policy").Attr<int32_t>("T").Attr<int32_t>("H").Attr<int32_t>("HH").Attr<int32_t>("t0")
        .Attr<int32_t>("env_t0").Attr<float>("lr_scale").Attr<bool>("decay").Attr<float>("w_std")
        .Attr<uint64_t>("seed").Attr<int64_t>("env_offset")
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    dlc_xla_lung_pid_rollout, LungPidRollout,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Attr<ffi::Span<const float>>("knots").Attr<ffi::Span<const float>>("sim_params").Attr<int32_t>("T")
        .Attr<int32_t>("sim").Attr<int32_t>("leak").Attr<int32_t>("delay")
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::S32>>().Ret<F32>()
        .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    dlc_xla_plan, Plan,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F32>().Arg<F32>()
        .Attr<ffi::Span<const float>>("env_true").Attr<ffi::Span<const float>>("env_sim")
        .Attr<ffi::Span<const float>>("cost").Attr<int32_t>("T").Attr<int32_t>("algo").Attr<bool>("backtracking")
        .Attr<bool>("fix_backtracking").Attr<int32_t>("n_alpha").Attr<int32_t>("w_is").Attr<float>("alpha0")
        .Attr<float>("lr")
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::U8>>());

#else   // no jaxlib headers on this machine: nothing to compile (see the header comment)
extern "C" int dlc_xla_ffi_unavailable(void) { return 1; }
#endif
