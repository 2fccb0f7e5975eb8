// The functional Env / Agent rollout of deluca/training/apg.py:39-68 on the classic envs, fused: per step
//     agent_state', action = agent(agent_state, obs)        deluca/agents/_pid.py:34-48 (generic PID on the observation
//                                                           vector), or linear state feedback u = -K x
//     env_state', obs'     = env(env_state, action)         deluca/envs/classic/*.py (the env consumes action[:m])
//     loss                 = loss_func(env_state, obs, action, env, t)   -- here: sum q (x - goal)^2 + sum r (u - goal_u)^2
// for N independent episodes, one thread per episode, all state in registers.  The kernel emits the three stacks of
// `Rollout` (agent states every `state_stride` steps, actions, losses) and, on request, d(sum_t loss)/d(gains) by
// forward-mode tangents carried next to the state: the agent leaves are 3 PID gains (or m*d feedback gains), so
// tangent propagation costs a few dozen registers and no checkpoints, and it yields what
// jax.value_and_grad(apg_loss) (apg.py:103-122,152-167) returns for those leaves.
#include "common.cuh"
#include "env_models.cuh"

namespace dlc {

struct EnvAgentArgs {
  DlcEnvAgentParams p;
  const float* gains;
  float *x_io, *agent_io, *loss_out, *action_out, *agent_stack_out, *x_stack_out;
  double* loss_sum_out;
  float* grad_out;
  int32_t* status_out;
};

template <int KIND, int AGENT, bool GRAD>
__global__ void __launch_bounds__(128) env_agent_kernel(const EnvAgentArgs a) {
  using Mdl = Model<KIND>;
  constexpr int D = Mdl::D, M = Mdl::M;
  constexpr int NG = AGENT == DLC_AGENT_PID ? 3 : M * D;      // gains (agent leaves)
  constexpr int DA = AGENT == DLC_AGENT_PID ? D : M;          // action vector length
  constexpr int NGA = GRAD ? NG : 1;
  const DlcEnvAgentParams& p = a.p;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const Mdl env(p.env);
  float g[NG];
  {
    const float* gp = a.gains + (p.gains_per_env ? n * NG : 0);
#pragma unroll
    for (int i = 0; i < NG; ++i) g[i] = gp[i];
  }
  float x[D], P[D], I[D], Dv[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    x[i] = a.x_io[n * D + i];
    if (AGENT == DLC_AGENT_PID) {
      P[i] = a.agent_io[(n * 3 + 0) * D + i];
      I[i] = a.agent_io[(n * 3 + 1) * D + i];
      Dv[i] = a.agent_io[(n * 3 + 2) * D + i];
    }
  }
  // tangents w.r.t. the gains
  float dx[D][NGA], dI[D][NGA], dD[D][NGA], dP[D][NGA], dloss[NGA];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      dloss[k] = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) { dx[i][k] = 0.f; dI[i][k] = 0.f; dD[i][k] = 0.f; dP[i][k] = 0.f; }
    }
  }
  const float decay = p.pid_decay;
  double lsum = 0.0;
  const int stride = p.state_stride > 0 ? p.state_stride : 0;
  const int64_t N = p.N;
  for (int t = 0; t < p.T; ++t) {
    float act[DA], dact[M][NGA];
    if (AGENT == DLC_AGENT_PID) {
      // P = obs ; I += decay (obs - I) ; D += decay (obs - P_prev - D) ; action = K_P P + K_I I + K_D D   _pid.py:40-48
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float obs = x[i];
        const float In = I[i] + decay * (obs - I[i]);
        const float Dn = Dv[i] + decay * ((obs - P[i]) - Dv[i]);
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < NG; ++k) {
            const float dIn = dI[i][k] + decay * (dx[i][k] - dI[i][k]);
            const float dDn = dD[i][k] + decay * ((dx[i][k] - dP[i][k]) - dD[i][k]);
            dI[i][k] = dIn; dD[i][k] = dDn; dP[i][k] = dx[i][k];
          }
        }
        P[i] = obs; I[i] = In; Dv[i] = Dn;
        act[i] = (g[0] * obs + g[1] * In) + g[2] * Dn;
      }
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < M; ++c)
#pragma unroll
          for (int k = 0; k < NG; ++k)
            dact[c][k] = (k == 0 ? P[c] : (k == 1 ? I[c] : Dv[c])) + ((g[0] * dP[c][k] + g[1] * dI[c][k]) + g[2] * dD[c][k]);
      }
    } else {
      // u = -K x (linear state feedback, deluca/agents/_lqr.py:62-72 on the env's state vector)
#pragma unroll
      for (int c = 0; c < M; ++c) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) s = fmaf(-g[c * D + j], x[j], s);
        act[c] = s;
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < NG; ++k) {
            float ds = (k / D == c) ? -x[k % D] : 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) ds = fmaf(-g[c * D + j], dx[j][k], ds);
            dact[c][k] = ds;
          }
        }
      }
    }
    // loss on the state the agent saw and the action it took                                 apg.py:54
    float loss = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) { const float e = x[i] - p.goal[i]; loss = fmaf(p.q[i] * e, e, loss); }
#pragma unroll
    for (int c = 0; c < M; ++c) { const float e = act[c] - p.goal_u[c]; loss = fmaf(p.r[c] * e, e, loss); }
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        float s = dloss[k];
#pragma unroll
        for (int i = 0; i < D; ++i) s = fmaf(2.f * p.q[i] * (x[i] - p.goal[i]), dx[i][k], s);
#pragma unroll
        for (int c = 0; c < M; ++c) s = fmaf(2.f * p.r[c] * (act[c] - p.goal_u[c]), dact[c][k], s);
        dloss[k] = s;
      }
    }
    lsum += (double)loss;
    if (a.loss_out) a.loss_out[(int64_t)t * N + n] = loss;
    if (a.action_out) {
#pragma unroll
      for (int i = 0; i < DA; ++i) a.action_out[((int64_t)t * N + n) * DA + i] = act[i];
    }
    // env step (the env consumes the first m components of the action)                       apg.py:53
    float xn[D];
    if (GRAD) {
      float Fx[D * D], Fu[D * M];
      env.jac(x, act, Fx, Fu);
      float dxn[D][NGA];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < D; ++j) s = fmaf(Fx[i * D + j], dx[j][k], s);
#pragma unroll
          for (int c = 0; c < M; ++c) s = fmaf(Fu[i * M + c], dact[c][k], s);
          dxn[i][k] = s;
        }
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int k = 0; k < NG; ++k) dx[i][k] = dxn[i][k];
    }
    env.step(x, act, xn);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
    if (stride && ((t + 1) % stride == 0)) {
      const int64_t ts = (t + 1) / stride - 1;
      if (a.x_stack_out) {
#pragma unroll
        for (int i = 0; i < D; ++i) a.x_stack_out[(ts * N + n) * D + i] = x[i];
      }
      if (AGENT == DLC_AGENT_PID && a.agent_stack_out) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          a.agent_stack_out[((ts * N + n) * 3 + 0) * D + i] = P[i];
          a.agent_stack_out[((ts * N + n) * 3 + 1) * D + i] = I[i];
          a.agent_stack_out[((ts * N + n) * 3 + 2) * D + i] = Dv[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    a.x_io[n * D + i] = x[i];
    if (AGENT == DLC_AGENT_PID) {
      a.agent_io[(n * 3 + 0) * D + i] = P[i];
      a.agent_io[(n * 3 + 1) * D + i] = I[i];
      a.agent_io[(n * 3 + 2) * D + i] = Dv[i];
    }
  }
  if (a.loss_sum_out) a.loss_sum_out[n] = lsum;
  if (GRAD && a.grad_out) {
#pragma unroll
    for (int k = 0; k < NG; ++k) a.grad_out[n * NG + k] = dloss[k];
  }
  if (a.status_out) a.status_out[n] = isfinite(lsum) ? 0 : DLC_STATUS_NONFINITE;
}

template <int KIND, int AGENT>
static void launch_env_agent(const EnvAgentArgs& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.p.N + 127) / 128);
  if (a.p.want_grad) env_agent_kernel<KIND, AGENT, true><<<grid, 128, 0, st>>>(a);
  else env_agent_kernel<KIND, AGENT, false><<<grid, 128, 0, st>>>(a);
}

}  // namespace dlc

using namespace dlc;

extern "C" int32_t dlc_env_agent_rollout_f32(const DlcEnvAgentParams* p, const float* gains, float* x_io,
                                             float* agent_io, float* loss_out, float* action_out,
                                             float* agent_stack_out, float* x_stack_out, double* loss_sum_out,
                                             float* grad_out, int32_t* status_out, void* stream) {
  if (!p) return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: params is NULL");
  if (p->N < 0 || p->T < 0) return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: negative N or T");
  if (p->env.kind < DLC_ENV_PENDULUM || p->env.kind > DLC_ENV_REACHER)
    return fail(DLC_ERR_SHAPE, "dlc_env_agent_rollout_f32: unknown env kind");
  if (p->agent != DLC_AGENT_PID && p->agent != DLC_AGENT_LINEAR)
    return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: unknown agent");
  if (p->state_stride < 0) return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: state_stride must be >= 0");
  if (p->N == 0) return DLC_OK;
  if (!gains || !x_io || (p->agent == DLC_AGENT_PID && !agent_io))
    return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: gains, x_io (and agent_io for PID) are required");
  if (p->want_grad && !grad_out) return fail(DLC_ERR_ARG, "dlc_env_agent_rollout_f32: want_grad needs grad_out");
  EnvAgentArgs a{};
  a.p = *p; a.gains = gains; a.x_io = x_io; a.agent_io = agent_io; a.loss_out = loss_out; a.action_out = action_out;
  a.agent_stack_out = agent_stack_out; a.x_stack_out = x_stack_out; a.loss_sum_out = loss_sum_out;
  a.grad_out = grad_out; a.status_out = status_out;
  cudaStream_t st = (cudaStream_t)stream;
#define EA(KIND_)                                                                     \
  if (p->env.kind == KIND_) {                                                         \
    if (p->agent == DLC_AGENT_PID) launch_env_agent<KIND_, DLC_AGENT_PID>(a, st);     \
    else launch_env_agent<KIND_, DLC_AGENT_LINEAR>(a, st);                            \
  }
  EA(DLC_ENV_PENDULUM) EA(DLC_ENV_CARTPOLE) EA(DLC_ENV_QUADROTOR) EA(DLC_ENV_REACHER)
#undef EA
  return cuda_status(cudaGetLastError(), "env_agent_kernel launch");
}
