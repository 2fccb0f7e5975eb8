/*
 * deluca_b200.h -- C ABI of the B200-native batched rollout library (libdeluca_b200.so).
 *
 * The reference (google/deluca v0.0.17) is pure Python on JAX and has no FFI of its own
 * (SURVEY.md F1).  Its "operator interface" for the hot path is a Python calling
 * convention: Env.__call__(state, action) / Agent.__call__(state, obs) stepped by
 * lax.scan and batched by jax.vmap (deluca/core.py:113-141, deluca/training/apg.py:39-100).
 * Each entry point below replaces the fused body of one such scan x vmap; the file:line
 * of the reference code it stands for is cited next to it.  INTEGRATION.md shows the
 * XLA-FFI / ctypes stubs a deluca maintainer would add on the reference side.
 *
 * Conventions (all entry points)
 *   - plain C: pointers, sizes, POD parameter structs.  No torch / JAX types.
 *   - every pointer is a DEVICE pointer unless its name ends in _host.
 *   - the caller allocates every buffer; the library allocates nothing, keeps no pointer
 *     after returning and only ENQUEUES work on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  No host synchronisation inside.
 *   - *_io buffers are updated in place, every other input is const.
 *   - return value: 0 = enqueued; < 0 = argument/shape error, nothing launched
 *     (DLC_ERR_*); > 0 = a cudaError_t.  dlc_last_error() returns a thread-local message.
 *   - arithmetic is fp32 (default JAX; x64 is never enabled in the reference).
 *   - batched leaves are env-major [N, ...] (the layout jax.vmap gives pytree leaves);
 *     per-step outputs are time-major [T, N] / [T, Ns, ...] so that a warp's stores for
 *     one step are contiguous.
 */
#ifndef DELUCA_B200_H_
#define DELUCA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLC_ABI_VERSION 2

#define DLC_OK 0
#define DLC_ERR_ARG (-1)         /* NULL where a buffer is required, negative size, ... */
#define DLC_ERR_SHAPE (-2)       /* dimension outside what the kernels support */
#define DLC_ERR_WORKSPACE (-3)   /* workspace too small / missing */
#define DLC_ERR_DEVICE (-4)      /* not an sm_100 device */

/* per-env status bits written by the kernels */
#define DLC_STATUS_NONFINITE 1   /* a NaN/Inf appeared in the env's state or cost */
#define DLC_STATUS_NOT_PD 2      /* Q_uu not positive definite in the Riccati backward pass */
#define DLC_STATUS_BAD_HINT 4    /* a caller's hint did not hold (DlcOfParams.flags: Phi != I, or |A^h B| > |B|): nothing computed, cost_sum = NaN */

int32_t dlc_abi_version(void);
/* sizeof() of the parameter structs as compiled, so a binding can verify its own layout:
 * which = 0 DlcLdsParams, 1 DlcLungParams, 2 DlcLungBuffers, 3 DlcPlanParams, 4 DlcSharedGpcParams, 5 DlcOfParams;
 * 0 for unknown */
size_t dlc_sizeof(int32_t which);
const char* dlc_last_error(void);
/* 0 if the current CUDA device can run this library (compute capability 10.x) */
int32_t dlc_check_device(void);

/* ------------------------------------------------------------------------------------
 * Counter-based Gaussian disturbance stream  (replaces GaussianDisturbance.__call__ =
 * jax.random.normal(key,(d,1))*std with a per-step split key, deluca/envs/_lds.py:35-41,
 * 102-106; Threefry cannot be reproduced without JAX).
 *   Philox4x32-10, key = (seed lo, seed hi), counter = (env, t, b/4, stream) -> Box-Muller.
 *   W[t, n, b] = std * z(env_offset + n, t0 + t, b).   Same device code as the in-kernel path
 *   of dlc_lds_rollout_f32 (W == NULL), so materialised and in-register streams agree bitwise.
 * ------------------------------------------------------------------------------------ */
int32_t dlc_gaussian_disturbance_f32(float* W /*[T,N,d]*/, int64_t N, int32_t T, int32_t d,
                                     uint64_t seed, int64_t env_offset, int32_t t0, float std,
                                     void* stream);

/* SinusDisturbance.__call__ (deluca/envs/_lds.py:44-52): W[t, n, i] = sin((t0 + t) * phases[i]) * amplitude for every
 * env n (the reference's phases ~ N(0,1) under PRNGKey(0); here the caller passes them, the mirror draws them from the
 * library stream with seed 0).  float32: the product t * phase and the sine are each rounded once. */
int32_t dlc_sinus_disturbance_f32(float* W /*[T,N,d]*/, int64_t N, int32_t T, int32_t d, const float* phases /*[d]*/,
                                  float amplitude, int32_t t0, void* stream);

/* ------------------------------------------------------------------------------------
 * LDS environment driven by LQR / GPC / BPC for T steps, fused.
 *
 * Replaces, per env and per step t:
 *     u_t   = agent(x_t)                GPC.__call__  deluca/agents/_gpc.py:130-183
 *                                       BPC.__call__  deluca/agents/_bpc.py:97-158
 *                                       LQR.__call__  deluca/agents/_lqr.py:62-72
 *     c_t   = quad_loss(x_t, u_t)       deluca/agents/_gpc.py:29-40
 *     x_t+1 = A x_t + B u_t + w_t       LDS.__call__  deluca/envs/_lds.py:102-111
 * and the scan x vmap around it (deluca/training/apg.py:39-100).
 * GPC's jit(grad(policy_loss)) (_gpc.py:109-128) is a hand-derived adjoint.
 * ------------------------------------------------------------------------------------ */
#define DLC_POLICY_LQR 0
#define DLC_POLICY_GPC 1
#define DLC_POLICY_BPC 2
#define DLC_POLICY_OPEN 3  /* u_t read from u_in[T,N,m]: LDS.__call__(u) on its own (_lds.py:102-111) */

#define DLC_MODE_CLOSED_LOOP 0 /* agent then env, T steps */
#define DLC_MODE_AGENT_ONLY 1  /* Agent.__call__ only: x_io is the observed state and is not advanced */

typedef struct DlcLdsParams {
  int64_t N;           /* envs in this call (this rank's shard) */
  int64_t env_offset;  /* global index of env 0 (keys the disturbance stream) */
  int32_t T;           /* steps */
  int32_t d, m;        /* state / action dim */
  int32_t H, HH;       /* GPC controller / system history (_gpc.py:53-54); BPC uses H only */
  int32_t policy;      /* DLC_POLICY_* */
  int32_t t0;          /* AGENT step counter at entry: lr decay (_gpc.py:161-162), BPC perturbation counter */
  int32_t decay;       /* GPC: lr = lr_scale/(t+1) (_gpc.py:161-162); BPC: /(t^0.75+1) (_bpc.py:128-129) */
  int32_t noise_uses_prev_action; /* 0 = reference as written (B u_t, _gpc.py:141-142,155); 1 = B u_{t-1} */
  int32_t shared_dynamics;        /* 1: A,B,K are [d,d],[d,m],[m,d] shared by all envs; 0: per env [N,...] */
  int32_t traj_stride; /* env n is sampled into x_out/u_out iff n % traj_stride == 0 (>=1) */
  float lr_scale;      /* _gpc.py:55 */
  float w_std;         /* std of the in-kernel Gaussian disturbance, used when W == NULL */
  float bpc_delta;     /* _bpc.py:53 */
  int32_t kernel_variant; /* 0 = auto; 1 = force the runtime-dims kernel (testing); 2..7 = force a launch shape
                             of the lane-split kernel: (128 thr x 3 CTA/SM), (64x5), (64x7), (128x2), (32x7), (32x8);
                             8..12 = launch shapes of the packed-FFMA2 kernel (128x2, 128x3, 64x5, 32x9, 256x1);
                             13..16 = launch shapes of the quad-split kernel for H = HH even (128x2, 64x4, 256x1,
                             128x1); 17 = the same kernel with incremental window products (measured slower, kept for
                             cross-checks) */
  int32_t mode;        /* DLC_MODE_*: which half of the loop runs (single-step Env/Agent calls use 1 and 2) */
  uint64_t seed;       /* disturbance / perturbation stream seed */
  int32_t env_t0;      /* ENV step counter at entry (LDS.t, _lds.py:109): counter of the in-kernel disturbance stream.
                          Separate from t0 so that a fresh agent on a continued env (or chunked LQR rollouts) neither
                          replays nor skips noise (ABI version 2) */
  int32_t pad_;
} DlcLdsParams;

/* bytes of `workspace` dlc_lds_rollout_f32 needs for these params: 0 when a compile-time-shape kernel takes the call,
 * otherwise the runtime-dims kernel's scratch vectors plus an [element][env] copy of every per-env array. */
size_t dlc_lds_workspace_bytes(const DlcLdsParams* p);

int32_t dlc_lds_rollout_f32(
    const DlcLdsParams* p,
    const float* A,       /* [N,d,d] or [d,d] */
    const float* B,       /* [N,d,m] or [d,m] */
    const float* K,       /* [N,m,d] or [m,d] */
    const float* W,       /* [T,N,d] disturbances, or NULL: generate in-kernel (Philox, w_std) */
    const float* u_in,    /* [T,N,m] actions for DLC_POLICY_OPEN (NULL otherwise) */
    float* x_io,          /* [N,d]   env state: x_t0 in, x_t0+T out */
    float* M_io,          /* [N,H,m,d]     GPC/BPC M        (NULL for LQR) */
    float* hist_io,       /* [N,H+HH,d]    noise_history, oldest -> newest (BPC: [N,H,d]) (NULL for LQR) */
    float* xprev_io,      /* [N,d]   agent's copy of the previous state (_gpc.py:101,167) (NULL for LQR) */
    float* uprev_io,      /* [N,m]   previous action; required iff noise_uses_prev_action */
    float* bpc_eps_io,    /* [N,H,H,m,d]   BPC perturbation ring (_bpc.py:80) (NULL otherwise) */
    const float* bpc_eps_new, /* [T,N,H,m,d] fresh unit-norm perturbations, or NULL: in-kernel Philox */
    float* bpc_cost_io,   /* [N] cost handed to BPC at the next call (_bpc.py:97): the cost realised one step
                             earlier, 0 before the first step (NULL otherwise) */
    float* cost_out,      /* [T,N]   per-step realised cost, or NULL */
    double* cost_sum_out, /* [N]     sum over the T steps (fp64 accumulator), or NULL */
    float* x_out,         /* [T,Ns,d] sampled x_t (the state the agent saw), Ns = ceil(N/traj_stride), or NULL */
    float* u_out,         /* [T,Ns,m] sampled u_t, or NULL */
    int32_t* status_out,  /* [N]     DLC_STATUS_* bit flags, or NULL */
    void* workspace, size_t workspace_bytes, void* stream);

/* The same closed-loop LDS + GPC rollout with a caller's cost_fn in the controller's update: GPC.__init__ takes
 * cost_fn (deluca/agents/_gpc.py:52,76) and learns from policy_loss = cost_fn(y_f, u_f) (:125,128).  The family taken here
 * is the diagonal quadratic cost_fn(x, u) = sum_k q_k x_k^2 + sum_c r_c u_c^2: cost_q [d] and cost_r [m] are device
 * arrays shared by all envs (NULL = ones, i.e. quad_loss), seeding the hand adjoint with dL/du_f = 2 r (.) u_f and
 * lam = 2 q (.) y_f - K' dL/du_f.  cost_out / cost_sum_out remain the realised quad_loss(x_t, u_t) (a re-weighted
 * report is the caller's loss_func, deluca/training/apg.py:54).  p->policy must be DLC_POLICY_GPC.  Runs the lane-group
 * kernel (closed loop, d <= 64, m <= 64) or the runtime-dims kernel (any shape, single Agent.__call__ steps), which
 * needs dlc_lds_workspace_bytes() of workspace -- allocate it as if p->kernel_variant were 1.  Other arguments as above. */
int32_t dlc_lds_rollout_cost_f32(const DlcLdsParams* p, const float* A, const float* B, const float* K,
                                 const float* W, float* x_io, float* M_io, float* hist_io, float* xprev_io,
                                 float* uprev_io, const float* cost_q, const float* cost_r, float* cost_out,
                                 double* cost_sum_out, float* x_out, float* u_out, int32_t* status_out,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Reverse mode through the LDS rollout: value and gradient of
 *     loss_n = sum_{t<T} quad_loss(x_t, u_t),  u_t = -K x_t,  x_{t+1} = A x_t + B u_t + w_t
 * with respect to the linear policy K (and optionally x_0) -- what jax.value_and_grad(apg_loss)
 * (deluca/training/apg.py:103-122,152-167) computes for a linear state-feedback agent on LDS
 * (deluca/envs/_lds.py:102-111, deluca/agents/_lqr.py:62-72, quad_loss deluca/agents/_gpc.py:29-40).
 * The forward pass checkpoints x_t in the workspace ([T][d][N] floats); the backward pass is the hand adjoint
 *     g_u = 2 u_t + B' lam_{t+1};  dK -= g_u x_t';  lam_t = 2 x_t + A' lam_{t+1} - K' g_u.
 * d <= 16, m <= 8.  Disturbances: W[T,N,d] or the library's Philox stream (W == NULL), as in dlc_lds_rollout_f32.
 * ------------------------------------------------------------------------------------ */
size_t dlc_lds_lqr_grad_workspace_bytes(int64_t N, int32_t T, int32_t d);
int32_t dlc_lds_lqr_grad_f32(int64_t N, int32_t T, int32_t d, int32_t m, int32_t shared_dynamics,
                             const float* A, const float* B, const float* K, /* [N,...] or unbatched */
                             const float* x0,      /* [N,d] */
                             const float* W,       /* [T,N,d] or NULL */
                             uint64_t seed, int64_t env_offset, int32_t t0, float w_std,
                             float* loss_out,      /* [N] */
                             float* gradK_out,     /* [N,m,d] d loss_n / d K (per env, also when K is shared) */
                             float* gradx0_out,    /* [N,d] or NULL */
                             double* sums_out,     /* [1 + m*d] += (sum loss, sum dK) or NULL; caller zeroes */
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Ventilator episode: lung PID + Expiratory controllers driving BalloonLung / DelayLung, fused.
 *
 * Replaces lax.scan(loop_over_tt) of run_controller_scan
 * (deluca/lung/utils/scripts/run_controller.py:111-141,144-220) batched over N breaths, i.e. per step
 *     u_in  = PID(c_state, obs)          deluca/lung/controllers/_pid.py:77-102
 *     u_out = Expiratory(e_state, obs)   deluca/lung/controllers/_expiratory.py:35-45
 *     state, obs = env(state,(u_in,u_out))   deluca/lung/envs/_balloon_lung.py:100-146
 *                                            deluca/lung/envs/_delay_lung.py:75-133
 * with BreathWaveform.at / is_in (deluca/lung/core.py:54-66) evaluated from a host-built knot table
 * (the sorted, period-padded table jnp.interp(..., period=) builds; float32 knots reproduce the
 * reference's 3 - 1e-8 == 3 collapse, SURVEY A-9).  All reference quirks are kept: steps = time + 1,
 * timestamp = dt*steps - dt, DelayLung's `time < delay` and pre-dynamics observation, flow = env.flow.
 * ------------------------------------------------------------------------------------ */
#define DLC_SIM_BALLOON 0
#define DLC_SIM_DELAY 1
#define DLC_LUNG_MODE_CLOSED_LOOP 0 /* PID + Expiratory + simulator (run_controller_scan) */
#define DLC_LUNG_MODE_CTRL_ONLY 1   /* Controller.__call__ only: obs_*_io is the observation, simulator untouched */
#define DLC_LUNG_MODE_ENV_ONLY 2    /* LungEnv.__call__ only: actions read from u_in_in/u_out_in [T,N] */
#define DLC_MAX_KNOTS 16
#define DLC_MAX_DELAY 64

typedef struct DlcLungParams {
  int64_t N;                 /* breaths (envs) */
  int32_t T;                 /* steps */
  int32_t sim;               /* DLC_SIM_* */
  int32_t nk;                /* knots in kx / per-env kf rows (<= DLC_MAX_KNOTS) */
  int32_t kf_per_env;        /* 1: kf is [N,nk]; 0: kf is [nk], shared */
  float kx[DLC_MAX_KNOTS];   /* knot abscissae, ascending, period-padded */
  float period;              /* 60 / bpm */
  float xp_in;               /* xp[in_bounds[1]]: is_in(t) = (t mod period) <= xp_in */
  float pid_decay;           /* dt / (dt + RC), _pid.py:83 */
  float default_dt;          /* DEFAULT_DT floor of the controllers' dt bookkeeping */
  float run_dt;              /* run_controller_scan's dt (timestamp = env.time(state) - dt) */
  float env_flow;            /* the env's static `flow` field, emitted as the flow series */
  /* BalloonLung statics (_balloon_lung.py:61-73); r0 = (3 min_volume / 4 pi)^(1/3) from the host */
  int32_t leak;
  float peep_valve, PC, P0, R, dt, min_volume, r0, r0_sq, leak_coef /* dt/(5+dt) */;
  /* DelayLung statics (_delay_lung.py:41-50) */
  int32_t delay;
  float d_R, d_C, inertia, control_gain;
  int32_t mode;              /* DLC_LUNG_MODE_* */
  int32_t kernel_variant;    /* 0 = auto; 1 = one breath per thread; 2 = two per thread wherever that instance applies
                                (even N, 8-byte aligned arrays, default waveform table, closed loop, all six series)
                                regardless of T.  The two produce identical numbers; 1 / 2 are for tests and
                                measurements */
} DlcLungParams;

typedef struct DlcLungBuffers {
  const float* K;            /* [N,3] PID gains (the Dense(1, no bias) kernel) */
  const float* kf;           /* [N,nk] or [nk] knot ordinates */
  const float *u_in_in, *u_out_in; /* [T,N] actions, DLC_LUNG_MODE_ENV_ONLY only */
  /* simulator state, one array per pytree leaf (SimulatorState), all [N] */
  float *steps_io, *time_io, *volume_io, *pressure_io, *target_io;
  float *pipe_pressure_io;   /* DelayLung only */
  float *in_history_io, *out_history_io; /* DelayLung only, [N,delay], index 0 = newest */
  float *obs_pressure_io, *obs_time_io;  /* Observation */
  /* PIDControllerState / ControllerState leaves, [N] */
  float *pid_p_io, *pid_i_io, *pid_d_io, *pid_time_io, *pid_dt_io;
  int32_t* pid_steps_io;
  float *exp_time_io, *exp_dt_io;
  int32_t* exp_steps_io;
  /* per-step series, each [T,N] or NULL */
  float *timestamp_out, *u_in_out, *u_out_out, *pressure_out, *flow_out, *target_out;
  int32_t* status_out;       /* [N] DLC_STATUS_* or NULL */
} DlcLungBuffers;

int32_t dlc_lung_pid_rollout_f32(const DlcLungParams* p, const DlcLungBuffers* b, void* stream);

/* ------------------------------------------------------------------------------------
 * Reverse mode through the ventilator episode: value and gradient of
 * train_controller.rollout (deluca/lung/utils/scripts/train_controller.py:33-67) with respect to the PID
 * gains, i.e. what jax.value_and_grad(rollout_parallel) (:151-153) computes for the lung PID controller on
 * BalloonLung (deluca/lung/envs/_balloon_lung.py:100-146) or DelayLung (_delay_lung.py:75-133).  Every breath starts
 * from sim.reset() / controller.init(waveform) as in the reference;
 *     loss_n = sum_i [u_out_i == 0] * loss_fn(waveform.at(tt[i]), pressure_i),  loss_fn = |x - y| or (x - y)^2.
 * Default: forward mode (the three gains' tangents ride along the episode, no workspace).  With
 * DLC_LUNG_GRAD_REVERSE=1 (BalloonLung only, kept for cross-checks) the forward pass stores six floats per step in the
 * workspace and the backward pass is a hand-derived adjoint.  Both are held to the reference's own
 * jax.value_and_grad(rollout_parallel) (tests/golden/ref_lung_train.npz).  p->T = len(tt); for DelayLung pass
 * reset_pressure = 0 (its reset() has no normalized peep, :62-74).
 * ------------------------------------------------------------------------------------ */
#define DLC_LOSS_L1 0
#define DLC_LOSS_L2 1
size_t dlc_lung_pid_grad_workspace_bytes(const DlcLungParams* p);
int32_t dlc_lung_pid_grad_f32(const DlcLungParams* p,
                              const float* K,        /* [N,3] PID gains */
                              const float* kf,       /* [N,nk] or [nk] knot ordinates (p->kf_per_env) */
                              const float* tt,       /* [T] loss time stamps (jnp.linspace(0, duration, T)) */
                              int32_t loss_kind,     /* DLC_LOSS_* */
                              float noise,           /* use_noise * (mean + std * z): added to the pressure each step */
                              float reset_pressure,  /* sim.reset_normalized_peep */
                              float* loss_out,       /* [N] */
                              float* gradK_out,      /* [N,3] d loss_n / d K_n */
                              double* sums_out,      /* [4] += (sum loss, sum dK0, sum dK1, sum dK2) or NULL; caller zeroes */
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Generic leaky PID agent (deluca/agents/_pid.py:20-48), T steps over a given error series:
 *     P = e; I += decay (e - I); D += decay (e - P_prev - D); u = K_P P + K_I I + K_D D,
 *     decay = dt / (dt + RC) (:30-32).
 * ------------------------------------------------------------------------------------ */
int32_t dlc_pid_f32(int64_t N, int32_t T, float decay,
                    const float* K_P, const float* K_I, const float* K_D, /* [N] gains (pytree leaves) */
                    const float* err,                                   /* [T,N] observations (= errors) */
                    float* P_io, float* I_io, float* D_io,              /* [N] PIDState */
                    float* u_out,                                       /* [T,N] actions */
                    void* stream);

/* ------------------------------------------------------------------------------------
 * deluca/igpc: trajectory rollout, linearisation, Riccati backward pass and the iLQR / iGPC / iLC
 * outer loops on the classic envs, batched over N independent trajectories.
 *
 *   dlc_env_rollout_f32        igpc.rollout           deluca/igpc/rollout.py:22-65
 *   dlc_linearize_f32          compute_ders           deluca/igpc/rollout.py:68-99   (closed-form Jacobians
 *                                                     of the env step instead of jax.jacfwd / jax.grad)
 *   dlc_riccati_backward_f32   LQR                    deluca/igpc/lqr_solver.py:40-86 (host NumPy there)
 *   dlc_igpc_rollout_f32       GPC_rollout            deluca/igpc/igpc.py:63-126 (+ counterfact_loss :22-60,
 *                                                     hand adjoint instead of jit(grad))
 *   dlc_plan_f32               iLQR / iGPC_closed / iLC_closed   ilqr.py:27-93, igpc.py:129-179, ilc.py:23-69
 *                              the whole outer loop incl. backtracking stays on the device; the accept
 *                              order of the reference's sequential `if cC <= c: break` is preserved.
 *                              When an iteration has 2..16 distinct candidates (and N <= 65,536), an 8-lane group per
 *                              trajectory evaluates them side by side (in rounds of 8) and adopts the first
 *                              accepted one (identical results, 1-2 rollouts of latency instead of up to
 *                              n_alpha; the workspace then holds 8 candidate buffers per trajectory).
 * Envs: Pendulum (deluca/envs/classic/_pendulum.py:54-73, as written) and Cartpole
 * (_cartpole.py:60-89 with the intent-fixing waiver SURVEY A-7) and PlanarQuadrotor (8(f) rank 2).
 * Cost: c(x,u) = sum q_i (x_i-goal_i)^2 + sum r_j (u_j-goal_u_j)^2 (the reference ships no cost for this path).
 * Layouts are the vmapped reference shapes: X[N,H+1,d], U[N,H,m], k[N,H,m], K[N,H,m,d].
 * ------------------------------------------------------------------------------------ */
#define DLC_ENV_PENDULUM 0 /* d = 3, m = 1; params: m, l, g, max_torque, dt */
#define DLC_ENV_CARTPOLE 1 /* d = 4, m = 1; params: m, M, l, g, dt, offset */
#define DLC_ENV_QUADROTOR 2 /* PlanarQuadrotor (deluca/envs/classic/_planar_quadrotor.py:51-104): d = 6, m = 2;
                               params: m, l, g, dt, wind, wind_kind (0 = dissipative, 1 = constant) */
#define DLC_ENV_REACHER 3 /* Reacher (deluca/envs/classic/_reacher.py:56-141): d = 6, m = 2;
                             params: m1, m2, l1, l2, g, dt, goal_coord[0], goal_coord[1] */
#define DLC_PLAN_MAXD 8
#define DLC_PLAN_MAXM 4
#define DLC_ALGO_ILQR 0
#define DLC_ALGO_IGPC 1
#define DLC_ALGO_ILC 2
#define DLC_W_DE 0
#define DLC_W_DEDE 1
#define DLC_W_DELIN 2

typedef struct DlcEnvSpec {
  int32_t kind;        /* DLC_ENV_* */
  float p[8];          /* static fields in the order listed above */
} DlcEnvSpec;

typedef struct DlcPlanParams {
  int64_t N;           /* trajectories */
  int32_t H;           /* env.H: steps per trajectory */
  int32_t T;           /* outer iterations (dlc_plan_f32) */
  DlcEnvSpec env_true; /* the env rolled out */
  DlcEnvSpec env_sim;  /* the model that is linearised (== env_true for iLQR) */
  float q[DLC_PLAN_MAXD], goal[DLC_PLAN_MAXD], r[DLC_PLAN_MAXM]; /* quadratic cost */
  float goal_u[DLC_PLAN_MAXM]; /* action target of the cost (e.g. PlanarQuadrotor.goal_action = hover thrust) */
  int32_t algo;        /* DLC_ALGO_* */
  int32_t backtracking;
  int32_t fix_backtracking; /* iLQR only: 0 = as written (candidates use alpha, ilqr.py:58-67), 1 = alphaC */
  int32_t n_alpha;     /* candidates per iteration: 10 (iLQR, iLC), 6 (iGPC) */
  int32_t w_is;        /* DLC_W_* (iGPC) */
  float alpha0;        /* iLQR alpha = 0.5; iGPC / iLC ref_alpha = 1.0 */
  float lr;            /* iGPC ref_lr */
} DlcPlanParams;

int32_t dlc_env_rollout_f32(const DlcPlanParams* p, int32_t use_sim_env,
                            const float* U_old,  /* [N,H,m] */
                            const float* k,      /* [N,H,m] or NULL: open loop (U = U_old) */
                            const float* K,      /* [N,H,m,d] */
                            const float* X_old,  /* [N,H+1,d] */
                            const float* alpha,  /* [N] or NULL (= 1.0) */
                            const float* x0,     /* [N,d] start states */
                            float* X_out, float* U_out, float* cost_out /* [N] */, void* stream);

int32_t dlc_linearize_f32(const DlcPlanParams* p, int32_t use_sim_env, const float* X, const float* U,
                          float* D_out /*[N,H,d]*/, float* Fx_out /*[N,H,d,d]*/, float* Fu_out /*[N,H,d,m]*/,
                          float* cx_out /*[N,H,d]*/, float* cu_out /*[N,H,m]*/, float* cxx_out /*[N,H,d,d]*/,
                          float* cuu_out /*[N,H,m,m]*/, void* stream);

int32_t dlc_riccati_backward_f32(int64_t N, int32_t H, int32_t d, int32_t m, const float* Fx, const float* Fu,
                                 const float* cx, const float* cu, const float* cxx, const float* cuu,
                                 float* k_out, float* K_out, int32_t* status_out /*[N] or NULL*/, void* stream);

/* LQG (deluca/igpc/lqr_solver.py:89-154) with isPD_and_invert (:157-163) and the increasing branch of
 * lambda_adjust (:21-37) on the device: regularised Riccati pass (reg_type 0 = "Q": Q_uu + lambda I; 1 = "V":
 * Q_uu + lambda f_u'f_u and Q_ux + lambda f_u'f_x), Cholesky PD test per step, up to 9 attempts with lambda raised
 * after each failure.  lambda_io / multiplier_io [N] carry reg_lambda and multiplier in and the values the
 * reference would return out; gains_out [N,2] = (gains_lin, gains_quad) or NaN when no lambda was found
 * (reference: None, None); status_out DLC_STATUS_NOT_PD in that case.  cux [N,H,m,d]. */
int32_t dlc_lqg_backward_f32(int64_t N, int32_t H, int32_t d, int32_t m, const float* Fx, const float* Fu,
                             const float* cx, const float* cu, const float* cxx, const float* cuu,
                             const float* cux, float* lambda_io, float* multiplier_io, float factor,
                             float lambda_min, int32_t reg_type, float* k_out, float* K_out,
                             float* gains_out /*[N,2] or NULL*/, int32_t* status_out /*[N] or NULL*/, void* stream);

int32_t dlc_igpc_rollout_f32(const DlcPlanParams* p, const float* U_old, const float* k, const float* K,
                             const float* X_old, const float* alpha /*[N]*/, const float* x0,
                             float* X_out, float* U_out, float* cost_out, void* stream);

size_t dlc_plan_workspace_bytes(const DlcPlanParams* p);
int32_t dlc_plan_f32(const DlcPlanParams* p,
                     float* U_io,        /* [N,H,m] initial controls in, final controls out */
                     const float* x0,    /* [N,d] */
                     float* X_out,       /* [N,H+1,d] */
                     float* k_out, float* K_out, /* last gains ([N,H,m], [N,H,m,d]) */
                     float* cost_out,    /* [N] */
                     float* alpha_out,   /* [N] */
                     int32_t* status_out,/* [N] or NULL */
                     int32_t* decision_out, /* [T,N] or NULL: per outer iteration the index j of the accepted
                                               backtracking candidate (alphaC = 1.1 alpha 1.1^(-j^2); ilqr.py:55-73,
                                               igpc.py:159-170, ilc.py:49-60), -1 when none passed */
                     float* cost_trace_out, /* [T,N] or NULL: the trajectory cost after each outer iteration */
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * The functional Env / Agent rollout on the classic envs: apg_rollout x vmap (deluca/training/apg.py:39-100) with
 *     agent = generic PID acting on the observation vector (deluca/agents/_pid.py:20-48; the env consumes
 *             action[:m], as Pendulum does with action[0], _pendulum.py:65) or linear state feedback u = -K x
 *             (deluca/agents/_lqr.py:62-72 on the env's state vector),
 *     env   = Pendulum / Cartpole (intent-fixed, SURVEY A-7) / PlanarQuadrotor / Reacher,
 *     loss  = sum_i q_i (x_i - goal_i)^2 + sum_j r_j (u_j - goal_u_j)^2 on (env_state.arr, action[:m]) of the state
 *             the agent saw (apg.py:54) -- the quadratic family; arbitrary Python loss functions have no CUDA path.
 * Emits Rollout(agent_states, actions, losses) (apg.py:68): losses [T,N], actions [T,N,da] (da = d for PID, m for linear
 * feedback), agent / env state stacks every `state_stride` steps, and -- with want_grad -- d(sum_t loss)/d(gains) per env
 * by forward-mode tangents: what jax.value_and_grad(apg_loss) (apg.py:103-122,152-167) yields for the agent's leaves.
 * ------------------------------------------------------------------------------------ */
#define DLC_AGENT_PID 0
#define DLC_AGENT_LINEAR 1

typedef struct DlcEnvAgentParams {
  int64_t N;           /* episodes */
  int32_t T;           /* steps (unroll_length) */
  int32_t agent;       /* DLC_AGENT_* */
  DlcEnvSpec env;
  int32_t gains_per_env; /* 1: gains is [N,ng]; 0: [ng] shared (ng = 3 for PID: K_P, K_I, K_D; m*d for linear feedback) */
  int32_t state_stride;  /* > 0: stack the agent / env state after every state_stride-th step; 0: final state only */
  int32_t want_grad;
  float pid_decay;     /* dt / (dt + RC), _pid.py:37-38 */
  float q[DLC_PLAN_MAXD], goal[DLC_PLAN_MAXD], r[DLC_PLAN_MAXM], goal_u[DLC_PLAN_MAXM];
} DlcEnvAgentParams;

int32_t dlc_env_agent_rollout_f32(const DlcEnvAgentParams* p,
                                  const float* gains,       /* [N,ng] or [ng] */
                                  float* x_io,              /* [N,d] env state (arr): start in, final out */
                                  float* agent_io,          /* PID: [N,3,d] (P, I, D); NULL for linear feedback */
                                  float* loss_out,          /* [T,N] or NULL */
                                  float* action_out,        /* [T,N,da] or NULL */
                                  float* agent_stack_out,   /* [T/state_stride,N,3,d] or NULL */
                                  float* x_stack_out,       /* [T/state_stride,N,d] or NULL */
                                  double* loss_sum_out,     /* [N] or NULL */
                                  float* grad_out,          /* [N,ng] d(sum_t loss)/d(gains), required with want_grad */
                                  int32_t* status_out,      /* [N] or NULL */
                                  void* stream);

/* ------------------------------------------------------------------------------------
 * Shared-dynamics LDS + LQR rollout on the tensor cores (tcgen05, TF32 x3 split, fp32 accumulate in TMEM).
 * When A, B, K are shared by all envs (jax.vmap in_axes=None) the batch is the N dimension of a GEMM:
 *     [x_{t+1} - w_t ; u_t] = [A - B K ; -K] X_t,   cost_t = x_t'x_t + u_t'u_t
 * Replaces LDS.__call__ (deluca/envs/_lds.py:102-111) + LQR.__call__ (deluca/agents/_lqr.py:62-72) +
 * quad_loss (deluca/agents/_gpc.py:29-40) under lax.scan x vmap.  Shapes: d % 16 == 0, 16 <= d <= 256,
 * d + m <= 384.  dlc_lds_shared_pack_f32 forms [A - B K ; -K] (fp64 accumulate), splits it into TF32
 * hi + lo and lays it out in the kernel's staging order; G_packed holds dlc_lds_shared_packed_floats(d, m)
 * floats and can be reused across rollouts.
 * ------------------------------------------------------------------------------------ */
size_t dlc_lds_shared_packed_floats(int32_t d, int32_t m);
int32_t dlc_lds_shared_pack_f32(const float* A /*[d,d]*/, const float* B /*[d,m]*/, const float* K /*[m,d]*/,
                                int32_t d, int32_t m, float* G_packed, void* stream);
int32_t dlc_lds_shared_lqr_rollout_f32(int64_t N, int32_t T, int32_t d, int32_t m, const float* G_packed,
                                       const float* W /*[T,N,d] or NULL (no disturbance)*/,
                                       float* x_io /*[N,d]*/, float* cost_out /*[T,N] or NULL*/,
                                       double* cost_sum_out /*[N] or NULL*/, int32_t* status_out /*[N] or NULL*/,
                                       void* stream);

/* ------------------------------------------------------------------------------------
 * Shared-dynamics LDS + GPC rollout with a per-env policy tensor M (BASELINE config 5: d = 256, m = 64,
 * H = HH = 16, M[H,m,d] = 1 MiB per env).  A, B, K are shared by all envs (jax.vmap in_axes=None); M, the noise
 * history, the state and the disturbances are per env.
 * Replaces GPC.__call__ = get_action + update with jit(grad(policy_loss)) (deluca/agents/_gpc.py:109-183),
 * LDS.__call__ (deluca/envs/_lds.py:102-111) and quad_loss (_gpc.py:29-40) under lax.scan x vmap, reference
 * semantics as written (the noise estimate subtracts B u_t with the action just computed, _gpc.py:141-142,155).
 * Two kernels alternate per step (2 T + 1 launches per call):
 *   gpc_m_stream_kernel      M_t = M_{t-1} - lr g (x) hist  and the H window products  V_t[h] = sum_i M_t[i] w[h+i]
 *                            in one pass over M (read once + written once per env-step: HBM-bound);
 *   gpc_shared_chain_kernel  env step, realised cost, counterfactual chain and its adjoint as dense FP32 GEMMs
 *                            [d+m x d] x [d x 64 envs] with the shared operators.
 * Built for (d, m, H = HH) in {(256, 64, 16), (64, 32, 4)}; other shapes -> DLC_ERR_SHAPE (the per-env kernels of
 * dlc_lds_rollout_f32 take any shape with shared_dynamics = 1).
 * ------------------------------------------------------------------------------------ */
typedef struct DlcSharedGpcParams {
  int64_t N;       /* envs in this call */
  int32_t T;       /* steps */
  int32_t d, m;    /* state / action dim */
  int32_t H, HH;   /* GPC controller / system history; HH must equal H */
  int32_t t0;      /* agent step counter at entry (lr decay) */
  int32_t decay;   /* lr = lr_scale/(t+1) (_gpc.py:161-162) */
  float lr_scale;  /* _gpc.py:55 */
  int32_t reserved;
} DlcSharedGpcParams;

size_t dlc_lds_shared_gpc_workspace_bytes(const DlcSharedGpcParams* p);
int32_t dlc_lds_shared_gpc_rollout_f32(const DlcSharedGpcParams* p,
                                       const float* A /*[d,d]*/, const float* B /*[d,m]*/, const float* K /*[m,d]*/,
                                       const float* W /*[T,N,d] or NULL (no disturbance)*/,
                                       float* x_io /*[N,d]*/, float* M_io /*[N,H,m,d]*/,
                                       float* hist_io /*[N,2H,d] oldest -> newest, or NULL = zeros, not returned*/,
                                       float* xprev_io /*[N,d] agent's previous state, or NULL = zeros, not returned*/,
                                       float* cost_out /*[T,N] or NULL*/, double* cost_sum_out /*[N] or NULL*/,
                                       int32_t* status_out /*[N] or NULL*/,
                                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Partially observed LDS (y = C x) driven by an output-feedback controller for T steps, fused
 * (SURVEY 8(f) rank 3).  Replaces, per env and per step,
 *     y_t   = C x_t                                   LDS.__call__   deluca/envs/_lds.py:102-111
 *     u_t   = agent(y_t)                              LQG.__call__   deluca/agents/_lqg.py:83-98
 *                                                     DRC.__call__   deluca/agents/_drc.py:131-203
 *                                                     GRC.__call__   deluca/agents/_grc.py:110-169
 *                                                     SFC.__call__   deluca/agents/_sfc.py:154-214
 *                                                     DSC.__call__   deluca/agents/_dsc.py:210-292
 *     c_t   = quad_loss(y_t, u_t)                     deluca/agents/_grc.py:27-39
 *     x_t+1 = A x_t + B u_t + w_t
 * GRC, SFC and DSC are one family: a(k) = sum_f Theta[f] (sum_o Phi[f,o] ynat[k+o]) with a fixed filter bank
 * Phi[F,L] (GRC: F = L = m, Phi = I; SFC: F = h+1, L = m+1; DSC: F = 1+2h+h^2, L = 2m+1) and a y_nat history of
 * L + m entries.  Theta stacks the reference's parameter blocks in that order (M | M_0, M | M_0, M_1, M_2, M_3).
 * jit(grad(policy_loss)) is a hand adjoint; the m-step evolve recursion uses the Markov operators C A^j B.
 * A group of 8, 16 or 32 lanes per env (chosen from the contraction widths), controller state resident in shared
 * memory: the per-env state (dlc_lds_of_smem_bytes) must fit in 227 KB, d, n, p <= 64.
 * ------------------------------------------------------------------------------------ */
#define DLC_OF_LQG 0
#define DLC_OF_DRC 1
#define DLC_OF_FGRC 2 /* GRC / SFC / DSC through the filter bank */

typedef struct DlcOfParams {
  int64_t N;           /* envs in this call */
  int64_t env_offset;  /* global index of env 0 (keys the in-kernel disturbance stream) */
  int32_t T;           /* steps */
  int32_t d, n, p;     /* state / action / observation dim (the agents' d, n, p) */
  int32_t agent;       /* DLC_OF_* */
  int32_t m;           /* controller history: GRC/SFC/DSC m, DRC m; 0 for LQG */
  int32_t h;           /* DRC: length of the Markov operator / action history (_drc.py:52) */
  int32_t F, L;        /* filter family: number of filters and filter length */
  int32_t t0;          /* AGENT step counter at entry (lr decay) */
  int32_t decay;       /* lr = lr/(t+1) (_grc.py:145); DRC without decay uses lr = 1.0 as written (_drc.py:184) */
  int32_t shared_dynamics; /* 1: A, B, C (and K, Lk) are shared by all envs; 0: per env [N,...] */
  int32_t mode;        /* DLC_MODE_CLOSED_LOOP, or DLC_MODE_AGENT_ONLY: Agent.__call__(y) with y from y_in */
  float lr;            /* GRC/SFC/DSC lr (_grc.py:52), DRC lr_scale (_drc.py:53) */
  float R_M;           /* projection radius R_M (_grc.py:49) / RM (_drc.py:55) */
  float w_std;         /* std of the in-kernel Gaussian disturbance, used when W == NULL */
  int32_t lanes_per_env; /* 0 = automatic; 8, 16 or 32 forces that many lanes per env (tuning / tests) */
  uint64_t seed;
  int32_t flags;       /* DLC_OF_FLAG_*: bit 0 = the caller asserts Phi == I (GRC), bit 1 = the caller asserts that the truncated
                          response decays, |A^h B|_F <= |B|_F (DRC: sum_i G[i] us[i] is then carried as the d-vector
                          recursion s <- A s + B u - A^h B u_old); each unlocks a register-resident kernel on the colab's
                          system and is verified on the device (DLC_STATUS_BAD_HINT); bits an agent has no use for are ignored */
  int32_t env_t0;      /* ENV step counter at entry: counter of the in-kernel disturbance stream (see DlcLdsParams.env_t0) */
} DlcOfParams;
#define DLC_OF_FLAG_PHI_IDENTITY 1
#define DLC_OF_FLAG_STABLE_A 2

/* shared memory one env needs (bytes), 0 for an invalid shape */
size_t dlc_lds_of_smem_bytes(const DlcOfParams* p);

int32_t dlc_lds_of_rollout_f32(
    const DlcOfParams* p,
    const float* A,      /* [N,d,d] or [d,d] */
    const float* B,      /* [N,d,n] or [d,n] */
    const float* C,      /* [N,p,d] or [p,d] */
    const float* K,      /* LQG: [N,n,d]; DRC: [N,n,p] or NULL (= 0, _drc.py:104); NULL for the filter family */
    const float* Lk,     /* LQG: Kalman gain [N,d,p]; NULL otherwise */
    const float* Phi,    /* filter family: [F,L] (shared by all envs); NULL otherwise */
    const float* W,      /* [T,N,d] disturbances, or NULL: in-kernel Philox stream (w_std) */
    const float* y_in,   /* [T,N,p] observations for DLC_MODE_AGENT_ONLY; NULL otherwise */
    float* x_io,         /* [N,d] env state (closed loop) */
    float* theta_io,     /* filter family: Theta[N,F,n,p]; DRC: M[N,m,n,p]; NULL for LQG */
    float* z_io,         /* filter family: internal model state z[N,d] (_grc.py:79); LQG: x_hat[N,d]; NULL for DRC */
    float* ynat_io,      /* filter family: ynat_history[N,L+m,p], oldest -> newest; DRC: y_nat[N,m,p], newest first */
    float* us_io,        /* DRC: us[N,h,n], newest first; NULL otherwise */
    float* cost_out,     /* [T,N] or NULL */
    double* cost_sum_out,/* [N] or NULL */
    float* u_out,        /* [T,N,n] or NULL */
    int32_t* status_out, /* [N] or NULL */
    void* stream);

/* ------------------------------------------------------------------------------------
 * The LDS environment on its own for T steps: given or random actions in, observations out.
 *     LDS.__call__                    deluca/envs/_lds.py:102-111   w_t = disturbance(t); x <- A x + B u_t + w_t; y_t = C x
 *     LDS.generate_random_trajectory  deluca/envs/_lds.py:117-140   u_t ~ N(0,1) [n,1], y_t = self(u_t)
 *   u_in == NULL: u_t = action_std * z with z from Philox stream 3, counter (env_offset + n, t0 + t, c/4, 3).
 *   W == NULL: Gaussian disturbance of std w_std from stream 0 (the numbers the controller kernels draw); w_std = 0
 *   gives ZeroDisturbance; any other disturbance (SinusDisturbance) is passed as W[T,N,d].
 *   observe0 = 1 additionally writes y = C x of the entry state into y_out[0] (what LDS.init / reset return); the
 *   steps then write y_out[1 + t].  y_out is [T + observe0, N, p].
 *   A[d,d] B[d,n] C[p,d] shared (shared_dynamics = 1) or per env [N,...]; d, n, p <= 64.
 * ------------------------------------------------------------------------------------ */
int32_t dlc_lds_open_rollout_f32(int64_t N, int32_t T, int32_t d, int32_t n, int32_t p, int32_t shared_dynamics,
                                 const float* A, const float* B, const float* C /* or NULL if y_out is NULL */,
                                 float* x_io /*[N,d]*/, const float* u_in /*[T,N,n] or NULL*/,
                                 const float* W /*[T,N,d] or NULL*/, uint64_t seed, int64_t env_offset, int32_t t0,
                                 float w_std, float action_std, int32_t observe0,
                                 float* y_out /*[T+observe0,N,p] or NULL*/, float* u_out /*[T,N,n] or NULL*/,
                                 void* stream);

/* ------------------------------------------------------------------------------------
 * Experience collection around the rollout.
 *
 *   dlc_gae_f32          gae_advantages                 deluca/training/ppo.py:44-84
 *                        + the tail of ac_rollout        deluca/training/ppo.py:131-134
 *       rewards = -losses;  A_{T-1} = r_{T-1} - v_{T-1};
 *       A_t = r_t + discount v_{t+1} mask_t - v_t + discount gae_param mask_t A_{t+1};
 *       returns = advantages + log_probs  -- as written (`advantages + t[4]`, t[4] is log_probs);
 *       returns_use_values = 1 gives the intended advantages + values.
 *       done_masks == NULL means all ones (ac_rollout emits the constant 1.0, ppo.py:113).
 *       Layout: time_major = 0 -> [N,T] (what vmap(ac_rollout) stacks; flattening it to the
 *       [N*T] experience batch of ac_get_experience, ppo.py:196-199, is a view);
 *       time_major = 1 -> [T,N] (gae_advantages' own documented shape).
 *       float32, each operation rounded separately in the reference's order.
 *
 *   dlc_env_collect_f32  Learner.generate_trajectories  deluca/learners/core.py:17-41
 *       N random-policy episodes (SimpleRandom: action ~ N(0,1)*action_std, deluca/agents/_random.py:39-53)
 *       of T steps on a classic env; writes obs[N,T,d], action[N,T,m], next_obs[N,T,d].
 *       Actions: component c of step t is normal number f = t*m + c of the trajectory's stream, i.e. element f % 4 of
 *       the Philox block with counter (env_offset + n, f / 4, 0, 2) and key = seed.
 * ------------------------------------------------------------------------------------ */
int32_t dlc_gae_f32(int64_t N, int32_t T, int32_t time_major,
                    const float* losses, const float* values,
                    const float* log_probs /* NULL allowed if returns_out == NULL or returns_use_values */,
                    const float* done_masks /* or NULL */,
                    float discount, float gae_param, int32_t returns_use_values,
                    float* advantages_out, float* returns_out /* or NULL */, void* stream);

int32_t dlc_env_collect_f32(const DlcEnvSpec* env, int64_t N, int32_t T, uint64_t seed, int64_t env_offset,
                            float action_std, const float* x0 /*[N,d]*/, float* obs_out /*[N,T,d]*/,
                            float* action_out /*[N,T,m]*/, float* next_obs_out /*[N,T,d]*/, void* stream);

/* ------------------------------------------------------------------------------------
 * Strided slice transfer (no reference counterpart: JAX owns its transfers).  Copies `rows` rows of `row_bytes` bytes
 * between pitched buffers on `stream` (one asynchronous 2-D copy): an env slice [T, n] of a time-major result lands in
 * columns [lo, lo + n) of the caller's pinned [T, N] host array (kind 2, device -> host), or the reverse (kind 1).
 * Used by the env-sliced end-to-end rollout (deluca_b200.fused.lds_rollout_host_to_host), whose per-slice uploads,
 * kernels and downloads overlap.  kind: 1 host -> device, 2 device -> host, 3 device -> device.
 * ------------------------------------------------------------------------------------ */
int32_t dlc_copy_rows_async(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes,
                            size_t row_bytes, size_t rows, int32_t kind, void* stream);

/* ------------------------------------------------------------------------------------
 * Pipe-rate microbenchmarks (no reference counterpart): the FP32 FMA peak that bounds the LDS+GPC
 * kernels is not in MEASURED_PEAKS.json, so bench.py measures it with these.
 *   which: 0 FFMA (72 FMA/thread/iter)   1 packed fma.f32x2 (72 FMA/thread/iter)
 *          2 shared loads (32/thread/iter)   3 shuffles (32/thread/iter)
 *          4 FFMA + shared loads mixed (80 FMA + 16 loads /thread/iter)
 *          5 / 6 the store pattern of the time-major series kernels: six series [iters, blocks*128], one 4-byte store
 *            per thread, series and step, with 8 / 100 dependent FMAs per step;   7 the same with one series
 *   256 threads per block (128 for 5..7); sink = at least 32 zero floats (5, 6: 6*iters*blocks*128 floats; 7: a sixth).
 * ------------------------------------------------------------------------------------ */
int32_t dlc_microbench_f32(int32_t which, float* sink, int32_t iters, int32_t blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DELUCA_B200_H_ */
