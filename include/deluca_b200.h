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

#define DLC_ABI_VERSION 1

#define DLC_OK 0
#define DLC_ERR_ARG (-1)         /* NULL where a buffer is required, negative size, ... */
#define DLC_ERR_SHAPE (-2)       /* dimension outside what the kernels support */
#define DLC_ERR_WORKSPACE (-3)   /* workspace too small / missing */
#define DLC_ERR_DEVICE (-4)      /* not an sm_100 device */

/* per-env status bits written by the kernels */
#define DLC_STATUS_NONFINITE 1   /* a NaN/Inf appeared in the env's state or cost */
#define DLC_STATUS_NOT_PD 2      /* Q_uu not positive definite in the Riccati backward pass */

int32_t dlc_abi_version(void);
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

typedef struct DlcLdsParams {
  int64_t N;           /* envs in this call (this rank's shard) */
  int64_t env_offset;  /* global index of env 0 (keys the disturbance stream) */
  int32_t T;           /* steps */
  int32_t d, m;        /* state / action dim */
  int32_t H, HH;       /* GPC controller / system history (_gpc.py:53-54); BPC uses H only */
  int32_t policy;      /* DLC_POLICY_* */
  int32_t t0;          /* agent step counter at entry (lr decay, disturbance counter) */
  int32_t decay;       /* GPC: lr = lr_scale/(t+1) (_gpc.py:161-162); BPC: /(t^0.75+1) (_bpc.py:128-129) */
  int32_t noise_uses_prev_action; /* 0 = reference as written (B u_t, _gpc.py:141-142,155); 1 = B u_{t-1} */
  int32_t shared_dynamics;        /* 1: A,B,K are [d,d],[d,m],[m,d] shared by all envs; 0: per env [N,...] */
  int32_t traj_stride; /* env n is sampled into x_out/u_out iff n % traj_stride == 0 (>=1) */
  float lr_scale;      /* _gpc.py:55 */
  float w_std;         /* std of the in-kernel Gaussian disturbance, used when W == NULL */
  float bpc_delta;     /* _bpc.py:53 */
  int32_t kernel_variant; /* 0 = auto; 1 = force the runtime-dims kernel (testing) */
  uint64_t seed;       /* disturbance / perturbation stream seed */
} DlcLdsParams;

size_t dlc_lds_workspace_bytes(const DlcLdsParams* p);

int32_t dlc_lds_rollout_f32(
    const DlcLdsParams* p,
    const float* A,       /* [N,d,d] or [d,d] */
    const float* B,       /* [N,d,m] or [d,m] */
    const float* K,       /* [N,m,d] or [m,d] */
    const float* W,       /* [T,N,d] disturbances, or NULL: generate in-kernel (Philox, w_std) */
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

/* ------------------------------------------------------------------------------------
 * Pipe-rate microbenchmarks (no reference counterpart): the FP32 FMA peak that bounds the LDS+GPC
 * kernels is not in MEASURED_PEAKS.json, so bench.py measures it with these.
 *   which: 0 FFMA (72 FMA/thread/iter)   1 packed fma.f32x2 (72 FMA/thread/iter)
 *          2 shared loads (32/thread/iter)   3 shuffles (32/thread/iter)
 *          4 FFMA + shared loads mixed (80 FMA + 16 loads /thread/iter)
 *   256 threads per block; sink = at least 32 zero floats.
 * ------------------------------------------------------------------------------------ */
int32_t dlc_microbench_f32(int32_t which, float* sink, int32_t iters, int32_t blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DELUCA_B200_H_ */
