// Argument block shared by the LDS rollout kernels.
#pragma once
#include "common.cuh"

namespace dlc {

struct LdsArgs {
  DlcLdsParams p;
  const float *A, *B, *K, *W, *U_in;
  float *x_io, *M_io, *hist_io, *xprev_io, *uprev_io, *eps_io;
  const float* eps_new;
  float* cost_out;
  double* cost_sum_out;
  float *x_out, *u_out;
  int32_t* status_out;
  float* ws;
  // GPC's cost_fn as a diagonal quadratic sum(q x^2) + sum(r u^2) (deluca/agents/_gpc.py:52,76,125); NULL = ones
  // (quad_loss).  Shared by all envs; only the lane-group kernel reads them (dlc_lds_rollout_cost_f32).
  const float *cost_q = nullptr, *cost_r = nullptr;
};

// packed-FFMA2 lane-split GPC kernel (lds_rollout_f2.cu).  Returns false if (d,m,H,HH) has no instantiation.
bool launch_gpc_f2(const LdsArgs& a, cudaStream_t st, int32_t* rc);

// quad-split GPC kernel for long even histories, H = HH (lds_gpc_quad.cu).  Returns false if it does not take the call.
bool launch_gpc_quad(const LdsArgs& a, cudaStream_t st, int32_t* rc);

// lane-split BPC kernel for compile-time shapes (lds_bpc.cu).  Returns false if (d,m,H) has no instantiation.
bool launch_bpc(const LdsArgs& a, float* bpc_cost_io, cudaStream_t st, int32_t* rc);
bool bpc_supported(const DlcLdsParams& p);

// lane-group GPC kernel for any (d, m <= 64, H, HH) (lds_gpc_group.cu).  Returns false if it does not take the call.
bool launch_gpc_group(const LdsArgs& a, cudaStream_t st, int32_t* rc);

// register-resident thread-per-env LQR kernel for (d, m) = (10, 3) (lds_lqr_reg.cu)
bool launch_lqr_reg(const LdsArgs& a, cudaStream_t st, int32_t* rc);
bool lqr_reg_supported(const DlcLdsParams& p);

}  // namespace dlc
