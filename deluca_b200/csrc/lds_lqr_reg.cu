// lds_lqr_reg.cu -- LDS + LQR (u = -K x, deluca/agents/_lqr.py:62-72; env deluca/envs/_lds.py:102-111) for small
// compile-time systems, one THREAD per env with A, B, K and the state in registers (170 floats at d = 10, m = 3).
// A step is 160 FFMA; the kernel is bound by the disturbance stream it reads (44 B per env-step), so W is fetched two
// steps ahead (two register buffers) to keep enough bytes in flight at 8 warps per SM -- or, when the rows are 16-byte
// tileable (RS > 0), through a per-warp shared-memory ring that cp.async fills RS steps ahead (row_ring.cuh).
#include <stdlib.h>

#include "common.cuh"
#include "lds_args.cuh"
#include "row_ring.cuh"

namespace dlc {
namespace {

// RS: 0 = disturbance rows through two register buffers (or drawn in the kernel), > 0 = stages of the cp.async ring.
// SAMPLE: sampled trajectories X / U are stored (a separate instance: the per-step test costs ~10 of ~300 instructions
// in a kernel that ncu shows ISSUE-bound at 8 warps per SM -- profiles/r02_lds_lqr_reg_ring_full.txt).
template <int D, int M, int RS, bool SAMPLE>
__global__ void __launch_bounds__(128) lds_lqr_reg_kernel(const LdsArgs a) {
  const DlcLdsParams& p = a.p;
  const int64_t nt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t nw0 = nt - lane;                       // first env of this warp
  if (nw0 >= p.N) return;                              // (warp-uniform)
  const bool live = nt < p.N;                          // idle lanes of the last warp shadow its last env, store nothing
  const int64_t n = live ? nt : p.N - 1;
  const int64_t np = p.shared_dynamics ? 0 : n;
  float A[D][D], B[D][M], K[M][D], x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + np * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      B[i][c] = __ldg(a.B + np * D * M + i * M + c);
      K[c][i] = __ldg(a.K + np * M * D + c * D + i);
    }
    x[i] = a.x_io[n * D + i];
  }
  const bool sampled = SAMPLE && live && (n % p.traj_stride == 0);
  const int64_t ns = n / p.traj_stride;
  const int64_t Ns = (p.N + p.traj_stride - 1) / p.traj_stride;
  const uint32_t env_id = (uint32_t)(p.env_offset + n);
  double csum = 0.0;
  // running output pointer (a 64-bit add per step instead of the index arithmetic)
  const bool store_cost = a.cost_out != nullptr && live;
  float* cost_p = a.cost_out + n;
  const int64_t cost_step = p.N;
  auto fetch = [&](int t, float (&w)[D]) {
    if (a.W) {
      const float* wr = a.W + ((int64_t)t * p.N + n) * D;
      if (D % 2 == 0) {
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(wr) + i);
          w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = __ldg(wr + i);
      }
    } else {
#pragma unroll
      for (int g = 0; g < (D + 3) / 4; ++g) {
        float z[4];
        normal4(p.seed, env_id, (uint32_t)(p.env_t0 + t), (uint32_t)g, 0u, z);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < D) w[4 * g + k] = __fmul_rn(p.w_std, z[k]);
      }
    }
  };
  auto advance = [&](int t, const float (&w)[D]) {
    float u[M];
#pragma unroll
    for (int c = 0; c < M; ++c) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(K[c][j], x[j], s);
      u[c] = -s;
    }
    float cost = 0.f, cu = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) cost = fmaf(x[i], x[i], cost);
#pragma unroll
    for (int c = 0; c < M; ++c) cu = fmaf(u[c], u[c], cu);
    cost += cu;
    csum += (double)cost;
    if (store_cost) *cost_p = cost;
    cost_p += cost_step;
    if (SAMPLE && sampled) {
      if (a.x_out)
#pragma unroll
        for (int i = 0; i < D; ++i) a.x_out[((int64_t)t * Ns + ns) * D + i] = x[i];
      if (a.u_out)
#pragma unroll
        for (int c = 0; c < M; ++c) a.u_out[((int64_t)t * Ns + ns) * M + c] = u[c];
    }
    float xn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(A[i][j], x[j], s);
#pragma unroll
      for (int c = 0; c < M; ++c) s = fmaf(B[i][c], u[c], s);
      xn[i] = s + w[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
  };
  if constexpr (RS > 0) {
    // disturbance rows through the warp's cp.async ring: RS steps (RS x 32 rows) in flight per warp.  The step loop is
    // unrolled by 4 so that stage offsets are compile-time constants within a group.
    static_assert(RS % 4 == 0, "the step loop is unrolled by 4");
    using Ring = WarpRowRing<D, RS>;
    extern __shared__ __align__(16) unsigned char ring_smem[];
    Ring ring;
    const int64_t left = p.N - nw0;
    ring.init(ring_smem + (size_t)(threadIdx.x >> 5) * Ring::kBytesPerWarp, a.W + nw0 * D, p.N * D,
              (int)(left < 32 ? left : 32), lane, live ? lane : 0);
    ring.prime(p.T);
    for (int tb = 0; tb < p.T; tb += 4) {
      const int sb = tb & (RS - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tb + k < p.T) {
          const uint32_t row = ring.acquire(sb + k);
          float w[D];
          if (D % 2 == 0) {
#pragma unroll
            for (int i = 0; i < D / 2; ++i) {
              const float2 v = lds_f2(row + 8u * i);
              w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < D; ++i) w[i] = lds_f1(row + 4u * i);
          }
          ring.release(sb + k, tb + k + RS < p.T);
          advance(tb + k, w);
        }
      }
    }
  } else {
    float w0[D], w1[D];
    if (p.T > 0) fetch(0, w0);
    if (p.T > 1) fetch(1, w1);
    for (int t = 0; t < p.T; t += 2) {
      {
        float w[D];
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = w0[i];
        if (t + 2 < p.T) fetch(t + 2, w0);
        advance(t, w);
      }
      if (t + 1 < p.T) {
        float w[D];
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = w1[i];
        if (t + 3 < p.T) fetch(t + 3, w1);
        advance(t + 1, w);
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < D; ++i) a.x_io[n * D + i] = x[i];
  if (a.cost_sum_out) a.cost_sum_out[n] = csum;
  if (a.status_out) a.status_out[n] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
}

}  // namespace

bool lqr_reg_supported(const DlcLdsParams& p) {
  return p.policy == DLC_POLICY_LQR && p.mode == DLC_MODE_CLOSED_LOOP && p.kernel_variant == 0 && p.d == 10 && p.m == 3;
}

template <int RS>
static int32_t launch_lqr_variant(const LdsArgs& a, unsigned grid, cudaStream_t st) {
  const bool sample = a.x_out || a.u_out;
  auto kern = sample ? lds_lqr_reg_kernel<10, 3, RS, true> : lds_lqr_reg_kernel<10, 3, RS, false>;
  size_t smem = 0;
  if (RS > 0) {
    smem = 4 * WarpRowRing<10, RS ? RS : 4>::kBytesPerWarp;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_lqr_reg_kernel)");
  }
  kern<<<grid, 128, smem, st>>>(a);
  return cuda_status(cudaGetLastError(), "lds_lqr_reg_kernel launch");
}

bool launch_lqr_reg(const LdsArgs& a, cudaStream_t st, int32_t* rc) {
  if (!lqr_reg_supported(a.p)) return false;
  const unsigned grid = (unsigned)((a.p.N + 127) / 128);
  // DLC_ROW_RING: stages of the cp.async disturbance ring (profiling switch; 0 = the register double buffer)
  const char* renv = getenv("DLC_ROW_RING");
  const int want = renv ? atoi(renv) : 4;   // measured (tools/bench_lqr.py): 3.63 ms without, 2.16 / 2.37 / 2.35 ms at 4 / 8 / 16
  if (want > 0 && a.p.T > 2 && WarpRowRing<10, 4>::usable(a.W, a.p.N))
    *rc = want == 8 ? launch_lqr_variant<8>(a, grid, st) : want == 16 ? launch_lqr_variant<16>(a, grid, st)
                                                                       : launch_lqr_variant<4>(a, grid, st);
  else
    *rc = launch_lqr_variant<0>(a, grid, st);
  return true;
}

}  // namespace dlc
