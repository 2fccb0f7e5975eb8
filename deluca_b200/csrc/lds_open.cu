// lds_open.cu -- the LDS environment on its own, fused over T steps: given or random actions in, observations out.
//
//   LDS.__call__                    deluca/envs/_lds.py:102-111   w_t = disturbance(t); x <- A x + B u + w_t; y = C x; t += 1
//   LDS.generate_random_trajectory  deluca/envs/_lds.py:117-140   u_t ~ N(0, 1) [n,1]; obs = self(u_t); keeps obs[0,0]
//   (and the open-loop rollouts every system-identification caller builds from LDS.__call__)
//
// Mapping: for any (d, n, p) <= 64 the lane-group layout of lds_of.cu -- GL = 8 / 16 / 32 lanes own an env, A, B, C and
// the state stay in the group's slice of shared memory for all T steps, lanes split the rows of each mat-vec; for the
// colab's system (d_hidden = 10, d_action = 3, d_obs = 2) a thread per env with everything in registers.
// Streams: disturbances are Philox stream 0 (the same numbers the controller kernels draw), actions stream 3.
#include <stdlib.h>

#include "common.cuh"
#include "row_ring.cuh"

namespace dlc {
namespace {

struct OpenArgs {
  int64_t N, env_offset;
  int T, d, n, p, shared, t0, observe0;
  const float *A, *B, *C, *u_in, *W;
  float *x_io, *y_out, *u_out;
  uint64_t seed;
  float w_std, action_std;
  int epc;  // envs per CTA
  int S;    // floats of shared memory per env
};

__device__ __forceinline__ float pick4o(const float z[4], int k) { return k == 0 ? z[0] : (k == 1 ? z[1] : (k == 2 ? z[2] : z[3])); }

// r[k] (+)= sum_c Mat[row][c] v[c] for row = gl + k GL < rows (rows <= 2 GL)
template <int GL, bool ACC>
__device__ __forceinline__ void open_matvec(const float* Mat, const float* v, int rows, int cols, int gl, float (&r)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (!ACC) r[k] = 0.f;
    if (k * GL < rows) {
      const int row = gl + k * GL;
      if (row < rows) {
        const float* mr = Mat + row * cols;
        float acc = r[k];
        for (int c = 0; c < cols; ++c) acc = fmaf(mr[c], v[c], acc);
        r[k] = acc;
      }
    }
  }
}

template <int GL, int SD, int SN, int SP>
__global__ void __launch_bounds__(256) lds_open_kernel(const OpenArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int gl = threadIdx.x & (GL - 1), grp = threadIdx.x / GL;
  const int d = SD ? SD : a.d, n = SD ? SN : a.n, p = SD ? SP : a.p;
  int64_t env = (int64_t)blockIdx.x * a.epc + grp;
  const bool live = env < a.N;
  if (!live) env = a.N - 1;   // idle groups shadow the last env: warp-wide __syncwarp stays whole
  float* s = smem + (size_t)grp * a.S;
  float *sA = s, *sB = sA + d * d, *sC = sB + d * n, *sx = sC + p * d, *su = sx + d;
  const size_t eD = a.shared ? 0 : (size_t)env;
  for (int i = gl; i < d * d; i += GL) sA[i] = a.A[eD * d * d + i];
  for (int i = gl; i < d * n; i += GL) sB[i] = a.B[eD * d * n + i];
  if (a.C) for (int i = gl; i < p * d; i += GL) sC[i] = a.C[eD * p * d + i];
  for (int i = gl; i < d; i += GL) sx[i] = a.x_io[(size_t)env * d + i];
  __syncwarp();
  const uint32_t genv = (uint32_t)(a.env_offset + env);
  if (a.observe0 && a.y_out) {   // y = C x of the state as it stands (LDS.init / reset return it)
    float r[2];
    open_matvec<GL, false>(sC, sx, p, d, gl, r);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (live && gl + k * GL < p) a.y_out[(size_t)env * p + gl + k * GL] = r[k];
  }
  for (int t = 0; t < a.T; ++t) {
    // ---- action: the caller's, or N(0, action_std^2) from stream 3
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = gl + k * GL;
      if (k * GL < n && c < n) {
        float u;
        if (a.u_in) {
          u = a.u_in[((size_t)t * a.N + env) * n + c];
        } else {
          float z4[4];
          normal4(a.seed, genv, (uint32_t)(a.t0 + t), (uint32_t)(c >> 2), 3u, z4);
          u = __fmul_rn(a.action_std, pick4o(z4, c & 3));
        }
        su[c] = u;
        if (a.u_out && live) a.u_out[((size_t)t * a.N + env) * n + c] = u;
      }
    }
    // ---- disturbance w_t (own rows)
    float w[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      w[k] = 0.f;
      const int row = gl + k * GL;
      if (k * GL < d && row < d) {
        if (a.W) {
          w[k] = a.W[((size_t)t * a.N + env) * d + row];
        } else if (a.w_std != 0.f) {
          float z4[4];
          normal4(a.seed, genv, (uint32_t)(a.t0 + t), (uint32_t)(row >> 2), 0u, z4);
          w[k] = __fmul_rn(a.w_std, pick4o(z4, row & 3));
        }
      }
    }
    __syncwarp();
    // ---- x <- A x + B u + w                                            _lds.py:107
    float r[2];
    open_matvec<GL, false>(sA, sx, d, d, gl, r);
    open_matvec<GL, true>(sB, su, d, n, gl, r);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (gl + k * GL < d) sx[gl + k * GL] = r[k] + w[k];
    __syncwarp();
    // ---- y = C x                                                        _lds.py:108
    if (a.y_out) {
      open_matvec<GL, false>(sC, sx, p, d, gl, r);
      float* yo = a.y_out + ((size_t)(t + a.observe0) * a.N + env) * p;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (live && gl + k * GL < p) yo[gl + k * GL] = r[k];
    }
  }
  if (live) for (int i = gl; i < d; i += GL) a.x_io[(size_t)env * d + i] = sx[i];
}

// Thread-per-env variant for small compile-time shapes: A, B, C and the state live in REGISTERS (150 + 20 floats at
// d = 10, n = 3, p = 2), a step is 150 FFMA with no shared memory and no warp synchronisation, and the next step's
// action / disturbance are fetched while the current one computes.  The lane-group kernel above spends its time in
// shared-memory mat-vecs and __syncwarp; this one is bound by HBM when the inputs come from memory.
// RS > 0: given actions and disturbances arrive through per-warp cp.async rings RS steps deep (row_ring.cuh) instead of
// the two register buffers; the kernel is issue-bound at 8 warps per SM, so the per-step index arithmetic is replaced by
// running pointers as well.
template <int D, int NA, int P, int RS>
__global__ void __launch_bounds__(128) lds_open_reg_kernel(const OpenArgs a) {
  const int64_t nt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t nw0 = nt - lane;                       // first env of this warp
  if (nw0 >= a.N) return;                              // (warp-uniform)
  const bool live = nt < a.N;                          // idle lanes of the last warp shadow its last env, store nothing
  const int64_t env = live ? nt : a.N - 1;
  const size_t eD = a.shared ? 0 : (size_t)env;
  float A[D][D], B[D][NA], C[P][D], x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) A[i][j] = __ldg(a.A + eD * D * D + i * D + j);
#pragma unroll
    for (int c = 0; c < NA; ++c) B[i][c] = __ldg(a.B + eD * D * NA + i * NA + c);
    x[i] = a.x_io[(size_t)env * D + i];
  }
  if (a.C) {
#pragma unroll
    for (int q = 0; q < P; ++q)
#pragma unroll
      for (int j = 0; j < D; ++j) C[q][j] = __ldg(a.C + eD * P * D + q * D + j);
  }
  const uint32_t genv = (uint32_t)(a.env_offset + env);
  auto observe = [&](float* yo) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(C[q][j], x[j], s);
      yo[q] = s;
    }
  };
  auto fetch = [&](int t, float (&u)[NA], float (&w)[D]) {
    if (a.u_in) {
#pragma unroll
      for (int c = 0; c < NA; ++c) u[c] = __ldg(a.u_in + ((size_t)t * a.N + env) * NA + c);
    } else {
#pragma unroll
      for (int g = 0; g < (NA + 3) / 4; ++g) {
        float z4[4];
        normal4(a.seed, genv, (uint32_t)(a.t0 + t), (uint32_t)g, 3u, z4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < NA) u[4 * g + k] = __fmul_rn(a.action_std, z4[k]);
      }
    }
    if (a.W) {
      const float* wr = a.W + ((size_t)t * a.N + env) * D;
      if (D % 2 == 0) {   // rows of an even d start on 8-byte boundaries: half the load instructions
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
          const float2 v = __ldg(reinterpret_cast<const float2*>(wr) + i);
          w[2 * i] = v.x; w[2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = __ldg(wr + i);
      }
    } else if (a.w_std != 0.f) {
#pragma unroll
      for (int g = 0; g < (D + 3) / 4; ++g) {
        float z4[4];
        normal4(a.seed, genv, (uint32_t)(a.t0 + t), (uint32_t)g, 0u, z4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < D) w[4 * g + k] = __fmul_rn(a.w_std, z4[k]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = 0.f;
    }
  };
  if (a.observe0 && a.y_out && live) observe(a.y_out + (size_t)env * P);
  const bool store_u = a.u_out != nullptr && live, store_y = a.y_out != nullptr && live;
  float* yo = a.y_out + ((size_t)a.observe0 * a.N + env) * P;   // running output pointers (not dereferenced when absent)
  float* uo = a.u_out + (size_t)env * NA;
  const int64_t y_step = a.N * P, u_step = a.N * NA;
  // one step of the env given its action and disturbance
  auto advance = [&](const float (&u)[NA], const float (&w)[D]) {
    if (store_u)
#pragma unroll
      for (int c = 0; c < NA; ++c) uo[c] = u[c];
    uo += u_step;
    float xn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {   // same accumulation order as the lane-group kernel: A x, then B u, then + w
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fmaf(A[i][j], x[j], s);
#pragma unroll
      for (int c = 0; c < NA; ++c) s = fmaf(B[i][c], u[c], s);
      xn[i] = s + w[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = xn[i];
    if (store_y) {
      float yv[P];
      observe(yv);
      if (P == 2) {
        *reinterpret_cast<float2*>(yo) = make_float2(yv[0], yv[P - 1]);
      } else {
#pragma unroll
        for (int q = 0; q < P; ++q) yo[q] = yv[q];
      }
    }
    yo += y_step;
  };
  if constexpr (RS > 0) {
    // both input streams through the warp's rings (one commit group per step holds the step's two tiles); the step
    // loop is unrolled by 4 so that stage offsets are compile-time constants within a group
    static_assert(RS % 4 == 0, "the step loop is unrolled by 4");
    using RingW = WarpRowRing<D, RS>;
    using RingU = WarpRowRing<NA, RS>;
    extern __shared__ __align__(16) unsigned char ring_smem[];
    unsigned char* ws = ring_smem + (size_t)(threadIdx.x >> 5) * (RingW::kBytesPerWarp + RingU::kBytesPerWarp);
    const int64_t left = a.N - nw0;
    const int nvalid = (int)(left < 32 ? left : 32), rl = live ? lane : 0;
    RingW rw;
    RingU ru;
    rw.init(ws, a.W + nw0 * D, a.N * D, nvalid, lane, rl);
    ru.init(ws + RingW::kBytesPerWarp, a.u_in + nw0 * NA, a.N * NA, nvalid, lane, rl);
#pragma unroll 1
    for (int s = 0; s < RS; ++s) {
      rw.fill(s, s < a.T);
      ru.fill(s, s < a.T);
      RingW::commit();
    }
    for (int tb = 0; tb < a.T; tb += 4) {
      const int sb = tb & (RS - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tb + k < a.T) {
          RingW::wait();
          const uint32_t wrow = rw.row_addr(sb + k), urow = ru.row_addr(sb + k);
          float u[NA], w[D];
          if (D % 2 == 0) {
#pragma unroll
            for (int i = 0; i < D / 2; ++i) {
              const float2 v = lds_f2(wrow + 8u * i);
              w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < D; ++i) w[i] = lds_f1(wrow + 4u * i);
          }
#pragma unroll
          for (int c = 0; c < NA; ++c) u[c] = lds_f1(urow + 4u * c);
          __syncwarp();
          const bool more = tb + k + RS < a.T;
          rw.fill(sb + k, more);
          ru.fill(sb + k, more);
          RingW::commit();
          advance(u, w);
        }
      }
    }
    if (live)
#pragma unroll
      for (int i = 0; i < D; ++i) a.x_io[(size_t)env * D + i] = x[i];
    return;
  }
  // inputs are fetched TWO steps ahead (two register buffers, the loop advances two steps per trip): with 8 warps per
  // SM one step's 52 bytes per thread in flight is well short of what the HBM latency needs
  float u0[NA], w0[D], u1[NA], w1[D];
  if (a.T > 0) fetch(0, u0, w0);
  if (a.T > 1) fetch(1, u1, w1);
  for (int t = 0; t < a.T; t += 2) {
    {
      float u[NA], w[D];
#pragma unroll
      for (int c = 0; c < NA; ++c) u[c] = u0[c];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = w0[i];
      if (t + 2 < a.T) fetch(t + 2, u0, w0);
      advance(u, w);
    }
    if (t + 1 < a.T) {
      float u[NA], w[D];
#pragma unroll
      for (int c = 0; c < NA; ++c) u[c] = u1[c];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = w1[i];
      if (t + 3 < a.T) fetch(t + 3, u1, w1);
      advance(u, w);
    }
  }
  if (live)
#pragma unroll
    for (int i = 0; i < D; ++i) a.x_io[(size_t)env * D + i] = x[i];
}

// SinusDisturbance.__call__ (deluca/envs/_lds.py:44-52): w_t[i] = sin(t * phases[i]) * amplitude, the same for every env
__global__ void __launch_bounds__(256) sinus_disturbance_kernel(float* __restrict__ W, int64_t N, int T, int d,
                                                                const float* __restrict__ phases, float amplitude,
                                                                int t0) {
  const int64_t total = (int64_t)T * N * d;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d);
    const int t = (int)(idx / ((int64_t)N * d));
    W[idx] = __fmul_rn(sinf(__fmul_rn((float)(t0 + t), phases[i])), amplitude);
  }
}

}  // namespace
}  // namespace dlc

using namespace dlc;

extern "C" int32_t dlc_sinus_disturbance_f32(float* W, int64_t N, int32_t T, int32_t d, const float* phases,
                                             float amplitude, int32_t t0, void* stream) {
  if (!W || !phases || N < 0 || T < 0 || d <= 0) return fail(DLC_ERR_ARG, "dlc_sinus_disturbance_f32: bad argument");
  if (N == 0 || T == 0) return DLC_OK;
  const int64_t total = (int64_t)T * N * d;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  sinus_disturbance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, N, T, d, phases, amplitude, t0);
  return cuda_status(cudaGetLastError(), "sinus_disturbance_kernel launch");
}

extern "C" int32_t dlc_lds_open_rollout_f32(int64_t N, int32_t T, int32_t d, int32_t n, int32_t p,
                                            int32_t shared_dynamics, const float* A, const float* B, const float* C,
                                            float* x_io, const float* u_in, const float* W, uint64_t seed,
                                            int64_t env_offset, int32_t t0, float w_std, float action_std,
                                            int32_t observe0, float* y_out, float* u_out, void* stream) {
  if (N < 0 || T < 0) return fail(DLC_ERR_ARG, "dlc_lds_open_rollout_f32: negative N or T");
  if (d < 1 || n < 1 || p < 1 || d > 64 || n > 64 || p > 64)
    return fail(DLC_ERR_SHAPE, "dlc_lds_open_rollout_f32: d, n, p must be in 1..64");
  if (N == 0 || (T == 0 && !observe0)) return DLC_OK;
  if (!A || !B || !x_io) return fail(DLC_ERR_ARG, "dlc_lds_open_rollout_f32: A, B and x_io are required");
  if (y_out && !C) return fail(DLC_ERR_ARG, "dlc_lds_open_rollout_f32: y_out needs C");
  if ((uint64_t)(env_offset + N) > 0xFFFFFFFFull)
    return fail(DLC_ERR_SHAPE, "dlc_lds_open_rollout_f32: env_offset + N exceeds the 32-bit stream counter");
  OpenArgs a{};
  a.N = N; a.env_offset = env_offset; a.T = T; a.d = d; a.n = n; a.p = p; a.shared = shared_dynamics; a.t0 = t0;
  a.observe0 = observe0 ? 1 : 0;
  a.A = A; a.B = B; a.C = C; a.u_in = u_in; a.W = W; a.x_io = x_io; a.y_out = y_out; a.u_out = u_out;
  a.seed = seed; a.w_std = w_std; a.action_std = action_std;
  if (d == 10 && n == 3 && p == 2) {   // the colab's system: register-resident thread-per-env kernel
    const unsigned grid = (unsigned)((N + 127) / 128);
    const char* renv = getenv("DLC_ROW_RING");   // profiling switch: 0 = the register double buffers
    if ((!renv || atoi(renv) > 0) && T > 2 && WarpRowRing<10, 4>::usable(W, N) && WarpRowRing<3, 4>::usable(u_in, N)) {
      auto kern = lds_open_reg_kernel<10, 3, 2, 4>;
      const size_t smem = 4 * (WarpRowRing<10, 4>::kBytesPerWarp + WarpRowRing<3, 4>::kBytesPerWarp);
      kern<<<grid, 128, smem, (cudaStream_t)stream>>>(a);
      return cuda_status(cudaGetLastError(), "lds_open_reg_kernel (row rings) launch");
    }
    lds_open_reg_kernel<10, 3, 2, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "lds_open_reg_kernel launch");
  }
  int widest = d > n ? d : n;
  if (p > widest) widest = p;
  const int GL = widest <= 16 ? 8 : (widest <= 32 ? 16 : 32);
  int S = d * d + d * n + p * d + d + n;
  S = (S + 3) & ~3;
  if (GL < 32) S += ((GL - S % 32) + 32) % 32;   // slices of one warp's groups on different banks
  a.S = S;
  int dev = 0, max_smem = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_status(e, "cudaGetDevice");
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int epc = 256 / GL;
  while (epc > 32 / GL && (size_t)S * 4 * epc > (size_t)max_smem / 4) epc >>= 1;
  a.epc = epc;
  const size_t smem = (size_t)S * 4 * epc;
  if (smem > (size_t)max_smem) return fail(DLC_ERR_SHAPE, "dlc_lds_open_rollout_f32: system does not fit shared memory");
  void (*kern)(const OpenArgs) = GL == 8 ? lds_open_kernel<8, 0, 0, 0> : (GL == 16 ? lds_open_kernel<16, 0, 0, 0>
                                                                                   : lds_open_kernel<32, 0, 0, 0>);
  if (d == 10 && n == 3 && p == 10) kern = lds_open_kernel<8, 10, 3, 10>;   // state feedback: C = I
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_open_kernel)");
  const unsigned grid = (unsigned)((N + epc - 1) / epc);
  kern<<<grid, GL * epc, smem, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "lds_open_kernel launch");
}
