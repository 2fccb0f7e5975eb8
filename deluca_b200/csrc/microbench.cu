// Pipe-rate microbenchmarks: the FP32 (non-tensor) FMA peak that MEASURED_PEAKS.json lacks
// (SURVEY F9), plus shared-memory load and shuffle issue rates that bound the lane-split kernel.
// Enqueue-only like every entry point; bench.py / tools time them with CUDA events.
#include "common.cuh"

namespace dlc {

// 8x8 register-resident mat-vec, repeated: the instruction mix of the rollout kernels
// (FFMA with three distinct register operands, one of them unique per instruction).
__global__ void __launch_bounds__(256) mb_ffma_kernel(float* sink, int iters) {
  float m[8][8], x[8], acc[8];
  const float seed = (float)(threadIdx.x & 7) * 1e-3f + 1.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = seed + i;
    acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) m[i][j] = 1e-3f * (float)(i * 8 + j) + sink[(i * 8 + j) & 31];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(m[i][j], x[j], acc[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(acc[i], 1e-9f, x[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i] + x[i];
  if (s == 123.456f) sink[0] = s;
}

// same arithmetic with packed fma.rn.f32x2 (two FMAs per instruction on 64-bit register pairs)
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__global__ void __launch_bounds__(256) mb_ffma2_kernel(float* sink, int iters) {
  unsigned long long m[4][8], x[8], acc[4];
  const float seed = (float)(threadIdx.x & 7) * 1e-3f + 1.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = pack2(seed + j, seed + j);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[i] = pack2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      m[i][j] = pack2(1e-3f * (float)(i * 16 + j) + sink[(i * 8 + j) & 31], 1e-3f * (float)(i * 16 + 8 + j));
  }
  const unsigned long long tiny = pack2(1e-9f, 1e-9f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fma2(m[i][j], x[j], acc[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = fma2(acc[i], tiny, x[i]);
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s ^= acc[i] ^ x[i];
  if (s == 0x1234567ull) sink[0] = 1.f;
}

// conflict-free 32-bit shared loads feeding FMAs (the [element][thread] layout of the rollout ring)
__global__ void __launch_bounds__(256) mb_lds_kernel(float* sink, int iters) {
  __shared__ float sm[32 * 256];
  for (int i = threadIdx.x; i < 32 * 256; i += 256) sm[i] = (float)i * 1e-6f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float* p = sm + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k & 7] += p[k * 256];
    asm volatile("" ::: "memory");
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123.456f) sink[0] = s;
}

// xor-shuffles feeding adds
__global__ void __launch_bounds__(256) mb_shfl_kernel(float* sink, int iters) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k & 7] += __shfl_xor_sync(0xffffffffu, acc[(k + 1) & 7], 1 + (k & 1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123.456f) sink[0] = s;
}

// FFMA and conflict-free LDS interleaved 4:1 (can the two pipes overlap?)
__global__ void __launch_bounds__(256) mb_mix_kernel(float* sink, int iters) {
  __shared__ float sm[16 * 256];
  for (int i = threadIdx.x; i < 16 * 256; i += 256) sm[i] = (float)i * 1e-6f;
  __syncthreads();
  float m[8][8], x[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = 1.0f + i;
    acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) m[i][j] = 1e-3f * (float)(i * 8 + j) + sink[(i * 8 + j) & 31];
  }
  const float* p = sm + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = p[(2 * j) * 256], b = p[(2 * j + 1) * 256];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(m[i][j], x[j], acc[i]);
      x[j] += a * 1e-9f + b * 1e-9f;
    }
    asm volatile("" ::: "memory");
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i] + x[i];
  if (s == 123.456f) sink[0] = s;
}

// The store pattern of the time-major series kernels (lung episode, rollouts that keep costs): NS series [T,N], one
// 4-byte store per thread, series and step; WORK dependent FMAs per step stand in for the per-step arithmetic.
// sink must hold NS * iters * blocks * 128 floats.  Measures what HBM takes from this pattern with nothing else going on.
template <int NS, int WORK>
__global__ void __launch_bounds__(128) mb_series_store_kernel(float* out, int T) {
  const long long N = (long long)gridDim.x * 128, n = (long long)blockIdx.x * 128 + threadIdx.x;
  float v = (float)n * 1e-6f;
  float* o[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) o[s] = out + (long long)s * T * N + n;
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int k = 0; k < WORK; ++k) v = fmaf(v, 0.999f, 1e-3f);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      *o[s] = v + (float)s;
      o[s] += N;
    }
  }
}

// The same traffic with VEC consecutive envs per thread (one 4*VEC-byte store per thread, series and step) and a
// cache hint (0 default, 1 st.global.cs, 2 st.global.wt): which store shapes HBM takes at full rate.
template <int NS, int VEC, int HINT, int TPB>
__global__ void __launch_bounds__(TPB) mb_series_store_vec_kernel(float* out, int T, long long N) {
  const long long n = ((long long)blockIdx.x * TPB + threadIdx.x) * VEC;
  if (n >= N) return;
  float v = (float)n * 1e-6f;
  float* o[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) o[s] = out + (long long)s * T * N + n;
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v = fmaf(v, 0.999f, 1e-3f);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float w = v + (float)s;
      if (VEC == 4) {
        if (HINT == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(o[s]), "f"(w) : "memory");
        else if (HINT == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(o[s]), "f"(w) : "memory");
        else *reinterpret_cast<float4*>(o[s]) = make_float4(w, w, w, w);
      } else if (VEC == 2) {
        *reinterpret_cast<float2*>(o[s]) = make_float2(w, w);
      } else {
        if (HINT == 1) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(o[s]), "f"(w) : "memory");
        else if (HINT == 2) asm volatile("st.global.wt.f32 [%0], %1;" ::"l"(o[s]), "f"(w) : "memory");
        else *o[s] = w;
      }
      o[s] += N;
    }
  }
}

}  // namespace dlc

// which: 0 FFMA (72 FMA/thread/iter), 1 FFMA2 (72 FMA/thread/iter as 36 packed), 2 LDS (32 loads),
//        3 SHFL (32 shuffles), 4 mix (64 FFMA + 16 LDS + 16 FFMA).  sink: >= 32 floats of zeros.
extern "C" int32_t dlc_microbench_f32(int32_t which, float* sink, int32_t iters, int32_t blocks,
                                      void* stream) {
  using namespace dlc;
  if (!sink || iters < 0 || blocks <= 0) return fail(DLC_ERR_ARG, "dlc_microbench_f32: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  switch (which) {
    case 0: mb_ffma_kernel<<<blocks, 256, 0, st>>>(sink, iters); break;
    case 1: mb_ffma2_kernel<<<blocks, 256, 0, st>>>(sink, iters); break;
    case 2: mb_lds_kernel<<<blocks, 256, 0, st>>>(sink, iters); break;
    case 3: mb_shfl_kernel<<<blocks, 256, 0, st>>>(sink, iters); break;
    case 4: mb_mix_kernel<<<blocks, 256, 0, st>>>(sink, iters); break;
    case 5: mb_series_store_kernel<6, 8><<<blocks, 128, 0, st>>>(sink, iters); break;     // sink: 6*iters*blocks*128 floats
    case 6: mb_series_store_kernel<6, 100><<<blocks, 128, 0, st>>>(sink, iters); break;
    case 7: mb_series_store_kernel<1, 8><<<blocks, 128, 0, st>>>(sink, iters); break;     // sink: iters*blocks*128 floats
    // 10..: six series over N = blocks * 128 envs in other store shapes (same bytes as 5)
    case 10: mb_series_store_vec_kernel<6, 4, 0, 128><<<blocks / 4, 128, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 11: mb_series_store_vec_kernel<6, 2, 0, 128><<<blocks / 2, 128, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 12: mb_series_store_vec_kernel<6, 1, 1, 128><<<blocks, 128, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 13: mb_series_store_vec_kernel<6, 1, 2, 128><<<blocks, 128, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 14: mb_series_store_vec_kernel<6, 4, 1, 128><<<blocks / 4, 128, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 15: mb_series_store_vec_kernel<6, 1, 0, 512><<<blocks / 4, 512, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 16: mb_series_store_vec_kernel<6, 1, 0, 32><<<blocks * 4, 32, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    case 17: mb_series_store_vec_kernel<6, 4, 0, 32><<<blocks, 32, 0, st>>>(sink, iters, (long long)blocks * 128); break;
    default: return fail(DLC_ERR_ARG, "dlc_microbench_f32: unknown benchmark");
  }
  return cuda_status(cudaGetLastError(), "microbench launch");
}
