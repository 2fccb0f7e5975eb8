// Shared-dynamics LDS + LQR rollout on the 5th-generation tensor cores (BASELINE config 5's
// dense part).  When every env shares A, B, K the batch becomes the N dimension of a genuine GEMM:
//
//     [ x_{t+1} - w_t ]   [ A - B K ]
//     [     u_t       ] = [   -K    ]  X_t          G = [(d+m) x d],  X_t = [d x N_envs]
//
// One CTA owns NT = 64 envs for all T steps.  X_t lives in shared memory as the B operand (K-major,
// no swizzle) split into TF32 hi + lo parts; G is streamed from L2 in k-chunks (pre-packed by
// dlc_lds_shared_pack_f32 in exactly the shared-memory layout, also hi + lo); three tcgen05.mma passes
// (hi*hi + hi*lo + lo*hi, "3xTF32") give fp32-level accuracy with fp32 accumulation in TMEM.  The
// epilogue reads the accumulators back with tcgen05.ld, adds the disturbance, forms the per-env cost
// x'x + u'u, re-splits X_{t+1} and writes it straight back into the operand buffer: states never leave
// the SM between steps; HBM traffic is W in, costs out.
//
// Replaces LDS.__call__ (deluca/envs/_lds.py:102-111) + LQR.__call__ (deluca/agents/_lqr.py:62-72) +
// quad_loss (deluca/agents/_gpc.py:29-40) under lax.scan x vmap for shared (in_axes=None) dynamics.
#include "common.cuh"

namespace dlc {

constexpr int kNT = 64;    // envs per CTA = MMA N
constexpr int kKC = 16;    // k-chunk of G staged per pipeline stage (two K=8 MMA steps)
constexpr int kThreads = 128;

struct SharedArgs {
  int64_t N;
  int T, d, m, MB;           // MB = 128-row blocks of G (rows padded with zeros)
  const float* Gp;           // packed G: [d/kKC chunks][2 (hi, lo)][kKC/4][MB*128][4]
  const float* W;            // [T,N,d] or NULL
  float* x_io;               // [N,d]
  float* cost_out;           // [T,N] or NULL
  double* cost_sum_out;      // [N] or NULL
  float* u_out;              // unused (reserved)
  int32_t* status_out;       // [N] or NULL
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave"): the tile is made of 8-row x 16-byte
// core matrices; SBO = byte stride between 8-row groups, LBO = byte stride between the two 16-byte K halves.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = kNT
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
               :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: a lost arrival traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  asm volatile("trap;\n");
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

// ------------------------------------------------------------------------------------ packing
// G = [A - B K ; -K ; 0] -> TF32 hi/lo, laid out per k-chunk exactly as the kernel's stage buffer.
__global__ void shared_pack_kernel(const float* A, const float* B, const float* K, int d, int m, int MB,
                                   float* Gp) {
  const int rows = MB * 128;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * d) return;
  const int r = idx / d, k = idx % d;
  double g = 0.0;
  if (r < d) {
    g = (double)A[r * d + k];
    for (int c = 0; c < m; ++c) g -= (double)B[r * m + c] * (double)K[c * d + k];
  } else if (r < d + m) {
    g = -(double)K[(r - d) * d + k];
  }
  float hi, lo;
  split_tf32((float)g, hi, lo);
  const int chunk = k / kKC, kk = k % kKC;
  const size_t per_part = (size_t)(kKC / 4) * rows * 4;                 // floats in one (hi or lo) part
  const size_t base = (size_t)chunk * 2 * per_part;
  const size_t off = ((size_t)(kk / 4) * rows + r) * 4 + (kk % 4);
  Gp[base + off] = hi;
  Gp[base + per_part + off] = lo;
}

// ------------------------------------------------------------------------------------ rollout
__global__ void __launch_bounds__(kThreads, 1) lds_shared_lqr_kernel(const SharedArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int d = a.d, MB = a.MB, rows = MB * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t env0 = (int64_t)blockIdx.x * kNT;
  // shared-memory carve-up
  float* sXhi = reinterpret_cast<float*>(smem);                         // [d/4][kNT][4]
  float* sXlo = sXhi + (size_t)kNT * d;
  const size_t part_floats = (size_t)(kKC / 4) * rows * 4;              // one hi (or lo) part of a stage
  float* sG0 = sXlo + (size_t)kNT * d;                                  // stage s: sG0 + s * 2 * part_floats
  const size_t stage_area = 4 * part_floats > (size_t)kNT * 129 + 31 ? 4 * part_floats : ((size_t)kNT * 129 + 31) / 32 * 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG0 + stage_area);       // [0,1] stage free, [2] accumulators ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float* scratch = sG0;                                                  // epilogue reduction scratch (stage 0)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n"
                 :: "r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // X_0 -> operand buffer (hi/lo split), thread = (env n = tid / 2, half of the coordinates)
  for (int idx = tid; idx < kNT * d; idx += kThreads) {
    const int n = idx / d, k = idx % d;
    const int64_t e = env0 + n;
    const float x = e < a.N ? a.x_io[e * d + k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    const size_t o = ((size_t)(k / 4) * kNT + n) * 4 + (k % 4);
    sXhi[o] = hi;
    sXlo[o] = lo;
  }
  float xsq = 0.f;  // thread n < kNT: |x_t|^2 of env n
  if (tid < kNT) {
    const int64_t e = env0 + tid;
    if (e < a.N)
      for (int k = 0; k < d; ++k) { const float x = a.x_io[e * d + k]; xsq = fmaf(x, x, xsq); }
  }
  double csum = 0.0;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_tf32(128, kNT);
  const int nchunk = d / kKC;
  const uint32_t lboA = (uint32_t)rows * 16u, lboB = (uint32_t)kNT * 16u;
  const int chunk_vec4 = (int)(2 * part_floats / 4);                    // uint4 loads per chunk
  uint32_t uses[2] = {0, 0};                                             // completed fills per stage

  for (int t = 0; t < a.T; ++t) {
    for (int c = 0; c < nchunk; ++c) {
      const int s = c & 1;
      if (uses[s] > 0) mbar_wait(&bars[s], (uses[s] - 1) & 1);          // MMAs of the previous fill retired
      float* stage = sG0 + (size_t)s * 2 * part_floats;
      const uint4* src = reinterpret_cast<const uint4*>(a.Gp + (size_t)c * 2 * part_floats);
      uint4* dst = reinterpret_cast<uint4*>(stage);
      for (int i = tid; i < chunk_vec4; i += kThreads) dst[i] = __ldg(src + i);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t ghi = smem_u32(stage), glo = smem_u32(stage + part_floats);
#pragma unroll
        for (int ks = 0; ks < kKC / 8; ++ks) {
          const uint32_t koffA = (uint32_t)ks * 2u * lboA;
          const uint32_t koffB = (uint32_t)(c * (kKC / 4) + ks * 2) * lboB;
          const uint64_t bhi = umma_desc(smem_u32(sXhi) + koffB, lboB, 128);
          const uint64_t blo = umma_desc(smem_u32(sXlo) + koffB, lboB, 128);
          for (int b = 0; b < MB; ++b) {
            const uint32_t roff = (uint32_t)b * 128u * 16u;
            const uint64_t ahi = umma_desc(ghi + koffA + roff, lboA, 128);
            const uint64_t alo = umma_desc(glo + koffA + roff, lboA, 128);
            const uint32_t td = tmem_base + (uint32_t)b * kNT;
            umma_tf32(td, ahi, bhi, idesc, (c | ks) != 0);
            umma_tf32(td, ahi, blo, idesc, 1);
            umma_tf32(td, alo, bhi, idesc, 1);
          }
        }
        umma_commit(&bars[s]);
        if (c == nchunk - 1) umma_commit(&bars[2]);
      }
      ++uses[s];
    }
    // ---- epilogue: accumulators -> registers
    mbar_wait(&bars[2], t & 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    float xs_new = 0.f, us = 0.f;
    for (int b = 0; b < MB; ++b) {
      const int row = b * 128 + tid;                                     // output row of G owned by this thread
      uint32_t v[kNT];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)b * kNT;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
          "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
          "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
            "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
            "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
            "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
            "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      const bool is_x = row < d, is_u = row >= d && row < d + a.m;
#pragma unroll
      for (int n = 0; n < kNT; ++n) {
        float val = __uint_as_float(v[n]);
        const int64_t e = env0 + n;
        if (is_x) {
          if (a.W && e < a.N) val += __ldg(a.W + ((int64_t)t * a.N + e) * d + row);   // x' = (A x + B u) + w_t
          float hi, lo;
          split_tf32(val, hi, lo);
          const size_t o = ((size_t)(row / 4) * kNT + n) * 4 + (row % 4);
          sXhi[o] = hi;
          sXlo[o] = lo;
          if (t == a.T - 1 && e < a.N) a.x_io[e * d + row] = val;
        }
        scratch[n * 129 + tid] = (is_x || is_u) ? val * val : 0.f;
      }
      __syncthreads();
      if (tid < kNT) {
        float sum = 0.f;
        for (int j = 0; j < kThreads; ++j) sum += scratch[tid * 129 + j];
        // rows [b*128, b*128+128) are all state rows, all action rows, or straddle d (d % 128 != 0)
        if ((b + 1) * 128 <= d) xs_new += sum;
        else if (b * 128 >= d) us += sum;
        else {  // straddling block: split the sum at row d
          float sx = 0.f;
          for (int j = 0; j < d - b * 128; ++j) sx += scratch[tid * 129 + j];
          xs_new += sx;
          us += sum - sx;
        }
      }
      __syncthreads();
    }
    if (tid < kNT) {
      const int64_t e = env0 + tid;
      const float cost = xsq + us;                                       // x_t'x_t + u_t'u_t
      if (e < a.N) {
        if (a.cost_out) a.cost_out[(int64_t)t * a.N + e] = cost;
        csum += (double)cost;
      }
      xsq = xs_new;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  if (tid < kNT) {
    const int64_t e = env0 + tid;
    if (e < a.N) {
      if (a.cost_sum_out) a.cost_sum_out[e] = csum;
      if (a.status_out) a.status_out[e] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(tmem_base) : "memory");
  }
}

}  // namespace dlc

using namespace dlc;

static int shared_blocks(int d, int m) { return (d + m + 127) / 128; }

extern "C" size_t dlc_lds_shared_packed_floats(int32_t d, int32_t m) {
  if (d <= 0 || m <= 0 || d % kKC) return 0;
  return (size_t)2 * shared_blocks(d, m) * 128 * d;
}

extern "C" int32_t dlc_lds_shared_pack_f32(const float* A, const float* B, const float* K, int32_t d, int32_t m,
                                           float* G_packed, void* stream) {
  if (!A || !B || !K || !G_packed) return fail(DLC_ERR_ARG, "dlc_lds_shared_pack_f32: NULL buffer");
  if (d < kKC || d % kKC || d > 256 || m < 1 || shared_blocks(d, m) > 3)
    return fail(DLC_ERR_SHAPE, "dlc_lds_shared_pack_f32: need d % 16 == 0, 16 <= d <= 256, d + m <= 384");
  const int MB = shared_blocks(d, m), total = MB * 128 * d;
  shared_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, B, K, d, m, MB, G_packed);
  return cuda_status(cudaGetLastError(), "shared_pack_kernel launch");
}

extern "C" int32_t dlc_lds_shared_lqr_rollout_f32(int64_t N, int32_t T, int32_t d, int32_t m, const float* G_packed,
                                                  const float* W, float* x_io, float* cost_out,
                                                  double* cost_sum_out, int32_t* status_out, void* stream) {
  if (N < 0 || T < 0) return fail(DLC_ERR_ARG, "dlc_lds_shared_lqr_rollout_f32: negative N or T");
  if (d < kKC || d % kKC || d > 256 || m < 1 || shared_blocks(d, m) > 3)
    return fail(DLC_ERR_SHAPE, "dlc_lds_shared_lqr_rollout_f32: need d % 16 == 0, 16 <= d <= 256, d + m <= 384");
  if (N == 0 || T == 0) return DLC_OK;
  if (!G_packed || !x_io) return fail(DLC_ERR_ARG, "dlc_lds_shared_lqr_rollout_f32: NULL buffer");
  SharedArgs a{};
  a.N = N; a.T = T; a.d = d; a.m = m; a.MB = shared_blocks(d, m);
  a.Gp = G_packed; a.W = W; a.x_io = x_io; a.cost_out = cost_out; a.cost_sum_out = cost_sum_out;
  a.status_out = status_out;
  const size_t part = (size_t)(kKC / 4) * a.MB * 128 * 4;
  const size_t stage_area = 4 * part > (size_t)kNT * 129 + 31 ? 4 * part : ((size_t)kNT * 129 + 31) / 32 * 32;
  const size_t smem = ((size_t)2 * kNT * d + stage_area) * sizeof(float) + 64;
  cudaError_t e = cudaFuncSetAttribute(lds_shared_lqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_shared_lqr)");
  const unsigned grid = (unsigned)((N + kNT - 1) / kNT);
  lds_shared_lqr_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "lds_shared_lqr_kernel launch");
}
