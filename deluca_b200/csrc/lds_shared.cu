// Shared-dynamics LDS + LQR rollout on the 5th-generation tensor cores (BASELINE config 5's
// dense part).  When every env shares A, B, K the batch becomes the N dimension of a genuine GEMM:
//
//     [ x_{t+1} - w_t ]   [ A - B K ]
//     [     u_t       ] = [   -K    ]  X_t          G = [(d+m) x d],  X_t = [d x N_envs]
//
// One CTA owns NT = 64 envs for all T steps.  X_t lives in shared memory as the B operand (K-major,
// no swizzle; each 4-coordinate block is padded by 16 bytes so that the epilogue's stores -- one
// coordinate per lane -- hit 32 distinct banks) split into TF32 hi + lo parts; G is streamed from L2 in k-chunks (pre-packed by
// dlc_lds_shared_pack_f32 in exactly the shared-memory layout, also hi + lo); three tcgen05.mma passes
// (hi*hi + hi*lo + lo*hi, "3xTF32") give fp32-level accuracy with fp32 accumulation in TMEM.  The
// epilogue reads the accumulators back with tcgen05.ld, adds the disturbance, forms the per-env cost
// x'x + u'u, re-splits X_{t+1} and writes it straight back into the operand buffer: states never leave
// the SM between steps; HBM traffic is W in, costs out.
//
// Replaces LDS.__call__ (deluca/envs/_lds.py:102-111) + LQR.__call__ (deluca/agents/_lqr.py:62-72) +
// quad_loss (deluca/agents/_gpc.py:29-40) under lax.scan x vmap for shared (in_axes=None) dynamics.
#include <stdlib.h>

#include "common.cuh"

namespace dlc {

constexpr int kNT = 64;    // envs per CTA = MMA N
constexpr int kKC = 8;     // k-chunk of G staged per pipeline stage (one K=8 MMA step)
constexpr int kNS = 4;     // pipeline stages (bulk copies in flight)
constexpr int kThreads = 256;  // 8 warps: warp w and w + 4 share TMEM lane quadrant w % 4 and split the 64 envs
constexpr int kHalf = kNT / 2; // envs (accumulator columns) handled by one thread in the epilogue

struct SharedArgs {
  int64_t N;
  int T, d, m, MB;           // MB = 128-row blocks of G (rows padded with zeros)
  int RP;                    // rows of G actually packed: d + m rounded up to 8 (the last 128-row MMA block reads past
                             // them; those accumulator rows are never used)
  int dbg;                   // profiling switches (DLC_SHARED_DBG): 1 = hi*hi pass only, 2 = never refill G stages, 4 = skip epilogue stores
  int C;                     // CTAs per cluster sharing each G chunk through TMA multicast (1, 2 or 4)
  const float* Gp;           // packed G: [d/kKC chunks][2 (hi, lo)][kKC/4][RP][4]
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// X operand (K-major, no swizzle): [k/4 blocks][kNT envs][4 coordinates], each block padded by 4 floats
constexpr uint32_t kXBlock = kNT * 4 + 4;   // floats per 4-coordinate block (LBO = 1040 B: banks rotate by 4 per block)
__device__ __forceinline__ uint32_t xoff(int n, int k) {
  return (uint32_t)(k >> 2) * kXBlock + (uint32_t)n * 4u + (uint32_t)(k & 3);
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
// the same arrival on the barrier at this offset in every CTA of `mask` (stage reuse is cluster-wide)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: a lost arrival traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
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
__global__ void shared_pack_kernel(const float* A, const float* B, const float* K, int d, int m, int rows,
                                   float* Gp) {
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
// tcgen05.ld of one accumulator row (this thread's TMEM lane), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_row32(uint32_t taddr, uint32_t (&v)[kHalf]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Column sums over the 32 lanes of a warp of a per-lane 32-vector (transposed butterfly: 31 shuffles).
// On return lane L holds the sum of column L in q[0].
__device__ __forceinline__ void warp_colsum32(float (&q)[kHalf], int lane) {
#pragma unroll
  for (int off = 16, len = 16; off >= 1; off >>= 1, len >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < len; ++i) {
      const float lo = q[i], hi = q[i + len];
      const float keep = up ? hi : lo, send = up ? lo : hi;
      q[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// NT envs per CTA (= MMA N), TH threads (TH = 4 NT: four threads stage one env; TH / 128 warps share each TMEM lane quadrant
// and split the NT accumulator columns into 32-wide halves), NS pipeline stages.  <64, 256, 4> is one CTA per SM (r01);
// <32, 128, 2> fits TWO CTAs per SM (111 KB of shared memory, 128 of the 512 TMEM columns, 214 x 128 registers each), so
// the epilogue of one CTA's step runs under the other CTA's MMAs without any change to the per-CTA program: the state
// feedback x_{t+1} = f(x_t) forbids overlapping a CTA's own step t+1 MMAs with its step t epilogue, two independent
// CTAs per SM do not care.  Each CTA streams the whole operator, so the per-env L2 traffic doubles; TMA multicast over
// a cluster of 2 / 4 CTAs takes it back.
template <int MB, int NT, int TH, int NS>
__global__ void __launch_bounds__(TH, TH == 256 ? 1 : 2) lds_shared_lqr_kernel(const SharedArgs a) {
  static_assert(TH == 4 * NT && (TH == 256 || TH == 128), "four staging threads per env; one or two warps per TMEM quadrant");
  constexpr uint32_t XB = NT * 4 + 4;             // floats per 4-coordinate block of the X operand (padded by 16 B)
  constexpr int NTc = NT;
  auto xoff = [](int n, int k) { return (uint32_t)(k >> 2) * XB + (uint32_t)n * 4u + (uint32_t)(k & 3); };
  extern __shared__ __align__(1024) unsigned char smem[];
  const int rows = a.RP;
  const uint32_t part_floats = (uint32_t)(kKC / 4) * rows * 4;          // one hi (or lo) part of a stage
  const uint32_t part_bytes = part_floats * 4, stage_bytes = 2 * part_bytes;
  const int d = a.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t env0 = (int64_t)blockIdx.x * NT;
  const int nvalid = (int)((a.N - env0) < NT ? (a.N - env0 > 0 ? a.N - env0 : 0) : NT);
  // shared-memory carve-up
  float* sXhi = reinterpret_cast<float*>(smem);                         // [d/4][NT (+1 pad)][4]
  float* sXlo = sXhi + (size_t)(d / 4) * XB;
  float* sG0 = sXlo + (size_t)(d / 4) * XB;                                  // stage s at sG0 + s * 2 * part_floats
  float* sSum = sG0 + 2 * NS * part_floats;                            // [2 (x, u)][4 warps][NT] column sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 512);            // (sSum rows are indexed which * 8 + warp: 16 x 32) full[NS], empty[NS], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  uint64_t *bar_full = bars, *bar_empty = bars + NS, *bar_accum = bars + 2 * NS;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"(smem_u32(tmem_slot)), "n"(NT * 4) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  const int C = a.C;
  uint32_t crank = 0;
  if (C > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&bars[i], 1);                // full: own producer's arrive.expect_tx
    for (int i = 0; i < NS; ++i) mbar_init(&bars[NS + i], C);          // empty: one commit per CTA of the cluster
    mbar_init(&bars[2 * NS], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // X_0 -> operand buffer (hi/lo split); |x_0|^2 per env through the column-sum scratch
  float xsq = 0.f;  // thread n < NT: |x_t|^2 of env n
  {
    float part = 0.f;  // this thread's share of env (tid / 4): a quarter of the coordinates
    const int n = tid >> 2, k0 = (tid & 3) * (d / 4);
    const bool ok = n < nvalid;
    const float* gx = a.x_io + (env0 + n) * d;
    for (int k = k0; k < k0 + d / 4; ++k) {
      const float x = ok ? gx[k] : 0.f;
      float hi, lo;
      split_tf32(x, hi, lo);
      const uint32_t o = xoff(n, k);
      sXhi[o] = hi;
      sXlo[o] = lo;
      part = fmaf(x, x, part);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if ((tid & 3) == 0) sSum[n] = part;
  }
  double csum = 0.0;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (tid < NT) xsq = sSum[tid];
  if (C > 1) cluster_sync_all();   // peers' barriers are initialised before anything is multicast at them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_tf32(128, NT);
  const int nchunk = d / kKC;
  const int64_t total = (int64_t)a.T * nchunk;                           // G chunks consumed over the rollout
  const uint32_t lboA = (uint32_t)rows * 16u;
  constexpr uint32_t lboB = XB * 4;      // bytes between the two 4-coordinate halves of one K = 8 MMA step

  // TMA producer (thread 0): one bulk copy per chunk; G is packed so that a chunk is contiguous
  auto issue_load = [&](int s, int c) {
    const uint32_t bar = smem_u32(&bar_full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(stage_bytes) : "memory");
    if (C == 1) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
          :: "r"(smem_u32(sG0) + (uint32_t)s * stage_bytes), "l"(a.Gp + (size_t)c * 2 * part_floats),
             "r"(stage_bytes), "r"(bar) : "memory");
    } else {
      // this CTA fetches its 1/C slice of the chunk and multicasts it into every CTA of the cluster; each
      // destination's full barrier (same offset) receives the bytes
      const uint32_t slice = stage_bytes / C;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
          :: "r"(smem_u32(sG0) + (uint32_t)s * stage_bytes + crank * slice),
             "l"(reinterpret_cast<const char*>(a.Gp) + (size_t)c * stage_bytes + crank * slice), "r"(slice), "r"(bar),
             "h"(cmask) : "memory");
    }
  };
  if (tid == 0) {
    for (int g = 0; g < NS && g < total; ++g) issue_load(g, g % nchunk);
  }
  // descriptor bases: per-MMA descriptors differ from these only in the 14-bit start-address field
  const uint64_t a_desc0 = umma_desc(smem_u32(sG0), lboA, 128);
  const uint64_t bhi_desc0 = umma_desc(smem_u32(sXhi), lboB, 128);
  const uint64_t blo_desc0 = umma_desc(smem_u32(sXlo), lboB, 128);
  // per-thread epilogue constants
  const int quad = warp & 3, half = warp >> 2;   // TMEM lane quadrant / which 32 of the 64 envs
  const uint32_t lane_taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)half * kHalf;
  const int nbase = half * kHalf;
  int fill = 0, fphase = 0;      // pipeline cursor of this thread's role (stage index, phase parity)
  int cnext = NS % nchunk;      // producer: chunk index of the next refill
  int64_t gnext = NS;

  for (int t = 0; t < a.T; ++t) {
    float wv[kHalf];
    auto load_w = [&](int b) {   // disturbance of this thread's row of block b for its 32 envs
      const int row = b * 128 + quad * 32 + lane;
      if (a.W != nullptr && row < d) {
        const float* wrow = a.W + ((int64_t)t * a.N + env0 + nbase) * d + row;
        if (nvalid == NT) {
#pragma unroll
          for (int n = 0; n < kHalf; ++n) wv[n] = __ldg(wrow + (uint32_t)n * (uint32_t)d);
        } else {
#pragma unroll
          for (int n = 0; n < kHalf; ++n) wv[n] = nbase + n < nvalid ? __ldg(wrow + (uint32_t)n * (uint32_t)d) : 0.f;
        }
      } else {
#pragma unroll
        for (int n = 0; n < kHalf; ++n) wv[n] = 0.f;
      }
    };
    if (warp == 0) {
      // ---- producer warp: refill each stage as soon as the MMAs that read it have retired
      for (int c = 0; c < nchunk; ++c) {
        if (a.dbg & 2) break;
        mbar_wait(&bar_empty[fill], (uint32_t)fphase);
        if (lane == 0 && gnext < total) issue_load(fill, cnext);
        ++gnext;
        cnext = cnext + 1 == nchunk ? 0 : cnext + 1;
        if (++fill == NS) { fill = 0; fphase ^= 1; }
      }
    } else if (warp == 1) {
      // ---- MMA warp: every lane tracks the pipeline (warp-uniform descriptors), one elected lane issues
      uint64_t bhi = bhi_desc0, blo = blo_desc0;
      for (int c = 0; c < nchunk; ++c) {
        if (!(a.dbg & 2) || (t == 0 && c < NS)) mbar_wait(&bar_full[fill], (uint32_t)fphase);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t ahi = a_desc0 + (uint64_t)((uint32_t)fill * (stage_bytes >> 4));
        const uint64_t alo = ahi + (part_bytes >> 4);
        const uint32_t acc = c != 0;
        if (elect_one()) {
#pragma unroll
          for (int b = 0; b < MB; ++b) {
            const uint32_t td = tmem_base + (uint32_t)b * NT;
            umma_tf32(td, ahi + (uint64_t)(b * 128), bhi, idesc, acc);     // 128 rows x 16 B = 2048 B >> 4
            if (!(a.dbg & 1)) {
              umma_tf32(td, ahi + (uint64_t)(b * 128), blo, idesc, 1);
              umma_tf32(td, alo + (uint64_t)(b * 128), bhi, idesc, 1);
            }
          }
          if (C == 1) umma_commit(&bar_empty[fill]); else umma_commit_mc(&bar_empty[fill], cmask);
          if (c == nchunk - 1) umma_commit(bar_accum);
        }
        __syncwarp();
        bhi += (2 * lboB) >> 4;
        blo += (2 * lboB) >> 4;
        if (++fill == NS) { fill = 0; fphase ^= 1; }
      }
    }
    load_w(0);
    // ---- epilogue: accumulators -> registers, x_{t+1} -> operand buffer, costs
    mbar_wait(bar_accum, (uint32_t)(t & 1));
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    float xs_new = 0.f, us = 0.f;
#pragma unroll
    for (int b = 0; b < MB; ++b) {
      const int row = b * 128 + quad * 32 + lane;                        // output row of G owned by this thread
      uint32_t v[kHalf];
      tmem_ld_row32(lane_taddr + (uint32_t)b * NT, v);
      const bool is_x = row < d, is_u = row >= d && row < d + a.m;
      float* xh = sXhi + (uint32_t)(row >> 2) * XB + (uint32_t)(row & 3) + (uint32_t)nbase * 4u;   // lane = coordinate: 32 banks
      float* xl = sXlo + (uint32_t)(row >> 2) * XB + (uint32_t)(row & 3) + (uint32_t)nbase * 4u;
      float qx[kHalf];
      if (is_x) {
#pragma unroll
        for (int n = 0; n < kHalf; ++n) {
          const float val = __uint_as_float(v[n]) + wv[n];               // x' = (A x + B u) + w_t
          float hi, lo;
          split_tf32(val, hi, lo);
          if (!(a.dbg & 4)) { xh[n * 4] = hi; xl[n * 4] = lo; }
          qx[n] = val * val;
        }
      } else {
#pragma unroll
        for (int n = 0; n < kHalf; ++n) {
          const float val = __uint_as_float(v[n]);
          qx[n] = is_u ? val * val : 0.f;
        }
      }
      if (b + 1 < MB) load_w(b + 1);                                     // next block's disturbance in flight
      // a 128-row block holds state rows, action rows, or both (when d % 128 != 0): reduce x and u parts;
      // sSum[part][warp][32]: the column sums of warp (quad, half) for its 32 envs
      const bool mixed = b * 128 < d && (b + 1) * 128 > d;
      if (mixed) {
        float qu[kHalf];
#pragma unroll
        for (int n = 0; n < kHalf; ++n) { qu[n] = is_u ? qx[n] : 0.f; qx[n] = is_x ? qx[n] : 0.f; }
        warp_colsum32(qu, lane);
        sSum[(8 + warp) * kHalf + lane] = qu[0];
      }
      warp_colsum32(qx, lane);
      const int which = (b * 128 >= d) ? 1 : 0;                          // 0: state block (or mixed: x part), 1: action block
      sSum[(which * 8 + warp) * kHalf + lane] = qx[0];
      __syncthreads();
      if (tid < NT) {                                                   // thread = env: add the four quadrants of its half
        const int hw = (tid >> 5) * 4, col = tid & 31;
        const float* ps = sSum + (which * 8 + hw) * kHalf + col;
        const float s4 = (ps[0] + ps[kHalf]) + (ps[2 * kHalf] + ps[3 * kHalf]);
        if (which == 0) xs_new += s4; else us += s4;
        if (mixed) {
          const float* pu = sSum + (8 + hw) * kHalf + col;
          us += (pu[0] + pu[kHalf]) + (pu[2 * kHalf] + pu[3 * kHalf]);
        }
      }
      __syncthreads();
    }
    if (tid < NT) {
      const float cost = xsq + us;                                       // x_t'x_t + u_t'u_t
      if (tid < nvalid) {
        if (a.cost_out) a.cost_out[(int64_t)t * a.N + env0 + tid] = cost;
        csum += (double)cost;
      }
      xsq = xs_new;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  // x_T back to global: hi + lo reconstructs the fp32 state exactly (lo = x - hi is exact)
  {
    const int n = tid >> 2, k0 = (tid & 3) * (d / 4);
    if (n < nvalid) {
      float* gx = a.x_io + (env0 + n) * d;
      for (int k = k0; k < k0 + d / 4; ++k) {
        const uint32_t o = xoff(n, k);
        gx[k] = sXhi[o] + sXlo[o];
      }
    }
  }
  if (tid < nvalid) {
    if (a.cost_sum_out) a.cost_sum_out[env0 + tid] = csum;
    if (a.status_out) a.status_out[env0 + tid] = isfinite(csum) ? 0 : DLC_STATUS_NONFINITE;
  }
  __syncthreads();
  if (C > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "n"(NT * 4) : "memory");
  }
}

}  // namespace dlc

using namespace dlc;

static int shared_blocks(int d, int m) { return (d + m + 127) / 128; }
static int shared_rows(int d, int m) { return (d + m + 7) / 8 * 8; }

extern "C" size_t dlc_lds_shared_packed_floats(int32_t d, int32_t m) {
  if (d <= 0 || m <= 0 || d % kKC) return 0;
  return (size_t)2 * shared_rows(d, m) * d;
}

extern "C" int32_t dlc_lds_shared_pack_f32(const float* A, const float* B, const float* K, int32_t d, int32_t m,
                                           float* G_packed, void* stream) {
  if (!A || !B || !K || !G_packed) return fail(DLC_ERR_ARG, "dlc_lds_shared_pack_f32: NULL buffer");
  if (d < kKC || d % kKC || d > 256 || m < 1 || shared_blocks(d, m) > 3)
    return fail(DLC_ERR_SHAPE, "dlc_lds_shared_pack_f32: need d % 16 == 0, 16 <= d <= 256, d + m <= 384");
  const int RP = shared_rows(d, m), total = RP * d;
  shared_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, B, K, d, m, RP, G_packed);
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
  a.N = N; a.T = T; a.d = d; a.m = m; a.MB = shared_blocks(d, m); a.RP = shared_rows(d, m);
  a.Gp = G_packed; a.W = W; a.x_io = x_io; a.cost_out = cost_out; a.cost_sum_out = cost_sum_out;
  a.status_out = status_out;
  const size_t part = (size_t)(kKC / 4) * a.RP * 4;
  // launch shape: default = one 64-env CTA per SM; DLC_SHARED_VARIANT=0 = two 32-env CTAs per SM (see the kernel's comment;
  // measured 7 % slower: N = 32 MMAs and twice the operator traffic cost what the overlap wins)
  const char* venv = getenv("DLC_SHARED_VARIANT");
  const int variant = venv ? atoi(venv) : 1;
  const int NT = variant == 1 ? 64 : 32, TH = 4 * NT, NS = variant == 1 ? 4 : 2;
  const size_t smem = ((size_t)2 * (d / 4) * (NT * 4 + 4) + 2 * NS * part + 2 * 4 * 64) * sizeof(float) + 128;
  void (*kern)(SharedArgs);
  if (variant == 1)
    kern = a.MB == 1 ? lds_shared_lqr_kernel<1, 64, 256, 4> : a.MB == 2 ? lds_shared_lqr_kernel<2, 64, 256, 4>
                                                                        : lds_shared_lqr_kernel<3, 64, 256, 4>;
  else
    kern = a.MB == 1 ? lds_shared_lqr_kernel<1, 32, 128, 2> : a.MB == 2 ? lds_shared_lqr_kernel<2, 32, 128, 2>
                                                                        : lds_shared_lqr_kernel<3, 32, 128, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(lds_shared_lqr)");
  unsigned grid = (unsigned)((N + NT - 1) / NT);
  // clusters of up to 4 CTAs share every G chunk through TMA multicast (L2 traffic / C); idle CTAs of a
  // partially filled last cluster run on zeros and store nothing
  const char* denv = getenv("DLC_SHARED_DBG");
  a.dbg = denv ? atoi(denv) : 0;
  const char* cenv = getenv("DLC_SHARED_CLUSTER");
  int C = cenv ? atoi(cenv) : 1;   // (measured, tools/bench_shared_lqr.py: 6.56 / 6.64 / 7.46 ms at C = 1 / 2 / 4 -- not L2-bound)
  if (C != 1 && C != 2 && C != 4) C = 1;
  while (C > 1 && grid < (unsigned)C) C >>= 1;
  a.C = C;
  grid = (grid + C - 1) / C * C;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TH);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a);
  return cuda_status(e, "lds_shared_lqr_kernel launch");
}
