// row_ring.cuh -- a per-WARP shared-memory ring of per-step input rows, filled S steps ahead with cp.async.
//
// The thread-per-env kernels keep their whole system in registers (~170 floats at d = 10, m = 3), so only 8 warps fit
// on an SM: they are ISSUE-bound (ncu: 295 instructions per step of which 160 are the FFMAs, profiles/
// r02_lds_lqr_reg_full.txt) and every register spent on prefetching disturbance rows is taken from the system.  The rows
// of one step for the 32 envs of a warp are contiguous in the time-major arrays (W[t, n0 : n0 + 32, :]), so the warp
// copies that tile -- 16 bytes per lane and instruction, straight to shared memory -- S steps ahead: S x 32 rows per
// warp in flight, no registers, three copy instructions per step and no cross-warp synchronisation.  (A first version
// used cp.async.bulk + mbarriers: ptxas wraps every bulk copy in a ~20-instruction uniform-register election loop, which
// an issue-bound kernel cannot afford -- 3.86 ms against 3.71 ms for the register double buffer.)
//
// Per step:   wait<S-1>  ->  __syncwarp  ->  each lane reads its row  ->  __syncwarp  ->  refill the stage, commit.
#pragma once
#include <stdint.h>

namespace dlc {

// make a value opaque to the optimiser, so that it is kept in a register instead of being recomputed from special
// registers at every use (ptxas rematerialises shared-window addresses under register pressure)
__device__ __forceinline__ uint32_t keep_u32(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

template <int ROWF, int S>   // floats per row, stages
struct WarpRowRing {
  static constexpr int kTileBytes = 32 * ROWF * 4;
  static constexpr int kChunks = kTileBytes / 16;                 // 16-byte chunks of a full tile
  static constexpr int kPerLane = (kChunks + 31) / 32;            // copy instructions per lane and step
  static constexpr size_t kBytesPerWarp = (size_t)S * kTileBytes;
  static_assert(kTileBytes % 16 == 0, "a warp's tile is a whole number of 16-byte chunks");

  uint32_t dst;          // shared-window address of this lane's first chunk in stage 0
  uint32_t row;          // shared-window address of this lane's row in stage 0
  const char* next;      // global address of this lane's first chunk of the tile the next refill asks for
  int64_t tstride;       // bytes between consecutive steps
  int nchunk;            // chunks of this warp's tile that exist (ragged last warp)
  int lane;

  // the global rows must be 16-byte aligned and a multiple of 16 bytes long for every warp and step
  static bool usable(const void* base, int64_t N) {
    return base != nullptr && (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (N * ROWF) % 4 == 0;
  }

  __device__ __forceinline__ void copy_tile(uint32_t sdst, const char* g) const {
#pragma unroll
    for (int c = 0; c < kPerLane; ++c)
      if (lane + 32 * c < nchunk)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sdst + 512u * c), "l"(g + 512 * c) : "memory");
  }

  // warp_smem: this warp's kBytesPerWarp slice (16-byte aligned); src: the warp's first row at step 0; nvalid: envs of
  // this warp inside the batch (1..32); row_lane: whose row this lane reads (an idle lane reads a live one's).
  // Several rings (one per input array) may share the commit groups: init() each, then prime() each stage of each
  // ring and commit() once per stage; per step wait(), read rows at row_addr(s), __syncwarp(), fill() each, commit().
  __device__ __forceinline__ void init(unsigned char* warp_smem, const float* src, int64_t tstride_floats, int nvalid,
                                       int lane_, int row_lane) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(warp_smem);
    lane = lane_;
    dst = keep_u32(base + 16u * lane);
    row = keep_u32(base + (uint32_t)(row_lane * ROWF * 4));
    tstride = tstride_floats * 4;
    nchunk = nvalid * ROWF / 4;
    next = reinterpret_cast<const char*>(src) + 16 * lane;
  }

  // copy the next tile of the source into stage s (if `more`: not past the horizon) and step the source
  __device__ __forceinline__ void fill(int s, bool more) {
    if (more) copy_tile(dst + (uint32_t)(s * kTileBytes), next);
    next += tstride;
  }
  static __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }

  // every lane: the oldest of the S groups in flight has landed and is visible to the whole warp
  static __device__ __forceinline__ void wait() {
    asm volatile("cp.async.wait_group %0;\n" :: "n"(S - 1) : "memory");
    __syncwarp();
  }
  __device__ __forceinline__ uint32_t row_addr(int s) const { return row + (uint32_t)(s * kTileBytes); }

  // ---- the single-ring protocol
  __device__ __forceinline__ void prime(int T) {
#pragma unroll 1
    for (int s = 0; s < S; ++s) {                                 // one group per stage, empty past the horizon
      fill(s, s < T);
      commit();
    }
  }
  __device__ __forceinline__ uint32_t acquire(int s) const {
    wait();
    return row_addr(s);
  }
  // every lane, after it has read its row from stage s: refill the stage with the tile S steps on (if `more`)
  __device__ __forceinline__ void release(int s, bool more) {
    __syncwarp();
    fill(s, more);
    commit();
  }
};

__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(saddr));
  return v;
}

}  // namespace dlc
