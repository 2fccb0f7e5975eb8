// Shared device/host helpers for libdeluca_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/deluca_b200.h"

namespace dlc {

// ---------------------------------------------------------------- error plumbing
char* last_error_buf();  // thread-local, 512 bytes (api.cu)

inline int32_t fail(int32_t code, const char* msg) {
  snprintf(last_error_buf(), 512, "%s", msg);
  return code;
}
inline int32_t cuda_status(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return DLC_OK;
  snprintf(last_error_buf(), 512, "%s: %s", where, cudaGetErrorString(e));
  return (int32_t)e;
}

// ---------------------------------------------------------------- Philox4x32-10
// Salmon et al., SC'11.  counter = (c0,c1,c2,c3), key = (k0,k1); oracle/philox.py restates it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// two uint32 -> two standard normals: Box-Muller on 24-bit uniforms, u1 in (0,1], u2 in [0,1).
//   rad = sqrt(max(-2 ln2 * lg2(u1), 0)),  ang = 2 pi (u2 - 1/2) in [-pi, pi)  (u2 - 1/2 is exact)
// Every step is a single pinned PTX instruction so that the in-register stream of the rollout
// kernels and the materialised stream of dlc_gaussian_disturbance_f32 agree bit for bit
// (the compiler may reassociate around __logf/__sincosf differently from one kernel to the next).
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const float u1 = __fmul_rn((float)((ra >> 8) + 1u), 5.9604644775390625e-08f);  // 2^-24
  const float u2 = __fmul_rn((float)(rb >> 8), 5.9604644775390625e-08f);
  float l2, rad, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  const float r2 = fmaxf(__fmul_rn(l2, -1.3862943611198906f), 0.0f);  // lg2.approx may round lg2(1-eps) above 0
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(r2));
  const float ang = __fmul_rn(u2 - 0.5f, 6.2831853071795864769f);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  z0 = __fmul_rn(rad, c);
  z1 = __fmul_rn(rad, s);
}

// four standard normals for (env, t, group j, stream s)
__device__ __forceinline__ void normal4(uint64_t seed, uint32_t env, uint32_t t, uint32_t j,
                                        uint32_t s, float z[4]) {
  const uint4 r = philox4x32_10(make_uint4(env, t, j, s), (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller(r.x, r.y, z[0], z[1]);
  box_muller(r.z, r.w, z[2], z[3]);
}

}  // namespace dlc
