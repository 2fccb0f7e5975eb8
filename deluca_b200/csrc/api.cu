// ABI plumbing shared by every entry point of libdeluca_b200.so.
#include "common.cuh"

namespace dlc {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace dlc

extern "C" int32_t dlc_abi_version(void) { return DLC_ABI_VERSION; }

extern "C" const char* dlc_last_error(void) { return dlc::last_error_buf(); }

extern "C" int32_t dlc_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return dlc::cuda_status(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return dlc::cuda_status(e, "cudaDeviceGetAttribute");
  if (major != 10) return dlc::fail(DLC_ERR_DEVICE, "libdeluca_b200 is built for sm_100a only");
  return DLC_OK;
}

extern "C" size_t dlc_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(DlcLdsParams);
    case 1: return sizeof(DlcLungParams);
    case 2: return sizeof(DlcLungBuffers);
    case 3: return sizeof(DlcPlanParams);
    case 4: return sizeof(DlcSharedGpcParams);
    case 5: return sizeof(DlcOfParams);
    default: return 0;
  }
}

extern "C" int32_t dlc_copy_rows_async(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes,
                                       size_t row_bytes, size_t rows, int32_t kind, void* stream) {
  if (kind < 1 || kind > 3) return dlc::fail(DLC_ERR_ARG, "dlc_copy_rows_async: kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
  if (row_bytes == 0 || rows == 0) return DLC_OK;
  if (!dst || !src) return dlc::fail(DLC_ERR_ARG, "dlc_copy_rows_async: NULL buffer");
  if (dst_pitch_bytes < row_bytes || src_pitch_bytes < row_bytes)
    return dlc::fail(DLC_ERR_ARG, "dlc_copy_rows_async: a pitch is smaller than the row");
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  return dlc::cuda_status(cudaMemcpy2DAsync(dst, dst_pitch_bytes, src, src_pitch_bytes, row_bytes, rows, k,
                                            (cudaStream_t)stream),
                          "cudaMemcpy2DAsync");
}
