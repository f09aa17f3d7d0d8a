// Error reporting and device queries of the octseg C-ABI.
#include <cstring>
#include "common.h"
#include "ptx_sm100.h"

namespace octseg {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

int encode_tensor_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr,
                           int swizzle_bytes, int l2_promotion_bytes, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(OCTSEG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapL2promotion promo = l2_promotion_bytes == 256   ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : l2_promotion_bytes == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                       : l2_promotion_bytes == 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                                   : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(estr),
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(OCTSEG_ECUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, static_cast<int>(r));
  return OCTSEG_OK;
}

}  // namespace octseg

extern "C" const char* octseg_last_error(void) { return octseg::last_error_buf(); }

extern "C" int octseg_abi_version(void) { return OCTSEG_ABI_VERSION; }

extern "C" int octseg_sm_count(void) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return octseg::fail(OCTSEG_ENODEV, "cudaGetDevice failed");
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return octseg::fail(OCTSEG_ENODEV, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return octseg::fail(OCTSEG_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                        prop.major, prop.minor);
  if (dev >= 0 && dev < 64) cached[dev] = prop.multiProcessorCount;
  return prop.multiProcessorCount;
}
