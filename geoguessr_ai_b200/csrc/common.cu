// Host helpers: error text, tensor-map encoding via the runtime's driver entry point.
#include <cudaTypedefs.h>

#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace gg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GG_ERR_CUDA;
  // The encode call is a DRIVER entry point: it needs a context current on the calling thread.
  // Threads that have only used the runtime lazily (e.g. torch's autograd worker threads) may not
  // have one bound yet; cudaFree(0) binds the device's primary context, once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    GG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  GG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, GG_ERR_ARG, "TMA base %p is not 16-byte aligned", base);
  GG_CHECK(pitch_bytes % 16 == 0, GG_ERR_ARG, "TMA row pitch %llu is not a multiple of 16 bytes",
           (unsigned long long)pitch_bytes);
  GG_CHECK(box_inner * 2 == 128 && box_rows >= 1 && box_rows <= 256, GG_ERR_ARG, "bad TMA box %u x %u", box_inner,
           box_rows);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_CHECK(r == CUDA_SUCCESS, GG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu)",
           (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
  return GG_OK;
}

// 64-byte swizzled boxes of 32 bf16 (the head GEMM's half-width logits staging)
int make_tmap_bf16_2d_sw64(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                           uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GG_ERR_CUDA;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    GG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  GG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, GG_ERR_ARG, "TMA base %p is not 16-byte aligned", base);
  GG_CHECK(pitch_bytes % 16 == 0, GG_ERR_ARG, "TMA row pitch %llu is not a multiple of 16 bytes",
           (unsigned long long)pitch_bytes);
  GG_CHECK(box_inner * 2 == 64 && box_rows >= 1 && box_rows <= 256, GG_ERR_ARG, "bad TMA box %u x %u", box_inner,
           box_rows);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_CHECK(r == CUDA_SUCCESS, GG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu)",
           (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
  return GG_OK;
}

int make_tmap_bf16_2d_plain(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                            uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GG_ERR_CUDA;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    GG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  GG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, GG_ERR_ARG, "TMA base %p is not 16-byte aligned", base);
  GG_CHECK(pitch_bytes % 16 == 0, GG_ERR_ARG, "TMA row pitch %llu is not a multiple of 16 bytes",
           (unsigned long long)pitch_bytes);
  GG_CHECK(box_inner >= 8 && box_inner <= 256 && box_inner % 8 == 0 && box_rows >= 1 && box_rows <= 256, GG_ERR_ARG,
           "bad TMA box %u x %u", box_inner, box_rows);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_CHECK(r == CUDA_SUCCESS, GG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu)",
           (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
  return GG_OK;
}

int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return GG_ERR_CUDA;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    GG_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  GG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, GG_ERR_ARG, "TMA base %p is not 16-byte aligned", base);
  GG_CHECK(pitch_bytes % 16 == 0, GG_ERR_ARG, "TMA row pitch %llu is not a multiple of 16 bytes",
           (unsigned long long)pitch_bytes);
  GG_CHECK(box_rows >= 1 && box_rows <= 256, GG_ERR_ARG, "bad TMA box 32 x %u", box_rows);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_CHECK(r == CUDA_SUCCESS, GG_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed with CUresult %d (inner=%llu outer=%llu pitch=%llu)",
           (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
  return GG_OK;
}

int set_max_dynamic_smem_once_impl(const void* kernel, size_t bytes) {
  struct Entry { const void* fn; unsigned long long devs; };
  static Entry table[64];
  static int n = 0;
  static std::mutex mu;
  int dev = 0;
  GG_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  Entry* e = nullptr;
  for (int i = 0; i < n; ++i)
    if (table[i].fn == kernel) e = &table[i];
  if (e && dev >= 0 && dev < 64 && ((e->devs >> dev) & 1ull)) return GG_OK;
  GG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  if (!e && n < 64) { e = &table[n++]; e->fn = kernel; e->devs = 0; }
  if (e && dev >= 0 && dev < 64) e->devs |= 1ull << dev;
  return GG_OK;
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

}  // namespace gg

extern "C" const char* gg_last_error(void) { return gg::g_err; }
extern "C" int gg_abi_version(void) { return GG_ABI_VERSION; }
