// Host-side helpers shared by the C-ABI entry points.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/geoguessr_b200.h"

namespace gg {

typedef __nv_bfloat16 bf16;

// Last error text, readable through gg_last_error().
void set_error(const char* fmt, ...);

#define GG_CHECK(cond, code, ...)    \
  do {                               \
    if (!(cond)) {                   \
      gg::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

#define GG_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      gg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GG_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define GG_LAUNCH_CHECK() GG_CUDA(cudaGetLastError())

// 2-D bf16 row-major tensor map with 128-byte swizzle.  inner = contiguous extent (elements),
// outer = rows, pitch in bytes (multiple of 16), box = {64 elements (=128 B), box_rows}.
// Out-of-bounds box elements are zero-filled.  Resolved through cudaGetDriverEntryPoint so the
// library does not link libcuda (it must load on a CPU-only box).
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_rows);

// Same without swizzle: box = {box_inner elements (<= 256, multiple of 8), box_rows}, rows land
// contiguously (box_inner * 2 bytes apart) in shared memory.
int make_tmap_bf16_2d_plain(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                            uint32_t box_inner, uint32_t box_rows);
int make_tmap_bf16_2d_sw64(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                            uint32_t box_inner, uint32_t box_rows);

// fp32 row-major tensor map, 128-byte swizzle: box = {32 floats (= 128 B), box_rows} (TMA stores of fp32 tiles)
int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_rows);

int device_sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): keeps the launch path free
// of driver calls (and legal inside CUDA-graph stream capture).
int set_max_dynamic_smem_once_impl(const void* kernel, size_t bytes);
template <typename K>
int set_max_dynamic_smem_once(K kernel, size_t bytes) {
  return set_max_dynamic_smem_once_impl(reinterpret_cast<const void*>(kernel), bytes);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace gg
