// Probe (GPU box only): how does cp.async.bulk.tensor.2d...tile::gather4 want its tensor map (box rows 1 or 4), and does
// a tile assembled from 32 gather4 loads equal the 128-row box load (128-byte swizzle) byte for byte?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu -lcuda && ./gather4_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm_box, const __grid_constant__ CUtensorMap tm_g, const int* rows,
                      int col0, uint8_t* out_box, uint8_t* out_gather) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a = base;            // 16 KB: box load of rows rows[0..128) must equal ...
  uint8_t* b = base + 16384;    // 16 KB: 32 gather4 loads
  uint64_t* bar = (uint64_t*)(base + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(16384) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                 "r"(smem_u32(a)), "l"((uint64_t)&tm_box), "r"(smem_u32(bar)), "r"(col0), "r"(rows[0]) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + 1)), "r"(16384) : "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(b + l * 512)), "l"((uint64_t)&tm_g), "r"(col0), "r"(rows[4 * l]), "r"(rows[4 * l + 1]),
        "r"(rows[4 * l + 2]), "r"(rows[4 * l + 3]), "r"(smem_u32(bar + 1)) : "memory");
  }
  for (int w = 0; w < 2; ++w) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar + w)) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("probe: barrier %d timed out\n", w); return; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) { out_box[i] = a[i]; out_gather[i] = b[i]; }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int R = 1000, D = 256;
  CK(cudaFree(0));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  __nv_bfloat16* h = (__nv_bfloat16*)malloc(sizeof(__nv_bfloat16) * R * D);
  for (int i = 0; i < R * D; ++i) h[i] = __float2bfloat16((float)(i % 3001) - 1500.0f);
  __nv_bfloat16* d;
  CK(cudaMalloc(&d, sizeof(__nv_bfloat16) * R * D));
  CK(cudaMemcpy(d, h, sizeof(__nv_bfloat16) * R * D, cudaMemcpyHostToDevice));
  uint8_t *ob, *og;
  CK(cudaMalloc(&ob, 16384)); CK(cudaMalloc(&og, 16384));
  int* drows; CK(cudaMalloc(&drows, 128 * 4));
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm_box;
  cuuint32_t box128[2] = {64, 128};
  CUresult r = enc(&tm_box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box128: %d\n", (int)r);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
    CUtensorMap tm_g;
    cuuint32_t boxg[2] = {64, (cuuint32_t)boxrows};
    r = enc(&tm_g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, boxg, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode gather map with box rows %d: %d\n", boxrows, (int)r);
    for (int mode = 0; mode < 2; ++mode) {  // 0: identity rows 100..227 (== box load), 1: permutation
      int rows[128];
      for (int i = 0; i < 128; ++i) rows[i] = mode == 0 ? 100 + i : (100 + (i * 37) % 128 * 7) % R;
      CK(cudaMemcpy(drows, rows, sizeof(rows), cudaMemcpyHostToDevice));
      CK(cudaMemset(og, 0xEE, 16384));
      probe<<<1, 128, 40000>>>(tm_box, tm_g, drows, 64, ob, og);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("box rows %d mode %d: kernel failed: %s\n", boxrows, mode, cudaGetErrorString(e)); return 0; }
      static uint8_t hb[16384], hg[16384];
      CK(cudaMemcpy(hb, ob, 16384, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hg, og, 16384, cudaMemcpyDeviceToHost));
      if (mode == 0) {
        printf("box rows %d identity: gather == box load: %s\n", boxrows, memcmp(hb, hg, 16384) == 0 ? "YES" : "no");
      } else {
        // expected: smem row i holds global row rows[i], columns 64..127, 16-byte pieces swizzled by (i & 7)
        int bad = 0;
        for (int i = 0; i < 128; ++i)
          for (int c = 0; c < 64; ++c) {
            const int piece = (c / 8) ^ (i & 7);
            __nv_bfloat16 v;
            memcpy(&v, hg + i * 128 + piece * 16 + (c % 8) * 2, 2);
            if (__bfloat162float(v) != __bfloat162float(h[rows[i] * D + 64 + c])) ++bad;
          }
        printf("box rows %d permutation: %d mismatching elements of 8192\n", boxrows, bad);
      }
    }
  }
  return 0;
}
