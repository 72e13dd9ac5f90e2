// Hardware probe, not part of the path (tools/symm_probe.py): how fast do stores issued by the SMs reach a peer's
// memory over NVLink, by the form of the store?  dst is any device-visible address (a symmetric-memory peer mapping).
//   mode 0  coalesced 16-byte st.global.cg, every thread of every CTA striding over the buffer
//   mode 1  cp.async.bulk (TMA, 1-D) shared -> global, `chunk` bytes per instruction, up to 4 bulk groups in flight per CTA
//   mode 2  cp.async.bulk.tensor 2-D: 32 x 32 fp32 boxes (128-byte row pieces, 4 KB apart) into a (rows, 1024) fp32
//           matrix -- the form gg_head_bwd's push epilogue uses -- two stores in flight per warp, four warps per CTA
#include "common.cuh"
#include "ptx.cuh"

namespace gg {

__global__ void __launch_bounds__(256) nvlink_store_plain_kernel(float4* __restrict__ dst, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) __stcg(dst + i, v);
}

__global__ void __launch_bounds__(128) nvlink_store_bulk_kernel(uint8_t* __restrict__ dst, size_t bytes, uint32_t chunk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16)
    *reinterpret_cast<float4*>(buf + i) = make_float4(1.f, 2.f, 3.f, 4.f);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t nchunks = bytes / chunk;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk),
                 "r"(smem_u32(buf)), "r"(chunk)
                 : "memory");
    tma_store_commit();
    tma_store_wait_read<3>();  // (the source never changes: this only bounds the groups in flight)
  }
  tma_store_wait_all<0>();
}

__global__ void __launch_bounds__(128) nvlink_store_tile_kernel(const __grid_constant__ CUtensorMap tm, int rows, int cols) {
  __shared__ __align__(1024) uint8_t buf[4][2][4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane * 16; i < 8192; i += 32 * 16) *reinterpret_cast<float4*>(&buf[warp][0][0] + i) = make_float4(1.f, 2.f, 3.f, 4.f);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane != 0) return;
  const int tiles_c = cols / 32, tiles = (rows / 32) * tiles_c;
  int ob = 0;
  for (int t = blockIdx.x * 4 + warp; t < tiles; t += gridDim.x * 4) {
    tma_store_wait_read<1>();
    tma_store_2d(&tm, buf[warp][ob], (t % tiles_c) * 32, (t / tiles_c) * 32);
    tma_store_commit();
    ob ^= 1;
  }
  tma_store_wait_all<0>();
}

}  // namespace gg

extern "C" int gg_debug_nvlink_store_probe(void* dst, size_t bytes, int mode, int chunk, int ctas, gg_stream_t stream) {
  GG_CHECK(dst && bytes % 16 == 0 && ctas > 0, GG_ERR_ARG, "gg_debug_nvlink_store_probe: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == 0) {
    gg::nvlink_store_plain_kernel<<<ctas, 256, 0, s>>>(static_cast<float4*>(dst), bytes / 16);
  } else if (mode == 2) {
    const int cols = 1024, rows = static_cast<int>(bytes / (cols * 4)) / 32 * 32;
    CUtensorMap tm;
    if (int e = gg::make_tmap_f32_2d(&tm, dst, cols, rows, static_cast<uint64_t>(cols) * 4, 32)) return e;
    gg::nvlink_store_tile_kernel<<<ctas, 128, 0, s>>>(tm, rows, cols);
  } else {
    GG_CHECK(chunk >= 16 && chunk % 16 == 0 && chunk <= 40 * 1024 && bytes % chunk == 0, GG_ERR_ARG,
             "gg_debug_nvlink_store_probe: chunk=%d", chunk);
    const size_t smem = static_cast<size_t>(chunk) + 128;  // (<= 48 KB: no opt-in needed)
    gg::nvlink_store_bulk_kernel<<<ctas, 128, smem, s>>>(static_cast<uint8_t*>(dst), bytes, static_cast<uint32_t>(chunk));
  }
  GG_LAUNCH_CHECK();
  return GG_OK;
}
