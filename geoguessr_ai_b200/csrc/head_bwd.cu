// Geocell head backward: dW = scale * dlogits^T x  (C x D, fp32) and db = scale * sum_b dlogits.
//
// Replaces autograd's addmm backward for models/super_guessr.py:354 (reached from
// main_coordinator_idun_s3.py:423).  Both operands are consumed in the layout the forward pass
// left them in -- dlogits (B, ldc) and x (B, D), batch-major -- i.e. as MN-major UMMA operands:
// the contraction index (batch) is the slow index of both.  TMA loads 64(k) x 64(mn) boxes with
// the 128-byte swizzle; a 128 x 256 x 64 stage is 2 + 4 such boxes.
//
// CTA pairs (clusters of two): the CTAs of a pair own vertically adjacent geocell blocks of the same 256
// embedding columns, so they contract against the SAME x tile.  Each loads two of its four boxes and TMA
// multicasts them into both CTAs' shared memory: 32 KB instead of 48 KB per CTA and k-block through the
// L2 -> SM fabric.  A stage is recycled when both CTAs' MMAs have read it (tcgen05.commit multicast).
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kWM = 128;  // geocells per tile (UMMA M)
constexpr int kWN = 256;  // embedding columns per tile (UMMA N)
constexpr int kWK = 64;   // batch rows per stage
constexpr int kWStages = 4;
constexpr int kBwdThreads = 192;
constexpr uint32_t kAtomBytes = kWK * 128;                 // 64 k-rows x 128 B
constexpr uint32_t kWStageA = (kWM / 64) * kAtomBytes;     // 16 KB
constexpr uint32_t kWStageB = (kWN / 64) * kAtomBytes;     // 32 KB

struct BwdSmem {
  uint8_t a[kWStages][kWStageA];
  uint8_t b[kWStages][kWStageB];
  uint64_t full[kWStages];
  uint64_t empty[kWStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBwdThreads, 1)
head_bwd_kernel(const __grid_constant__ CUtensorMap tm_g,  // dlogits: inner C, rows B
                const __grid_constant__ CUtensorMap tm_x,  // x:       inner D, rows B
                float* __restrict__ dW, int C, int D, int Bk, float scale_in,
                const float* __restrict__ grad_scale) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (C + 2 * kWM - 1) / (2 * kWM);  // pairs of geocell blocks
  const int num_n = (D + kWN - 1) / kWN;
  const int num_tiles = num_m * num_n;              // pair-tiles
  const int crank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_k = (Bk + kWK - 1) / kWK;
  const float scale = grad_scale ? scale_in * __ldg(grad_scale) : scale_in;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 2);  // this CTA's MMAs and the partner's
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.acc_full[a], 1);
      mbar_init(&sm.acc_empty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync();  // the partner's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  // tile order: the num_n column tiles of one geocell block are adjacent, so the CTAs that run
  // concurrently share the dlogits tile through L2 and dlogits is streamed from HBM once.
  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues
    int s = 0;
    uint32_t ph = 0;
    for (int t = pair; t < num_tiles; t += npairs) {
      const int m0 = (2 * (t / num_n) + crank) * kWM, n0 = (t % num_n) * kWN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&sm.empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.full[s], kWStageA + kWStageB);
#pragma unroll
          for (int i = 0; i < kWM / 64; ++i)
            tma_load_2d(sm.a[s] + i * kAtomBytes, &tm_g, &sm.full[s], m0 + 64 * i, kb * kWK);
#pragma unroll
          for (int i = 0; i < kWN / 128; ++i) {  // my half of the x tile's boxes, delivered to both CTAs
            const int atom = crank * (kWN / 128) + i;
            tma_load_2d_multicast(sm.b[s] + atom * kAtomBytes, &tm_x, &sm.full[s], n0 + 64 * atom, kb * kWK, 0x3);
          }
        }
        __syncwarp();
        if (++s == kWStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16(kWM, kWN, 1, 1);
    // 16 k-rows of 128 B per MMA; atoms (64 mn-elements) kAtomBytes apart; 8-row groups 1024 B apart
    const uint64_t da_base = umma_desc_sw128(smem_u32(sm.a[0]), kAtomBytes, 1024);
    const uint64_t db_base = umma_desc_sw128(smem_u32(sm.b[0]), kAtomBytes, 1024);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&sm.acc_empty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kWN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&sm.full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = da_base + static_cast<uint64_t>(s * (kWStageA >> 4));
          const uint64_t db = db_base + static_cast<uint64_t>(s * (kWStageB >> 4));
          umma_f16(d_tmem, da, db, idesc, kb != 0);
#pragma unroll
          for (int k = 1; k < kWK / 16; ++k) umma_f16_acc(d_tmem, da + k * (2048 >> 4), db + k * (2048 >> 4), idesc);
          umma_commit_multicast(&sm.empty[s], 0x3);
          if (kb == num_k - 1) umma_commit(&sm.acc_full[acc]);
        }
        __syncwarp();
        if (++s == kWStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    int it = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++it) {
      const int m0 = (2 * (t / num_n) + crank) * kWM, n0 = (t % num_n) * kWN;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(&sm.acc_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kWN;
#pragma unroll 1
      for (int c = 0; c < kWN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < C) {
          float* dst = dW + static_cast<size_t>(row) * D + col0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col0 + 4 * q + 4 <= D) {
              float4 o;
              o.x = __uint_as_float(r[4 * q + 0]) * scale;
              o.y = __uint_as_float(r[4 * q + 1]) * scale;
              o.z = __uint_as_float(r[4 * q + 2]) * scale;
              o.w = __uint_as_float(r[4 * q + 3]) * scale;
              *reinterpret_cast<float4*>(dst + 4 * q) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.acc_empty[acc]);
    }
  }
  tc_fence_before();
  cluster_sync();  // the partner may still be multicasting into / arriving on this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// db: column sums of dlogits.  Block (32 x 8): each thread owns 8 adjacent columns (one 16-byte
// load per row), 8 row lanes; grid.y row slices write partial[slice][C], reduced by a second kernel
// (deterministic, no atomics).
constexpr int kDbSlices = 32;
__global__ void db_partial_kernel(const bf16* __restrict__ g, int ldc, int B, int C, float* __restrict__ partial) {
  __shared__ float red[8][32][8];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < ldc) {
    for (int r = blockIdx.y * 8 + threadIdx.y; r < B; r += 8 * gridDim.y) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + static_cast<size_t>(r) * ldc + c0));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][i];
      if (c0 + i < C) partial[static_cast<size_t>(blockIdx.y) * C + c0 + i] = s;
    }
  }
}
// Block (32 columns x 8 slice lanes): lane y sums slices y, y+8, ... (coalesced 128-byte rows), then a
// fixed-order reduction over y in shared memory: deterministic, and wide enough to hide the latency.
__global__ void db_final_kernel(const float* __restrict__ partial, int C, int ld, int slices, float scale_in,
                                const float* __restrict__ grad_scale, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    int i = threadIdx.y;
    for (; i + 8 < slices; i += 16) {
      s0 += partial[static_cast<size_t>(i) * ld + c];
      s1 += partial[static_cast<size_t>(i + 8) * ld + c];
    }
    if (i < slices) s0 += partial[static_cast<size_t>(i) * ld + c];
  }
  red[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    const float scale = grad_scale ? scale_in * __ldg(grad_scale) : scale_in;
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x];
    db[c] = s * scale;
  }
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_head_bwd_workspace_bytes(int C) { return static_cast<size_t>(kDbSlices) * C * sizeof(float); }

extern "C" int gg_head_bwd(const void* dlogits_bf16, int ldc, const void* x_bf16, int x_ld, int B, int C, int D,
                           float scale, const float* grad_scale, float* dW, float* db, const float* db_partials,
                           int db_parts, int db_ld, void* workspace, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && D > 0, GG_ERR_ARG, "gg_head_bwd: empty problem B=%d C=%d D=%d", B, C, D);
  GG_CHECK(dlogits_bf16 && x_bf16 && dW, GG_ERR_ARG, "gg_head_bwd: null pointer");
  GG_CHECK(ldc >= C && ldc % 8 == 0, GG_ERR_ARG, "gg_head_bwd: ldc=%d must be >= C and a multiple of 8", ldc);
  GG_CHECK(D % 8 == 0 && x_ld >= D && x_ld % 8 == 0, GG_ERR_ARG, "gg_head_bwd: D=%d / x_ld=%d must be multiples of 8", D, x_ld);
  GG_CHECK(!db || workspace || db_partials, GG_ERR_ARG, "gg_head_bwd: db needs the workspace or db_partials");
  GG_CHECK(!db_partials || (db_parts > 0 && db_ld >= C), GG_ERR_ARG, "gg_head_bwd: bad db_partials shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_g, tm_x;
  int rc = make_tmap_bf16_2d(&tm_g, dlogits_bf16, C, B, static_cast<uint64_t>(ldc) * 2, 64, kWK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, x_bf16, D, B, static_cast<uint64_t>(x_ld) * 2, 64, kWK);
  if (rc) return rc;
  const int pair_tiles = ceil_div(C, 2 * kWM) * ceil_div(D, kWN);
  const int pairs = std::max(1, std::min(pair_tiles, device_sm_count() / 2));
  const size_t smem = sizeof(BwdSmem) + 1024;
  if (int e = set_max_dynamic_smem_once(head_bwd_kernel, smem)) return e;
  // cluster shape (2,1,1) is compiled into the kernel
  head_bwd_kernel<<<2 * pairs, kBwdThreads, smem, s>>>(tm_g, tm_x, dW, C, D, B, scale, grad_scale);
  GG_LAUNCH_CHECK();
  if (db && db_partials) {  // column sums already accumulated by the loss kernel (one row per CTA)
    db_final_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, s>>>(db_partials, C, db_ld, db_parts, scale, grad_scale, db);
    GG_LAUNCH_CHECK();
  } else if (db) {
    float* partial = static_cast<float*>(workspace);
    dim3 blk(32, 8), grd(ceil_div(ldc, 256), kDbSlices);
    db_partial_kernel<<<grd, blk, 0, s>>>(static_cast<const bf16*>(dlogits_bf16), ldc, B, C, partial);
    GG_LAUNCH_CHECK();
    db_final_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, s>>>(partial, C, C, kDbSlices, scale, grad_scale, db);
    GG_LAUNCH_CHECK();
  }
  return GG_OK;
}
