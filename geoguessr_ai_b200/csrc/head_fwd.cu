// Geocell head forward: logits = x W^T + b as a persistent, warp-specialised tcgen05 GEMM.
//
// Replaces  models/super_guessr.py:354 (nn.Linear), :355 (softmax), :358-361 (argmax + centroid
// gather) and :365 (top-k) of the reference.  Per 128x256 output tile the accumulator lives in
// TMEM (2 x 256 fp32 columns, double buffered); the epilogue warps add the bias, optionally
// write the bf16 logits (training only) and keep, per row, an online (max, sum-exp) and the top-K
// logits of the tile, so that in serving the (B, C) logit / probability matrices never reach HBM.
// A small merge kernel combines the per-tile partials into top-k probabilities and indices,
// argmax, predicted centroid and the row log-sum-exp.
//
// Layout: x (M=B, K=D) bf16 row-major; W (N=C, K=D) bf16 row-major (both K-major operands, 128 B
// swizzled TMA boxes of 64 K-elements); logits (B, ldc) bf16, ldc >= C padded to a multiple of 64.
#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kBM = 128;        // rows of x per tile (UMMA M)
constexpr int kBN = 256;        // geocells per tile   (UMMA N)
constexpr int kBK = 64;         // K elements per stage (128 B of bf16 = one swizzle span)
constexpr int kStages = 4;      // 4 x (16 KB + 32 KB) = 192 KB
constexpr int kFwdThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr uint32_t kStageBytesA = kBM * kBK * 2;
constexpr uint32_t kStageBytesB = kBN * kBK * 2;
constexpr float kLog2e = 1.4426950408889634f;

struct FwdSmem {
  uint8_t a[kStages][kStageBytesA];
  uint8_t b[kStages][kStageBytesB];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

template <int KTOP>
__device__ __forceinline__ void topk_insert(float (&tv)[KTOP], int (&ti)[KTOP], float v, int idx) {
  // list sorted descending; strict '>' keeps the lower index on ties (columns are visited ascending)
  if (v > tv[KTOP - 1]) {
    tv[KTOP - 1] = v;
    ti[KTOP - 1] = idx;
#pragma unroll
    for (int j = KTOP - 1; j > 0; --j) {
      if (tv[j] > tv[j - 1]) {
        float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
        int iv = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = iv;
      }
    }
  }
}

template <int KTOP, bool WRITE_LOGITS>
__global__ void __launch_bounds__(kFwdThreads, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                const float* __restrict__ bias_pad, bf16* __restrict__ logits, int ldc,
                float* __restrict__ pmax, float* __restrict__ psum, float* __restrict__ ptopv,
                int* __restrict__ ptopi, int M, int N, int K, int Mpad) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + kBM - 1) / kBM;
  const int num_n = (N + kBN - 1) / kBN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % num_m) * kBM, n0 = (t / num_m) * kBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&sm.full[s], kStageBytesA + kStageBytesB);
          tma_load_2d(sm.a[s], &tm_x, &sm.full[s], kb * kBK, m0);
          tma_load_2d(sm.b[s], &tm_w, &sm.full[s], kb * kBK, n0);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&sm.acc_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sm.a[s]), b0 = smem_u32(sm.b[s]);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a0 + k * 32, 16, 1024);
            const uint64_t db = umma_desc_sw128(b0 + k * 32, 16, 1024);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&sm.empty[s]);  // smem slot reusable once these MMAs have read it
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&sm.acc_full[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps (128 threads, one row each) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int mb = t % num_m, nb = t / num_m;
      const int m0 = mb * kBM, n0 = nb * kBN;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int row = m0 + row_in_tile;
      mbar_wait(&sm.acc_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBN;

      float run_max = -INFINITY, run_sum = 0.f;
      float tv[KTOP];
      int ti[KTOP];
#pragma unroll
      for (int j = 0; j < KTOP; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }

#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        float v[32];
        const float4* bp = reinterpret_cast<const float4*>(bias_pad + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = __ldg(bp + q);
          v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
          v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
          v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
          v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
        }
        if (WRITE_LOGITS) {
          if (row < M) {
            bf16* dst = logits + static_cast<size_t>(row) * ldc + col0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (col0 + 8 * q + 8 <= ldc) {
                uint4 pk;
                pk.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
                pk.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                pk.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
                pk.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                *reinterpret_cast<uint4*>(dst + 8 * q) = pk;
              }
            }
          }
        }
        if (col0 + 32 > N) {  // tail tile: geocells >= C do not exist
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i >= N) v[i] = -INFINITY;
        }
        float cmax = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) cmax = fmaxf(cmax, v[i]);
        if (cmax > run_max) {
          run_sum *= ex2_approx((run_max - cmax) * kLog2e);  // run_max=-inf -> 0 * 0
          run_max = cmax;
        }
        if (run_max > -INFINITY) {
          const float ms = run_max * kLog2e;
          float acc_s = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc_s += ex2_approx(fmaf(v[i], kLog2e, -ms));
          run_sum += acc_s;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) topk_insert<KTOP>(tv, ti, v[i], col0 + i);
      }
      // TMEM accumulator drained -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&sm.acc_empty[acc]);

      if (row < M) {
        const size_t p = static_cast<size_t>(nb) * Mpad + row;
        pmax[p] = run_max;
        psum[p] = run_sum;
#pragma unroll
        for (int j = 0; j < KTOP; ++j) {
          const size_t pj = (static_cast<size_t>(nb) * KTOP + j) * Mpad + row;
          ptopv[pj] = tv[j];
          ptopi[pj] = ti[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Per row: merge the per-tile (max, sumexp, top-K) partials.
//   topk_val = softmax probabilities exp(l - max) / sum   (super_guessr.py:355,365)
//   pred_cell = argmax (:358), pred_llh = centroids[pred_cell] (:359-361), lse = max + log(sum)
template <int KTOP>
__global__ void head_merge_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                  const float* __restrict__ ptopv, const int* __restrict__ ptopi, int num_n,
                                  int Mpad, int M, int k, const float* __restrict__ centroids,
                                  float* __restrict__ topk_val, long long* __restrict__ topk_idx,
                                  long long* __restrict__ pred_cell, float* __restrict__ pred_llh,
                                  float* __restrict__ lse) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float gmax = -INFINITY;
  for (int nb = 0; nb < num_n; ++nb) gmax = fmaxf(gmax, pmax[static_cast<size_t>(nb) * Mpad + row]);
  float gsum = 0.f;
  float tv[KTOP];
  int ti[KTOP];
#pragma unroll
  for (int j = 0; j < KTOP; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  for (int nb = 0; nb < num_n; ++nb) {
    const size_t p = static_cast<size_t>(nb) * Mpad + row;
    const float m = pmax[p];
    if (m > -INFINITY) gsum += psum[p] * expf(m - gmax);
#pragma unroll
    for (int j = 0; j < KTOP; ++j) {
      const size_t pj = (static_cast<size_t>(nb) * KTOP + j) * Mpad + row;
      const float v = ptopv[pj];
      if (!(v > tv[KTOP - 1])) break;  // partial lists are sorted descending
      topk_insert<KTOP>(tv, ti, v, ptopi[pj]);
    }
  }
  const float inv = 1.f / gsum;
  for (int j = 0; j < k; ++j) {
    topk_val[static_cast<size_t>(row) * k + j] = expf(tv[j] - gmax) * inv;
    topk_idx[static_cast<size_t>(row) * k + j] = ti[j];
  }
  const int best = ti[0];
  if (pred_cell) pred_cell[row] = best;
  if (pred_llh) {
    pred_llh[2 * row + 0] = centroids[2 * best + 0];
    pred_llh[2 * row + 1] = centroids[2 * best + 1];
  }
  if (lse) lse[row] = gmax + logf(gsum);
}

static size_t fwd_smem_bytes() { return sizeof(FwdSmem) + 1024; }

template <int KTOP>
static int launch_head_fwd(const void* x, const void* W, const float* bias_pad, int B, int C, int D, void* logits,
                           int ldc, int k, void* workspace, const float* centroids, float* topk_val,
                           long long* topk_idx, long long* pred_cell, float* pred_llh, float* lse,
                           cudaStream_t stream) {
  CUtensorMap tm_x, tm_w;
  int rc = make_tmap_bf16_2d(&tm_x, x, D, B, static_cast<uint64_t>(D) * 2, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_w, W, D, C, static_cast<uint64_t>(D) * 2, kBK, kBN);
  if (rc) return rc;
  const int num_m = ceil_div(B, kBM), num_n = ceil_div(C, kBN);
  const int Mpad = num_m * kBM;
  float* pmax = static_cast<float*>(workspace);
  float* psum = pmax + static_cast<size_t>(num_n) * Mpad;
  float* ptopv = psum + static_cast<size_t>(num_n) * Mpad;
  int* ptopi = reinterpret_cast<int*>(ptopv + static_cast<size_t>(num_n) * KTOP * Mpad);
  const int grid = std::min(num_m * num_n, device_sm_count());
  const size_t smem = fwd_smem_bytes();
  if (logits) {
    auto kern = head_fwd_kernel<KTOP, true>;
    GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kFwdThreads, smem, stream>>>(tm_x, tm_w, bias_pad, static_cast<bf16*>(logits), ldc, pmax, psum,
                                              ptopv, ptopi, B, C, D, Mpad);
  } else {
    auto kern = head_fwd_kernel<KTOP, false>;
    GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kFwdThreads, smem, stream>>>(tm_x, tm_w, bias_pad, nullptr, ldc, pmax, psum, ptopv, ptopi, B, C,
                                              D, Mpad);
  }
  GG_LAUNCH_CHECK();
  head_merge_kernel<KTOP><<<ceil_div(B, 128), 128, 0, stream>>>(pmax, psum, ptopv, ptopi, num_n, Mpad, B, k,
                                                               centroids, topk_val, topk_idx, pred_cell,
                                                               pred_llh, lse);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_head_fwd_workspace_bytes(int B, int C, int k) {
  const int ktop = k <= 5 ? 5 : 8;
  const size_t num_n = ceil_div(C, kBN), Mpad = static_cast<size_t>(ceil_div(B, kBM)) * kBM;
  return num_n * Mpad * (2 + 2 * ktop) * sizeof(float);
}

extern "C" int gg_head_logits_ld(int C) { return ceil_div(C, 64) * 64; }
extern "C" int gg_head_bias_pad(int C) { return ceil_div(C, kBN) * kBN; }

extern "C" int gg_head_fwd(const void* x_bf16, const void* w_bf16, const float* bias_pad, int B, int C, int D,
                           void* logits_bf16, int ldc, int k, void* workspace, const float* centroids,
                           float* topk_val, long long* topk_idx, long long* pred_cell, float* pred_llh, float* lse,
                           gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && D > 0, GG_ERR_ARG, "gg_head_fwd: empty problem B=%d C=%d D=%d", B, C, D);
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "gg_head_fwd: embed dim D=%d must be a multiple of 8 (16-byte TMA pitch)", D);
  GG_CHECK(k >= 1 && k <= 8 && k <= C, GG_ERR_ARG, "gg_head_fwd: num_candidates k=%d must be in [1, min(8, C)]", k);
  GG_CHECK(!logits_bf16 || (ldc >= C && ldc % 8 == 0), GG_ERR_ARG, "gg_head_fwd: ldc=%d must be >= C and a multiple of 8", ldc);
  GG_CHECK(x_bf16 && w_bf16 && bias_pad && workspace && topk_val && topk_idx, GG_ERR_ARG, "gg_head_fwd: null pointer");
  GG_CHECK(!pred_llh || centroids, GG_ERR_ARG, "gg_head_fwd: pred_llh needs the centroid table");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (k <= 5)
    return launch_head_fwd<5>(x_bf16, w_bf16, bias_pad, B, C, D, logits_bf16, ldc, k, workspace, centroids, topk_val,
                              topk_idx, pred_cell, pred_llh, lse, s);
  return launch_head_fwd<8>(x_bf16, w_bf16, bias_pad, B, C, D, logits_bf16, ldc, k, workspace, centroids, topk_val,
                            topk_idx, pred_cell, pred_llh, lse, s);
}
