// Plain tcgen05 GEMM for the two dense contractions next to the head that are NOT on the BASELINE configs' step:
//
//   gg_head_dx      dx = dlogits . W  (autograd of models/super_guessr.py:354 w.r.t. its input, reached from
//                   main_coordinator_idun_s3.py:423 whenever the encoder is trained: TinyViT's last stage, CLIP's last
//                   layer, super_guessr.py:127-153), times the mean's 1/V broadcast to the V headings (:347).
//                   A = dlogits (B, C) K-major as the loss kernel left it; B operand = the forward's bf16 W (C, D)
//                   consumed MN-major (geocell = contraction index is its slow index): no transposed copy.
//   gg_linear_bf16  y = a W^T + b for nn.Linear-shaped weights (N, K), both operands K-major: the projections of the
//                   `hierarchical=True` heading fusion (super_guessr.py:89-99,340-345), fed with hi/mid/lo split
//                   operands (hier.cu) so that the fp32 module is reproduced to fp32 accuracy.
//
// One 128 x 256 output tile per CTA (grid = tiles: both problems are a few hundred tiles at most), warp-specialised:
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (thread = output row).  4 stages of 48 KB.
#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kGM = 128, kGN = 256, kGK = 64, kGStages = 4;
constexpr int kGemmThreads = 192;
constexpr uint32_t kGStageA = kGM * kGK * 2;   // 16 KB
constexpr uint32_t kGStageB = kGN * kGK * 2;   // 32 KB
constexpr uint32_t kGAtom = kGK * 128;         // MN-major B: [64 k][64 n] atoms of 8 KB

struct GemmSmem {
  uint8_t a[kGStages][kGStageA];
  uint8_t b[kGStages][kGStageB];
  uint64_t full[kGStages];
  uint64_t empty[kGStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// out[(row * V + v) * ldo + col] = (acc + bias[col]) * scale   for v < V
template <bool B_MN_MAJOR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tile_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const float* __restrict__ bias, int M, int N, int K, float scale_in,
                 const float* __restrict__ grad_scale, float* __restrict__ out, int ldo, int V) {
  extern __shared__ uint8_t smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = (N + kGN - 1) / kGN;
  const int m0 = (blockIdx.x / num_n) * kGM, n0 = (blockIdx.x % num_n) * kGN;
  const int num_k = (K + kGK - 1) / kGK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < kGStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sm.full[s], kGStageA + kGStageB);
        tma_load_2d(sm.a[s], &tm_a, &sm.full[s], kb * kGK, m0);
        if (B_MN_MAJOR) {
#pragma unroll
          for (int i = 0; i < kGN / 64; ++i) tma_load_2d(sm.b[s] + i * kGAtom, &tm_b, &sm.full[s], n0 + 64 * i, kb * kGK);
        } else {
          tma_load_2d(sm.b[s], &tm_b, &sm.full[s], kb * kGK, n0);
        }
        if (++s == kGStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGM, kGN, 0, B_MN_MAJOR ? 1 : 0);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&sm.full[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.a[s]), b0 = smem_u32(sm.b[s]);
#pragma unroll
        for (int k = 0; k < kGK / 16; ++k) {
          const uint64_t da = umma_desc_sw128(a0 + k * 32, 16, 1024);
          // MN-major: 16 k-rows of 128 B per MMA, 64-column atoms kGAtom apart, 8-row groups 1024 B apart
          const uint64_t db = B_MN_MAJOR ? umma_desc_sw128(b0 + k * 2048, kGAtom, 1024) : umma_desc_sw128(b0 + k * 32, 16, 1024);
          umma_f16(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&sm.empty[s]);
        if (++s == kGStages) { s = 0; ph ^= 1; }
      }
      umma_commit(&sm.acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    const float scale = grad_scale ? scale_in * __ldg(grad_scale) : scale_in;
    mbar_wait(&sm.acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < kGN / 32; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= N) break;  // CTA-uniform
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = col0 + 4 * q;
          if (col + 4 <= N) {
            float4 o;
            o.x = __uint_as_float(r[4 * q + 0]);
            o.y = __uint_as_float(r[4 * q + 1]);
            o.z = __uint_as_float(r[4 * q + 2]);
            o.w = __uint_as_float(r[4 * q + 3]);
            if (bias) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + col));
              o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
            }
            o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
            for (int v = 0; v < V; ++v)
              *reinterpret_cast<float4*>(out + (static_cast<size_t>(row) * V + v) * ldo + col) = o;
          } else {
            for (int i = 0; i < 4 && col + i < N; ++i) {
              const float val = (__uint_as_float(r[4 * q + i]) + (bias ? __ldg(bias + col + i) : 0.f)) * scale;
              for (int v = 0; v < V; ++v) out[(static_cast<size_t>(row) * V + v) * ldo + col + i] = val;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <bool B_MN_MAJOR>
static int launch_gemm(const void* a, int lda, const void* b, int ldb, const float* bias, int M, int N, int K, float scale,
                       const float* grad_scale, float* out, int ldo, int V, cudaStream_t s) {
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap_bf16_2d(&tm_a, a, K, M, static_cast<uint64_t>(lda) * 2, kGK, kGM);
  if (rc) return rc;
  if (B_MN_MAJOR)  // b is (K, N) row-major: inner = N, rows = K, 64 x 64 boxes
    rc = make_tmap_bf16_2d(&tm_b, b, N, K, static_cast<uint64_t>(ldb) * 2, 64, kGK);
  else             // b is (N, K) row-major
    rc = make_tmap_bf16_2d(&tm_b, b, K, N, static_cast<uint64_t>(ldb) * 2, kGK, kGN);
  if (rc) return rc;
  const int tiles = ceil_div(M, kGM) * ceil_div(N, kGN);
  const size_t smem = sizeof(GemmSmem) + 1024;
  auto kern = gemm_tile_kernel<B_MN_MAJOR>;
  if (int e = set_max_dynamic_smem_once(kern, smem)) return e;
  kern<<<tiles, kGemmThreads, smem, s>>>(tm_a, tm_b, bias, M, N, K, scale, grad_scale, out, ldo, V);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg

using namespace gg;

extern "C" int gg_head_dx(const void* dlogits_bf16, int ldc, const void* w_bf16, int w_ld, int B, int C, int D, float scale,
                          const float* grad_scale, int V, float* demb, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && D > 0 && V >= 1, GG_ERR_ARG, "gg_head_dx: bad sizes B=%d C=%d D=%d V=%d", B, C, D, V);
  GG_CHECK(dlogits_bf16 && w_bf16 && demb, GG_ERR_ARG, "gg_head_dx: null pointer");
  GG_CHECK(ldc >= C && ldc % 8 == 0 && D % 8 == 0 && w_ld >= D && w_ld % 8 == 0, GG_ERR_ARG,
           "gg_head_dx: ldc=%d / D=%d / w_ld=%d must be multiples of 8", ldc, D, w_ld);
  // K extent = C: dlogits' pad columns (>= C) and W's rows beyond C are zero-filled by the TMA engine
  return launch_gemm<true>(dlogits_bf16, ldc, w_bf16, w_ld, nullptr, B, D, C, scale / static_cast<float>(V), grad_scale,
                           demb, D, V, static_cast<cudaStream_t>(stream));
}

extern "C" int gg_linear_bf16(const void* a_bf16, int lda, const void* w_bf16, int ldw, const float* bias, int M, int N,
                              int K, float* out, int ldo, gg_stream_t stream) {
  GG_CHECK(M > 0 && N > 0 && K > 0, GG_ERR_ARG, "gg_linear_bf16: bad sizes M=%d N=%d K=%d", M, N, K);
  GG_CHECK(a_bf16 && w_bf16 && out, GG_ERR_ARG, "gg_linear_bf16: null pointer");
  GG_CHECK(K % 8 == 0 && lda >= K && lda % 8 == 0 && ldw >= K && ldw % 8 == 0 && ldo >= N && N % 4 == 0 && ldo % 4 == 0,
           GG_ERR_ARG, "gg_linear_bf16: K / lda / ldw multiples of 8, N / ldo multiples of 4");
  return launch_gemm<false>(a_bf16, lda, w_bf16, ldw, bias, M, N, K, 1.0f, nullptr, out, ldo, 1,
                            static_cast<cudaStream_t>(stream));
}
