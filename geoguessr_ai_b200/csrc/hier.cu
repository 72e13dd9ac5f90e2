// `hierarchical=True` heading fusion, the reference's alternative to the heading mean:
//   models/super_guessr.py:340-345    layer_input = pos_encoder(layer_input); output = self_attn(x, x, x)[0][:, 0]
//   models/layers/positional_encoder.py:21-44    z[b, t] = x[b, t] + PE[b]   (the table is indexed by the BATCH row)
//   models/super_guessr.py:89-99      nn.MultiheadAttention(D, 16, dropout=0.1, batch_first=True), eval mode
// The three projections are dense contractions and run on the tensor cores (gg_linear_bf16, gemm.cu).  The module is
// fp32, so its operands are split into three bf16 terms (hi + mid + lo carries 24 mantissa bits) and contracted as six
// products -- hh, hm, mh, hl, lh, mm; what is dropped is below 2^-26 relative -- by laying the terms side by side along
// K: activations [h|h|m|h|l|m], weights [h|m|h|l|h|m] (K' = 6 D).  This file holds the two small kernels around the
// GEMMs: the split (with the positional offset folded in) and the attention over the V headings for token 0.
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// role 0 (activation): [h|h|m|h|l|m]; role 1 (weight): [h|m|h|l|h|m].  One thread per element pair.
__device__ __forceinline__ void store_split3(bf16* __restrict__ dst_row, int D, int col, int role, float x0, float x1) {
  const float h0 = bf16_round(x0), h1 = bf16_round(x1);
  const float r0 = x0 - h0, r1 = x1 - h1;
  const float m0 = bf16_round(r0), m1 = bf16_round(r1);
  const float l0 = r0 - m0, l1 = r1 - m1;
  const uint32_t H = pack_bf16x2(h0, h1), Mv = pack_bf16x2(m0, m1), L = pack_bf16x2(l0, l1);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst_row + col);
  const int s = D / 2;  // section stride in 32-bit words
  if (role == 0) { d[0] = H; d[s] = H; d[2 * s] = Mv; d[3 * s] = H; d[4 * s] = L; d[5 * s] = Mv; }
  else           { d[0] = H; d[s] = Mv; d[2 * s] = H; d[3 * s] = L; d[4 * s] = H; d[5 * s] = Mv; }
}

__global__ void split3_kernel(const float* __restrict__ src, long long rows, int D, int role, const float* __restrict__ pe,
                              int V, bf16* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // element pair
  const int d2 = D / 2;
  if (i >= rows * d2) return;
  const long long row = i / d2;
  const int col = static_cast<int>(i - row * d2) * 2;
  float2 v = *reinterpret_cast<const float2*>(src + row * D + col);
  if (pe) {  // positional_encoder.py:44: token_embedding + pos_encoding[:B] -- row b of the table for every heading of b
    const float2 p = *reinterpret_cast<const float2*>(pe + (row / V) * D + col);
    v.x += p.x;
    v.y += p.y;
  }
  store_split3(dst + row * 6 * D, D, col, role, v.x, v.y);
}

// One warp per (sample b, head h): scores of token 0 against the V headings, softmax, weighted sum of the values;
// written as the split activation operand of the output projection.  qkv (B*V, 3D) fp32 = [q | k | v] per token.
__global__ void hier_attention_kernel(const float* __restrict__ qkv, int B, int V, int D, int heads,
                                      bf16* __restrict__ ctx_split) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * heads) return;
  const int b = w / heads, h = w - b * heads;
  const int dh = D / heads;
  const float* q0 = qkv + static_cast<size_t>(b) * V * 3 * D + h * dh;
  const float inv_sqrt = rsqrtf(static_cast<float>(dh));
  constexpr int kMaxV = 8;
  float s[kMaxV];
  float mx = -CUDART_INF_F;
  for (int t = 0; t < V; ++t) {
    const float* kt = qkv + (static_cast<size_t>(b) * V + t) * 3 * D + D + h * dh;
    float acc = 0.f;
    for (int i = lane; i < dh; i += 32) acc = fmaf(q0[i], kt[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    s[t] = acc * inv_sqrt;
    mx = fmaxf(mx, s[t]);
  }
  float den = 0.f;
  for (int t = 0; t < V; ++t) { s[t] = expf(s[t] - mx); den += s[t]; }
  const float inv = 1.0f / den;
  for (int i = 2 * lane; i < dh; i += 64) {  // dh is even (D % 32 == 0)
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < V; ++t) {
      const float* vt = qkv + (static_cast<size_t>(b) * V + t) * 3 * D + 2 * D + h * dh;
      a0 = fmaf(s[t] * inv, vt[i], a0);
      a1 = fmaf(s[t] * inv, vt[i + 1], a1);
    }
    store_split3(ctx_split + static_cast<size_t>(b) * 6 * D, D, h * dh + i, 0, a0, a1);
  }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_split3_bf16(const float* src, long long rows, int D, int role, const float* pos_encoding, int V,
                              void* dst_bf16, gg_stream_t stream) {
  GG_CHECK(src && dst_bf16 && rows > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG, "gg_split3_bf16: rows=%lld D=%d (D a multiple of 8)",
           rows, D);
  GG_CHECK((role == 0 || role == 1) && V >= 1, GG_ERR_ARG, "gg_split3_bf16: role=%d V=%d", role, V);
  const long long n = rows * (D / 2);
  split3_kernel<<<static_cast<unsigned int>(ceil_div_ll(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, rows, D, role, pos_encoding, V, static_cast<bf16*>(dst_bf16));
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_hier_attention(const float* qkv, int B, int V, int D, int heads, void* ctx_split_bf16, gg_stream_t stream) {
  GG_CHECK(qkv && ctx_split_bf16 && B > 0 && D > 0 && heads > 0, GG_ERR_ARG, "gg_hier_attention: bad arguments");
  GG_CHECK(V >= 1 && V <= 8, GG_ERR_UNSUPPORTED, "gg_hier_attention: V=%d headings (<= 8)", V);
  GG_CHECK(D % heads == 0 && (D / heads) % 2 == 0 && D % 8 == 0, GG_ERR_ARG,
           "gg_hier_attention: D=%d must split into %d heads of even width", D, heads);
  const int warps = B * heads;
  hier_attention_kernel<<<ceil_div(warps, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      qkv, B, V, D, heads, static_cast<bf16*>(ctx_split_bf16));
  GG_LAUNCH_CHECK();
  return GG_OK;
}
