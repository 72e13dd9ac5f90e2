// Bandwidth-bound helpers around the GEMMs: heading fusion and operand preparation.
#include "common.cuh"
#include "ptx.cuh"

namespace gg {

// Heading fusion, models/super_guessr.py:347 (`layer_input.mean(dim=1)`, (N,4,C) -> (N,C)) and
// proto_refiner.py:150-151, fused with the bf16 cast that feeds the tensor-core operands.
// emb (B,V,D) fp32 -> x (B, ld) bf16.  split = 0: x[:, :D] = bf16(mean).  split = 1 ("bf16x3",
// fp32-faithful): x = [hi | hi | lo] with hi = bf16(mean), lo = bf16(mean - hi), ld = 3D, to be
// contracted against W' = [hi | lo | hi] so that x.W ~= hi.hi + hi.lo + lo.hi.
// Optionally also writes ||hi||^2 per row (prototype retrieval needs the query norms).
template <int SPLIT>
__device__ __forceinline__ void fuse_headings_row(const float* __restrict__ emb, bf16* __restrict__ x, int row, int V,
                                                  int D, int ld, float* __restrict__ sqnorm) {
  const float inv = 1.0f / static_cast<float>(V);
  const float4* src = reinterpret_cast<const float4*>(emb + static_cast<size_t>(row) * V * D);
  const int d4 = D >> 2;
  float nrm = 0.f;
  for (int i = threadIdx.x; i < d4; i += blockDim.x) {
    float4 acc = __ldcs(src + i);
    for (int v = 1; v < V; ++v) {
      const float4 t = __ldcs(src + static_cast<size_t>(v) * d4 + i);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (V > 1) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
    uint2 hi;
    hi.x = pack_bf16x2(acc.x, acc.y);
    hi.y = pack_bf16x2(acc.z, acc.w);
    bf16* dst = x + static_cast<size_t>(row) * ld + 4 * i;
    *reinterpret_cast<uint2*>(dst) = hi;
    const float h0 = __uint_as_float(hi.x << 16), h1 = __uint_as_float(hi.x & 0xffff0000u);
    const float h2 = __uint_as_float(hi.y << 16), h3 = __uint_as_float(hi.y & 0xffff0000u);
    nrm += h0 * h0 + h1 * h1 + h2 * h2 + h3 * h3;
    if (SPLIT) {
      uint2 lo;
      lo.x = pack_bf16x2(acc.x - h0, acc.y - h1);
      lo.y = pack_bf16x2(acc.z - h2, acc.w - h3);
      *reinterpret_cast<uint2*>(dst + D) = hi;
      *reinterpret_cast<uint2*>(dst + 2 * D) = lo;
    }
  }
  if (sqnorm) {
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) sqnorm[row] = v;
    }
  }
}
template <int SPLIT>
__global__ void fuse_headings_kernel(const float* __restrict__ emb, bf16* __restrict__ x, int B, int V, int D, int ld,
                                     float* __restrict__ sqnorm) {
  fuse_headings_row<SPLIT>(emb, x, blockIdx.x, V, D, ld, sqnorm);
}

// W (C,D) fp32 -> bf16 operand.  split = 0: (C,D).  split = 1: (C,3D) = [hi | lo | hi].
// Blocks past the weight matrix zero-pad the bias (b (C) -> bias_pad (Cpad)); b == nullptr: weights only.
template <int SPLIT>
__device__ __forceinline__ void cast_weight_block(int block, const float* __restrict__ w, bf16* __restrict__ out,
                                                  long long n4, int d4, const float* __restrict__ b,
                                                  float* __restrict__ bias_pad, int C, int Cpad, int w_blocks) {
  if (block >= w_blocks) {
    const int i = (block - w_blocks) * blockDim.x + threadIdx.x;
    if (i < Cpad) bias_pad[i] = i < C ? b[i] : 0.f;
    return;
  }
  const long long i = static_cast<long long>(block) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldcs(reinterpret_cast<const float4*>(w) + i);
  uint2 hi;
  hi.x = pack_bf16x2(v.x, v.y);
  hi.y = pack_bf16x2(v.z, v.w);
  if (!SPLIT) {
    reinterpret_cast<uint2*>(out)[i] = hi;
  } else {
    const long long row = i / d4;
    const int c = static_cast<int>(i - row * d4);
    uint2 lo;
    lo.x = pack_bf16x2(v.x - __uint_as_float(hi.x << 16), v.y - __uint_as_float(hi.x & 0xffff0000u));
    lo.y = pack_bf16x2(v.z - __uint_as_float(hi.y << 16), v.w - __uint_as_float(hi.y & 0xffff0000u));
    uint2* o = reinterpret_cast<uint2*>(out) + row * 3 * d4;
    o[c] = hi;
    o[d4 + c] = lo;
    o[2 * d4 + c] = hi;
  }
}
template <int SPLIT>
__global__ void cast_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, long long n4, int d4,
                                   const float* __restrict__ b, float* __restrict__ bias_pad, int C, int Cpad,
                                   int w_blocks) {
  cast_weight_block<SPLIT>(blockIdx.x, w, out, n4, d4, b, bias_pad, C, Cpad, w_blocks);
}

// Heading fusion and weight cast of one training step in ONE launch: two short HBM-bound kernels back to back
// each pay their own ramp and tail (37 us for 153 MB, profiles/r01g); as one grid the cast blocks fill in behind
// the fusion blocks.  Blocks [0, B) fuse one row each, the rest cast 256 float4 of W each / pad the bias.
template <int SPLIT>
__global__ void __launch_bounds__(256)
fuse_and_cast_kernel(const float* __restrict__ emb, bf16* __restrict__ x, int B, int V, int D, int ld,
                     const float* __restrict__ w, bf16* __restrict__ w16, long long n4, const float* __restrict__ b,
                     float* __restrict__ bias_pad, int C, int Cpad, int w_blocks) {
  if (static_cast<int>(blockIdx.x) < B) fuse_headings_row<SPLIT>(emb, x, blockIdx.x, V, D, ld, nullptr);
  else cast_weight_block<SPLIT>(blockIdx.x - B, w, w16, n4, D / 4, b, bias_pad, C, Cpad, w_blocks);
}

// squared L2 norm of each bf16 row (prototype bank), one warp per row
__global__ void row_sqnorm_bf16_kernel(const bf16* __restrict__ m, long long rows, int D, float* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(m + row * D);
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < (D >> 3); i += 32) {
    const uint4 v = __ldg(src + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = __uint_as_float(w[j] << 16), b = __uint_as_float(w[j] & 0xffff0000u);
      acc = fmaf(a, a, acc);
      acc = fmaf(b, b, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) out[row] = acc;
}


// Prototype bank builder: one prototype per cluster = mean over the cluster's member locations of
// the location's mean over its V headings -- the reference's Embeddings.generate_embeddings
// (models/proto_refiner.py:461-517: per member `vec.mean(dim=0)`, running fp32 sum in member order,
// `sum / count`; members out of range or without finite coordinates are skipped (:467-472); no valid member ->
// zero vector (:499-515)), with the encoder replaced by the stored embeddings.  One CTA per cluster, threads
// across D (16-byte loads), members walked in list order so the fp32 sum has the reference's order.
__global__ void __launch_bounds__(256)
build_prototypes_kernel(const float* __restrict__ emb, long long L, int V, int D, const long long* __restrict__ member_off,
                        const int* __restrict__ members, const unsigned char* __restrict__ valid,
                        bf16* __restrict__ bank, float* __restrict__ bank_f32, int* __restrict__ count_out) {
  const long long p = blockIdx.x;
  const long long m0 = member_off[p], m1 = member_off[p + 1];
  const int d4 = D >> 2;
  const float inv_v = 1.0f / static_cast<float>(V);
  int count = 0;
  for (int i = threadIdx.x; i < d4; i += blockDim.x) {
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    int n = 0;
    for (long long m = m0; m < m1; ++m) {
      const long long idx = members[m];
      if (idx < 0 || idx >= L || (valid && !valid[idx])) continue;  // block-uniform
      const float4* src = reinterpret_cast<const float4*>(emb + idx * static_cast<long long>(V) * D) + i;
      float4 acc = __ldg(src);
      for (int v = 1; v < V; ++v) {
        const float4 t = __ldg(src + static_cast<size_t>(v) * d4);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      if (V > 1) { acc.x *= inv_v; acc.y *= inv_v; acc.z *= inv_v; acc.w *= inv_v; }
      sum.x += acc.x; sum.y += acc.y; sum.z += acc.z; sum.w += acc.w;
      ++n;
    }
    if (n > 0) {
      const float c = static_cast<float>(n);
      sum.x /= c; sum.y /= c; sum.z /= c; sum.w /= c;
    }
    count = n;
    uint2 o;
    o.x = pack_bf16x2(sum.x, sum.y);
    o.y = pack_bf16x2(sum.z, sum.w);
    *reinterpret_cast<uint2*>(bank + p * D + 4 * i) = o;
    if (bank_f32) *reinterpret_cast<float4*>(bank_f32 + p * D + 4 * i) = sum;
  }
  if (count_out && threadIdx.x == 0) {
    if (d4 == 0) count = 0;
    count_out[p] = count;
  }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_fuse_headings(const float* emb, void* x_bf16, int B, int V, int D, int split, float* sqnorm,
                                gg_stream_t stream) {
  GG_CHECK(emb && x_bf16 && B > 0 && V > 0 && D > 0, GG_ERR_ARG, "gg_fuse_headings: bad arguments");
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "gg_fuse_headings: D=%d must be a multiple of 8", D);
  GG_CHECK((reinterpret_cast<uintptr_t>(emb) & 15) == 0, GG_ERR_ARG, "gg_fuse_headings: emb must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = D >= 1024 ? 256 : 128;
  if (split)
    fuse_headings_kernel<1><<<B, threads, 0, s>>>(emb, static_cast<bf16*>(x_bf16), B, V, D, 3 * D, sqnorm);
  else
    fuse_headings_kernel<0><<<B, threads, 0, s>>>(emb, static_cast<bf16*>(x_bf16), B, V, D, D, sqnorm);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_prepare_head_weights(const float* w, const float* b, void* w_bf16, float* bias_pad, int C, int D,
                                       int split, gg_stream_t stream) {
  GG_CHECK(w && b && w_bf16 && bias_pad && C > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG,
           "gg_prepare_head_weights: bad arguments (D must be a multiple of 8)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n4 = static_cast<long long>(C) * D / 4;
  const int blocks = static_cast<int>(ceil_div_ll(n4, 256));
  const int Cpad = gg_head_bias_pad(C);
  const int grid = blocks + ceil_div(Cpad, 256);  // the last blocks pad the bias
  if (split)
    cast_weight_kernel<1><<<grid, 256, 0, s>>>(w, static_cast<bf16*>(w_bf16), n4, D / 4, b, bias_pad, C, Cpad, blocks);
  else
    cast_weight_kernel<0><<<grid, 256, 0, s>>>(w, static_cast<bf16*>(w_bf16), n4, D / 4, b, bias_pad, C, Cpad, blocks);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_fuse_and_prepare(const float* emb, void* x_bf16, int B, int V, int D, const float* w, const float* b,
                                   void* w_bf16, float* bias_pad, int C, int split, gg_stream_t stream) {
  GG_CHECK(emb && x_bf16 && w && b && w_bf16 && bias_pad && B > 0 && V > 0 && C > 0 && D > 0, GG_ERR_ARG,
           "gg_fuse_and_prepare: bad arguments");
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "gg_fuse_and_prepare: D=%d must be a multiple of 8", D);
  GG_CHECK((reinterpret_cast<uintptr_t>(emb) & 15) == 0, GG_ERR_ARG, "gg_fuse_and_prepare: emb must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n4 = static_cast<long long>(C) * D / 4;
  const int w_blocks = static_cast<int>(ceil_div_ll(n4, 256));
  const int Cpad = gg_head_bias_pad(C);
  const int grid = B + w_blocks + ceil_div(Cpad, 256);
  if (split)
    fuse_and_cast_kernel<1><<<grid, 256, 0, s>>>(emb, static_cast<bf16*>(x_bf16), B, V, D, 3 * D, w,
                                                 static_cast<bf16*>(w_bf16), n4, b, bias_pad, C, Cpad, w_blocks);
  else
    fuse_and_cast_kernel<0><<<grid, 256, 0, s>>>(emb, static_cast<bf16*>(x_bf16), B, V, D, D, w,
                                                 static_cast<bf16*>(w_bf16), n4, b, bias_pad, C, Cpad, w_blocks);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_cast_bf16(const float* src, void* dst_bf16, long long n, gg_stream_t stream) {
  GG_CHECK(src && dst_bf16 && n > 0 && n % 4 == 0, GG_ERR_ARG, "gg_cast_bf16: n must be a positive multiple of 4");
  const int blocks = static_cast<int>(ceil_div_ll(n / 4, 256));
  cast_weight_kernel<0><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<bf16*>(dst_bf16), n / 4, 1,
                                                                               nullptr, nullptr, 0, 0, blocks);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_row_sqnorm_bf16(const void* m_bf16, long long rows, int D, float* out, gg_stream_t stream) {
  GG_CHECK(m_bf16 && out && rows > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG, "gg_row_sqnorm_bf16: bad arguments");
  row_sqnorm_bf16_kernel<<<static_cast<int>(ceil_div_ll(rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(m_bf16), rows, D, out);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_build_prototypes(const float* emb, long long L, int V, int D, const long long* member_off,
                                   const int* members, const unsigned char* valid, long long P, void* bank_bf16,
                                   float* bank_f32, int* count, gg_stream_t stream) {
  GG_CHECK(emb && member_off && members && bank_bf16, GG_ERR_ARG, "gg_build_prototypes: null pointer");
  GG_CHECK(L > 0 && V >= 1 && D > 0 && D % 4 == 0, GG_ERR_ARG, "gg_build_prototypes: L=%lld V=%d D=%d (D a multiple of 4)", L,
           V, D);
  GG_CHECK(P >= 0 && P <= 0x7fffffffLL, GG_ERR_ARG, "gg_build_prototypes: P=%lld clusters", P);
  if (P == 0) return GG_OK;
  build_prototypes_kernel<<<static_cast<unsigned int>(P), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      emb, L, V, D, member_off, members, valid, static_cast<bf16*>(bank_bf16), bank_f32, count);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
