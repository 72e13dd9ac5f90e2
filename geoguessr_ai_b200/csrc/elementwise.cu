// Bandwidth-bound helpers around the GEMMs: heading fusion and operand preparation.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

// Heading fusion, models/super_guessr.py:347 (`layer_input.mean(dim=1)`, (N,4,C) -> (N,C)) and
// proto_refiner.py:150-151, fused with the bf16 cast that feeds the tensor-core operands.
// emb (B,V,D) fp32 (or, opt-in, bf16 / fp16: half the bytes over PCIe) -> x (B, ld) bf16.  split = 0:
// x[:, :D] = bf16(mean).  split = 1 ("bf16x3", fp32-faithful): x = [hi | hi | lo] with hi = bf16(mean),
// lo = bf16(mean - hi), ld = 3D, to be contracted against W' = [hi | lo | hi] so that x.W ~= hi.hi + hi.lo + lo.hi.
// Optionally also writes ||hi||^2 per row (prototype retrieval needs the query norms).
//
// An item = 4 consecutive elements of one row.  The kernels are pure streams, bound by how many bytes each SM
// keeps in flight: every thread issues ALL loads of its items (V headings x kFuseItems items, compile-time
// unrolled) before the first add.
constexpr int kInF32 = 0, kInBF16 = 1, kInF16 = 2;
constexpr int kFuseItems = 2;  // items per thread of the flat kernels (x V loads each)
constexpr int kCastItems = 4;  // float4 per thread of the weight cast

template <int IN> struct InVec;
template <> struct InVec<kInF32> {
  typedef float4 raw;
  static __device__ __forceinline__ raw ld(const void* base, size_t item) { return __ldcs(static_cast<const float4*>(base) + item); }
  static __device__ __forceinline__ float4 cvt(const raw& r) { return r; }
};
template <> struct InVec<kInBF16> {
  typedef uint2 raw;
  static __device__ __forceinline__ raw ld(const void* base, size_t item) { return __ldcs(static_cast<const uint2*>(base) + item); }
  static __device__ __forceinline__ float4 cvt(const raw& r) {
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
  }
};
template <> struct InVec<kInF16> {
  typedef uint2 raw;
  static __device__ __forceinline__ raw ld(const void* base, size_t item) { return __ldcs(static_cast<const uint2*>(base) + item); }
  static __device__ __forceinline__ float4 cvt(const raw& r) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

// mean over the headings of item `i` (of d4 per heading row) of sample `row`: V > 0 compile-time, V == 0 run-time
template <int V, int IN>
__device__ __forceinline__ float4 heading_mean(const void* __restrict__ emb, size_t row, int i, int d4, int v_rt) {
  typedef InVec<IN> L;
  const size_t base = row * static_cast<size_t>(V > 0 ? V : v_rt) * d4 + i;
  float4 acc;
  if (V > 0) {
    typename L::raw r[V > 0 ? V : 1];
#pragma unroll
    for (int v = 0; v < V; ++v) r[v] = L::ld(emb, base + static_cast<size_t>(v) * d4);
    acc = L::cvt(r[0]);
#pragma unroll
    for (int v = 1; v < V; ++v) {
      const float4 t = L::cvt(r[v]);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  } else {
    acc = L::cvt(L::ld(emb, base));
    for (int v = 1; v < v_rt; ++v) {
      const float4 t = L::cvt(L::ld(emb, base + static_cast<size_t>(v) * d4));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
  }
  const int vv = V > 0 ? V : v_rt;
  if (vv > 1) {
    const float inv = 1.0f / static_cast<float>(vv);
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
  }
  return acc;
}

// store one fused item; returns ||hi||^2 of its 4 elements
template <int SPLIT>
__device__ __forceinline__ float store_fused(bf16* __restrict__ x, size_t row, int i, int D, int ld, const float4& acc) {
  uint2 hi;
  hi.x = pack_bf16x2(acc.x, acc.y);
  hi.y = pack_bf16x2(acc.z, acc.w);
  bf16* dst = x + row * ld + 4 * i;
  *reinterpret_cast<uint2*>(dst) = hi;
  const float h0 = __uint_as_float(hi.x << 16), h1 = __uint_as_float(hi.x & 0xffff0000u);
  const float h2 = __uint_as_float(hi.y << 16), h3 = __uint_as_float(hi.y & 0xffff0000u);
  if (SPLIT) {
    uint2 lo;
    lo.x = pack_bf16x2(acc.x - h0, acc.y - h1);
    lo.y = pack_bf16x2(acc.z - h2, acc.w - h3);
    *reinterpret_cast<uint2*>(dst + D) = hi;
    *reinterpret_cast<uint2*>(dst + 2 * D) = lo;
    // the split operand stands for hi + lo: that vector's norm
    const float e0 = h0 + __uint_as_float(lo.x << 16), e1 = h1 + __uint_as_float(lo.x & 0xffff0000u);
    const float e2 = h2 + __uint_as_float(lo.y << 16), e3 = h3 + __uint_as_float(lo.y & 0xffff0000u);
    return e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
  }
  return h0 * h0 + h1 * h1 + h2 * h2 + h3 * h3;
}

// flat form (no row norms): block `block` of 256 threads handles items [block * 256 * kFuseItems, +256 * kFuseItems)
template <int V, int SPLIT, int IN>
__device__ __forceinline__ void fuse_items_block(int block, const void* __restrict__ emb, bf16* __restrict__ x,
                                                 long long n_items, int v_rt, int D, int ld) {
  const int d4 = D >> 2;
  const unsigned int i0 = static_cast<unsigned int>(block) * (256 * kFuseItems) + threadIdx.x;  // n_items < 2^31 (host)
  const unsigned int n = static_cast<unsigned int>(n_items), ud4 = static_cast<unsigned int>(d4);
  float4 acc[kFuseItems];
  unsigned int row[kFuseItems], col[kFuseItems];
#pragma unroll
  for (int u = 0; u < kFuseItems; ++u) {
    const unsigned int idx = i0 + u * 256;
    row[u] = idx / ud4;
    col[u] = idx - row[u] * ud4;
    if (idx < n) acc[u] = heading_mean<V, IN>(emb, row[u], static_cast<int>(col[u]), d4, v_rt);
  }
#pragma unroll
  for (int u = 0; u < kFuseItems; ++u)
    if (i0 + u * 256 < n) store_fused<SPLIT>(x, row[u], static_cast<int>(col[u]), D, ld, acc[u]);
}
template <int V, int SPLIT, int IN>
__global__ void __launch_bounds__(256)
fuse_flat_kernel(const void* __restrict__ emb, bf16* __restrict__ x, long long n_items, int v_rt, int D, int ld) {
  fuse_items_block<V, SPLIT, IN>(blockIdx.x, emb, x, n_items, v_rt, D, ld);
}

// row form (with ||x||^2): one CTA per row
template <int V, int SPLIT, int IN>
__global__ void __launch_bounds__(256)
fuse_rows_kernel(const void* __restrict__ emb, bf16* __restrict__ x, int v_rt, int D, int ld, float* __restrict__ sqnorm) {
  const int row = blockIdx.x, d4 = D >> 2;
  float nrm = 0.f;
  for (int i = threadIdx.x; i < d4; i += blockDim.x)
    nrm += store_fused<SPLIT>(x, row, i, D, ld, heading_mean<V, IN>(emb, row, i, d4, v_rt));
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

// W (C,D) fp32 -> bf16 operand.  split = 0: (C,D).  split = 1: (C,3D) = [hi | lo | hi].
// Block `block` casts float4 [block * 256 * kCastItems, +256 * kCastItems); blocks past the weight matrix zero-pad
// the bias (b (C) -> bias_pad (Cpad)); b == nullptr: weights only.
template <int SPLIT>
__device__ __forceinline__ void cast_weight_block(int block, const float* __restrict__ w, bf16* __restrict__ out,
                                                  long long n4, int d4, const float* __restrict__ b,
                                                  float* __restrict__ bias_pad, int C, int Cpad, int w_blocks) {
  if (block >= w_blocks) {
    const int i = (block - w_blocks) * blockDim.x + threadIdx.x;
    if (i < Cpad) bias_pad[i] = i < C ? b[i] : 0.f;
    return;
  }
  const long long i0 = static_cast<long long>(block) * (256 * kCastItems) + threadIdx.x;
  float4 v[kCastItems];
#pragma unroll
  for (int u = 0; u < kCastItems; ++u)
    if (i0 + u * 256 < n4) v[u] = __ldcs(reinterpret_cast<const float4*>(w) + i0 + u * 256);
#pragma unroll
  for (int u = 0; u < kCastItems; ++u) {
    const long long i = i0 + u * 256;
    if (i >= n4) continue;
    uint2 hi;
    hi.x = pack_bf16x2(v[u].x, v[u].y);
    hi.y = pack_bf16x2(v[u].z, v[u].w);
    if (!SPLIT) {
      reinterpret_cast<uint2*>(out)[i] = hi;
    } else {
      const long long row = i / d4;
      const int c = static_cast<int>(i - row * d4);
      uint2 lo;
      lo.x = pack_bf16x2(v[u].x - __uint_as_float(hi.x << 16), v[u].y - __uint_as_float(hi.x & 0xffff0000u));
      lo.y = pack_bf16x2(v[u].z - __uint_as_float(hi.y << 16), v[u].w - __uint_as_float(hi.y & 0xffff0000u));
      uint2* o = reinterpret_cast<uint2*>(out) + row * 3 * d4;
      o[c] = hi;
      o[d4 + c] = lo;
      o[2 * d4 + c] = hi;
    }
  }
}
template <int SPLIT>
__global__ void __launch_bounds__(256)
cast_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, long long n4, int d4,
                   const float* __restrict__ b, float* __restrict__ bias_pad, int C, int Cpad, int w_blocks) {
  cast_weight_block<SPLIT>(blockIdx.x, w, out, n4, d4, b, bias_pad, C, Cpad, w_blocks);
}

// Heading fusion and weight cast of one training step in ONE launch: two short HBM-bound kernels back to back
// each pay their own ramp and tail; as one grid the cast blocks fill in behind the fusion blocks.  Blocks
// [0, fuse_blocks) fuse 256 * kFuseItems items each, the rest cast 256 * kCastItems float4 of W each / pad the bias.
template <int V, int SPLIT, int IN>
__global__ void __launch_bounds__(256)
fuse_and_cast_kernel(const void* __restrict__ emb, bf16* __restrict__ x, long long n_items, int v_rt, int D, int ld,
                     int fuse_blocks, const float* __restrict__ w, bf16* __restrict__ w16, long long n4,
                     const float* __restrict__ b, float* __restrict__ bias_pad, int C, int Cpad, int w_blocks) {
  if (static_cast<int>(blockIdx.x) < fuse_blocks) fuse_items_block<V, SPLIT, IN>(blockIdx.x, emb, x, n_items, v_rt, D, ld);
  else cast_weight_block<SPLIT>(blockIdx.x - fuse_blocks, w, w16, n4, D / 4, b, bias_pad, C, Cpad, w_blocks);
}

// squared L2 norm of each bf16 row (prototype bank), one warp per row.  split: the row is [hi | lo | hi] (3D
// entries) and stands for hi + lo.
__global__ void row_sqnorm_bf16_kernel(const bf16* __restrict__ m, long long rows, int D, int split,
                                       float* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(m + row * (split ? 3 * D : D));
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < (D >> 3); i += 32) {
    const uint4 v = __ldg(src + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t l[4] = {0u, 0u, 0u, 0u};
    if (split) {
      const uint4 lo = __ldg(src + (D >> 3) + i);
      l[0] = lo.x; l[1] = lo.y; l[2] = lo.z; l[3] = lo.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = __uint_as_float(w[j] << 16) + __uint_as_float(l[j] << 16);
      const float b = __uint_as_float(w[j] & 0xffff0000u) + __uint_as_float(l[j] & 0xffff0000u);
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

// dispatch on (V in {1, 4, other}, split, input dtype)
#define GG_FUSE_DISPATCH(LAUNCH)                                                      \
  do {                                                                                \
    if (in_dtype == kInF32) { GG_FUSE_V(LAUNCH, kInF32); }                            \
    else if (in_dtype == kInBF16) { GG_FUSE_V(LAUNCH, kInBF16); }                     \
    else { GG_FUSE_V(LAUNCH, kInF16); }                                               \
  } while (0)
#define GG_FUSE_V(LAUNCH, IN)                                                         \
  do {                                                                                \
    if (V == 4) { if (split) LAUNCH(4, 1, IN); else LAUNCH(4, 0, IN); }               \
    else if (V == 1) { if (split) LAUNCH(1, 1, IN); else LAUNCH(1, 0, IN); }          \
    else { if (split) LAUNCH(0, 1, IN); else LAUNCH(0, 0, IN); }                      \
  } while (0)

static int check_fuse_args(const char* who, const void* emb, int in_dtype, int B, int V, int D) {
  GG_CHECK(emb && B > 0 && V > 0 && D > 0, GG_ERR_ARG, "%s: bad arguments", who);
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "%s: D=%d must be a multiple of 8", who, D);
  GG_CHECK(in_dtype >= 0 && in_dtype <= 2, GG_ERR_ARG, "%s: in_dtype=%d (0 fp32, 1 bf16, 2 fp16)", who, in_dtype);
  GG_CHECK((reinterpret_cast<uintptr_t>(emb) & 15) == 0, GG_ERR_ARG, "%s: emb must be 16-byte aligned", who);
  GG_CHECK(static_cast<long long>(B) * (D / 4) < (1ll << 31), GG_ERR_ARG, "%s: B * D / 4 overflows int32", who);
  return GG_OK;
}

extern "C" int gg_fuse_headings(const void* emb, int in_dtype, void* x_bf16, int B, int V, int D, int split,
                                float* sqnorm, gg_stream_t stream) {
  if (int e = check_fuse_args("gg_fuse_headings", emb, in_dtype, B, V, D)) return e;
  GG_CHECK(x_bf16, GG_ERR_ARG, "gg_fuse_headings: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* x = static_cast<bf16*>(x_bf16);
  const int ld = split ? 3 * D : D;
  if (sqnorm) {
    const int threads = D >= 1024 ? 256 : 128;
#define GG_LAUNCH_ROWS(VV, SP, IN) fuse_rows_kernel<VV, SP, IN><<<B, threads, 0, s>>>(emb, x, V, D, ld, sqnorm)
    GG_FUSE_DISPATCH(GG_LAUNCH_ROWS);
#undef GG_LAUNCH_ROWS
  } else {
    const long long n_items = static_cast<long long>(B) * (D / 4);
    const int blocks = static_cast<int>(ceil_div_ll(n_items, 256 * kFuseItems));
#define GG_LAUNCH_FLAT(VV, SP, IN) fuse_flat_kernel<VV, SP, IN><<<blocks, 256, 0, s>>>(emb, x, n_items, V, D, ld)
    GG_FUSE_DISPATCH(GG_LAUNCH_FLAT);
#undef GG_LAUNCH_FLAT
  }
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_prepare_head_weights(const float* w, const float* b, void* w_bf16, float* bias_pad, int C, int D,
                                       int split, gg_stream_t stream) {
  GG_CHECK(w && b && w_bf16 && bias_pad && C > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG,
           "gg_prepare_head_weights: bad arguments (D must be a multiple of 8)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n4 = static_cast<long long>(C) * D / 4;
  const int blocks = static_cast<int>(ceil_div_ll(n4, 256 * kCastItems));
  const int Cpad = gg_head_bias_pad(C);
  const int grid = blocks + ceil_div(Cpad, 256);  // the last blocks pad the bias
  if (split)
    cast_weight_kernel<1><<<grid, 256, 0, s>>>(w, static_cast<bf16*>(w_bf16), n4, D / 4, b, bias_pad, C, Cpad, blocks);
  else
    cast_weight_kernel<0><<<grid, 256, 0, s>>>(w, static_cast<bf16*>(w_bf16), n4, D / 4, b, bias_pad, C, Cpad, blocks);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_fuse_and_prepare(const void* emb, int in_dtype, void* x_bf16, int B, int V, int D, const float* w,
                                   const float* b, void* w_bf16, float* bias_pad, int C, int split,
                                   gg_stream_t stream) {
  if (int e = check_fuse_args("gg_fuse_and_prepare", emb, in_dtype, B, V, D)) return e;
  GG_CHECK(x_bf16 && w && b && w_bf16 && bias_pad && C > 0, GG_ERR_ARG, "gg_fuse_and_prepare: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* x = static_cast<bf16*>(x_bf16);
  bf16* w16 = static_cast<bf16*>(w_bf16);
  const int ld = split ? 3 * D : D;
  const long long n_items = static_cast<long long>(B) * (D / 4);
  const int fuse_blocks = static_cast<int>(ceil_div_ll(n_items, 256 * kFuseItems));
  const long long n4 = static_cast<long long>(C) * D / 4;
  const int w_blocks = static_cast<int>(ceil_div_ll(n4, 256 * kCastItems));
  const int Cpad = gg_head_bias_pad(C);
  const int grid = fuse_blocks + w_blocks + ceil_div(Cpad, 256);
#define GG_LAUNCH_BOTH(VV, SP, IN)                                                                                   \
  fuse_and_cast_kernel<VV, SP, IN><<<grid, 256, 0, s>>>(emb, x, n_items, V, D, ld, fuse_blocks, w, w16, n4, b, bias_pad, \
                                                        C, Cpad, w_blocks)
  GG_FUSE_DISPATCH(GG_LAUNCH_BOTH);
#undef GG_LAUNCH_BOTH
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_cast_bf16(const float* src, void* dst_bf16, long long rows, int D, int split, gg_stream_t stream) {
  GG_CHECK(src && dst_bf16 && rows > 0 && D > 0 && D % 4 == 0, GG_ERR_ARG,
           "gg_cast_bf16: rows=%lld D=%d (D a positive multiple of 4)", rows, D);
  const long long n4 = rows * D / 4;
  const int blocks = static_cast<int>(ceil_div_ll(n4, 256 * kCastItems));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (split)
    cast_weight_kernel<1><<<blocks, 256, 0, s>>>(src, static_cast<bf16*>(dst_bf16), n4, D / 4, nullptr, nullptr, 0, 0, blocks);
  else
    cast_weight_kernel<0><<<blocks, 256, 0, s>>>(src, static_cast<bf16*>(dst_bf16), n4, D / 4, nullptr, nullptr, 0, 0, blocks);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_row_sqnorm_bf16(const void* m_bf16, long long rows, int D, int split, float* out, gg_stream_t stream) {
  GG_CHECK(m_bf16 && out && rows > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG, "gg_row_sqnorm_bf16: bad arguments");
  row_sqnorm_bf16_kernel<<<static_cast<int>(ceil_div_ll(rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(m_bf16), rows, D, split, out);
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
