// Gradient averaging for data-parallel head training over NVLink / NVSwitch peer memory.
//
// The reference leaves this step to DDP / Accelerate (an NCCL ring all-reduce of dW (C,D) and db (C) after
// loss.backward(), main_coordinator_idun_s3.py:423-424).  Here every rank keeps its [dW | db] gradient in a
// symmetric-memory buffer mapped into all peers, and ONE kernel per rank does a two-shot all-reduce in place:
//   shot 1  rank r reads slice r of every rank's buffer over NVLink (16-byte loads, many in flight per thread)
//           and sums them in rank order -- the same order on every rank and for every element, so the result
//           is deterministic and identical everywhere -- times 1 / world;
//   shot 2  it writes the averaged slice r back into every rank's buffer (posted 16-byte peer stores).
// Slice r is read and written by rank r only, so the exchange needs no staging copy.  Per rank (N - 1) / N of
// the buffer crosses NVLink in each direction, once: 26 MB each way for the 51.8 MB head gradient at N = 2,
// 45 MB at N = 8, against the 2 (N - 1) / N a ring moves in 2 (N - 1) latency-bound steps.  The caller brackets
// the launch with symmetric-memory barriers (all gradients written / all slices delivered).
#include <stdio.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kP2PMaxWorld = 8;
constexpr int kP2PThreads = 512;

struct PeerPtrs {
  float4* p[kP2PMaxWorld];
};

// L2-only accesses (no L1 allocation): the lines are written by other GPUs between launches
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4& v) { __stcg(p, v); }

// Scheduling fence: the value is "produced" by an (empty) volatile asm, and volatile asms keep their program
// order.  Placed after a batch of loads it stops the compiler from sinking arithmetic in between them, where the
// in-order issue would stall on the first NVLink round trip with most of the batch still unissued.
__device__ __forceinline__ void pin(float4& v) { asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)); }

template <int WORLD>
__global__ void __launch_bounds__(kP2PThreads)
p2p_allreduce_avg_kernel(const __grid_constant__ PeerPtrs peers, size_t lo, size_t hi, float inv_world) {
  constexpr int kUnroll = WORLD >= 4 ? 2 : 4;  // WORLD x unroll 16-byte loads in flight per thread
  const size_t stride = static_cast<size_t>(gridDim.x) * kP2PThreads;
  size_t base = lo + static_cast<size_t>(blockIdx.x) * kP2PThreads + threadIdx.x;
  // whole rounds: every load of the round first, then the sums (rank order: identical on every rank), then the stores
  for (; base + (kUnroll - 1) * stride < hi; base += stride * kUnroll) {
    float4 v[kUnroll][WORLD];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int r = 0; r < WORLD; ++r) v[u][r] = ld_peer(peers.p[r] + base + u * stride);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int r = 0; r < WORLD; ++r) pin(v[u][r]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float4 acc = v[u][0];
#pragma unroll
      for (int r = 1; r < WORLD; ++r) {
        acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w;
      }
      acc.x *= inv_world; acc.y *= inv_world; acc.z *= inv_world; acc.w *= inv_world;
#pragma unroll
      for (int r = 0; r < WORLD; ++r) st_peer(peers.p[r] + base + u * stride, acc);
    }
  }
  for (; base < hi; base += stride) {  // ragged end of the slice
    float4 acc = ld_peer(peers.p[0] + base);
#pragma unroll
    for (int r = 1; r < WORLD; ++r) {
      const float4 t = ld_peer(peers.p[r] + base);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    acc.x *= inv_world; acc.y *= inv_world; acc.z *= inv_world; acc.w *= inv_world;
#pragma unroll
    for (int r = 0; r < WORLD; ++r) st_peer(peers.p[r] + base, acc);
  }
}

// Same exchange through the NVSwitch's multicast object (NVLS): one multimem.ld_reduce pulls slice r from every
// GPU and adds the copies INSIDE the switch, one multimem.st broadcasts the averaged slice to every GPU.  Each
// GPU then sends and receives about one buffer's worth of bytes whatever the number of ranks (the two-shot
// peer version moves 2 (N - 1) / N of it per direction), and the reduced value is computed once -- by the
// switch, in its fixed port order -- so every rank still ends up with identical bits.
__device__ __forceinline__ float4 multimem_ld_add(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kP2PThreads)
nvls_allreduce_avg_kernel(float4* __restrict__ mc, size_t lo, size_t hi, float inv_world) {
  constexpr int kUnroll = 8;
  const size_t stride = static_cast<size_t>(gridDim.x) * kP2PThreads;
  size_t base = lo + static_cast<size_t>(blockIdx.x) * kP2PThreads + threadIdx.x;
  for (; base + (kUnroll - 1) * stride < hi; base += stride * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = multimem_ld_add(mc + base + u * stride);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) pin(v[u]);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      v[u].x *= inv_world; v[u].y *= inv_world; v[u].z *= inv_world; v[u].w *= inv_world;
      multimem_st(mc + base + u * stride, v[u]);
    }
  }
  for (; base < hi; base += stride) {
    float4 v = multimem_ld_add(mc + base);
    v.x *= inv_world; v.y *= inv_world; v.z *= inv_world; v.w *= inv_world;
    multimem_st(mc + base, v);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Exchange fused into the dW GEMM (gg_head_bwd's push mode + gg_grad_exchange).  The two kernels above run AFTER the
// GEMM between two host-issued barriers and move the whole buffer then.  Here dW is cut into blocks of 128 geocells
// (+ the block's 128 db entries); block b is reduced by rank b % world, and the GEMM's epilogue stores every tile
// straight into the reducer's staging slab for its source rank (posted stores over NVLink, tile by tile under the
// GEMM; see head_bwd.cu) and counts every delivered tile on the reducer's `ready` counter of the block.  What is left for after
// the GEMM is this kernel: per owned block wait for the last announcements, add the `world` staged copies in rank
// order from LOCAL memory (deterministic, identical everywhere), scale, and write the average into every rank's
// gradient buffer -- multimem.st through the NVSwitch when the multicast mapping is given, else posted peer stores
// -- then tell every rank; it returns when every block of every reducer has landed in this rank's gradient.
// No peer loads anywhere, no host barrier: counters only ever grow (expected values are multiples of a per-launch
// epoch kept in the control region), nothing is reset while a peer could still add.
//
// (A first version averaged the blocks with a second kernel running NEXT TO the GEMM and peer loads; measured on
// 2 x B200 it slowed the GEMM 2-4x -- the co-resident CTAs' system-scope fences and polling stall the SM's memory
// pipeline -- and is gone.)
//
// Control region (symmetric memory, GG_GRAD_CTRL_BYTES per rank, zeroed once): u32 words
//   [0, 1024)     (reserved)
//   [1024, 2048)  ready      delivered column tiles per block (remote adds by every rank's gg_head_bwd: ceil(D / 256) per rank)
//   [2048]        done       exchanged units landed in this rank's copy (remote adds by the reducers)
//   [2049]        exit ticket, [2050] epoch (local)
constexpr int kCtrlReady = 1024, kCtrlDone = 2048, kCtrlExit = 2049, kCtrlEpoch = 2050;
constexpr int kGradBlockRows = 128;
constexpr int kGradMaxParts = 16;  // CTAs that share one block: chosen per launch (grad_parts) so that one wave covers all units
constexpr int kGradThreads = 256;

struct GradPeers {
  float4* grad[kP2PMaxWorld];        // every rank's [dW | db | pad] buffer
  unsigned int* ctrl[kP2PMaxWorld];  // every rank's control region
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_red_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until *p has reached `target` (wrap-safe); traps instead of hanging the GPU if a peer never arrives
__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target, const char* what, int a, int b) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_sys(p) - target) < 0) {
    __nanosleep(100);
    if (clock64() - t0 > 10000000000LL) {
      printf("gg: grad exchange timeout waiting for %s (rank %d, index %d): have %u want %u\n", what, a, b,
             ld_acquire_sys(p), target);
      __trap();
    }
  }
}

// dst[i] (every rank's copy, float4 index dst0 + i) = inv_world * sum_src stage[src * slab4 + src0 + i], i in [0, n)
template <int WORLD, bool NVLS>
__device__ __forceinline__ void reduce_range(const GradPeers& peers, float4* mc, const float4* __restrict__ stage,
                                             size_t slab4, size_t src0, size_t dst0, size_t n, float inv_world) {
  constexpr int kUnroll = WORLD >= 8 ? 2 : (WORLD >= 4 ? 4 : 8);  // WORLD x unroll local 16-byte loads in flight
  size_t i = threadIdx.x;
  auto emit = [&](size_t at, const float4& v) {
    if (NVLS) {
      multimem_st(mc + dst0 + at, v);
    } else {
#pragma unroll
      for (int r = 0; r < WORLD; ++r) st_peer(peers.grad[r] + dst0 + at, v);
    }
  };
  for (; i + (kUnroll - 1) * kGradThreads < n; i += kUnroll * kGradThreads) {
    float4 v[kUnroll][WORLD];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
      for (int r = 0; r < WORLD; ++r) v[u][r] = __ldcs(stage + r * slab4 + src0 + i + u * kGradThreads);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float4 acc = v[u][0];
#pragma unroll
      for (int r = 1; r < WORLD; ++r) {  // rank order: identical on every rank
        acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w;
      }
      acc.x *= inv_world; acc.y *= inv_world; acc.z *= inv_world; acc.w *= inv_world;
      emit(i + u * kGradThreads, acc);
    }
  }
  for (; i < n; i += kGradThreads) {
    float4 acc = __ldcs(stage + src0 + i);
#pragma unroll
    for (int r = 1; r < WORLD; ++r) {
      const float4 t = __ldcs(stage + r * slab4 + src0 + i);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    acc.x *= inv_world; acc.y *= inv_world; acc.z *= inv_world; acc.w *= inv_world;
    emit(i, acc);
  }
}

// ---- sharded AdamW inside the exchange (gg_grad_exchange_adamw) ------------------------------------------------
// The averaged gradient of a block exists only in the reducer's registers: the reducer applies AdamW to ITS blocks of
// the fp32 master weights (torch.optim.AdamW's update, main_coordinator_idun_s3.py:286-291,424) and broadcasts what
// the next forward needs -- the bf16 operand rows (half the bytes of the fp32 gradient) and the fp32 bias -- into
// every rank's operand buffer.  The separate optimizer pass over all of W on every rank and the per-step fp32 -> bf16
// recast of W disappear; each rank keeps moments and current master weights only for the blocks it reduces.
struct AdamwArgs {
  float* w; float* m; float* v;      // master weight (C, D) and moments: rows of this rank's blocks are current
  float* b; float* mb; float* vb;    // bias (C) and moments
  const float* hyper;                // device: lr, beta1, beta2, eps, weight_decay
  long long* step;                   // device: optimizer steps taken so far (incremented by the kernel)
  uint4* w16[kP2PMaxWorld];          // every rank's bf16 operand (C, D)
  float4* bias[kP2PMaxWorld];        // every rank's padded fp32 bias
  uint4* w16_mc;                     // multicast addresses of the two (NVLS)
  float4* bias_mc;
};
struct AdamwStep {
  float lr, beta1, beta2, eps, lr_wd, step_size, inv_bc2_sqrt;
};
__device__ __forceinline__ void adamw_update(float& w, float& m, float& v, float g, const AdamwStep& h) {
  w -= h.lr_wd * w;                       // decoupled weight decay
  m += (1.f - h.beta1) * (g - m);         // lerp(m, g, 1 - beta1)
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) * h.inv_bc2_sqrt + h.eps;
  w -= h.step_size * m / denom;
}
__device__ __forceinline__ void st_peer16(uint4* p, const uint4& v) { __stcg(p, v); }
__device__ __forceinline__ void multimem_st16(uint4* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w) : "memory");
}

// n8 items of 8 floats: staged floats [8 (src8 + i), +8) of every slab -> averaged gradient -> AdamW on master floats
// [8 (dst8 + i), +8) -> 8 bf16 into every rank's operand
template <int WORLD, bool NVLS>
__device__ __forceinline__ void adamw_range(const AdamwArgs& aw, const AdamwStep& h, const float4* __restrict__ stage,
                                            size_t slab4, size_t src8, size_t dst8, size_t n8, float inv_world) {
  constexpr int kU = WORLD >= 8 ? 1 : 2;  // items in flight per thread: (2 WORLD + 6) kU 16-byte loads
  float4* const w4 = reinterpret_cast<float4*>(aw.w);
  float4* const m4 = reinterpret_cast<float4*>(aw.m);
  float4* const v4 = reinterpret_cast<float4*>(aw.v);
  for (size_t i0 = threadIdx.x; i0 < n8; i0 += kU * kGradThreads) {
    float4 a[kU][WORLD], c[kU][WORLD], w0[kU], w1[kU], m0[kU], m1[kU], v0[kU], v1[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const size_t i = i0 + u * kGradThreads;
      if (i < n8) {
#pragma unroll
        for (int r = 0; r < WORLD; ++r) {
          a[u][r] = __ldcs(stage + r * slab4 + 2 * (src8 + i));
          c[u][r] = __ldcs(stage + r * slab4 + 2 * (src8 + i) + 1);
        }
        const size_t at = 2 * (dst8 + i);
        w0[u] = __ldcs(w4 + at); w1[u] = __ldcs(w4 + at + 1);
        m0[u] = __ldcs(m4 + at); m1[u] = __ldcs(m4 + at + 1);
        v0[u] = __ldcs(v4 + at); v1[u] = __ldcs(v4 + at + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const size_t i = i0 + u * kGradThreads;
      if (i >= n8) break;
      float4 g0 = a[u][0], g1 = c[u][0];
#pragma unroll
      for (int r = 1; r < WORLD; ++r) {  // rank order: identical on every rank
        g0.x += a[u][r].x; g0.y += a[u][r].y; g0.z += a[u][r].z; g0.w += a[u][r].w;
        g1.x += c[u][r].x; g1.y += c[u][r].y; g1.z += c[u][r].z; g1.w += c[u][r].w;
      }
      adamw_update(w0[u].x, m0[u].x, v0[u].x, g0.x * inv_world, h); adamw_update(w0[u].y, m0[u].y, v0[u].y, g0.y * inv_world, h);
      adamw_update(w0[u].z, m0[u].z, v0[u].z, g0.z * inv_world, h); adamw_update(w0[u].w, m0[u].w, v0[u].w, g0.w * inv_world, h);
      adamw_update(w1[u].x, m1[u].x, v1[u].x, g1.x * inv_world, h); adamw_update(w1[u].y, m1[u].y, v1[u].y, g1.y * inv_world, h);
      adamw_update(w1[u].z, m1[u].z, v1[u].z, g1.z * inv_world, h); adamw_update(w1[u].w, m1[u].w, v1[u].w, g1.w * inv_world, h);
      const size_t at = 2 * (dst8 + i);
      __stcs(w4 + at, w0[u]); __stcs(w4 + at + 1, w1[u]);
      __stcs(m4 + at, m0[u]); __stcs(m4 + at + 1, m1[u]);
      __stcs(v4 + at, v0[u]); __stcs(v4 + at + 1, v1[u]);
      uint4 o;
      o.x = pack_bf16x2(w0[u].x, w0[u].y); o.y = pack_bf16x2(w0[u].z, w0[u].w);
      o.z = pack_bf16x2(w1[u].x, w1[u].y); o.w = pack_bf16x2(w1[u].z, w1[u].w);
      if (NVLS) {
        multimem_st16(aw.w16_mc + dst8 + i, o);
      } else {
#pragma unroll
        for (int r = 0; r < WORLD; ++r) st_peer16(aw.w16[r] + dst8 + i, o);
      }
    }
  }
}

// the block's bias entries [c0, c0 + 128) (entries >= C: the pad, kept at zero); staged at float4 index src4
template <int WORLD, bool NVLS>
__device__ __forceinline__ void adamw_bias(const AdamwArgs& aw, const AdamwStep& h, const float4* __restrict__ stage,
                                           size_t slab4, size_t src4, int c0, int C, float inv_world) {
  const int i = threadIdx.x;
  if (i >= kGradBlockRows / 4 || c0 + 4 * i >= C) return;
  float4 g = __ldcs(stage + src4 + i);
#pragma unroll
  for (int r = 1; r < WORLD; ++r) {
    const float4 t = __ldcs(stage + r * slab4 + src4 + i);
    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
  }
  const float gs[4] = {g.x * inv_world, g.y * inv_world, g.z * inv_world, g.w * inv_world};
  float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = c0 + 4 * i + e;
    if (c < C) {
      float w = aw.b[c], m = aw.mb[c], v = aw.vb[c];
      adamw_update(w, m, v, gs[e], h);
      aw.b[c] = w; aw.mb[c] = m; aw.vb[c] = v;
      o[e] = w;
    }
  }
  const float4 ov = make_float4(o[0], o[1], o[2], o[3]);
  const size_t at = static_cast<size_t>(c0) / 4 + i;
  if (NVLS) {
    multimem_st(aw.bias_mc + at, ov);
  } else {
#pragma unroll
    for (int r = 0; r < WORLD; ++r) st_peer(aw.bias[r] + at, ov);
  }
}

template <int WORLD, bool NVLS, bool ADAMW>
__global__ void __launch_bounds__(kGradThreads)
grad_exchange_kernel(const __grid_constant__ GradPeers peers, float4* __restrict__ mc_grad,
                     unsigned int* __restrict__ mc_ctrl, const float4* __restrict__ stage, int rank, int C, int D,
                     int nblk, int own_max, float inv_world, int no_wait, int parts,
                     const __grid_constant__ AdamwArgs aw) {
  unsigned int* const ctrl = peers.ctrl[rank];
  __shared__ unsigned int s_epoch;
  __shared__ AdamwStep s_h;
  if (threadIdx.x == 0) {
    s_epoch = ld_acquire_sys(ctrl + kCtrlEpoch);
    if (ADAMW) {
      // every CTA reads the step count before the last one out (below) advances it
      const double t = static_cast<double>(*reinterpret_cast<volatile long long*>(aw.step) + 1);
      const float lr = aw.hyper[0], b1 = aw.hyper[1], b2 = aw.hyper[2];
      s_h.lr = lr; s_h.beta1 = b1; s_h.beta2 = b2; s_h.eps = aw.hyper[3];
      s_h.lr_wd = lr * aw.hyper[4];
      s_h.step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(b1), t)));
      s_h.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(b2), t)));
    }
  }
  __syncthreads();
  const AdamwStep h = s_h;
  const unsigned int epoch = s_epoch + 1u;  // announcements / landed units expected so far = epoch x (per-launch count)
  const int own = nblk > rank ? (nblk - rank + WORLD - 1) / WORLD : 0;  // blocks rank, rank + WORLD, ...
  const size_t row4 = static_cast<size_t>(D) / 4;                       // float4 per geocell row
  const size_t db4 = static_cast<size_t>(C) * row4;                     // float4 offset of db in the gradient buffer
  const size_t rows = static_cast<size_t>(own_max) * kGradBlockRows;    // staged rows per slab
  const size_t slab4 = rows * (static_cast<size_t>(D) + 1) / 4;         // float4 per slab (dW rows, then db entries)
  for (int u = blockIdx.x; u < own * parts; u += gridDim.x) {
    const int j = u / parts, part = u % parts;
    const int b = rank + j * WORLD;
    // every rank's gg_head_bwd adds one per column tile (256 embedding columns) of the block it has delivered
    if (threadIdx.x == 0)
      spin_until(ctrl + kCtrlReady + b, epoch * static_cast<unsigned int>(WORLD * ((D + 255) / 256)), "block", rank, b);
    __syncthreads();
    const int r0 = b * kGradBlockRows, r1 = min(C, r0 + kGradBlockRows);
    const size_t n4 = static_cast<size_t>(r1 - r0) * row4;
    if (ADAMW) {
      const size_t n8 = n4 / 2;  // D is a multiple of 8 in this mode
      const size_t lo = n8 * part / parts, hi = n8 * (part + 1) / parts;
      adamw_range<WORLD, NVLS>(aw, h, stage, slab4, static_cast<size_t>(j) * kGradBlockRows * (row4 / 2) + lo,
                               static_cast<size_t>(r0) * (row4 / 2) + lo, hi - lo, inv_world);
      if (part == 0)
        adamw_bias<WORLD, NVLS>(aw, h, stage, slab4, rows * row4 + static_cast<size_t>(j) * (kGradBlockRows / 4), r0, C,
                                inv_world);
    } else {
      const size_t lo = n4 * part / parts, hi = n4 * (part + 1) / parts;
      reduce_range<WORLD, NVLS>(peers, mc_grad, stage, slab4, static_cast<size_t>(j) * kGradBlockRows * row4 + lo,
                                static_cast<size_t>(r0) * row4 + lo, hi - lo, inv_world);
      if (part == 0)  // the block's db entries (128 floats; the last block's ragged end runs into the pad)
        reduce_range<WORLD, NVLS>(peers, mc_grad, stage, slab4, rows * row4 + static_cast<size_t>(j) * (kGradBlockRows / 4),
                                  db4 + static_cast<size_t>(r0) / 4, (static_cast<size_t>(r1 - r0) + 3) / 4, inv_world);
    }
    __syncthreads();  // every thread's stores are issued
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (NVLS) {
        multimem_red_release_sys(mc_ctrl + kCtrlDone, 1u);
      } else {
#pragma unroll
        for (int r = 0; r < WORLD; ++r) red_release_sys(peers.ctrl[r] + kCtrlDone, 1u);
      }
    }
  }
  // every unit of every reducer has landed in this rank's copy
  if (threadIdx.x == 0) {
    if (!no_wait)
      spin_until(ctrl + kCtrlDone, epoch * static_cast<unsigned int>(nblk * parts), "landed units", rank, -1);
    __threadfence_system();
    if (atomicAdd(ctrl + kCtrlExit, 1u) == gridDim.x - 1) {  // last CTA out: next launch's epoch
      ctrl[kCtrlExit] = 0u;
      if (ADAMW) *aw.step += 1;
      __threadfence();
      atomicExch(ctrl + kCtrlEpoch, epoch);
    }
  }
}

}  // namespace gg

using namespace gg;

// Slice of rank r, in 16-byte units: equal parts of ceil(n4 / world), the last one short.
extern "C" void gg_p2p_slice(size_t n_floats, int world, int rank, size_t* lo4, size_t* hi4) {
  const size_t n4 = n_floats / 4;
  const size_t per = (n4 + world - 1) / world;
  const size_t lo = std::min(n4, per * static_cast<size_t>(rank));
  *lo4 = lo;
  *hi4 = std::min(n4, lo + per);
}

extern "C" int gg_p2p_allreduce_avg(const unsigned long long* peer_ptrs, int world, int rank, size_t n_floats,
                                    gg_stream_t stream) {
  GG_CHECK(peer_ptrs && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, GG_ERR_ARG,
           "gg_p2p_allreduce_avg: world=%d rank=%d (1 <= world <= %d)", world, rank, kP2PMaxWorld);
  GG_CHECK(world == 1 || world == 2 || world == 4 || world == 8, GG_ERR_UNSUPPORTED,
           "gg_p2p_allreduce_avg: world=%d (1, 2, 4 or 8 GPUs of one NVSwitch domain)", world);
  GG_CHECK(n_floats % 4 == 0, GG_ERR_ARG, "gg_p2p_allreduce_avg: n_floats=%zu must be a multiple of 4", n_floats);
  PeerPtrs peers = {};
  for (int r = 0; r < world; ++r) {
    GG_CHECK(peer_ptrs[r] != 0 && (peer_ptrs[r] & 15) == 0, GG_ERR_ARG, "gg_p2p_allreduce_avg: peer buffer %d (%p) missing or not 16-byte aligned",
             r, reinterpret_cast<void*>(peer_ptrs[r]));
    peers.p[r] = reinterpret_cast<float4*>(peer_ptrs[r]);
  }
  if (world == 1 || n_floats == 0) return GG_OK;
  size_t lo, hi;
  gg_p2p_slice(n_floats, world, rank, &lo, &hi);
  if (hi <= lo) return GG_OK;
  const size_t work = hi - lo;
  const int sms = device_sm_count();
  const size_t per_cta = static_cast<size_t>(kP2PThreads) * (world >= 4 ? 2 : 4);
  const int grid = static_cast<int>(std::min<size_t>(static_cast<size_t>(2 * sms), (work + per_cta - 1) / per_cta));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float inv = 1.0f / static_cast<float>(world);
  if (world == 2) p2p_allreduce_avg_kernel<2><<<grid, kP2PThreads, 0, s>>>(peers, lo, hi, inv);
  else if (world == 4) p2p_allreduce_avg_kernel<4><<<grid, kP2PThreads, 0, s>>>(peers, lo, hi, inv);
  else p2p_allreduce_avg_kernel<8><<<grid, kP2PThreads, 0, s>>>(peers, lo, hi, inv);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_nvls_allreduce_avg(void* multicast_ptr, int world, int rank, size_t n_floats, gg_stream_t stream) {
  GG_CHECK(multicast_ptr && (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0, GG_ERR_ARG,
           "gg_nvls_allreduce_avg: multicast address %p missing or not 16-byte aligned", multicast_ptr);
  GG_CHECK(world >= 1 && rank >= 0 && rank < world, GG_ERR_ARG, "gg_nvls_allreduce_avg: world=%d rank=%d", world, rank);
  GG_CHECK(n_floats % 4 == 0, GG_ERR_ARG, "gg_nvls_allreduce_avg: n_floats=%zu must be a multiple of 4", n_floats);
  if (n_floats == 0) return GG_OK;
  size_t lo, hi;
  gg_p2p_slice(n_floats, world, rank, &lo, &hi);
  if (hi <= lo) return GG_OK;
  const size_t per_cta = static_cast<size_t>(kP2PThreads) * 8;
  const int grid = static_cast<int>(
      std::min<size_t>(static_cast<size_t>(2 * device_sm_count()), (hi - lo + per_cta - 1) / per_cta));
  nvls_allreduce_avg_kernel<<<grid, kP2PThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float4*>(multicast_ptr), lo, hi, 1.0f / static_cast<float>(world));
  GG_LAUNCH_CHECK();
  return GG_OK;
}

// CTAs per block: as many as keep every unit of the rank with the most blocks in ONE wave of resident CTAs (a second,
// partly filled wave doubled the kernel's time at 2 ranks: 400 units on 296 CTAs).  Every rank must arrive at the same
// number (the `done` counter counts units of all ranks): it depends on the problem, the world size and the
// instance's occupancy only.
static int grad_parts(int nblk, int world, int resident) {
  const int own_max = (nblk + world - 1) / world;
  return std::max(1, std::min(kGradMaxParts, resident / std::max(1, own_max)));
}

extern "C" size_t gg_grad_ctrl_bytes(void) { return GG_GRAD_CTRL_BYTES; }

extern "C" int gg_grad_exchange(const unsigned long long* grad_ptrs, const unsigned long long* ctrl_ptrs, void* grad_mc,
                                void* ctrl_mc, const void* stage, int world, int rank, int C, int D, int flags,
                                gg_stream_t stream) {
  GG_CHECK(grad_ptrs && ctrl_ptrs && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, GG_ERR_ARG,
           "gg_grad_exchange: world=%d rank=%d (1 <= world <= %d)", world, rank, kP2PMaxWorld);
  GG_CHECK(world == 1 || world == 2 || world == 4 || world == 8, GG_ERR_UNSUPPORTED,
           "gg_grad_exchange: world=%d (1, 2, 4 or 8 GPUs of one NVSwitch domain)", world);
  GG_CHECK(C > 0 && D > 0 && D % 4 == 0, GG_ERR_ARG, "gg_grad_exchange: C=%d D=%d (D a multiple of 4)", C, D);
  const int nblk = (C + kGradBlockRows - 1) / kGradBlockRows;
  GG_CHECK(nblk <= kCtrlReady, GG_ERR_UNSUPPORTED, "gg_grad_exchange: C=%d exceeds %d geocells", C, kCtrlReady * kGradBlockRows);
  GG_CHECK((grad_mc == nullptr) == (ctrl_mc == nullptr), GG_ERR_ARG, "gg_grad_exchange: both multicast addresses or none");
  if (world == 1) return GG_OK;
  GG_CHECK(stage && (reinterpret_cast<uintptr_t>(stage) & 15) == 0, GG_ERR_ARG, "gg_grad_exchange: staging region missing");
  GradPeers peers = {};
  for (int r = 0; r < world; ++r) {
    GG_CHECK(grad_ptrs[r] != 0 && (grad_ptrs[r] & 15) == 0 && ctrl_ptrs[r] != 0 && (ctrl_ptrs[r] & 15) == 0, GG_ERR_ARG,
             "gg_grad_exchange: buffers of rank %d missing or not 16-byte aligned", r);
    peers.grad[r] = reinterpret_cast<float4*>(grad_ptrs[r]);
    peers.ctrl[r] = reinterpret_cast<unsigned int*>(ctrl_ptrs[r]);
  }
  const int own = nblk > rank ? (nblk - rank + world - 1) / world : 0;
  const int own_max = (nblk + world - 1) / world;
  // every CTA must be resident (each waits for units other CTAs of this grid deliver): <= 2 per SM
  const int resident = 2 * device_sm_count();  // (<= 128 registers in every instance of the plain exchange)
  const int parts = grad_parts(nblk, world, resident);
  const int grid = std::max(1, std::min(resident, own * parts));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float inv = 1.0f / static_cast<float>(world);
  float4* mg = static_cast<float4*>(grad_mc);
  unsigned int* mcc = static_cast<unsigned int*>(ctrl_mc);
  const float4* st = static_cast<const float4*>(stage);
  const int nw = (flags & GG_GRAD_NO_WAIT) ? 1 : 0;
#define GG_GX(W)                                                                                                   \
  do {                                                                                                             \
    if (mg) grad_exchange_kernel<W, true, false><<<grid, kGradThreads, 0, s>>>(peers, mg, mcc, st, rank, C, D, nblk, own_max, inv, nw, parts, AdamwArgs{});  \
    else grad_exchange_kernel<W, false, false><<<grid, kGradThreads, 0, s>>>(peers, mg, mcc, st, rank, C, D, nblk, own_max, inv, nw, parts, AdamwArgs{});    \
  } while (0)
  if (world == 2) GG_GX(2);
  else if (world == 4) GG_GX(4);
  else GG_GX(8);
#undef GG_GX
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_grad_exchange_adamw(const unsigned long long* w16_ptrs, const unsigned long long* bias_ptrs,
                                      const unsigned long long* ctrl_ptrs, void* w16_mc, void* bias_mc, void* ctrl_mc,
                                      const void* stage, int world, int rank, int C, int D, float* master_w,
                                      float* master_b, float* m_w, float* v_w, float* m_b, float* v_b,
                                      const float* hyper, long long* step, int flags, gg_stream_t stream) {
  GG_CHECK(w16_ptrs && bias_ptrs && ctrl_ptrs && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world,
           GG_ERR_ARG, "gg_grad_exchange_adamw: world=%d rank=%d (1 <= world <= %d)", world, rank, kP2PMaxWorld);
  GG_CHECK(world == 1 || world == 2 || world == 4 || world == 8, GG_ERR_UNSUPPORTED,
           "gg_grad_exchange_adamw: world=%d (1, 2, 4 or 8 GPUs of one NVSwitch domain)", world);
  GG_CHECK(C > 0 && D > 0 && D % 8 == 0, GG_ERR_ARG, "gg_grad_exchange_adamw: C=%d D=%d (D a multiple of 8)", C, D);
  GG_CHECK(master_w && master_b && m_w && v_w && m_b && v_b && hyper && step, GG_ERR_ARG,
           "gg_grad_exchange_adamw: optimizer state missing");
  const int nblk = (C + kGradBlockRows - 1) / kGradBlockRows;
  GG_CHECK(nblk <= kCtrlReady, GG_ERR_UNSUPPORTED, "gg_grad_exchange_adamw: C=%d exceeds %d geocells", C,
           kCtrlReady * kGradBlockRows);
  const bool nvls = w16_mc != nullptr;
  GG_CHECK(nvls == (bias_mc != nullptr) && nvls == (ctrl_mc != nullptr), GG_ERR_ARG,
           "gg_grad_exchange_adamw: all three multicast addresses or none");
  GG_CHECK(stage && (reinterpret_cast<uintptr_t>(stage) & 15) == 0, GG_ERR_ARG, "gg_grad_exchange_adamw: staging region missing");
  GradPeers peers = {};
  AdamwArgs aw = {};
  for (int r = 0; r < world; ++r) {
    GG_CHECK(w16_ptrs[r] != 0 && (w16_ptrs[r] & 15) == 0 && bias_ptrs[r] != 0 && (bias_ptrs[r] & 15) == 0 &&
                 ctrl_ptrs[r] != 0 && (ctrl_ptrs[r] & 15) == 0,
             GG_ERR_ARG, "gg_grad_exchange_adamw: buffers of rank %d missing or not 16-byte aligned", r);
    peers.ctrl[r] = reinterpret_cast<unsigned int*>(ctrl_ptrs[r]);
    aw.w16[r] = reinterpret_cast<uint4*>(w16_ptrs[r]);
    aw.bias[r] = reinterpret_cast<float4*>(bias_ptrs[r]);
  }
  aw.w = master_w; aw.m = m_w; aw.v = v_w;
  aw.b = master_b; aw.mb = m_b; aw.vb = v_b;
  aw.hyper = hyper; aw.step = step;
  aw.w16_mc = static_cast<uint4*>(w16_mc);
  aw.bias_mc = static_cast<float4*>(bias_mc);
  const int own = nblk > rank ? (nblk - rank + world - 1) / world : 0;
  const int own_max = (nblk + world - 1) / world;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float inv = 1.0f / static_cast<float>(world);
  unsigned int* mcc = static_cast<unsigned int*>(ctrl_mc);
  const float4* st = static_cast<const float4*>(stage);
  const int nw = (flags & GG_GRAD_NO_WAIT) ? 1 : 0;
  // every CTA must be resident (each waits for units other CTAs of this grid deliver): as many as the instance fits
#define GG_GXA_ONE(W, NV)                                                                                          \
  do {                                                                                                             \
    auto kern = grad_exchange_kernel<W, NV, true>;                                                                 \
    int per_sm = 0;                                                                                                \
    GG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGradThreads, 0));                        \
    const int resident = std::max(1, per_sm) * device_sm_count();                                                  \
    const int parts = grad_parts(nblk, world, resident);                                                           \
    const int g = std::max(1, std::min(resident, own * parts));                                                    \
    kern<<<g, kGradThreads, 0, s>>>(peers, nullptr, mcc, st, rank, C, D, nblk, own_max, inv, nw, parts, aw);       \
  } while (0)
#define GG_GXA(W)                                                                                                  \
  do {                                                                                                             \
    if (nvls) GG_GXA_ONE(W, true);                                                                                 \
    else GG_GXA_ONE(W, false);                                                                                     \
  } while (0)
  if (world == 1) GG_GXA(1);
  else if (world == 2) GG_GXA(2);
  else if (world == 4) GG_GXA(4);
  else GG_GXA(8);
#undef GG_GXA
#undef GG_GXA_ONE
  GG_LAUNCH_CHECK();
  return GG_OK;
}
