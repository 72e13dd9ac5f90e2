// Fused haversine label-smoothed geocell cross-entropy, forward + gradient.
//
// Replaces, per training step, the reference's
//   haversine_matrix(labels, centroids.t())        models/utils.py:39-57   (B x C distances)
//   smooth_labels(d) = exp(-(d - rowmin d)/65)      models/utils.py:20-32
//   t = s / max(sum_c s, 1e-12)                     models/super_guessr.py:377
//   loss_b = -sum_c t * log_softmax(logits)         models/super_guessr.py:379-380
//   dlogits = softmax(logits) - t                   (autograd, main_coordinator_idun_s3.py:423)
// without materialising any B x C distance / target / probability matrix: the only B x C HBM
// traffic is one read of the bf16 logits and one write of the bf16 gradient.
//
// Distance: with unit vectors u (label) and v_c (centroid), haversine's a = |u - v_c|^2 / 4 exactly,
// so q = |u - v_c|^2 (3 sub + 3 fma, no per-element trig, no cancellation at small distances) and
// d = 2R asin(sqrt(q)/2).  q is monotone in d, so row minimum and "can this cell carry target mass"
// run on q; sqrt/asin/exp are evaluated only for cells with d < dmin + far_km (the caller's cut-off:
// beyond it exp(-(d-dmin)/65) is below the chosen fraction of the nearest cell's weight; inf = off).
//
// Kernels:
//  (T) centroid table (once per table): unit vectors in class order, plus a copy sorted along a
//      Morton curve in 64-cell spatial groups with one bounding cap (centre, angular radius) each.
//  (A) gg_hav_row_stats, no B x C data: label unit vectors (fp64 trig, one thread per row), then one
//      warp per row bounds every spatial group's distance from its cap, evaluates exact q only in
//      the groups that can hold the nearest cell / a near cell, and emits {u, q_thr, dmin, 1/sum s}
//      plus the bitmask of exactly the 64-class groups that hold a near cell (and argmin_c d as a
//      by-product).
//  (B) hav_ce_stream_kernel, the HBM-bound pass.  A warp owns 256 adjacent classes (4 groups = one
//      nibble of the near mask), walks a block of rows and feeds itself through a private 3-stage TMA
//      ring (no producer warp, no block-wide synchronisation); a lane owns one bf16 pair in each of
//      the four groups (unit vectors and bias-gradient column sums in registers), so every warp-level
//      load/store is one full 128-byte line and a near group costs every lane exactly two target
//      evaluations -- no divergence.  Per 8-row stage: a branch-free pass p = exp(l - lse) over all
//      rows, then the rows with a near group redo the flagged groups as p - t (packed fp32x2 math) and
//      collect sum_c t*l, reduced once per stage and added to a per-row 2^-32 fixed-point accumulator
//      (integer atomics commute exactly: deterministic).
//  (C) tail of (B): the last strip CTA of a row block turns the accumulators into loss_b = lse_b -
//      sum_c t*l; the last row block sums the block partials in block order (batch mean).  Tickets and
//      accumulators live in the row-statistics buffer and are left zeroed for the next launch.
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr float kEarthRadiusKm = 6378.137f;  // models/utils.py:55 (6378137 m) / 1000
constexpr float kLog2eF = 1.4426950408889634f;
constexpr int kCellsPerGroup = 64;   // spatial group (one cap) and class group (one near-mask bit)
constexpr int kColsPerWarp = 256;    // kernel B: 4 class groups = one nibble of the near mask
constexpr int kMaxSlots = 2;         // kernel A keeps one group bound per thread per slot: <= 256 groups
constexpr int kMaxMaskWords = 8;     // near mask words per row: 256 class groups
constexpr float kCapMargin = 1.0e-5f;  // rad (64 m): absorbs fp32 rounding in the cap tests

struct RowRec {  // 32 bytes per row, written by (A), read by (B)/(C)
  float ux, uy, uz, q_thr;
  float off, inv_s, valid, dmin;
};

// Centroid table layout, in 4-byte words (see (T)):
//   x[Cpad] y[Cpad] z[Cpad]                       class order, Cpad = C rounded up to 256
//   scell[Cs] = {x, y, z, class index (int bits)}  Morton order, Cs = C rounded up to 64 (pad: 1e9, INT_MAX)
//   caps[Cs/64] = {cx, cy, cz, radius}             radius 4 = anywhere, < 0 = empty group
struct TableView {
  int Cpad, Cs, ngroups;
  const float *x, *y, *z;
  const float4* scell;
  const float4* caps;
};
__host__ __device__ inline int table_cpad(int C) { return (C + 255) / 256 * 256; }
__host__ __device__ inline int table_cs(int C) { return (C + 63) / 64 * 64; }
__host__ __device__ inline size_t table_words(int C) {
  const size_t Cpad = table_cpad(C), Cs = table_cs(C);
  return 3 * Cpad + 4 * Cs + 4 * (Cs / kCellsPerGroup);
}
__host__ __device__ inline TableView view_table(const float* t, int C) {
  TableView v;
  v.Cpad = table_cpad(C);
  v.Cs = table_cs(C);
  v.ngroups = v.Cs / kCellsPerGroup;
  v.x = t;
  v.y = v.x + v.Cpad;
  v.z = v.y + v.Cpad;
  v.scell = reinterpret_cast<const float4*>(v.z + v.Cpad);  // 3 * Cpad words: 16-byte aligned (Cpad % 256 == 0)
  v.caps = v.scell + v.Cs;
  return v;
}

// asin(r) for r <= 0.5, z = r*r.  Cephes-style (abs err ~1e-7).
__device__ __forceinline__ float asin_poly(float r, float z) {
  float p = 4.2163199048e-2f;
  p = fmaf(p, z, 2.4181311049e-2f);
  p = fmaf(p, z, 4.5470025998e-2f);
  p = fmaf(p, z, 7.4953002686e-2f);
  p = fmaf(p, z, 1.6666752422e-1f);
  return fmaf(p * z, r, r);
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// theta = 2 asin(sqrt(q)/2), q = squared chord in [0, 4]; IEEE sqrt (row minima, cap radii)
__device__ __forceinline__ float theta_from_q(float q) {
  const float h = fminf(0.25f * q, 1.0f);  // = sin^2(theta/2) = haversine 'a'
  if (h <= 0.25f) {
    return 2.0f * asin_poly(sqrtf(h), h);
  }
  const float x = sqrtf(h);
  const float z = 0.5f * (1.0f - x);
  return CUDART_PI_F - 4.0f * asin_poly(sqrtf(z), z);
}
// same, branch-free with the approximate square root (cap bounds: their margins absorb the error)
__device__ __forceinline__ float theta_from_q_fast(float q) {
  const float h = fminf(0.25f * q, 1.0f);
  const float x = sqrt_approx(h);
  const float z = 0.5f * (1.0f - x);
  const float near_half = 2.0f * asin_poly(x, h);
  const float far_half = CUDART_PI_F - 4.0f * asin_poly(sqrt_approx(z), z);
  return h <= 0.25f ? near_half : far_half;
}
// Unnormalised target s = exp(-(d - dmin)/tau) = 2^(neg_rk2 * theta + off) of a cell at squared chord
// q, or 0 when q >= q_thr.  (A) sums it and (B) applies it: both call exactly this function (the same
// bits), so the targets of a row sum to one.  Branch-free.
__device__ __forceinline__ float target_weight(float q, float q_thr, float neg_rk2, float off) {
  const float s = ex2_approx(fmaf(neg_rk2, theta_from_q_fast(q), off));
  return q < q_thr ? s : 0.f;
}
// Same bits as target_weight for q < q_thr <= 1 (d < 6672 km: the near half of theta_from_q_fast, evaluated
// by the same operations), 0 otherwise: lets (A) skip the far-half polynomial on almost every row.
__device__ __forceinline__ float target_weight_narrow(float q, float q_thr, float neg_rk2, float off) {
  const float h = fminf(0.25f * q, 1.0f);
  const float s = ex2_approx(fmaf(neg_rk2, 2.0f * asin_poly(sqrt_approx(h), h), off));
  return q < q_thr ? s : 0.f;
}
// squared chord; (A) and (B) must evaluate it identically (same threshold decisions)
__device__ __forceinline__ float chord2(float ux, float uy, float uz, float x, float y, float z) {
  const float dx = ux - x, dy = uy - y, dz = uz - z;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Packed (fp32x2) forms for the streaming pass: every half goes through exactly the operations of the scalar
// functions above (FADD2 / FMUL2 / FFMA2 round each half like FADD / FMUL / FFMA), so the bits are those (A) summed.
__device__ __forceinline__ float2 chord2x2(float2 ux, float2 uy, float2 uz, float2 x, float2 y, float2 z) {
  const float2 dx = __fadd2_rn(ux, make_float2(-x.x, -x.y)), dy = __fadd2_rn(uy, make_float2(-y.x, -y.y)),
               dz = __fadd2_rn(uz, make_float2(-z.x, -z.y));
  return __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
}
__device__ __forceinline__ float2 asin_poly2(float2 r, float2 z) {
  float2 p = make_float2(4.2163199048e-2f, 4.2163199048e-2f);
  p = __ffma2_rn(p, z, make_float2(2.4181311049e-2f, 2.4181311049e-2f));
  p = __ffma2_rn(p, z, make_float2(4.5470025998e-2f, 4.5470025998e-2f));
  p = __ffma2_rn(p, z, make_float2(7.4953002686e-2f, 7.4953002686e-2f));
  p = __ffma2_rn(p, z, make_float2(1.6666752422e-1f, 1.6666752422e-1f));
  return __ffma2_rn(__fmul2_rn(p, z), r, r);
}
__device__ __forceinline__ float2 target_weight_narrow2(float2 q, float q_thr, float2 neg_rk2, float2 off) {
  const float2 h4 = __fmul2_rn(make_float2(0.25f, 0.25f), q);
  const float2 h = make_float2(fminf(h4.x, 1.0f), fminf(h4.y, 1.0f));
  const float2 a = asin_poly2(make_float2(sqrt_approx(h.x), sqrt_approx(h.y)), h);
  const float2 e = __ffma2_rn(neg_rk2, __fmul2_rn(make_float2(2.0f, 2.0f), a), off);
  return make_float2(q.x < q_thr ? ex2_approx(e.x) : 0.f, q.y < q_thr ? ex2_approx(e.y) : 0.f);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ (T) centroid table
__device__ __forceinline__ uint32_t spread16(uint32_t v) {  // abcd -> 0a0b0c0d
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
// class-order unit vectors (fp64 trig rounded to fp32; pad cells far outside the unit sphere, so
// their q is huge: "infinitely far") and Morton keys of (lat, lng)
__global__ void table_unit_kernel(const float* __restrict__ centroids, float* __restrict__ table, int C,
                                  uint32_t* __restrict__ keys) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int Cpad = table_cpad(C);
  if (c >= Cpad) return;
  float x = 1.0e9f, y = 1.0e9f, z = 1.0e9f;
  if (c < C) {
    const float lngf = centroids[2 * c], latf = centroids[2 * c + 1];
    const double lng = static_cast<double>(lngf) * (CUDART_PI / 180.0);
    const double lat = static_cast<double>(latf) * (CUDART_PI / 180.0);
    x = static_cast<float>(cos(lat) * cos(lng));
    y = static_cast<float>(cos(lat) * sin(lng));
    z = static_cast<float>(sin(lat));
    const float fx = fminf(fmaxf((lngf + 180.0f) * (1.0f / 360.0f), 0.f), 1.f);
    const float fy = fminf(fmaxf((latf + 90.0f) * (1.0f / 180.0f), 0.f), 1.f);
    const uint32_t qx = static_cast<uint32_t>(fx * 65535.0f), qy = static_cast<uint32_t>(fy * 65535.0f);
    keys[c] = (fx == fx && fy == fy) ? (spread16(qx) | (spread16(qy) << 1)) : 0xffffffffu;
  }
  table[c] = x;
  table[Cpad + c] = y;
  table[2 * static_cast<size_t>(Cpad) + c] = z;
}
// rank of every class along the Morton curve (O(C^2) compares, C ~ 1e4, once per table)
__global__ void table_rank_kernel(const uint32_t* __restrict__ keys, int C, float4* __restrict__ scell) {
  __shared__ uint32_t tile[256];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mine = c < C ? keys[c] : 0u;
  int rank = 0;
  for (int base = 0; base < C; base += 256) {
    const int n = min(256, C - base);
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < n) tile[threadIdx.x] = keys[base + threadIdx.x];
    __syncthreads();
    for (int i = 0; i < n; ++i) {
      const uint32_t k = tile[i];
      rank += (k < mine || (k == mine && base + i < c)) ? 1 : 0;
    }
  }
  if (c < C) scell[rank].w = __int_as_float(c);
}
// sorted copy, one warp per spatial group (2 cells per lane): centre = normalised mean of the
// members, radius = largest member angle + margin
__global__ void table_group_kernel(float* __restrict__ table, int C) {
  const TableView tv = view_table(table, C);
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= tv.ngroups) return;
  float4* scell = const_cast<float4*>(tv.scell);
  float x[2], y[2], z[2];
  int idx[2];
  float n = 0.f, mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = g * kCellsPerGroup + lane + 32 * h;
    x[h] = y[h] = z[h] = 1.0e9f;
    idx[h] = 0x7fffffff;
    if (p < C) {
      idx[h] = __float_as_int(scell[p].w);
      x[h] = tv.x[idx[h]]; y[h] = tv.y[idx[h]]; z[h] = tv.z[idx[h]];
      n += 1.f; mx += x[h]; my += y[h]; mz += z[h];
    }
    scell[p] = make_float4(x[h], y[h], z[h], __int_as_float(idx[h]));
  }
  n = warp_sum(n); mx = warp_sum(mx); my = warp_sum(my); mz = warp_sum(mz);
  const float len = sqrtf(mx * mx + my * my + mz * mz);
  float rad;
  if (n == 0.f) {
    mx = 0.f; my = 0.f; mz = 1.f; rad = -1.f;
  } else if (len < 1.0e-3f * n || !(len == len)) {
    mx = 0.f; my = 0.f; mz = 1.f; rad = 4.f;
  } else {
    mx /= len; my /= len; mz /= len;
    float a = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (idx[h] != 0x7fffffff) a = fmaxf(a, theta_from_q(chord2(mx, my, mz, x[h], y[h], z[h])));
    rad = warp_max(a) * (1.0f + 1.0e-6f) + kCapMargin;
  }
  if (lane == 0) const_cast<float4*>(tv.caps)[g] = make_float4(mx, my, mz, rad);
}

// ------------------------------------------------------------------ (A) per-row statistics
// Unit vectors of the labels: (B,4) = {x, y, z, valid}.  Non-finite labels give valid = 0 (the
// reference's nan_to_num turns such a row's targets into zeros, utils.py:31).
__global__ void label_xyz_kernel(const float* __restrict__ labels, float4* __restrict__ out, int B,
                                 unsigned int* __restrict__ counters, int ncounters,
                                 unsigned long long* __restrict__ row_acc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < ncounters) counters[b] = 0u;  // tickets of the loss kernel's finishing steps (grid >= ncounters threads)
  if (b >= B) return;
  row_acc[b] = 0ull;  // the loss kernel's per-row fixed-point accumulator of sum_c t*l
  const float lngf = labels[2 * b], latf = labels[2 * b + 1];
  float4 o = make_float4(0.f, 0.f, 1.f, 0.f);
  if (isfinite(lngf) && isfinite(latf)) {
    const double lng = static_cast<double>(lngf) * (CUDART_PI / 180.0);
    const double lat = static_cast<double>(latf) * (CUDART_PI / 180.0);
    o = make_float4(static_cast<float>(cos(lat) * cos(lng)), static_cast<float>(cos(lat) * sin(lng)),
                    static_cast<float>(sin(lat)), 1.f);
  }
  out[b] = o;
}

// One CTA (4 warps) per row.  Thread t owns the caps of spatial groups 4*(t%32) + t/32 + 128*s, so
// Morton-adjacent groups belong to different warps and each warp walks only its own groups.
template <int SLOTS>
__global__ void __launch_bounds__(128)
hav_row_stats_kernel(const float4* __restrict__ lab_xyz, const float* __restrict__ table, int C, int B, float k2,
                     float cos_half_far, float sin_half_far, float phi, RowRec* __restrict__ rec,
                     uint32_t* __restrict__ near, int nwp, long long* __restrict__ nearest_cell,
                     float* __restrict__ nearest_km) {
  __shared__ float s_f[4];
  __shared__ int s_i[4];
  __shared__ uint32_t s_near[kMaxMaskWords];  // the row's near mask: bit = 64-class group holding a cell with q < q_thr
  const int row = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kMaxMaskWords) s_near[threadIdx.x] = 0u;  // (ordered before phase 3 by the barriers of phases 1-2)
  const TableView tv = view_table(table, C);
  const float4 u = __ldg(lab_xyz + row);
  const float ux = u.x, uy = u.y, uz = u.z;
  const bool valid = u.w != 0.f;

  // phase 1: distance bounds of my spatial groups from their caps
  float lo[SLOTS];
  float lo_min = CUDART_INF_F;  // (smallest a + radius: the group that guarantees the closest cell)
  int g_best = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    const int g = 4 * lane + warp + 128 * k;
    lo[k] = CUDART_INF_F;
    if (g < tv.ngroups) {
      const float4 cap = __ldg(tv.caps + g);
      if (cap.w >= 0.f) {
        const float a = theta_from_q_fast(chord2(ux, uy, uz, cap.x, cap.y, cap.z));
        lo[k] = fmaxf(a - cap.w, 0.f);
        if (a + cap.w < lo_min) { lo_min = a + cap.w; g_best = g; }
      }
    }
  }
  // scanning that group gives a real cell distance = a tight upper bound of the row minimum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ol = __shfl_xor_sync(0xffffffffu, lo_min, o);
    const int og = __shfl_xor_sync(0xffffffffu, g_best, o);
    if (ol < lo_min || (ol == lo_min && og < g_best)) { lo_min = ol; g_best = og; }
  }
  if (lane == 0) { s_f[warp] = lo_min; s_i[warp] = g_best; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float ol = s_f[w];
    const int og = s_i[w];
    if (ol < lo_min || (ol == lo_min && og < g_best)) { lo_min = ol; g_best = og; }
  }
  __syncthreads();  // s_f / s_i are reused below

  float qmin = CUDART_INF_F;
  int imin = 0x7fffffff;
  auto scan_cells = [&](const float4& c0, const float4& c1) {
    const float q0 = chord2(ux, uy, uz, c0.x, c0.y, c0.z), q1 = chord2(ux, uy, uz, c1.x, c1.y, c1.z);
    const int i0 = __float_as_int(c0.w), i1 = __float_as_int(c1.w);
    if (q0 < qmin || (q0 == qmin && i0 < imin)) { qmin = q0; imin = i0; }  // first class index on ties
    if (q1 < qmin || (q1 == qmin && i1 < imin)) { qmin = q1; imin = i1; }
  };
  {
    const float4* p = tv.scell + g_best * kCellsPerGroup + lane;
    scan_cells(__ldg(p), __ldg(p + 32));
  }
  const float ub = theta_from_q_fast(warp_min(qmin)) + kCapMargin;

  // phase 2: exact nearest cell among my groups that can still contain it
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    uint32_t m = __ballot_sync(0xffffffffu, lo[k] <= ub && 4 * lane + warp + 128 * k != g_best);
#pragma unroll 1
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const float4* p = tv.scell + (4 * j + warp + 128 * k) * kCellsPerGroup + lane;
      scan_cells(__ldg(p), __ldg(p + 32));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oq = __shfl_xor_sync(0xffffffffu, qmin, o);
    const int oi = __shfl_xor_sync(0xffffffffu, imin, o);
    if (oq < qmin || (oq == qmin && oi < imin)) { qmin = oq; imin = oi; }
  }
  if (lane == 0) { s_f[warp] = qmin; s_i[warp] = imin; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float oq = s_f[w];
    const int oi = s_i[w];
    if (oq < qmin || (oq == qmin && oi < imin)) { qmin = oq; imin = oi; }
  }
  const float theta_min = theta_from_q(qmin);
  const float dmin = kEarthRadiusKm * theta_min;
  // near test on q: d < dmin + far  <=>  q < 4 sin^2((theta_min + phi)/2), phi = far / R, expanded
  // with sin(theta_min/2) = sqrt(qmin)/2 so that no trig runs per row; everything is "near" once
  // theta_min + phi reaches pi.
  float q_thr = CUDART_INF_F;
  {
    const float hmin = fminf(0.25f * qmin, 1.0f);
    const float sa = sqrtf(hmin), ca = sqrtf(1.0f - hmin);
    if (ca * cos_half_far - sa * sin_half_far > 0.f) {
      const float sh = fmaf(sa, cos_half_far, ca * sin_half_far);
      q_thr = 4.0f * sh * sh;
    }
  }
  const float a_thr = q_thr < CUDART_INF_F ? theta_min + phi + kCapMargin : CUDART_INF_F;
  const float off = dmin * k2;  // s = 2^((dmin - d) * k2)
  const float neg_rk2 = -kEarthRadiusKm * k2;
  const bool narrow = q_thr <= 1.0f;  // CTA-uniform: every near cell is on the first asin branch

  // phase 3: sum of the unnormalised targets over my near groups; lane w accumulates word w of the
  // class-group mask
  float ssum = 0.f;
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    uint32_t m = valid ? __ballot_sync(0xffffffffu, lo[k] < a_thr) : 0u;  // lo = inf for absent groups
#pragma unroll 1
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const int g = 4 * j + warp + 128 * k;
      const float4* p = tv.scell + g * kCellsPerGroup + lane;
      const float4 c0 = __ldg(p), c1 = __ldg(p + 32);
      const float q0 = chord2(ux, uy, uz, c0.x, c0.y, c0.z), q1 = chord2(ux, uy, uz, c1.x, c1.y, c1.z);
      // pad cells sit at 1e9 (q ~ 3e18) with index INT_MAX: cut by index when q_thr = inf
      const bool ok0 = __float_as_int(c0.w) < C, ok1 = __float_as_int(c1.w) < C;
      float s0, s1;
      if (narrow) {
        s0 = target_weight_narrow(q0, q_thr, neg_rk2, off);
        s1 = target_weight_narrow(q1, q_thr, neg_rk2, off);
      } else {
        s0 = target_weight(q0, q_thr, neg_rk2, off);
        s1 = target_weight(q1, q_thr, neg_rk2, off);
      }
      s0 = ok0 ? s0 : 0.f;
      s1 = ok1 ? s1 : 0.f;
      ssum += s0 + s1;
      // Exactly the class groups of the near cells (not every class group the spatial group touches: that
      // doubled the (row, slice) pairs the streaming pass had to treat as near).
      if (ok0 && q0 < q_thr) {
        const int cls = __float_as_int(c0.w);
        atomicOr(&s_near[cls >> 11], 1u << ((cls >> 6) & 31));
      }
      if (ok1 && q1 < q_thr) {
        const int cls = __float_as_int(c1.w);
        atomicOr(&s_near[cls >> 11], 1u << ((cls >> 6) & 31));
      }
    }
  }
  ssum = warp_sum(ssum);
  __syncthreads();  // every warp has read s_f / s_i
  if (lane == 0) s_f[warp] = ssum;
  __syncthreads();
  if (warp != 0) return;
  if (lane < nwp)
    near[static_cast<size_t>(row) * nwp + lane] =
        lane < kMaxMaskWords ? s_near[lane] : 0u;
  if (lane == 0) {
    ssum = (s_f[0] + s_f[1]) + (s_f[2] + s_f[3]);
    RowRec r;
    r.ux = ux; r.uy = uy; r.uz = uz; r.q_thr = q_thr;
    r.off = off;
    r.inv_s = valid ? 1.0f / fmaxf(ssum, 1e-12f) : 0.f;  // s / max(sum, 1e-12) (super_guessr.py:377)
    r.valid = valid ? 1.f : 0.f;
    r.dmin = dmin;
    rec[row] = r;
    if (nearest_cell) nearest_cell[row] = imin;
    if (nearest_km) nearest_km[row] = dmin;
  }
}

// ------------------------------------------------------------------ (B) streaming pass
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

constexpr int kStripWarps = 4;                       // warps per CTA: a strip of 1024 classes
constexpr int kStageRows = 8;                        // rows per pipeline stage
constexpr int kStreamStages = 3;
constexpr int kStreamThreads = 32 * kStripWarps;
constexpr int kStreamCtasPerSm = 4;
constexpr uint32_t kSliceTileBytes = kStageRows * kColsPerWarp * 2;  // one TMA box: 8 rows x 512 B
constexpr uint32_t kStageBytes = kStripWarps * kSliceTileBytes;      // 16 KB

struct StreamSmem {
  uint8_t tile[kStreamStages][kStageBytes];
  float4 rec[kStripWarps][kStageRows][2];   // per warp: the current stage's row records (RowRec)
  float lse2[kStripWarps][kStageRows];      // row log-sum-exp * log2(e)
  uint32_t near[kStripWarps][kStageRows];   // the warp's 4 near bits of the row
  float slp[kStripWarps][kStageRows][32];   // per lane partials of a near row's sum_c t*l (reduced once per stage)
  uint64_t full[kStripWarps][kStreamStages];
};

__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// CTA = (row block rb, strip of 4 warp slices); the four warps are independent pipelines.  Warp w owns classes
// [256 (4 strip + w), +256) and streams its own slice of the logits through a private 3-stage TMA -> shared
// memory ring (8 rows x 256 classes = 4 KB per stage; 48 KB in flight per CTA, four CTAs per SM): lane 0 refills
// a stage as soon as the warp has consumed it, so there is no producer warp polling for free slots -- which
// used to load one scheduler of every SM with four spinning warps (profiles/r01c) -- and no block-wide
// synchronisation.  Lanes 0..7 fetch the 8 rows' scalars (record, lse, near bits) one stage ahead into
// registers and publish them to the warp's own shared-memory slot, so the row loop reads them as broadcasts.
// Lane l holds the bf16 pairs at 64 j + 2 l + {0,1}, j = 0..3 -- one pair in each of the slice's four 64-class
// groups -- so a near group costs every lane exactly two target evaluations (no divergence), shared-memory
// reads are conflict-free and every global store is one full 128-byte line.  Rows are padded to a multiple of
// 256 columns (ldc >= Cpad): nothing in the row loop is predicated; pad columns carry p only and are never read
// downstream.  Pairs are processed with the packed fp32x2 pipe (fma / add on both halves at once).
template <bool WANT_DB>
__global__ void __launch_bounds__(kStreamThreads, kStreamCtasPerSm)
hav_ce_stream_kernel(const __grid_constant__ CUtensorMap tm_logits, int ldc, const float* __restrict__ lse,
                     const RowRec* __restrict__ rec, const uint32_t* __restrict__ near, int nwp,
                     const float* __restrict__ table, int B, int C, int rows_per_block, int nslices, int nstrips,
                     float neg_rk2, bf16* __restrict__ dlogits, unsigned long long* __restrict__ row_acc,
                     float* __restrict__ db_part, unsigned int* __restrict__ counters, float* __restrict__ loss_rows,
                     float* __restrict__ block_part, float mean_scale, float* __restrict__ loss_mean) {
  extern __shared__ uint8_t smem_raw[];
  StreamSmem& sm = *reinterpret_cast<StreamSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x / nstrips, strip = blockIdx.x - rb * nstrips;
  const int nactive = min(kStripWarps, nslices - strip * kStripWarps);  // warps with a slice
  const int row0 = rb * rows_per_block, row1 = min(B, row0 + rows_per_block);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_logits);
    for (int w = 0; w < kStripWarps; ++w)
      for (int st = 0; st < kStreamStages; ++st) mbar_init(&sm.full[w][st], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp >= nactive) return;

  const int ws = strip * kStripWarps + warp;
  const int Cpad = table_cpad(C);
  const int cbase = ws * kColsPerWarp + 2 * lane;  // + 64 j
  const int near_word = ws >> 3;                   // the warp's 4 near bits: nibble (ws & 7) of this mask word
  const uint32_t near_shift = 4 * (ws & 7);

  // this warp's ring: lane 0 arms the stage's barrier and issues the 4 KB box (rows past B are zero-filled)
  uint8_t* const my_tiles = sm.tile[0] + warp * kSliceTileBytes;
  auto issue = [&](int stage_idx, int row) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&sm.full[warp][stage_idx], kSliceTileBytes);
      tma_load_2d_hint(my_tiles + stage_idx * kStageBytes, &tm_logits, &sm.full[warp][stage_idx], ws * kColsPerWarp,
                       row, kPolicyEvictFirst);
    }
  };
#pragma unroll
  for (int st = 0; st < kStreamStages; ++st)
    if (row0 + st * kStageRows < row1) issue(st, row0 + st * kStageRows);

  // lanes 0..7: one row's scalars each, fetched a stage ahead
  // (raw values only: any arithmetic on them here would stall the warp on the loads it has just issued)
  float4 pr0 = make_float4(0.f, 0.f, 0.f, 0.f), pr1 = pr0;
  float plse = 0.f;
  uint32_t pnw = 0u;
  auto fetch = [&](int row) {
    pr0 = make_float4(0.f, 0.f, 0.f, 0.f); pr1 = pr0; plse = 0.f; pnw = 0u;
    if (lane < kStageRows && row + lane < row1) {
      const float4* rp = reinterpret_cast<const float4*>(rec + row + lane);
      pr0 = __ldg(rp);
      pr1 = __ldg(rp + 1);
      plse = __ldg(lse + row + lane);
      pnw = __ldg(near + static_cast<size_t>(row + lane) * nwp + near_word);
    }
  };
  fetch(row0);

  float2 vx[4], vy[4], vz[4];  // unit vectors of this lane's pair in each of the slice's four 64-class groups
  float2 db[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    vx[j] = __ldg(reinterpret_cast<const float2*>(table + cbase + 64 * j));
    vy[j] = __ldg(reinterpret_cast<const float2*>(table + Cpad + cbase + 64 * j));
    vz[j] = __ldg(reinterpret_cast<const float2*>(table + 2 * static_cast<size_t>(Cpad) + cbase + 64 * j));
    db[j] = make_float2(0.f, 0.f);
  }
  // Pad classes (column >= C, last slice only) sit at 1e9 in the table (q ~ 3e18): capping the row's threshold
  // below that keeps them without target mass even when q_thr = inf, and changes no decision on a real cell
  // (q <= 4), so (A) and (B) still agree.
  constexpr float kPadCut = 1.0e18f;

  const size_t pitch = static_cast<size_t>(ldc);
  uint32_t* gptr = reinterpret_cast<uint32_t*>(dlogits + static_cast<size_t>(row0) * pitch + cbase);
  const size_t pitch_w = pitch / 2;  // row pitch in 32-bit words (ldc is even)
  const uint32_t tile0 = smem_u32(my_tiles) + 4 * lane;
  const uint32_t rec0 = smem_u32(&sm.rec[warp][0][0]), lse0 = smem_u32(&sm.lse2[warp][0]),
                 near0 = smem_u32(&sm.near[warp][0]);
  const float2 log2e2 = make_float2(kLog2eF, kLog2eF);
  int stage = 0;
  uint32_t phase = 0;

  for (int rbase = row0; rbase < row1; rbase += kStageRows) {
    const int nr = min(kStageRows, row1 - rbase);
    __syncwarp();  // every lane is done with the previous stage's scalars
    if (lane < kStageRows) {
      sm.rec[warp][lane][0] = pr0;
      sm.rec[warp][lane][1] = pr1;
      sm.lse2[warp][lane] = plse * kLog2eF;
      sm.near[warp][lane] = (pnw >> near_shift) & 0xfu;
    }
    // rows of this stage with a near group among this warp's classes (bit r = row r; rows past the batch fetch 0)
    const uint32_t near_rows = __ballot_sync(0xffffffffu, lane < kStageRows && ((pnw >> near_shift) & 0xfu) != 0u);
    fetch(rbase + kStageRows);  // next stage's scalars: in flight while this stage is processed
    __syncwarp();
    mbar_wait(&sm.full[warp][stage], phase);
    const uint32_t tile = tile0 + stage * kStageBytes;
    float my_sl = 0.f;  // lane r keeps row r's sum_c t*l of this warp's classes

    // ---- pass 1, branch-free: p = 2^(l log2e - lse log2e) for every row of the stage, stored and added to the
    // bias-gradient column sums.  Nothing data dependent in here, so the rows' chains interleave freely.
    auto row_probs = [&](int rr, float2 (&g)[4]) {
      uint32_t cur[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = lds_u32(tile + rr * (kColsPerWarp * 2) + 128 * j);
      const float nlse2 = -__uint_as_float(lds_u32(lse0 + rr * 4));
      const float2 nl = make_float2(nlse2, nlse2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // both halves of the pair at once
        const float2 a = __ffma2_rn(make_float2(__uint_as_float(cur[j] << 16), __uint_as_float(cur[j] & 0xffff0000u)),
                                    log2e2, nl);
        g[j] = make_float2(ex2_approx(a.x), ex2_approx(a.y));
      }
    };
    if (nr == kStageRows) {
#pragma unroll
      for (int rr = 0; rr < kStageRows; ++rr) {
        float2 g[4];
        row_probs(rr, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcs(gptr + rr * pitch_w + 32 * j, pack_bf16x2(g[j].x, g[j].y));
        if (WANT_DB) {
#pragma unroll
          for (int j = 0; j < 4; ++j) db[j] = __fadd2_rn(db[j], g[j]);
        }
      }
    } else {
      for (int rr = 0; rr < nr; ++rr) {  // ragged last stage of the batch
        float2 g[4];
        row_probs(rr, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) __stcs(gptr + rr * pitch_w + 32 * j, pack_bf16x2(g[j].x, g[j].y));
        if (WANT_DB) {
#pragma unroll
          for (int j = 0; j < 4; ++j) db[j] = __fadd2_rn(db[j], g[j]);
        }
      }
    }

    // ---- pass 2, rows with a near group in this warp's 256 classes (about one in seven): the flagged groups
    // are redone as p - t from the tile still in shared memory and stored over pass 1's lines (they merge in
    // L2), t leaves the column sums, and the row's sum_c t*l is collected.
    for (uint32_t rows = near_rows; rows != 0u; rows &= rows - 1u) {  // warp-uniform
      const int rr = __ffs(rows) - 1;
      const uint32_t nb = lds_u32(near0 + rr * 4);
      const float nlse2 = -__uint_as_float(lds_u32(lse0 + rr * 4));
      const float2 nl = make_float2(nlse2, nlse2);
      const uint4 ua = lds_u128(rec0 + rr * 32);
      const uint4 ub = lds_u128(rec0 + rr * 32 + 16);
      const float ux = __uint_as_float(ua.x), uy = __uint_as_float(ua.y), uz = __uint_as_float(ua.z);
      const float q_thr = __uint_as_float(ua.w), off = __uint_as_float(ub.x), inv_s = __uint_as_float(ub.y);
      const float q_cut = fminf(q_thr, kPadCut);
      const float2 ux2 = make_float2(ux, ux), uy2 = make_float2(uy, uy), uz2 = make_float2(uz, uz);
      const float2 nrk = make_float2(neg_rk2, neg_rk2), off2 = make_float2(off, off), inv2 = make_float2(inv_s, inv_s);
      float2 sl2 = make_float2(0.f, 0.f);
      auto near_groups = [&](auto narrow_tag) {
        constexpr bool kNarrow = decltype(narrow_tag)::value;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((nb >> j) & 1u) {  // warp-uniform
            const uint32_t cur = lds_u32(tile + rr * (kColsPerWarp * 2) + 128 * j);
            const float2 l2 = make_float2(__uint_as_float(cur << 16), __uint_as_float(cur & 0xffff0000u));
            const float2 a = __ffma2_rn(l2, log2e2, nl);
            const float2 pr = make_float2(ex2_approx(a.x), ex2_approx(a.y));  // the bits pass 1 computed
            const float2 q = chord2x2(ux2, uy2, uz2, vx[j], vy[j], vz[j]);
            // same bits either way (see target_weight_narrow); (A) summed exactly these values
            const float2 w = kNarrow ? target_weight_narrow2(q, q_cut, nrk, off2)
                                     : make_float2(target_weight(q.x, q_cut, neg_rk2, off),
                                                   target_weight(q.y, q_cut, neg_rk2, off));
            const float2 t = __fmul2_rn(w, inv2);
            const float2 nt = make_float2(-t.x, -t.y);
            const float2 gg2 = __fadd2_rn(pr, nt);
            __stcs(gptr + rr * pitch_w + 32 * j, pack_bf16x2(gg2.x, gg2.y));
            if (WANT_DB) db[j] = __fadd2_rn(db[j], nt);
            sl2 = __ffma2_rn(t, l2, sl2);
          }
        }
      };
      // row-uniform: with q_thr <= 1 every near cell is on the first asin branch
      if (q_thr <= 1.0f) near_groups(std::true_type{});
      else near_groups(std::false_type{});
      sm.slp[warp][rr][lane] = sl2.x + sl2.y;
    }
    if (near_rows != 0u) {  // warp-uniform
      // one transposed reduction per stage instead of a shuffle tree per near row: lane l sums a quarter of
      // row l / 4's partials, two shuffle rounds finish the row, and lane r fetches row r's total
      __syncwarp();
      float part = 0.f;
      if ((near_rows >> (lane >> 2)) & 1u) {
        const float4* sp = reinterpret_cast<const float4*>(&sm.slp[warp][lane >> 2][(lane & 3) * 8]);
        const float4 a = sp[0], b = sp[1];
        part = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      my_sl = __shfl_sync(0xffffffffu, part, (lane * 4) & 31);  // lanes >= 8 read rows that do not exist: unused
    }
    gptr += kStageRows * pitch_w;
    __syncwarp();  // all lanes have read the stage's tile: lane 0 may hand the slot back to the TMA engine
    if (rbase + kStreamStages * kStageRows < row1) issue(stage, rbase + kStreamStages * kStageRows);
    // Row sums of t*l across the 50 warp slices: 2^-32 fixed point in a 64-bit integer (|sum| <= max |logit|),
    // so the atomic adds commute exactly -- deterministic in any arrival order, and exact where an fp32 sum
    // would round.  Most (row, slice) pairs hold no near cell and add nothing.
    if (lane < nr && my_sl != 0.f)
      atomicAdd(row_acc + rbase + lane, static_cast<unsigned long long>(__float2ll_rn(my_sl * 4294967296.0f)));
    if (++stage == kStreamStages) { stage = 0; phase ^= 1; }
  }
  if (WANT_DB) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float2*>(db_part + static_cast<size_t>(rb) * Cpad + cbase + 64 * j) = db[j];
  }

  // ---- (C) folded in: whichever strip CTA of this row block finishes last turns the rows' accumulators into
  // loss_b = lse_b - sum_c t*l, and whichever row block finishes last sums the row-block partials in block
  // order: deterministic, and no second launch.
  __syncwarp();
  named_bar_sync(1, nactive * 32);  // the warps that own a slice (idle warps of the last strip have returned)
  if (warp != 0) return;
  unsigned int ticket = 0;
  if (lane == 0) {
    // the CTA's row_acc adds are performed device-wide before it takes its ticket: ONE fence, cumulative through the
    // barrier (a membar in each of the 128 threads stalled the memory pipeline the SM's other three CTAs stream through)
    __threadfence();
    ticket = atomicAdd(counters + 1 + rb, 1u);
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != static_cast<unsigned int>(nstrips - 1)) return;
  __threadfence();
  float acc = 0.f;
  for (int row = row0 + lane; row < row1; row += 32) {  // always warp 0, 32 lanes: the same order whoever is last
    const float sl = static_cast<float>(static_cast<double>(static_cast<long long>(__ldcg(row_acc + row))) *
                                        (1.0 / 4294967296.0));
    row_acc[row] = 0ull;  // ready for the next launch on the same row statistics
    const float v = __ldg(&rec[row].valid) != 0.f ? __ldg(lse + row) - sl : 0.f;
    loss_rows[row] = v;
    acc += v;
  }
  if (lane == 0) counters[1 + rb] = 0u;  // ready for the next launch on the same row statistics
  if (loss_mean == nullptr) return;
  acc = warp_sum(acc);
  const int nrb = gridDim.x / nstrips;
  unsigned int last = 0;
  if (lane == 0) {
    block_part[rb] = acc;
    __threadfence();
    last = atomicAdd(counters, 1u) == static_cast<unsigned int>(nrb - 1) ? 1u : 0u;
  }
  if (__shfl_sync(0xffffffffu, last, 0) == 0u) return;
  __threadfence();
  float tot = 0.f;
  for (int i = lane; i < nrb; i += 32) tot += __ldcg(block_part + i);
  tot = warp_sum(tot);
  if (lane == 0) {
    loss_mean[0] = tot * mean_scale;
    counters[0] = 0u;
  }
}

// ------------------------------------------------------------------ hard-label CE, batch mean
// Hard-label cross entropy (super_guessr.py:383, nn.CrossEntropyLoss) and its gradient
// p - onehot, for should_smooth_labels=False or labels=None.  One CTA per row.
__global__ void hard_ce_kernel(const bf16* __restrict__ logits, int ldc, const float* __restrict__ lse,
                               const long long* __restrict__ labels_clf, int B, int C, bf16* __restrict__ dlogits,
                               float* __restrict__ loss_rows) {
  const int nchunks = (C + 3) >> 2;
  for (int row = blockIdx.x; row < B; row += gridDim.x) {
    const bf16* lrow = logits + static_cast<size_t>(row) * ldc;
    bf16* grow = dlogits + static_cast<size_t>(row) * ldc;
    const float lse2 = lse[row] * kLog2eF;
    const long long y = labels_clf[row];
    for (int g = threadIdx.x; g < nchunks; g += blockDim.x) {
      const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(lrow) + g);
      float l[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                    __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u)};
      float gr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gr[i] = ex2_approx(fmaf(l[i], kLog2eF, -lse2));
        if (4 * g + i == y) {
          gr[i] -= 1.0f;
          loss_rows[row] = lse[row] - l[i];
        }
      }
      uint2 o;
      o.x = pack_bf16x2(gr[0], gr[1]);
      o.y = pack_bf16x2(gr[2], gr[3]);
      __stcs(reinterpret_cast<uint2*>(grow) + g, o);
    }
  }
}

// Deterministic mean of the per-row losses (single CTA; B is a few thousand to 64k).
__global__ void loss_mean_kernel(const float* __restrict__ loss_rows, int B, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) acc += loss_rows[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v * scale;
  }
}

// ------------------------------------------------------------------ host-side planning
struct HavPlan {
  int Cpad, nslices, nstrips, rows_per_block, nrb;
  size_t off_block_part, bytes;
};
static HavPlan make_hav_plan(int B, int C) {
  HavPlan p;
  p.Cpad = table_cpad(C);
  p.nslices = p.Cpad / kColsPerWarp;
  p.nstrips = ceil_div(p.nslices, kStripWarps);
  // one resident wave of CTAs: row blocks x strips <= SMs x CTAs per SM, row blocks of whole stages
  // (GG_HAV_WAVES > 1: that many times more, smaller row blocks -- the hardware block scheduler then evens out
  // the near-cell work, which differs by +-50 % between strips; tuning knob, read once)
  static const int waves = [] {
    const char* e = getenv("GG_HAV_WAVES");
    const int w = e ? atoi(e) : 1;
    return w >= 1 && w <= 16 ? w : 1;
  }();
  const int capacity = std::max(1, waves * device_sm_count() * kStreamCtasPerSm / p.nstrips);
  p.rows_per_block = ceil_div(std::max(kStageRows, ceil_div(B, capacity)), kStageRows) * kStageRows;
  p.nrb = ceil_div(B, p.rows_per_block);
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 255) & ~size_t(255); return r; };
  p.off_block_part = take(sizeof(float) * p.nrb);
  p.bytes = o;
  return p;
}
// Row statistics buffer (gg_hav_row_stats -> gg_hav_ce_fwd_bwd):
//   rec[B] | near[B * nwp] | counters[1 + row blocks of the loss kernel] | lab_xyz[B] | row_acc[B] (u64)
struct StatsPlan {
  int nwp, ncounters;
  size_t off_rec, off_near, off_counter, off_lab, off_acc, bytes;
};
static StatsPlan make_stats_plan(int B, int C) {
  StatsPlan p;
  const int Cpad = table_cpad(C);
  const int class_groups = Cpad / kCellsPerGroup;
  p.nwp = ceil_div(ceil_div(class_groups, 32), 4) * 4;
  p.ncounters = 1 + make_hav_plan(B, C).nrb;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 255) & ~size_t(255); return r; };
  p.off_rec = take(sizeof(RowRec) * static_cast<size_t>(B));
  p.off_near = take(sizeof(uint32_t) * static_cast<size_t>(B) * p.nwp);
  p.off_counter = take(sizeof(unsigned int) * p.ncounters);
  p.off_lab = take(sizeof(float4) * static_cast<size_t>(B));
  p.off_acc = take(sizeof(unsigned long long) * static_cast<size_t>(B));
  p.bytes = o;
  return p;
}

}  // namespace gg

using namespace gg;

extern "C" int gg_hav_cpad(int C) { return table_cpad(C); }
extern "C" size_t gg_centroid_table_floats(int C) { return table_words(C); }
extern "C" size_t gg_centroid_table_workspace_bytes(int C) { return sizeof(uint32_t) * static_cast<size_t>(C); }

extern "C" int gg_centroid_unit_vectors(const float* centroids, float* cent_table, int C, void* workspace,
                                        gg_stream_t stream) {
  GG_CHECK(centroids && cent_table && workspace && C > 0, GG_ERR_ARG, "gg_centroid_unit_vectors: bad arguments");
  GG_CHECK(table_cpad(C) / kCellsPerGroup <= 32 * kMaxMaskWords, GG_ERR_UNSUPPORTED,
           "gg_centroid_unit_vectors: C=%d exceeds the %d geocells the loss kernels cover", C,
           32 * kMaxMaskWords * kCellsPerGroup);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* keys = static_cast<uint32_t*>(workspace);
  const TableView tv = view_table(cent_table, C);
  table_unit_kernel<<<ceil_div(tv.Cpad, 256), 256, 0, s>>>(centroids, cent_table, C, keys);
  GG_LAUNCH_CHECK();
  table_rank_kernel<<<ceil_div(C, 256), 256, 0, s>>>(keys, C, const_cast<float4*>(tv.scell));
  GG_LAUNCH_CHECK();
  table_group_kernel<<<ceil_div(tv.ngroups, 8), 256, 0, s>>>(cent_table, C);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" size_t gg_hav_row_stats_bytes(int B, int C) { return make_stats_plan(B, C).bytes; }

extern "C" int gg_hav_row_stats(const float* labels, const float* cent_table, int B, int C, float tau, float far_km,
                                void* row_stats, long long* nearest_cell, float* nearest_km, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && labels && cent_table && row_stats, GG_ERR_ARG, "gg_hav_row_stats: bad arguments");
  GG_CHECK(tau > 0.f, GG_ERR_ARG, "gg_hav_row_stats: tau must be positive");
  GG_CHECK(far_km >= 1.0f, GG_ERR_ARG, "gg_hav_row_stats: far_km must be >= 1 km (inf disables the skip)");
  const TableView tv = view_table(cent_table, C);
  GG_CHECK(tv.ngroups <= 128 * kMaxSlots && tv.Cpad / kCellsPerGroup <= 32 * kMaxMaskWords, GG_ERR_UNSUPPORTED,
           "gg_hav_row_stats: C=%d exceeds the %d geocells the row-statistics kernel covers", C,
           128 * kMaxSlots * kCellsPerGroup);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const StatsPlan p = make_stats_plan(B, C);
  uint8_t* base = static_cast<uint8_t*>(row_stats);
  RowRec* rec = reinterpret_cast<RowRec*>(base + p.off_rec);
  uint32_t* near = reinterpret_cast<uint32_t*>(base + p.off_near);
  unsigned int* counter = reinterpret_cast<unsigned int*>(base + p.off_counter);
  float4* lab = reinterpret_cast<float4*>(base + p.off_lab);
  label_xyz_kernel<<<ceil_div(std::max(B, p.ncounters), 128), 128, 0, s>>>(
      labels, lab, B, counter, p.ncounters, reinterpret_cast<unsigned long long*>(base + p.off_acc));
  GG_LAUNCH_CHECK();
  // phi = far / R; far = inf (skip disabled) or >= pi R  ->  cos(phi/2) <= 0  ->  everything is near
  double phi = static_cast<double>(far_km) / 6378.137;
  if (!(phi < 3.141592653589793)) phi = 3.141592653589793;
  const float k2 = (1.0f / tau) * kLog2eF;
  const float chf = static_cast<float>(cos(0.5 * phi)), shf = static_cast<float>(sin(0.5 * phi));
  const int slots = ceil_div(tv.ngroups, 128);
  const int grid_a = B;
#define GG_HAV_A(S)                                                                                          \
  hav_row_stats_kernel<S><<<grid_a, 128, 0, s>>>(lab, cent_table, C, B, k2, chf, shf, static_cast<float>(phi), \
                                                 rec, near, p.nwp, nearest_cell, nearest_km)
  if (slots <= 1) GG_HAV_A(1);
  else GG_HAV_A(2);
#undef GG_HAV_A
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" size_t gg_hav_ce_workspace_bytes(int B, int C) { return make_hav_plan(B, C).bytes; }
// db_partials is (gg_hav_ce_db_parts(B, C), gg_hav_cpad(C)) fp32: one row of column sums per row block
extern "C" int gg_hav_ce_db_parts(int B, int C) { return make_hav_plan(B, C).nrb; }

extern "C" int gg_hav_ce_fwd_bwd(const void* logits_bf16, int ldc, const float* lse, const void* row_stats,
                                 const float* cent_table, int B, int C, float tau, void* dlogits_bf16,
                                 float* loss_rows, float* db_partials, void* workspace, float* loss_mean,
                                 float mean_scale, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: empty problem B=%d C=%d", B, C);
  GG_CHECK(logits_bf16 && lse && row_stats && cent_table && dlogits_bf16 && loss_rows && workspace, GG_ERR_ARG,
           "gg_hav_ce_fwd_bwd: null pointer");
  const HavPlan p = make_hav_plan(B, C);
  const StatsPlan sp = make_stats_plan(B, C);
  GG_CHECK(ldc >= p.Cpad && ldc % 8 == 0, GG_ERR_ARG,
           "gg_hav_ce_fwd_bwd: ldc=%d must be >= %d (C rounded up to 256) and a multiple of 8", ldc, p.Cpad);
  GG_CHECK(tau > 0.f, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: tau must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* sb = static_cast<const uint8_t*>(row_stats);
  const RowRec* rec = reinterpret_cast<const RowRec*>(sb + sp.off_rec);
  const uint32_t* near = reinterpret_cast<const uint32_t*>(sb + sp.off_near);
  unsigned int* counter = reinterpret_cast<unsigned int*>(const_cast<uint8_t*>(sb) + sp.off_counter);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  unsigned long long* row_acc = reinterpret_cast<unsigned long long*>(const_cast<uint8_t*>(sb) + sp.off_acc);
  float* block_part = reinterpret_cast<float*>(ws + p.off_block_part);
  const float neg_rk2 = -kEarthRadiusKm * (1.0f / tau) * kLog2eF;

  CUtensorMap tm_logits;
  int rc = make_tmap_bf16_2d_plain(&tm_logits, logits_bf16, static_cast<uint64_t>(p.Cpad), static_cast<uint64_t>(B),
                                   static_cast<uint64_t>(ldc) * 2, kColsPerWarp, kStageRows);
  if (rc) return rc;
  const int grid_b = p.nrb * p.nstrips;
  const size_t smem = sizeof(StreamSmem) + 128;
  if (db_partials) {
    auto kern = hav_ce_stream_kernel<true>;
    if (int e = set_max_dynamic_smem_once(kern, smem)) return e;
    kern<<<grid_b, kStreamThreads, smem, s>>>(tm_logits, ldc, lse, rec, near, sp.nwp, cent_table, B, C, p.rows_per_block,
                                              p.nslices, p.nstrips, neg_rk2, static_cast<bf16*>(dlogits_bf16),
                                              row_acc, db_partials, counter, loss_rows, block_part, mean_scale,
                                              loss_mean);
  } else {
    auto kern = hav_ce_stream_kernel<false>;
    if (int e = set_max_dynamic_smem_once(kern, smem)) return e;
    kern<<<grid_b, kStreamThreads, smem, s>>>(tm_logits, ldc, lse, rec, near, sp.nwp, cent_table, B, C, p.rows_per_block,
                                              p.nslices, p.nstrips, neg_rk2, static_cast<bf16*>(dlogits_bf16),
                                              row_acc, nullptr, counter, loss_rows, block_part, mean_scale,
                                              loss_mean);
  }
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_hard_ce_fwd_bwd(const void* logits_bf16, int ldc, const float* lse, const long long* labels_clf,
                                  int B, int C, void* dlogits_bf16, float* loss_rows, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && logits_bf16 && lse && labels_clf && dlogits_bf16 && loss_rows, GG_ERR_ARG,
           "gg_hard_ce_fwd_bwd: bad arguments");
  GG_CHECK(ldc >= C && ldc % 4 == 0, GG_ERR_ARG, "gg_hard_ce_fwd_bwd: ldc=%d must be >= C and a multiple of 4", ldc);
  const int grid = std::min(B, 4 * device_sm_count());
  hard_ce_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(logits_bf16), ldc, lse,
                                                                     labels_clf, B, C,
                                                                     static_cast<bf16*>(dlogits_bf16), loss_rows);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_loss_mean(const float* loss_rows, int B, float scale, float* loss_out, gg_stream_t stream) {
  GG_CHECK(loss_rows && loss_out && B > 0, GG_ERR_ARG, "gg_loss_mean: bad arguments");
  loss_mean_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(loss_rows, B, scale, loss_out);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
