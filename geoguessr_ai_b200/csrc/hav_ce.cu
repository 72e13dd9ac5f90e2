// Fused haversine label-smoothed geocell cross-entropy, forward + gradient in one kernel.
//
// Replaces, per training step, the reference's
//   haversine_matrix(labels, centroids.t())        models/utils.py:39-57   (B x C distances)
//   smooth_labels(d) = exp(-(d - rowmin d)/65)      models/utils.py:20-32
//   t = s / max(sum_c s, 1e-12)                     models/super_guessr.py:377
//   loss_b = -sum_c t * log_softmax(logits)         models/super_guessr.py:379-380
//   dlogits = softmax(logits) - t                   (autograd, main_coordinator_idun_s3.py:423)
// without materialising any B x C distance / target / probability matrix: per row the only HBM
// traffic is one read of the bf16 logits and one write of the bf16 gradient.
//
// Distance: with unit vectors u (label) and v_c (centroid), haversine's a = |u - v_c|^2 / 4 exactly,
// so q = |u - v_c|^2 (3 sub + 3 fma from registers / shared memory, no per-element trig, no
// cancellation at small distances) and d = 2R asin(sqrt(q)/2).  q is monotone in d, so the row
// minimum and the "can this cell carry any target mass" test run on q; sqrt/asin/exp are only
// evaluated for cells with d < dmin + far_km, where far_km = 65 km * ln(2^40) by default: beyond
// it exp(-(d-dmin)/65) < 2^-40 relative to the nearest cell's weight of 1 (far_km = inf disables
// the skip).  Geocells are ordered by (country, admin1, id), so chunks of 4 x 32 consecutive
// cells are geographically coherent and the skip is close to warp-uniform.
//
// One persistent CTA per SM keeps the whole centroid unit-vector table (3 x C fp32 = 152 KB at
// C = 12 647) resident in shared memory and walks rows b = blockIdx.x, +gridDim.x, ...; a thread
// owns the same 4-cell chunks in every pass, so q / s stay in its registers.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr float kEarthRadiusKm = 6378.137f;  // models/utils.py:55 (6378137 m) / 1000
constexpr float kLog2eF = 1.4426950408889634f;

// theta = 2 asin(sqrt(q)/2), q = squared chord in [0, 4].  Cephes-style asinf (abs err ~1e-7).
__device__ __forceinline__ float asin_poly(float r, float z) {  // asin(r) for r <= 0.5, z = r*r
  float p = 4.2163199048e-2f;
  p = fmaf(p, z, 2.4181311049e-2f);
  p = fmaf(p, z, 4.5470025998e-2f);
  p = fmaf(p, z, 7.4953002686e-2f);
  p = fmaf(p, z, 1.6666752422e-1f);
  return fmaf(p * z, r, r);
}
__device__ __forceinline__ float theta_from_q(float q) {
  const float h = fminf(0.25f * q, 1.0f);  // = sin^2(theta/2) = haversine 'a'
  if (h <= 0.25f) {
    return 2.0f * asin_poly(sqrtf(h), h);
  }
  const float x = sqrtf(h);
  const float z = 0.5f * (1.0f - x);
  return CUDART_PI_F - 4.0f * asin_poly(sqrtf(z), z);
}

// Unit vectors of the geocell centroids, SoA [x(0..Cpad) | y | z], fp64 trig rounded to fp32.
// Pad entries sit far outside the unit sphere so that their q is huge ("infinitely far").
__global__ void centroid_xyz_kernel(const float* __restrict__ centroids, float* __restrict__ xyz, int C, int Cpad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  float x = 1.0e9f, y = 1.0e9f, z = 1.0e9f;
  if (c < C) {
    const double lng = static_cast<double>(centroids[2 * c]) * (CUDART_PI / 180.0);
    const double lat = static_cast<double>(centroids[2 * c + 1]) * (CUDART_PI / 180.0);
    x = static_cast<float>(cos(lat) * cos(lng));
    y = static_cast<float>(cos(lat) * sin(lng));
    z = static_cast<float>(sin(lat));
  }
  xyz[c] = x;
  xyz[Cpad + c] = y;
  xyz[2 * Cpad + c] = z;
}

// Unit vectors of the labels: (B,4) = {x, y, z, valid}.  Non-finite labels give valid = 0
// (the reference's nan_to_num turns such a row's targets into zeros, utils.py:31).
__global__ void label_xyz_kernel(const float* __restrict__ labels, float4* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
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

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int SLOTS>
__global__ void __launch_bounds__(1024, 1)
hav_ce_kernel(const bf16* __restrict__ logits, int ldc, const float* __restrict__ lse,
              const float4* __restrict__ lab_xyz, const float* __restrict__ cent_xyz, int Cpad, int B, int C,
              float inv_tau, float cos_half_far, float sin_half_far, bf16* __restrict__ dlogits,
              float* __restrict__ loss_rows, long long* __restrict__ nearest_cell,
              float* __restrict__ nearest_km, float* __restrict__ db_partials) {
  extern __shared__ float4 smem_f4[];
  const int nchunks = Cpad >> 2;
  float4* cx = smem_f4;
  float4* cy = cx + nchunks;
  float4* cz = cy + nchunks;
  float4* dbs = cz + nchunks;  // per-CTA column sums of the gradient (bias gradient), owned per thread
  float* red_min = reinterpret_cast<float*>(dbs + nchunks);  // [32]
  float* red_sum = red_min + 32;                      // [32]
  float* red_loss = red_sum + 32;                     // [2][32]
  int* s_argmin = reinterpret_cast<int*>(red_loss + 64);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;

  {  // centroid table -> shared memory (L2 hits after the first CTA)
    const float4* gx = reinterpret_cast<const float4*>(cent_xyz);
    for (int i = tid; i < 3 * nchunks; i += nthr) smem_f4[i] = __ldg(gx + i);
    for (int i = tid; i < nchunks; i += nthr) dbs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) *s_argmin = 0x7fffffff;
  __syncthreads();

  const float k2 = inv_tau * kLog2eF;  // s = 2^((dmin - d) * k2)
  int parity = 0;
  int prev_row = -1;
  float prev_lse = 0.f, prev_valid = 0.f;

  // software pipeline over rows: the next row's logits (the only HBM read), label vector and lse are
  // requested while the current row is processed
  uint2 lraw_next[SLOTS];
  float4 u_next = make_float4(0.f, 0.f, 1.f, 0.f);
  float lse_next = 0.f;
  auto fetch_row = [&](int r) {
    const uint2* lrow = reinterpret_cast<const uint2*>(logits + static_cast<size_t>(r) * ldc);
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int g = j * nthr + tid;
      lraw_next[j] = (j < SLOTS - 1 || g < nchunks) ? __ldcs(lrow + g) : make_uint2(0u, 0u);
    }
    u_next = __ldg(lab_xyz + r);
    lse_next = __ldg(lse + r);
  };
  if (blockIdx.x < B) fetch_row(blockIdx.x);

  for (int row = blockIdx.x; row < B; row += gridDim.x, parity ^= 1) {
    bf16* grow = dlogits + static_cast<size_t>(row) * ldc;
    uint2 lraw[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) lraw[j] = lraw_next[j];
    const float4 u = u_next;
    const float row_lse = lse_next;
    if (row + gridDim.x < B) fetch_row(row + gridDim.x);

    // ---- pass 1a: squared chords to every centroid, row minimum
    float qmin = CUDART_INF_F;
    float4 qs[SLOTS];  // q, then s for the near chunks; a thread only ever touches its own chunks
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int g = j * nthr + tid;
      qs[j] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
      if (j < SLOTS - 1 || g < nchunks) {
        const float4 x = cx[g], y = cy[g], z = cz[g];
        float4 q;
        float dx, dy, dz;
        dx = u.x - x.x; dy = u.y - y.x; dz = u.z - z.x; q.x = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        dx = u.x - x.y; dy = u.y - y.y; dz = u.z - z.y; q.y = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        dx = u.x - x.z; dy = u.y - y.z; dz = u.z - z.z; q.z = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        dx = u.x - x.w; dy = u.y - y.w; dz = u.z - z.w; q.w = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        qs[j] = q;
        qmin = fminf(qmin, fminf(fminf(q.x, q.y), fminf(q.z, q.w)));
      }
    }
    qmin = warp_min(qmin);
    if (lane == 0) red_min[warp] = qmin;
    __syncthreads();  // (1)
    {
      float m = red_min[lane < nwarps ? lane : 0];
      qmin = warp_min(m);
    }
    // previous row's loss: its partials were published before barrier (1)
    if (tid < 32 && prev_row >= 0) {
      float sl = red_loss[(parity ^ 1) * 32 + (lane < nwarps ? lane : 0)];
      if (lane >= nwarps) sl = 0.f;
      sl = warp_sum(sl);
      if (lane == 0) loss_rows[prev_row] = prev_valid != 0.f ? prev_lse - sl : 0.f;
    }
    const float dmin = kEarthRadiusKm * theta_from_q(qmin);
    // near test on q: d < dmin + far  <=>  q < 4 sin^2((theta_min + phi)/2), phi = far / R, expanded
    // with sin(theta_min/2) = sqrt(qmin)/2 so that no trig runs per row; everything is "near"
    // once theta_min + phi reaches pi.
    float q_thr = CUDART_INF_F;
    {
      const float hmin = fminf(0.25f * qmin, 1.0f);
      const float sa = sqrtf(hmin), ca = sqrtf(1.0f - hmin);
      if (ca * cos_half_far - sa * sin_half_far > 0.f) {
        const float sh = fmaf(sa, cos_half_far, ca * sin_half_far);
        q_thr = 4.0f * sh * sh;
      }
    }

    // ---- pass 1b: unnormalised targets s for the cells that can carry mass, and their sum
    uint32_t near_mask = 0;
    float ssum = 0.f;
    const float off = dmin * k2;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int g = j * nthr + tid;
      {
        const float4 q = qs[j];  // +inf for slots past the table
        if (fminf(fminf(q.x, q.y), fminf(q.z, q.w)) < q_thr) {
          float4 s;
          s.x = ex2_approx(fmaf(-kEarthRadiusKm * k2, theta_from_q(q.x), off));
          s.y = ex2_approx(fmaf(-kEarthRadiusKm * k2, theta_from_q(q.y), off));
          s.z = ex2_approx(fmaf(-kEarthRadiusKm * k2, theta_from_q(q.z), off));
          s.w = ex2_approx(fmaf(-kEarthRadiusKm * k2, theta_from_q(q.w), off));
          if (4 * g + 3 >= C) {  // pad cells of the last chunk
            if (4 * g + 0 >= C) s.x = 0.f;
            if (4 * g + 1 >= C) s.y = 0.f;
            if (4 * g + 2 >= C) s.z = 0.f;
            s.w = 0.f;
          }
          if (nearest_cell != nullptr) {
            int hit = 0x7fffffff;
            if (q.w == qmin) hit = 4 * g + 3;
            if (q.z == qmin) hit = 4 * g + 2;
            if (q.y == qmin) hit = 4 * g + 1;
            if (q.x == qmin) hit = 4 * g + 0;
            if (hit != 0x7fffffff) atomicMin(s_argmin, hit);
          }
          qs[j] = s;
          near_mask |= 1u << j;
          ssum += (s.x + s.y) + (s.z + s.w);
        }
      }
    }
    ssum = warp_sum(ssum);
    if (lane == 0) red_sum[warp] = ssum;
    __syncthreads();  // (2)
    {
      float m = red_sum[lane < nwarps ? lane : 0];
      if (lane >= nwarps) m = 0.f;
      ssum = warp_sum(m);
    }
    // s / max(sum, 1e-12) (super_guessr.py:377); invalid label -> zero targets
    const float inv_s = u.w != 0.f ? 1.0f / fmaxf(ssum, 1e-12f) : 0.f;
    if (tid == 0 && nearest_cell != nullptr) {
      nearest_cell[row] = *s_argmin;
      if (nearest_km) nearest_km[row] = dmin;
      *s_argmin = 0x7fffffff;  // next atomicMin on it happens after the next barrier (1)
    }

    // ---- pass 2: p = exp(l - lse), gradient p - t, loss partial sum_c t * l
    const float lse2 = row_lse * kLog2eF;
    float sl = 0.f;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int g = j * nthr + tid;
      if (j < SLOTS - 1 || g < nchunks) {
        float l0 = __uint_as_float(lraw[j].x << 16), l1 = __uint_as_float(lraw[j].x & 0xffff0000u);
        float l2 = __uint_as_float(lraw[j].y << 16), l3 = __uint_as_float(lraw[j].y & 0xffff0000u);
        float g0 = ex2_approx(fmaf(l0, kLog2eF, -lse2));
        float g1 = ex2_approx(fmaf(l1, kLog2eF, -lse2));
        float g2 = ex2_approx(fmaf(l2, kLog2eF, -lse2));
        float g3 = ex2_approx(fmaf(l3, kLog2eF, -lse2));
        if (near_mask & (1u << j)) {
          const float4 s = qs[j];
          if (4 * g + 3 >= C) {  // pad logits are never defined
            if (4 * g + 0 >= C) l0 = 0.f;
            if (4 * g + 1 >= C) l1 = 0.f;
            if (4 * g + 2 >= C) l2 = 0.f;
            l3 = 0.f;
          }
          const float t0 = s.x * inv_s, t1 = s.y * inv_s, t2 = s.z * inv_s, t3 = s.w * inv_s;
          sl = fmaf(t0, l0, sl); sl = fmaf(t1, l1, sl); sl = fmaf(t2, l2, sl); sl = fmaf(t3, l3, sl);
          g0 -= t0; g1 -= t1; g2 -= t2; g3 -= t3;
        }
        uint2 o;
        o.x = pack_bf16x2(g0, g1);
        o.y = pack_bf16x2(g2, g3);
        __stcs(reinterpret_cast<uint2*>(grow) + g, o);
        if (db_partials != nullptr) {  // sum the bf16-rounded values the dW GEMM will see
          float4 a = dbs[g];
          a.x += __uint_as_float(o.x << 16); a.y += __uint_as_float(o.x & 0xffff0000u);
          a.z += __uint_as_float(o.y << 16); a.w += __uint_as_float(o.y & 0xffff0000u);
          dbs[g] = a;
        }
      }
    }
    sl = warp_sum(sl);
    if (lane == 0) red_loss[parity * 32 + warp] = sl;
    prev_row = row;
    prev_lse = row_lse;
    prev_valid = u.w;
  }
  __syncthreads();
  if (tid < 32 && prev_row >= 0) {
    float sl = red_loss[(parity ^ 1) * 32 + (lane < nwarps ? lane : 0)];
    if (lane >= nwarps) sl = 0.f;
    sl = warp_sum(sl);
    if (lane == 0) loss_rows[prev_row] = prev_valid != 0.f ? prev_lse - sl : 0.f;
  }
  if (db_partials != nullptr) {  // each chunk was only ever touched by its owning thread
    float4* out = reinterpret_cast<float4*>(db_partials + static_cast<size_t>(blockIdx.x) * Cpad);
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const int g = j * nthr + tid;
      if (j < SLOTS - 1 || g < nchunks) out[g] = dbs[g];
    }
  }
}

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

}  // namespace gg

using namespace gg;

extern "C" int gg_hav_cpad(int C) { return ceil_div(C, 4) * 4; }

extern "C" int gg_centroid_unit_vectors(const float* centroids, float* cent_xyz, int C, gg_stream_t stream) {
  GG_CHECK(centroids && cent_xyz && C > 0, GG_ERR_ARG, "gg_centroid_unit_vectors: bad arguments");
  const int Cpad = gg_hav_cpad(C);
  centroid_xyz_kernel<<<ceil_div(Cpad, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(centroids, cent_xyz, C, Cpad);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" size_t gg_hav_ce_workspace_bytes(int B) { return static_cast<size_t>(B) * sizeof(float4); }
// db_partials is (gg_hav_ce_db_parts(B), gg_hav_cpad(C)) fp32: one row of column sums per CTA
extern "C" int gg_hav_ce_db_parts(int B) { return std::min(B, device_sm_count()); }

template <int SLOTS>
static int launch_hav(const void* logits, int ldc, const float* lse, const float4* lab, const float* cent_xyz,
                      int Cpad, int B, int C, float tau, float far_km, void* dlogits, float* loss_rows,
                      long long* nearest_cell, float* nearest_km, float* db_partials, int nthr, size_t smem,
                      cudaStream_t s) {
  auto kern = hav_ce_kernel<SLOTS>;
  GG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = std::min(B, device_sm_count());
  // phi/2 = far / (2R); far = inf (skip disabled) or >= pi R  ->  cos <= 0  ->  everything is near
  double half_phi = 0.5 * static_cast<double>(far_km) / 6378.137;
  if (!(half_phi < 1.5707963267948966)) half_phi = 1.5707963267948966;
  kern<<<grid, nthr, smem, s>>>(static_cast<const bf16*>(logits), ldc, lse, lab, cent_xyz, Cpad, B, C, 1.0f / tau,
                                static_cast<float>(cos(half_phi)), static_cast<float>(sin(half_phi)),
                                static_cast<bf16*>(dlogits), loss_rows, nearest_cell, nearest_km, db_partials);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_hav_ce_fwd_bwd(const void* logits_bf16, int ldc, const float* lse, const float* labels,
                                 const float* cent_xyz, int B, int C, float tau, float far_km, void* dlogits_bf16,
                                 float* loss_rows, long long* nearest_cell, float* nearest_km, float* db_partials,
                                 void* workspace, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: empty problem B=%d C=%d", B, C);
  GG_CHECK(logits_bf16 && lse && labels && cent_xyz && dlogits_bf16 && loss_rows && workspace, GG_ERR_ARG,
           "gg_hav_ce_fwd_bwd: null pointer");
  GG_CHECK(ldc >= C && ldc % 4 == 0, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: ldc=%d must be >= C and a multiple of 4", ldc);
  GG_CHECK(tau > 0.f, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: tau must be positive");
  GG_CHECK(far_km >= 1.0f, GG_ERR_ARG, "gg_hav_ce_fwd_bwd: far_km must be >= 1 km (inf disables the skip)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Cpad = gg_hav_cpad(C);
  const int nchunks = Cpad / 4;
  const size_t smem = static_cast<size_t>(nchunks) * 4 * sizeof(float4) + 160 * sizeof(float);
  GG_CHECK(smem <= 227 * 1024, GG_ERR_UNSUPPORTED,
           "gg_hav_ce_fwd_bwd: C=%d needs %zu B of shared memory for the resident centroid table (max 232448); "
           "geocell tables above ~14.5k cells are not supported yet", C, smem);
  float4* lab = static_cast<float4*>(workspace);
  label_xyz_kernel<<<ceil_div(B, 256), 256, 0, s>>>(labels, lab, B);
  GG_LAUNCH_CHECK();
  int slots = ceil_div(nchunks, 1024);
  int nthr = ceil_div(ceil_div(nchunks, slots), 32) * 32;
  if (nthr < 128) nthr = 128;
  switch (slots) {
    case 1: return launch_hav<1>(logits_bf16, ldc, lse, lab, cent_xyz, Cpad, B, C, tau, far_km, dlogits_bf16, loss_rows, nearest_cell, nearest_km, db_partials, nthr, smem, s);
    case 2: return launch_hav<2>(logits_bf16, ldc, lse, lab, cent_xyz, Cpad, B, C, tau, far_km, dlogits_bf16, loss_rows, nearest_cell, nearest_km, db_partials, nthr, smem, s);
    case 3: return launch_hav<3>(logits_bf16, ldc, lse, lab, cent_xyz, Cpad, B, C, tau, far_km, dlogits_bf16, loss_rows, nearest_cell, nearest_km, db_partials, nthr, smem, s);
    default: return launch_hav<4>(logits_bf16, ldc, lse, lab, cent_xyz, Cpad, B, C, tau, far_km, dlogits_bf16, loss_rows, nearest_cell, nearest_km, db_partials, nthr, smem, s);
  }
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
