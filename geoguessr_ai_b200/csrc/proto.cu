// ProtoRefiner hot path: top-k geocell -> prototype retrieval + coordinate refinement.
//
// Replaces the Python double loop of models/proto_refiner.py:165-228 (per query, per candidate
// cell: host->device copy of the cell's prototypes, torch.cdist, max/argmax with .item() syncs).
//
// Stage 0 (grouping)  : counting-sort the B*topk (query, candidate) pairs by geocell and gather
//                       the query rows into cell order (Qs), so that every geocell sees its
//                       queries as one contiguous bf16 matrix.
// Stage 1 (retrieval) : persistent tcgen05 kernel.  Work item = (geocell c, chunk of <=128 pairs):
//                       D[128 pairs, 256 prototypes] = Qs_chunk . bank_c^T over K = D, operands
//                       streamed by TMA (the bank tile is read from HBM exactly once per chunk --
//                       the algorithmic traffic of the path), accumulator in TMEM (2 x 256 cols).
//                       Epilogue thread = one pair: dist^2 = |q|^2 + |p|^2 - 2 q.p, running arg-min
//                       over the cell's prototypes (proto_refiner.py:190-194: max of -cdist), then
//                       one 16-byte record {score, lng, lat, prototype id} per pair.
// Stage 2 (refinement): per query temperature softmax over the k scores (no max-subtraction,
//                       :378-389) x candidate probabilities (:210), arg-max, 1000 km guard
//                       (:216-223, geo_utils.py:39-54), output coords / geocell (:225-228).
// Multi-GPU: the bank is sharded by contiguous geocell ranges [cell_lo, cell_hi); a rank leaves
// score = -inf in records of pairs whose cell it does not own; after an all-gather of the
// record arrays stage 2 selects, per pair, the owning rank's record.
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kPM = 128;  // pairs per work item (UMMA M)
constexpr int kPN = 256;  // prototypes per accumulation unit (UMMA N)
constexpr int kPK = 64;
constexpr int kPBox = 32;  // granularity of a bank load: a unit fetches its group's prototypes rounded up to 32 rows
struct BankMaps {          // the bank (P_local, D) through boxes of 32, 64, ... 256 rows: one TMA per stage whatever the size
  CUtensorMap m[kPN / kPBox];
};
constexpr int kPStages = 4;
constexpr int kProtoThreads = 192;
constexpr uint32_t kPStageA = kPM * kPK * 2;
constexpr uint32_t kPStageB = kPN * kPK * 2;
constexpr float kMissingScore = -100000.0f;  // proto_refiner.py:185

// ------------------------------------------------------------------ stage 0: grouping
// Geocells are packed into GROUPS when the bank is installed (gg_proto_group_cells): consecutive cells whose
// prototypes fit one 256-row accumulation unit share it (a mean cell of the 1 M bank has 79 prototypes: alone it
// would fill a third of the unit and its neighbours' rows would be loaded -- and thrown away -- with it); a cell
// with more than 256 prototypes is a group of its own.  A work item is (group, chunk of <= 128 pairs); the pairs of
// a group are contiguous because pairs are sorted by cell, and every pair carries its own cell's prototype range,
// which masks the unit's columns in the epilogue.
//
// rec[p] initialised to "missing" (owned cell) or "not mine" (-inf); counts per owned cell.
__global__ void proto_count_kernel(const long long* __restrict__ cand, int cand_ld, int B, int topk, int cell_lo,
                                   int cell_hi, int* __restrict__ cnt, float4* __restrict__ rec) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B * topk) return;
  const long long c = cand[static_cast<size_t>(p / topk) * cand_ld + (p % topk)];
  const bool mine = c >= cell_lo && c < cell_hi;
  rec[p] = make_float4(mine ? kMissingScore : -CUDART_INF_F, 0.f, 0.f, __int_as_float(-1));
  if (mine) atomicAdd(&cnt[c - cell_lo], 1);
}

// Block-wide exclusive scan of one int per thread (1024 threads): warp shuffles + one pass over the 32 warp totals.
// Returns the exclusive prefix of `v`; *total receives the block sum.  `red` = 33 ints of shared scratch.
__device__ __forceinline__ int block_exclusive_scan(int v, int* red, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // (scratch reuse between calls)
  if (lane == 31) red[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = red[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    red[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int base = warp > 0 ? red[warp - 1] : 0;
  *total = red[31];
  return base + inc - v;
}

// Single CTA.  Pass 1: exclusive scan of the pair counts over the owned cells -> pair_off (ncell+1).  Pass 2: scan
// over the groups -> work list (group, chunk) for every group that has both pairs and prototypes.
// meta = {work items, pairs, accumulation units (work items x 256-prototype blocks of the group), 32-row bank boxes
// the retrieval kernel loads per k-block (work items x ceil(group size / 32))}.
__global__ void __launch_bounds__(1024)
proto_scan_kernel(const int* __restrict__ cnt, const int* __restrict__ cell_off, int ncell,
                  const int* __restrict__ group_off, int ngroups, int* __restrict__ pair_off,
                  int* __restrict__ cursor, int* __restrict__ work_group, int* __restrict__ work_chunk,
                  int* __restrict__ meta) {
  __shared__ int red[33];
  const int tid = threadIdx.x;
  {
    const int per = (ncell + 1023) / 1024;
    const int c0 = min(ncell, tid * per), c1 = min(ncell, c0 + per);
    int np = 0;
    for (int c = c0; c < c1; ++c) np += __ldg(cnt + c);
    int total;
    int pbase = block_exclusive_scan(np, red, &total);
    for (int c = c0; c < c1; ++c) {
      pair_off[c] = pbase;
      cursor[c] = 0;
      pbase += __ldg(cnt + c);
    }
    if (tid == 0) {
      pair_off[ncell] = total;
      meta[1] = total;
    }
    __syncthreads();  // pair_off complete (block-scope visibility of the global writes)
  }
  const int per = (ngroups + 1023) / 1024;
  const int g0 = min(ngroups, tid * per), g1 = min(ngroups, g0 + per);
  auto group_chunks = [&](int g, int& units, int& boxes) {
    const int c0 = __ldg(group_off + g), c1 = __ldg(group_off + g + 1);
    const int n = pair_off[c1] - pair_off[c0], np = __ldg(cell_off + c1) - __ldg(cell_off + c0);
    if (n <= 0 || np <= 0) { units = boxes = 0; return 0; }
    const int chunks = (n + kPM - 1) / kPM;
    units = chunks * ((np + kPN - 1) / kPN);
    boxes = chunks * ((np + kPBox - 1) / kPBox);
    return chunks;
  };
  int nt = 0, nu = 0, nb = 0;
  for (int g = g0; g < g1; ++g) {
    int u, b;
    nt += group_chunks(g, u, b);
    nu += u;
    nb += b;
  }
  int total_t, total_u, total_b;
  int tbase = block_exclusive_scan(nt, red, &total_t);
  block_exclusive_scan(nu, red, &total_u);
  block_exclusive_scan(nb, red, &total_b);
  for (int g = g0; g < g1; ++g) {
    int u, b;
    const int chunks = group_chunks(g, u, b);
    for (int m = 0; m < chunks; ++m) {
      work_group[tbase + m] = g;
      work_chunk[tbase + m] = m;
    }
    tbase += chunks;
  }
  if (tid == 0) {
    meta[0] = total_t;
    meta[2] = total_u;
    meta[3] = total_b;
  }
}

// slot = position of the pair in cell order: its pair id, its query row and its cell's prototype range
__global__ void proto_scatter_kernel(const long long* __restrict__ cand, int cand_ld, int B, int topk, int cell_lo,
                                     int cell_hi, const int* __restrict__ pair_off, int* __restrict__ cursor,
                                     const int* __restrict__ cell_off, int* __restrict__ pair_ids,
                                     int* __restrict__ slot_q, int2* __restrict__ slot_range) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B * topk) return;
  const long long c = cand[static_cast<size_t>(p / topk) * cand_ld + (p % topk)];
  if (c < cell_lo || c >= cell_hi) return;
  const int lc = static_cast<int>(c - cell_lo);
  const int slot = pair_off[lc] + atomicAdd(&cursor[lc], 1);
  pair_ids[slot] = p;
  slot_q[slot] = p / topk;
  slot_range[slot] = make_int2(cell_off[lc], cell_off[lc + 1]);
}

// One warp per grouped slot: copy the query row (bf16, D) into cell order (only when the retrieval kernel does
// not gather the rows itself).
__global__ void proto_gather_kernel(const bf16* __restrict__ q, const int* __restrict__ slot_q,
                                    const int* __restrict__ meta, int D, bf16* __restrict__ qs) {
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= meta[1]) return;
  const int qi = slot_q[slot];
  const uint4* src = reinterpret_cast<const uint4*>(q + static_cast<size_t>(qi) * D);
  uint4* dst = reinterpret_cast<uint4*>(qs + static_cast<size_t>(slot) * D);
  for (int i = threadIdx.x & 31; i < (D >> 3); i += 32) dst[i] = __ldg(src + i);
}

// ------------------------------------------------------------------ stage 1: retrieval
struct ProtoSmem {
  uint8_t a[kPStages][kPStageA];
  uint8_t b[kPStages][kPStageB];
  float pn[2][kPN];
  uint64_t full[kPStages];
  uint64_t empty[kPStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

// 4 rows of a 2-D bf16 tensor, picked by index, into 4 consecutive 128-byte rows of a swizzled tile
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

// GATHER4: the A operand (a work item's <= 128 query rows) is gathered straight from the query matrix by the TMA
// engine, four rows per instruction, one instruction per producer lane and k-block -- the rows never make the
// round trip through a cell-ordered copy in HBM.  Otherwise tm_q maps that copy (proto_gather_kernel).
// COSINE: score = q.p / (|q| |p|) (models/proto_refiner.py:347-362) instead of -|q - p| (:364-376).
template <bool GATHER4, bool COSINE>
__global__ void __launch_bounds__(kProtoThreads, 1)
proto_retrieve_kernel(const __grid_constant__ CUtensorMap tm_q,     // GATHER4: q (B, D), 1-row boxes; else Qs (slots, D)
                      const __grid_constant__ BankMaps tm_banks,    // bank (P_local, D), see BankMaps
                      const int* __restrict__ meta, const int* __restrict__ work_group,
                      const int* __restrict__ work_chunk, const int* __restrict__ group_off,
                      const int* __restrict__ pair_off, const int* __restrict__ pair_ids,
                      const int* __restrict__ slot_q, const int2* __restrict__ slot_range,
                      const int* __restrict__ cell_off, const float* __restrict__ q_n,
                      const float* __restrict__ pnorm, const float* __restrict__ pcoords, int proto_base, int D,
                      float4* __restrict__ rec) {
  extern __shared__ uint8_t smem_raw[];
  ProtoSmem& sm = *reinterpret_cast<ProtoSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (D + kPK - 1) / kPK;
  const int n_work = meta[0];

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_banks.m[kPN / kPBox - 1]);
    for (int s = 0; s < kPStages; ++s) {
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
    // ===== producer (whole warp: with GATHER4 every lane issues one 4-row gather per k-block) =====
    int s = 0;
    uint32_t ph = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int g = work_group[w];
      const int c0 = group_off[g], c1 = group_off[g + 1];
      const int a_row0 = pair_off[c0] + work_chunk[w] * kPM;
      const int a_last = pair_off[c1] - 1;  // last slot of the group: rows past it repeat this one (results unused)
      const int p0 = cell_off[c0], p1 = cell_off[c1];
      // Only the rows that hold pairs are loaded (a work item of the 1 M bank holds ~65 pairs, of the 10 M bank ~26):
      // 4-row groups when the TMA engine gathers them, 32-row boxes from the cell-ordered copy.  The MMA still spans
      // 128 rows; the rest of the tile keeps stale shared memory whose accumulator rows nobody reads.
      const int nq = min(kPM, pair_off[c1] - a_row0);
      const int granules = GATHER4 ? (nq + 3) >> 2 : (nq + 31) >> 5;
      const uint32_t a_bytes = static_cast<uint32_t>(granules) * (GATHER4 ? 4 : 32) * (kPK * 2);
      int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      const bool my_rows = GATHER4 && lane < granules;
      if (my_rows) {
        r0 = __ldg(slot_q + min(a_row0 + 4 * lane + 0, a_last));
        r1 = __ldg(slot_q + min(a_row0 + 4 * lane + 1, a_last));
        r2 = __ldg(slot_q + min(a_row0 + 4 * lane + 2, a_last));
        r3 = __ldg(slot_q + min(a_row0 + 4 * lane + 3, a_last));
      }
      for (int n0 = p0; n0 < p1; n0 += kPN) {
        // the bank side the same way: one box reaching up to the group's last prototype (rounded up to 32 rows) -- a
        // group of the 1 M bank holds ~200 prototypes, and a full 256-row box would fetch the next group's rows for
        // nothing.  (One TMA per stage through the map of that height: eight 32-row boxes instead cost more in TMA
        // issue slots than the bytes saved.)
        const int nbox = (min(kPN, p1 - n0) + kPBox - 1) / kPBox;
        const uint32_t b_bytes = static_cast<uint32_t>(nbox) * (kPBox * kPK * 2);
        const CUtensorMap* tm_bank = &tm_banks.m[nbox - 1];
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.empty[s], ph ^ 1);
          if (lane == 0) {
            mbar_arrive_expect_tx(&sm.full[s], a_bytes + b_bytes);
            tma_load_2d_hint(sm.b[s], tm_bank, &sm.full[s], kb * kPK, n0, kPolicyEvictFirst);
            if (!GATHER4) {
              for (int gq = 0; gq < granules; ++gq)
                tma_load_2d(sm.a[s] + gq * (32 * kPK * 2), &tm_q, &sm.full[s], kb * kPK, a_row0 + 32 * gq);
            }
          }
          if (GATHER4) {
            __syncwarp();  // the barrier is armed before any lane's bytes can land
            if (my_rows) tma_gather4(sm.a[s] + lane * 512, &tm_q, &sm.full[s], kb * kPK, r0, r1, r2, r3);
          }
          if (++s == kPStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kPM, kPN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int g = work_group[w];
        const int p0 = cell_off[group_off[g]], p1 = cell_off[group_off[g + 1]];
        for (int n0 = p0; n0 < p1; n0 += kPN, ++it) {
          const int acc = it & 1;
          const uint32_t acc_ph = (it >> 1) & 1;
          mbar_wait(&sm.acc_empty[acc], acc_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kPN;
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(&sm.full[s], ph);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sm.a[s]), b0 = smem_u32(sm.b[s]);
#pragma unroll
            for (int k = 0; k < kPK / 16; ++k) {
              const uint64_t da = umma_desc_sw128(a0 + k * 32, 16, 1024);
              const uint64_t db = umma_desc_sw128(b0 + k * 32, 16, 1024);
              umma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            umma_commit(&sm.empty[s]);
            if (++s == kPStages) { s = 0; ph ^= 1; }
          }
          umma_commit(&sm.acc_full[acc]);
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;        // pair slot within the chunk == TMEM lane
    const int et = threadIdx.x - 64;       // 0..127 among the epilogue threads
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int g = work_group[w];
      const int c0 = group_off[g], c1 = group_off[g + 1];
      const int a_row0 = pair_off[c0] + work_chunk[w] * kPM;
      const int nq = min(kPM, pair_off[c1] - a_row0);
      const int p0 = cell_off[c0], p1 = cell_off[c1];
      // this pair's cell: only its prototypes [lo, hi) of the unit's columns compete
      int2 rng = make_int2(0, 0);
      if (r < nq) rng = __ldg(slot_range + a_row0 + r);
      float best = CUDART_INF_F;
      int best_i = -1;
      for (int n0 = p0; n0 < p1; n0 += kPN, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        // prototype norms of this unit -> smem (pad with +inf: prototypes past the group / the bank);
        // cosine: 1 / |p| (0 for the pad and for zero vectors: their score is 0, never NaN)
        for (int i = et; i < kPN; i += 128) {
          float v = CUDART_INF_F;
          if (n0 + i < p1) v = __ldg(pnorm + n0 + i);
          if (COSINE) v = (n0 + i < p1 && v > 0.f) ? rsqrtf(v) : 0.f;
          sm.pn[acc][i] = v;
        }
        named_bar_sync(1, 128);
        mbar_wait(&sm.acc_full[acc], acc_ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kPN;
#pragma unroll 1
        for (int cc = 0; cc < kPN / 32; ++cc) {
          const int base = n0 + cc * 32;
          if (base >= p1) break;  // uniform across the CTA
          // columns of this chunk inside my cell's range (bit i = column base + i)
          const int lo = max(rng.x - base, 0), hi = min(rng.y - base, 32);
          const uint32_t msk = lo < hi ? ((hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo)) : 0u;
          if (!__any_sync(0xffffffffu, msk != 0u)) continue;  // (pairs are sorted by cell: most chunks are one warp's)
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + cc * 32, v);
          tmem_ld_wait();
          const float4* pn4 = reinterpret_cast<const float4*>(&sm.pn[acc][cc * 32]);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 pn = pn4[q4];
            float d0, d1, d2, d3;
            if (COSINE) {  // minimise -cos |q|
              d0 = -__uint_as_float(v[4 * q4 + 0]) * pn.x;
              d1 = -__uint_as_float(v[4 * q4 + 1]) * pn.y;
              d2 = -__uint_as_float(v[4 * q4 + 2]) * pn.z;
              d3 = -__uint_as_float(v[4 * q4 + 3]) * pn.w;
            } else {       // minimise |p|^2 - 2 q.p
              d0 = fmaf(-2.f, __uint_as_float(v[4 * q4 + 0]), pn.x);
              d1 = fmaf(-2.f, __uint_as_float(v[4 * q4 + 1]), pn.y);
              d2 = fmaf(-2.f, __uint_as_float(v[4 * q4 + 2]), pn.z);
              d3 = fmaf(-2.f, __uint_as_float(v[4 * q4 + 3]), pn.w);
            }
            const int j = base + 4 * q4;
            const uint32_t m4 = msk >> (4 * q4);
            if ((m4 & 1u) && d0 < best) { best = d0; best_i = j; }
            if ((m4 & 2u) && d1 < best) { best = d1; best_i = j + 1; }
            if ((m4 & 4u) && d2 < best) { best = d2; best_i = j + 2; }
            if ((m4 & 8u) && d3 < best) { best = d3; best_i = j + 3; }
          }
        }
        tc_fence_before();
        mbar_arrive(&sm.acc_empty[acc]);
      }
      if (r < nq && best_i >= 0) {
        const float qn = __ldg(q_n + __ldg(slot_q + a_row0 + r));
        float score;
        if (COSINE) score = qn > 0.f ? -best * rsqrtf(qn) : 0.f;
        else score = -sqrtf(fmaxf(best + qn, 0.f));
        const int p = __ldg(pair_ids + a_row0 + r);
        rec[p] = make_float4(score, __ldg(pcoords + 2 * static_cast<size_t>(best_i)),
                             __ldg(pcoords + 2 * static_cast<size_t>(best_i) + 1),
                             __int_as_float(proto_base + best_i));
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

// ------------------------------------------------------------------ stage 2: refinement
__device__ __forceinline__ float haversine_km(float lng1, float lat1, float lng2, float lat2) {
  // preprocessing/geo_utils.py:39-54 in fp32 (the fp64 radius is a 0-dim tensor: no promotion)
  const float d2r = 0.017453292519943295f;
  const float x0 = lng1 * d2r, x1 = lat1 * d2r, y0 = lng2 * d2r, y1 = lat2 * d2r;
  const float s1 = sinf((y1 - x1) * 0.5f), s0 = sinf((y0 - x0) * 0.5f);
  const float a = s1 * s1 + cosf(x1) * cosf(y1) * s0 * s0;
  const float c = 2.0f * asinf(sqrtf(a));
  return (6378137.0f * c) / 1000.0f;
}

// torch.argmax semantics: NaN counts as the maximum, first occurrence wins.
__device__ __forceinline__ int argmax_nan_first(const float* v, int k) {
  int bi = 0;
  float bv = v[0];
  if (bv != bv) return 0;
  for (int j = 1; j < k; ++j) {
    const float x = v[j];
    if (x != x) return j;
    if (x > bv) { bv = x; bi = j; }
  }
  return bi;
}

constexpr int kMaxTopk = 16;

__global__ void proto_take_image_coords_kernel(float4* __restrict__ rec, const float4* __restrict__ rec_img, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 im = rec_img[i];
  if (__float_as_int(im.w) >= 0) {
    float4 r = rec[i];
    r.y = im.y;
    r.z = im.z;
    rec[i] = r;
  }
}
__global__ void proto_record_ids_kernel(const float4* __restrict__ rec, long long n, long long* __restrict__ ids) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = __float_as_int(rec[i].w);
}

// main_coordinator_idun_s3.py:399-408, one CTA, fixed summation order
__global__ void topk_accuracy_kernel(const long long* __restrict__ topk_idx, int k, const long long* __restrict__ targets,
                                     int B, float* __restrict__ acc) {
  __shared__ int s1[32], sk[32];
  int c1 = 0, ck = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const long long t = targets[i];
    c1 += topk_idx[static_cast<size_t>(i) * k] == t ? 1 : 0;
    bool any = false;
    for (int j = 0; j < k; ++j) any |= topk_idx[static_cast<size_t>(i) * k + j] == t;
    ck += any ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    ck += __shfl_xor_sync(0xffffffffu, ck, o);
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = c1; sk[threadIdx.x >> 5] = ck; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t1 = 0, tk = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { t1 += s1[w]; tk += sk[w]; }
    acc[0] = static_cast<float>(t1) / static_cast<float>(B);
    acc[1] = static_cast<float>(tk) / static_cast<float>(B);
  }
}

__global__ void proto_refine_kernel(const float4* __restrict__ rec, int nranks, long long rank_stride,
                                    const float* __restrict__ cprobs, int cprobs_ld,
                                    const long long* __restrict__ cand, int cand_ld,
                                    const float* __restrict__ initial, int B, int topk, float temperature,
                                    float max_refinement_km, float* __restrict__ out_llh,
                                    long long* __restrict__ out_cell, int* __restrict__ out_guess,
                                    float* __restrict__ out_score, int* __restrict__ out_proto) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float score[kMaxTopk], lng[kMaxTopk], lat[kMaxTopk], cp[kMaxTopk], fin[kMaxTopk];
  float sum = 0.f;
  for (int j = 0; j < topk; ++j) {
    float4 best = make_float4(-CUDART_INF_F, 0.f, 0.f, __int_as_float(-1));
    for (int rk = 0; rk < nranks; ++rk) {
      const float4 t = rec[rk * rank_stride + static_cast<size_t>(i) * topk + j];
      if (t.x > best.x) best = t;  // exactly one rank owns the cell
    }
    if (!(best.x > -CUDART_INF_F)) best = make_float4(kMissingScore, 0.f, 0.f, __int_as_float(-1));
    score[j] = best.x;
    lng[j] = best.y;
    lat[j] = best.z;
    if (out_score) out_score[static_cast<size_t>(i) * topk + j] = best.x;
    if (out_proto) out_proto[static_cast<size_t>(i) * topk + j] = __float_as_int(best.w);
    cp[j] = cprobs ? cprobs[static_cast<size_t>(i) * cprobs_ld + j] : (j == 0 ? 1.f : 0.f);
    fin[j] = expf(score[j] / temperature);  // :387, no max-subtraction
    sum += fin[j];
  }
  for (int j = 0; j < topk; ++j) fin[j] = cp[j] * (fin[j] / sum);  // :389, :210 (0/0 -> NaN as in the reference)
  const int refined = argmax_nan_first(fin, topk);                // :211
  const float dist = haversine_km(initial[2 * i], initial[2 * i + 1], lng[refined], lat[refined]);
  int final_id = refined;
  if (dist > max_refinement_km) final_id = argmax_nan_first(cp, topk);  // :220-225
  out_llh[2 * i] = lng[final_id];
  out_llh[2 * i + 1] = lat[final_id];
  out_cell[i] = cand[static_cast<size_t>(i) * cand_ld + final_id];
  if (out_guess) out_guess[i] = final_id;
}

}  // namespace gg

using namespace gg;

// workspace layout (ints unless noted), ncell = cell_hi - cell_lo, npair = B * topk:
//   cnt[ncell] | pair_off[ncell+1] | cursor[ncell] | work_group[ncell + npair/128 + 1] | work_chunk[same] |
//   meta[4] | pair_ids[npair] | slot_q[npair] | slot_range[npair] (int2) | pad to 256 B | Qs[(npair + 128) * D] (bf16)
struct ProtoWs {
  int *cnt, *pair_off, *cursor, *work_group, *work_chunk, *meta, *pair_ids, *slot_q;
  int2* slot_range;
  bf16* qs;
  size_t bytes;
};
static ProtoWs carve_proto_ws(void* base, int ncell, long long npair, int D) {
  ProtoWs w;
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t nbytes) {
    uint8_t* r = p;
    p += (nbytes + 255) & ~size_t(255);
    return r;
  };
  const size_t nwork = static_cast<size_t>(ncell) + npair / kPM + 1;
  w.meta = reinterpret_cast<int*>(take(sizeof(int) * 4));  // first: gg_proto_retrieve_meta reads it back
  w.cnt = reinterpret_cast<int*>(take(sizeof(int) * ncell));
  w.pair_off = reinterpret_cast<int*>(take(sizeof(int) * (ncell + 1)));
  w.cursor = reinterpret_cast<int*>(take(sizeof(int) * ncell));
  w.work_group = reinterpret_cast<int*>(take(sizeof(int) * nwork));
  w.work_chunk = reinterpret_cast<int*>(take(sizeof(int) * nwork));
  w.pair_ids = reinterpret_cast<int*>(take(sizeof(int) * npair));
  w.slot_q = reinterpret_cast<int*>(take(sizeof(int) * npair));
  w.slot_range = reinterpret_cast<int2*>(take(sizeof(int2) * npair));
  w.qs = reinterpret_cast<bf16*>(take(sizeof(bf16) * static_cast<size_t>(npair + kPM) * D));
  w.bytes = static_cast<size_t>(p - static_cast<uint8_t*>(base));
  return w;
}

extern "C" size_t gg_proto_retrieve_workspace_bytes(int B, int topk, int D, int ncell) {
  return carve_proto_ws(nullptr, ncell, static_cast<long long>(B) * topk, D).bytes;
}

// HOST: pack consecutive geocells into groups of <= 256 prototypes (a larger cell is its own group).
// cell_off: (ncell+1) host ints; group_off: (ncell+1) host ints to fill; returns the number of groups.
extern "C" int gg_proto_group_cells(const int* cell_off, int ncell, int* group_off) {
  int ng = 0, c = 0;
  group_off[0] = 0;
  while (c < ncell) {
    int e = c + 1;
    while (e < ncell && cell_off[e + 1] - cell_off[c] <= kPN) ++e;
    group_off[++ng] = e;
    c = e;
  }
  return ng;
}

extern "C" int gg_proto_retrieve(const void* q_bf16, const float* q_sqnorm, int B, int D, const long long* cand,
                                 int cand_ld, int topk, const void* bank_bf16, const float* bank_sqnorm,
                                 const float* bank_coords, long long n_protos, const int* cell_off, int cell_lo,
                                 int cell_hi, const int* group_off, int ngroups, int proto_base, int metric, int flags,
                                 void* rec_out, void* workspace, gg_stream_t stream) {
  GG_CHECK(B > 0 && D > 0 && topk >= 1 && topk <= kMaxTopk, GG_ERR_ARG,
           "gg_proto_retrieve: bad sizes B=%d D=%d topk=%d (topk <= %d)", B, D, topk, kMaxTopk);
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "gg_proto_retrieve: D=%d must be a multiple of 8", D);
  GG_CHECK(cand_ld >= topk, GG_ERR_ARG, "gg_proto_retrieve: topk=%d exceeds the %d candidate columns", topk, cand_ld);
  GG_CHECK(cell_hi > cell_lo && cell_lo >= 0, GG_ERR_ARG, "gg_proto_retrieve: empty cell range [%d, %d)", cell_lo, cell_hi);
  GG_CHECK(q_bf16 && q_sqnorm && cand && cell_off && rec_out && workspace, GG_ERR_ARG, "gg_proto_retrieve: null pointer");
  GG_CHECK(n_protos >= 0 && n_protos < (1ll << 31), GG_ERR_ARG, "gg_proto_retrieve: n_protos out of range");
  GG_CHECK(n_protos == 0 || (bank_bf16 && bank_sqnorm && bank_coords), GG_ERR_ARG, "gg_proto_retrieve: null bank pointer");
  GG_CHECK(n_protos == 0 || (group_off && ngroups >= 1 && ngroups <= cell_hi - cell_lo), GG_ERR_ARG,
           "gg_proto_retrieve: group table missing (gg_proto_group_cells)");
  GG_CHECK(metric == GG_METRIC_L2 || metric == GG_METRIC_COSINE, GG_ERR_ARG, "gg_proto_retrieve: metric=%d", metric);
  GG_CHECK(static_cast<long long>(B) * topk < (1ll << 31), GG_ERR_ARG, "gg_proto_retrieve: B*topk overflows int32");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ncell = cell_hi - cell_lo;
  const long long npair = static_cast<long long>(B) * topk;
  ProtoWs w = carve_proto_ws(workspace, ncell, npair, D);
  float4* rec = static_cast<float4*>(rec_out);
  const bool gather4 = (flags & GG_RETRIEVE_NO_GATHER4) == 0;

  GG_CUDA(cudaMemsetAsync(w.cnt, 0, sizeof(int) * ncell, s));
  const int pb = static_cast<int>(ceil_div_ll(npair, 256));
  proto_count_kernel<<<pb, 256, 0, s>>>(cand, cand_ld, B, topk, cell_lo, cell_hi, w.cnt, rec);
  GG_LAUNCH_CHECK();
  if (n_protos == 0) {  // this rank owns cells but no prototypes: every owned pair stays "missing"
    GG_CUDA(cudaMemsetAsync(w.meta, 0, sizeof(int) * 4, s));
    return GG_OK;
  }
  proto_scan_kernel<<<1, 1024, 0, s>>>(w.cnt, cell_off, ncell, group_off, ngroups, w.pair_off, w.cursor, w.work_group,
                                        w.work_chunk, w.meta);
  GG_LAUNCH_CHECK();
  proto_scatter_kernel<<<pb, 256, 0, s>>>(cand, cand_ld, B, topk, cell_lo, cell_hi, w.pair_off, w.cursor, cell_off,
                                          w.pair_ids, w.slot_q, w.slot_range);
  GG_LAUNCH_CHECK();

  CUtensorMap tm_q;
  BankMaps tm_banks;
  int rc;
  if (gather4) {
    rc = make_tmap_bf16_2d(&tm_q, q_bf16, D, static_cast<uint64_t>(B), static_cast<uint64_t>(D) * 2, kPK, 1);
  } else {
    proto_gather_kernel<<<static_cast<int>(ceil_div_ll(npair, 8)), 256, 0, s>>>(static_cast<const bf16*>(q_bf16), w.slot_q,
                                                                             w.meta, D, w.qs);
    GG_LAUNCH_CHECK();
    rc = make_tmap_bf16_2d(&tm_q, w.qs, D, static_cast<uint64_t>(npair), static_cast<uint64_t>(D) * 2, kPK, 32);
  }
  if (rc) return rc;
  {  // the bank's maps only change when another bank is installed: keep the last set per host thread
    struct Key { const void* p; long long n; int D; };
    thread_local Key last{nullptr, 0, 0};
    thread_local BankMaps last_maps;
    if (last.p != bank_bf16 || last.n != n_protos || last.D != D) {
      for (int i = 0; i < kPN / kPBox; ++i) {
        rc = make_tmap_bf16_2d(&last_maps.m[i], bank_bf16, D, static_cast<uint64_t>(n_protos),
                               static_cast<uint64_t>(D) * 2, kPK, kPBox * (i + 1));
        if (rc) { last.p = nullptr; return rc; }
      }
      last = Key{bank_bf16, static_cast<long long>(n_protos), D};
    }
    tm_banks = last_maps;
  }
  const size_t smem = sizeof(ProtoSmem) + 1024;
  const int grid = device_sm_count();
#define GG_RETRIEVE(G4, COS)                                                                                       \
  do {                                                                                                             \
    auto kern = proto_retrieve_kernel<G4, COS>;                                                                    \
    if (int e = set_max_dynamic_smem_once(kern, smem)) return e;                                                   \
    kern<<<grid, kProtoThreads, smem, s>>>(tm_q, tm_banks, w.meta, w.work_group, w.work_chunk, group_off, w.pair_off, \
                                           w.pair_ids, w.slot_q, w.slot_range, cell_off, q_sqnorm, bank_sqnorm,    \
                                           bank_coords, proto_base, D, rec);                                       \
  } while (0)
  if (gather4) {
    if (metric == GG_METRIC_COSINE) GG_RETRIEVE(true, true);
    else GG_RETRIEVE(true, false);
  } else {
    if (metric == GG_METRIC_COSINE) GG_RETRIEVE(false, true);
    else GG_RETRIEVE(false, false);
  }
#undef GG_RETRIEVE
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_proto_refine(const void* rec, int nranks, long long rank_stride, const float* cand_probs,
                               int cand_probs_ld, const long long* cand, int cand_ld, const float* initial_llh, int B,
                               int topk, float temperature, float max_refinement_km, float* out_llh,
                               long long* out_cell, int* out_guess, float* out_score, int* out_proto,
                               gg_stream_t stream) {
  GG_CHECK(B > 0 && topk >= 1 && topk <= kMaxTopk && nranks >= 1, GG_ERR_ARG, "gg_proto_refine: bad sizes");
  GG_CHECK(rec && cand && initial_llh && out_llh && out_cell, GG_ERR_ARG, "gg_proto_refine: null pointer");
  GG_CHECK(cand_ld >= topk && (!cand_probs || cand_probs_ld >= topk), GG_ERR_ARG, "gg_proto_refine: topk exceeds candidate columns");
  proto_refine_kernel<<<ceil_div(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float4*>(rec), nranks, rank_stride, cand_probs, cand_probs_ld, cand, cand_ld, initial_llh, B,
      topk, temperature, max_refinement_km, out_llh, out_cell, out_guess, out_score, out_proto);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_proto_take_image_coords(void* rec, const void* rec_img, long long n, gg_stream_t stream) {
  GG_CHECK(rec && rec_img && n > 0, GG_ERR_ARG, "gg_proto_take_image_coords: bad arguments");
  proto_take_image_coords_kernel<<<static_cast<unsigned int>(ceil_div_ll(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float4*>(rec), static_cast<const float4*>(rec_img), n);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_proto_record_ids(const void* rec, long long n, long long* ids, gg_stream_t stream) {
  GG_CHECK(rec && ids && n > 0, GG_ERR_ARG, "gg_proto_record_ids: bad arguments");
  proto_record_ids_kernel<<<static_cast<unsigned int>(ceil_div_ll(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float4*>(rec), n, ids);
  GG_LAUNCH_CHECK();
  return GG_OK;
}

extern "C" int gg_topk_accuracy(const long long* topk_idx, int k, const long long* targets, int B, float* acc,
                                gg_stream_t stream) {
  GG_CHECK(topk_idx && targets && acc && B > 0 && k >= 1, GG_ERR_ARG, "gg_topk_accuracy: bad arguments");
  topk_accuracy_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(topk_idx, k, targets, B, acc);
  GG_LAUNCH_CHECK();
  return GG_OK;
}
