// Geocell head forward: logits = x W^T + b as a persistent, warp-specialised tcgen05 GEMM.
//
// Replaces  models/super_guessr.py:354 (nn.Linear), :355 (softmax), :358-361 (argmax + centroid
// gather) and :365 (top-k) of the reference.  Per 128x256 output tile the accumulator lives in
// TMEM (2 x 256 fp32 columns, double buffered); sixteen epilogue warps (four per TMEM lane quadrant,
// 64 columns each) add the bias, optionally write the bf16 logits (training only) and keep, per
// row, an online (max, sum-exp) and the running top-K logits, so that in serving the (B, C) logit /
// probability matrices never reach HBM.
//
// CTA pairs (tcgen05 cta_group::2): the kernel runs as clusters of two CTAs on the two SMs of a TPC, which
// execute ONE 256 x 256 x 16 MMA together: each CTA holds its own 128 rows of x and HALF of the W tile (128
// geocells) in shared memory, and its 128 x 256 half of the accumulator in its own TMEM.  Per CTA and k-block
// 32 KB instead of 48 KB cross the L2 -> SM fabric and sit in shared memory, which is what lets the tensor
// pipe run ahead of operand delivery.  The leader CTA (cluster rank 0) issues the MMAs; both CTAs' TMA loads
// signal the leader's `full` barrier; tcgen05.commit (multicast) releases the stage / publishes the
// accumulator in both CTAs; both CTAs' epilogue warps hand the accumulator back on the leader's barrier.
//
// Tile schedule: pair-tiles (256 rows x 256 geocells) are ordered geocell-tile-fastest within a 256-row
// block and every pair owns a CONTIGUOUS range of that order, i.e. it sweeps many geocell tiles of the
// same rows back to back.
// The per-row state therefore stays in registers across tiles ("run") and is flushed once per
// (CTA, row block): after the first tile the running 5th-best logit rejects almost every 32-column
// chunk by its maximum alone (which the softmax needs anyway), so the top-k costs ~1 compare per
// chunk.  The few partials of a row are combined into top-k probabilities and indices, argmax, predicted
// centroid and the row log-sum-exp by whichever CTA flushes the LAST partial of a 128-row block (tickets in the
// workspace): no second launch, and only that block's short merge trails the GEMM.
// num_candidates > 8 (torch.topk has no limit, super_guessr.py:365): the GEMM is run again per further 8
// ranks with a per-row "ceiling" (the previous pass's last entry): only logits ordered after it compete, so
// every pass is exact on the same fp32 accumulators.
//
// Layout: x (M=B, K=D) bf16 row-major; W (N=C, K=D) bf16 row-major (both K-major operands, 128 B
// swizzled TMA boxes of 64 K-elements); logits (B, ldc) bf16, ldc >= C padded to a multiple of 64.
#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kBM = 128;        // rows of x per tile (UMMA M)
constexpr int kBN = 256;        // geocells per tile   (UMMA N)
constexpr int kBK = 64;         // K elements per stage (128 B of bf16 = one swizzle span)
// Operand ring: 32 KB per stage and CTA.  The MMA warp was waiting on operands about half of the time with four
// stages (profiles/r01c): the ring is latency bound, so it gets all the shared memory the epilogue can spare --
// five stages next to 32 KB of logits staging in training, six in serving (no staging).
template <bool WRITE_LOGITS> struct FwdCfg { static constexpr int kStages = WRITE_LOGITS ? 5 : 6; };
constexpr int kEpiWarps = 16;
constexpr int kColGroups = kEpiWarps / 4;       // column groups of a tile (one warp per TMEM lane quadrant each)
constexpr int kColsPerEpiWarp = kBN / kColGroups;
constexpr int kFwdThreads = 64 + 32 * kEpiWarps;  // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr uint32_t kStageBytesA = kBM * kBK * 2;
constexpr uint32_t kStageBytesB = (kBN / 2) * kBK * 2;  // this CTA's half of the W tile
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kThrSlots = 8;    // shared top-k thresholds are kept per run, in a ring of run slots
constexpr int kKeyMin = static_cast<int>(0x80000000u);

template <bool WRITE_LOGITS>
struct FwdSmem {
  static constexpr int kStages = FwdCfg<WRITE_LOGITS>::kStages;
  uint8_t a[kStages][kStageBytesA];
  uint8_t b[kStages][kStageBytesB];
  // per epilogue warp: 32 rows x 32 bf16 logits (one 32-column chunk), 64-byte swizzled (TMA store box)
  uint8_t out[WRITE_LOGITS ? kEpiWarps : 1][32 * 64];
  int thr[kThrSlots][kBM];           // per row: best known lower bound of the run's k-th largest logit (ordered key)
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t is_last;
};

// Static schedule shared by the kernel, its merge step and the host.  Units are CTA PAIRS and
// pair-tiles: num_m = 256-row blocks, grid = number of pairs (clusters); CTA 2c + r of pair c works on
// row block 2 * mb + r.  Pair c owns the contiguous tiles [first[c], first[c + 1]).
//
// The ranges are balanced by COST, not by tile count: the first tiles of a run over a new row block are slow in
// the epilogue (the row's top-k starts empty, so whole chunks go through the insertion path until the thresholds
// have risen; measured with tools/head_fwd_timeline.py: a pair whose range crosses a row-block boundary finishes
// 8-10 us = ~1.4 tiles later than its neighbours), and the slowest pair ends the kernel.
constexpr int kMaxPairs = 96;
constexpr float kRunPenaltyTiles = 1.4f;  // extra cost of every run after a pair's first, in tiles
struct FwdSched {
  int num_m, num_n, tiles, grid, runs;  // runs = max row blocks a pair can touch
  int first[kMaxPairs + 1];
  __host__ __device__ int start(int c) const { return first[c]; }
  __host__ __device__ int owner(int t) const {  // (cold paths only)
    int lo = 0, hi = grid - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (first[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
  }
};
static FwdSched make_sched_uncached(int M, int N, int sms);
// (the bisection walks the tile list a few dozen times: memoised, the launch path must stay cheap)
static FwdSched make_sched(int M, int N, int sms) {
  struct Entry { int M, N, sms; FwdSched s; };
  static Entry cache[8];
  static int used = 0, next = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (cache[i].M == M && cache[i].N == N && cache[i].sms == sms) return cache[i].s;
  Entry& e = cache[next];
  next = (next + 1) % 8;
  if (used < 8) ++used;
  e.M = M; e.N = N; e.sms = sms;
  e.s = make_sched_uncached(M, N, sms);
  return e.s;
}
static FwdSched make_sched_uncached(int M, int N, int sms) {
  FwdSched s;
  s.num_m = ceil_div(M, 2 * kBM);
  s.num_n = ceil_div(N, kBN);
  s.tiles = s.num_m * s.num_n;
  s.grid = std::max(1, std::min(std::min(s.tiles, sms / 2), kMaxPairs));
  // greedy fill with a cost limit T; the smallest T that needs <= grid pairs (bisection)
  auto fill = [&](float T, int* first) {
    int c = 0, t = 0;
    while (t < s.tiles) {
      if (first) first[c] = t;
      float cost = 0.f;
      int n = 0;
      while (t < s.tiles) {
        const float add = 1.0f + ((n > 0 && t % s.num_n == 0) ? kRunPenaltyTiles : 0.f);
        if (n > 0 && cost + add > T) break;
        cost += add;
        ++t;
        ++n;
      }
      ++c;
      if (c > s.grid && !first) return c;
      if (c >= s.grid && t < s.tiles && first) {  // (cannot happen for a feasible T; keep the table consistent)
        t = s.tiles;
      }
    }
    if (first) {
      for (int i = c; i <= s.grid; ++i) first[i] = s.tiles;
    }
    return c;
  };
  float lo = static_cast<float>(s.tiles) / s.grid, hi = static_cast<float>(s.tiles) * (1.0f + kRunPenaltyTiles) + 1.0f;
  for (int it = 0; it < 40; ++it) {
    const float mid = 0.5f * (lo + hi);
    if (fill(mid, nullptr) <= s.grid) hi = mid; else lo = mid;
  }
  fill(hi, s.first);
  s.runs = 1;
  for (int c = 0; c < s.grid; ++c) {
    if (s.first[c + 1] > s.first[c])
      s.runs = std::max(s.runs, (s.first[c + 1] - 1) / s.num_n - s.first[c] / s.num_n + 1);
  }
  return s;
}

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

// v[i] for a run-time i without spilling v[] to local memory: 5-level select tree (31 FSEL).
__device__ __forceinline__ float select32(const float (&v)[32], int i) {
  float a[16], b[8], c[4], d[2];
  const bool b0 = i & 1, b1 = i & 2, b2 = i & 4, b3 = i & 8, b4 = i & 16;
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = b0 ? v[2 * j + 1] : v[2 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = b1 ? a[2 * j + 1] : a[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = b2 ? b[2 * j + 1] : b[2 * j];
#pragma unroll
  for (int j = 0; j < 2; ++j) d[j] = b3 ? c[2 * j + 1] : c[2 * j];
  return b4 ? d[1] : d[0];
}

// float <-> int key with the same ordering (involution)
__device__ __forceinline__ int ordered_key(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_to_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// ------------------------------------------------------------------ merge of a 128-row block's partials
//   topk_val = softmax probabilities exp(l - max) / sum   (super_guessr.py:355,365)
//   pred_cell = argmax (:358), pred_llh = centroids[pred_cell] (:359-361), lse = max + log(sum)
// The partial arrays are [partial][value][row in the 128-row tile] (lane = row: every load is one full
// 128-byte line).  Ties keep the lower geocell index first.
struct FwdOut {
  const float* centroids;
  float* topk_val;
  long long* topk_idx;
  long long* pred_cell;
  float* pred_llh;
  float* lse;
  int k;      // row pitch of topk_val / topk_idx
  int k_off;  // first rank this pass writes
  int k_cnt;  // ranks this pass writes (<= KTOP)
  float* ceil_v;        // (B) last (value, index) of this pass = ceiling of the next one; null when this is the last
  int* ceil_i;
  const float* lse_in;  // ceiling passes: the row log-sum-exp of pass 0
};

template <int KTOP>
__device__ __forceinline__ void merge_insert(float (&tv)[KTOP], int (&ti)[KTOP], float v, int id) {
  if (v > tv[KTOP - 1] || (v == tv[KTOP - 1] && v > -INFINITY && id < ti[KTOP - 1])) {
    tv[KTOP - 1] = v;
    ti[KTOP - 1] = id;
#pragma unroll
    for (int q = KTOP - 1; q > 0; --q) {
      if (tv[q] > tv[q - 1] || (tv[q] == tv[q - 1] && ti[q] < ti[q - 1])) {
        float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
        int iv = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = iv;
      }
    }
  }
}
__device__ __forceinline__ void lse_merge(float& lmax, float& lsum, float m, float s) {
  if (m > -INFINITY) {
    if (m > lmax) { lsum = lsum * expf(lmax - m) + s; lmax = m; }
    else lsum += s * expf(m - lmax);
  }
}
constexpr uint32_t kEpiBarrier = 1;  // named barrier of the 16 epilogue warps

// Optional per-CTA timeline (tools/head_fwd_timeline.py): 64 globaltimer stamps per CTA when a buffer is set
// (16 named points; then for each of the CTA's first 16 tiles: 16+i accumulator seen complete by epilogue warp 0,
// 32+i that warp ready for it -- the gap is its wait --, 48+i the tile's last MMA issued).
__device__ __forceinline__ void stamp(long long* tl, int slot) {
  if (tl != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[static_cast<size_t>(blockIdx.x) * 64 + slot] = t;
  }
}

// ---- branch-free merging of sorted top-k lists (the final merge of a row block) ------------------------------
// The merge runs on 512 threads for 128 rows x ~24 lists: with insertion code every list costs ~200 divergent
// instructions (any lane inserting stalls the warp) and the merge took 10-17 us at the very end of the kernel
// (tools/head_fwd_timeline.py).  Here a candidate is one 64-bit key -- ordered logit bits above, inverted geocell
// index below, so that a plain unsigned compare orders by value and breaks ties towards the LOWER index -- and
// merging two descending lists is max(A[i], B[K-1-i]) (exactly the K largest of the union) followed by a fixed
// sorting network: ~100 straight-line instructions per list, no divergence.
typedef unsigned long long key_t64;
__device__ __forceinline__ key_t64 make_key(float v, int idx) {
  return (static_cast<key_t64>(static_cast<uint32_t>(ordered_key(v)) ^ 0x80000000u) << 32) | static_cast<uint32_t>(~idx);
}
__device__ __forceinline__ float key_value(key_t64 k) { return key_to_float(static_cast<int>(static_cast<uint32_t>(k >> 32) ^ 0x80000000u)); }
__device__ __forceinline__ int key_index(key_t64 k) { return static_cast<int>(~static_cast<uint32_t>(k)); }
__device__ __forceinline__ void key_cx(key_t64& hi, key_t64& lo) {  // compare-exchange: hi >= lo afterwards
  const key_t64 a = hi, b = lo;
  const bool sw = a < b;
  hi = sw ? b : a;
  lo = sw ? a : b;
}
template <int K> __device__ __forceinline__ void sort_keys_desc(key_t64 (&c)[K]);
template <> __device__ __forceinline__ void sort_keys_desc<5>(key_t64 (&c)[5]) {  // optimal 9-comparator network
  key_cx(c[0], c[1]); key_cx(c[3], c[4]); key_cx(c[2], c[4]); key_cx(c[2], c[3]); key_cx(c[1], c[4]);
  key_cx(c[0], c[3]); key_cx(c[0], c[2]); key_cx(c[1], c[3]); key_cx(c[1], c[2]);
}
template <> __device__ __forceinline__ void sort_keys_desc<8>(key_t64 (&c)[8]) {  // bitonic merge (the input is bitonic)
#pragma unroll
  for (int d = 4; d >= 1; d >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((i & d) == 0) key_cx(c[i], c[i + d]);
  }
}
// a <- the K largest of a U b, descending (both inputs descending)
template <int K>
__device__ __forceinline__ void merge_keys(key_t64 (&a)[K], const key_t64 (&b)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = a[i] > b[K - 1 - i] ? a[i] : b[K - 1 - i];
  sort_keys_desc<K>(a);
}

// (max, sum-exp) pairs combined with the hardware exponential (2 ulp: far below the bf16 logits' own rounding)
__device__ __forceinline__ void lse_merge_fast(float& lmax, float& lsum, float m, float s) {
  const float nm = fmaxf(lmax, m);
  const float a = nm > -INFINITY ? ex2_approx((lmax - nm) * kLog2e) : 0.f;
  const float b = nm > -INFINITY ? ex2_approx((m - nm) * kLog2e) : 0.f;
  lsum = lsum * a + s * b;
  lmax = nm;
}

// Called by all 512 epilogue threads of the CTA that flushed the last partial of row block (mb, crank).  Four
// threads share a row: thread `sub` folds the partials of column group `sub` of every contributing CTA -- the next
// contributor's loads are in flight while the current one is merged -- then two shuffle rounds combine the four
// threads (no shared memory, no barrier) and sub 0 writes the row's results.  Fixed order: deterministic.
template <int KTOP, bool HAS_CEIL>
__device__ __forceinline__ void merge_block(const float* __restrict__ pmax, const float* __restrict__ psum,
                                            const key_t64* __restrict__ pkeys, const FwdSched& sc, int mb, int crank,
                                            int M, const FwdOut& o, long long* timeline) {
  const int te = static_cast<int>(threadIdx.x) - 64;  // 0..511
  const int rit = te >> 2, sub = te & 3;
  const int row = (2 * mb + crank) * kBM + rit;
  const int c_lo = sc.owner(mb * sc.num_n), c_hi = sc.owner((mb + 1) * sc.num_n - 1);

  float lmax = -INFINITY, lsum = 0.f;
  key_t64 best[KTOP];
#pragma unroll
  for (int j = 0; j < KTOP; ++j) best[j] = make_key(-INFINITY, 0x7fffffff);

  float m = -INFINITY, s = 0.f;
  key_t64 nxt[KTOP];
  auto load = [&](int c) {  // L2 loads: the partials were written by other SMs during this launch
    const int run = mb - sc.start(c) / sc.num_n;
    const size_t p = (static_cast<size_t>(2 * c + crank) * sc.runs + run) * kColGroups + sub;
    if (!HAS_CEIL) {
      m = __ldcg(pmax + p * kBM + rit);
      s = __ldcg(psum + p * kBM + rit);
    }
#pragma unroll
    for (int j = 0; j < KTOP; ++j) nxt[j] = __ldcg(pkeys + (p * KTOP + j) * kBM + rit);
  };
  if (te == 0) stamp(timeline, 12);
  load(c_lo);
#pragma unroll 1
  for (int c = c_lo; c <= c_hi; ++c) {
    const float cm = m, cs = s;
    key_t64 cur[KTOP];
#pragma unroll
    for (int j = 0; j < KTOP; ++j) cur[j] = nxt[j];
    if (c < c_hi) load(c + 1);
    if (!HAS_CEIL) lse_merge_fast(lmax, lsum, cm, cs);
    merge_keys<KTOP>(best, cur);
    if (te == 0 && c == c_lo) stamp(timeline, 13);
  }
  if (te == 0) stamp(timeline, 14);
#pragma unroll 1
  for (int x = 1; x <= 2; x <<= 1) {  // the row's four threads are adjacent lanes
    const float om = __shfl_xor_sync(0xffffffffu, lmax, x), os = __shfl_xor_sync(0xffffffffu, lsum, x);
    key_t64 other[KTOP];
#pragma unroll
    for (int j = 0; j < KTOP; ++j) other[j] = __shfl_xor_sync(0xffffffffu, best[j], x);
    if (!HAS_CEIL) lse_merge_fast(lmax, lsum, om, os);
    merge_keys<KTOP>(best, other);
  }
  if (sub != 0 || row >= M) return;
  float lse_row, inv = 1.f;
  if (!HAS_CEIL) {
    inv = 1.f / lsum;
    lse_row = lmax + logf(lsum);
  } else {
    lse_row = o.lse_in[row];
    lmax = lse_row;
  }
#pragma unroll
  for (int j = 0; j < KTOP; ++j) {
    if (j < o.k_cnt) {
      o.topk_val[static_cast<size_t>(row) * o.k + o.k_off + j] = expf(key_value(best[j]) - lmax) * inv;
      o.topk_idx[static_cast<size_t>(row) * o.k + o.k_off + j] = key_index(best[j]);
    }
  }
  if (o.ceil_v) {
    o.ceil_v[row] = key_value(best[KTOP - 1]);
    o.ceil_i[row] = key_index(best[KTOP - 1]);
  }
  if (HAS_CEIL) return;
  const int best0 = key_index(best[0]);
  if (o.pred_cell) o.pred_cell[row] = best0;
  if (o.pred_llh) {
    o.pred_llh[2 * row + 0] = o.centroids[2 * best0 + 0];
    o.pred_llh[2 * row + 1] = o.centroids[2 * best0 + 1];
  }
  if (o.lse) o.lse[row] = lse_row;
}

template <int KTOP, bool WRITE_LOGITS, bool HAS_CEIL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ bias_pad,
                float* __restrict__ pmax, float* __restrict__ psum, key_t64* __restrict__ pkeys,
                unsigned int* __restrict__ tickets, const float* __restrict__ ceil_in_v,
                const int* __restrict__ ceil_in_i, int M, int N, int K, FwdSched sc, FwdOut out,
                long long* __restrict__ timeline) {
  extern __shared__ uint8_t smem_raw[];
  using Smem = FwdSmem<WRITE_LOGITS>;
  constexpr int kStages = Smem::kStages;
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k = (K + kBK - 1) / kBK;
  const int crank = static_cast<int>(cluster_ctarank());  // == blockIdx.x & 1
  const int pair = blockIdx.x >> 1;
  const int t_begin = sc.start(pair), t_end = sc.start(pair + 1);

  if (threadIdx.x == 0) {
    stamp(timeline, 0);
    if (timeline != nullptr) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      timeline[static_cast<size_t>(blockIdx.x) * 64 + 15] = smid;
    }
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    if (WRITE_LOGITS) tma_prefetch_desc(&tm_out);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);   // leader's: its own arrive.expect_tx, bytes from both CTAs' loads
      mbar_init(&sm.empty[s], 1);  // one multicast commit per round
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.acc_full[a], 1);
      mbar_init(&sm.acc_empty[a], 2 * kEpiWarps);  // leader's: one arrive per epilogue warp of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // collective over the pair: one warp in each CTA
    tmem_alloc_pair(&sm.tmem_base, 512);
    tmem_relinquish_pair();
  }
  for (int i = threadIdx.x; i < kThrSlots * kBM; i += kFwdThreads) (&sm.thr[0][0])[i] = kKeyMin;
  tc_fence_before();
  cluster_sync();  // the partner's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (threadIdx.x == 0) stamp(timeline, 1);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t leader_full = mapa_u32(smem_u32(&sm.full[0]), 0);
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int m0 = (2 * (t / sc.num_n) + crank) * kBM, n0 = (t % sc.num_n) * kBN + crank * (kBN / 2);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&sm.empty[s], ph ^ 1);
        if (elect_one()) {
          if (crank == 0) mbar_arrive_expect_tx(&sm.full[s], 2 * (kStageBytesA + kStageBytesB));
          tma_load_2d_pair(sm.a[s], &tm_x, leader_full + 8 * s, kb * kBK, m0);  // my 128 rows of x
          tma_load_2d_pair(sm.b[s], &tm_w, leader_full + 8 * s, kb * kBK, n0);  // my 128 of the 256 geocells
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only; warp-uniform loop, one elected lane issues) =======
    if (crank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, kBN, 0, 0);
      const uint64_t da_base = umma_desc_sw128(smem_u32(sm.a[0]), 16, 1024);
      const uint64_t db_base = umma_desc_sw128(smem_u32(sm.b[0]), 16, 1024);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&sm.acc_empty[acc], acc_ph ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            // descriptors differ only in the start-address field: stage s, then 32 bytes per 16-wide k step
            const uint64_t da = da_base + static_cast<uint64_t>(s * (kStageBytesA >> 4));
            const uint64_t db = db_base + static_cast<uint64_t>(s * (kStageBytesB >> 4));
            umma_pair_f16(d_tmem, da, db, idesc, kb != 0);
#pragma unroll
            for (int k = 1; k < kBK / 16; ++k) umma_pair_f16_acc(d_tmem, da + 2 * k, db + 2 * k, idesc);
            umma_pair_commit(&sm.empty[s], 0x3);  // slot reusable in both CTAs once these MMAs have read it
            if (kb == num_k - 1) umma_pair_commit(&sm.acc_full[acc], 0x3);  // accumulator complete (both halves)
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (it == 0 && lane == 0) stamp(timeline, 2);  // first tile's MMAs issued
        if (it < 16 && lane == 0) stamp(timeline, 48 + it);
      }
      if (lane == 0) stamp(timeline, 3);  // last MMA issued
    }
  } else {
    // ===================== epilogue warps (512 threads: row = TMEM lane, 64 columns each) ==========
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;   // which 64 columns of the 256-wide tile
    const int row_in_tile = quad * 32 + lane;
    uint8_t* const my_out = sm.out[WRITE_LOGITS ? warp - 2 : 0];
    const uint32_t leader_acc_empty = mapa_u32(smem_u32(&sm.acc_empty[0]), 0);
    const uint32_t my_out_row = smem_u32(my_out) + lane * 64;

    float run_max = -INFINITY, run_sum = 0.f;
    float tv[KTOP];
    int ti[KTOP];
#pragma unroll
    for (int j = 0; j < KTOP; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
    const int mb_first = t_begin / sc.num_n;
    int run_id = 0;
    sm.thr[4][row_in_tile] = kKeyMin;  // slot of run 4 (see the flush below); slots 0..3 start clean
    // ceiling passes: only logits ordered after (ceil_v, ceil_i) -- lower value, or equal value and higher index
    float ceil_v = INFINITY;
    int ceil_i = -1;
    int ceil_mb = -1;

    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int mb = t / sc.num_n, nb = t % sc.num_n;
      const int m0 = (2 * mb + crank) * kBM, n0 = nb * kBN + cg * kColsPerEpiWarp;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const uint32_t thr_slot = smem_u32(&sm.thr[run_id & (kThrSlots - 1)][row_in_tile]);
      if (HAS_CEIL && mb != ceil_mb) {
        const int row = m0 + row_in_tile;
        ceil_v = row < M ? __ldg(ceil_in_v + row) : -INFINITY;
        ceil_i = row < M ? __ldg(ceil_in_i + row) : 0x7fffffff;
        ceil_mb = mb;
      }
      if (threadIdx.x == 64 && it < 16) stamp(timeline, 32 + it);
      mbar_wait(&sm.acc_full[acc], acc_ph);
      tc_fence_after();
      if (threadIdx.x == 64) {
        if (it == 0) stamp(timeline, 4);            // first accumulator complete
        if (it < 16) stamp(timeline, 16 + it);      // every tile's accumulator
        if (t == t_end - 1) stamp(timeline, 5);     // last accumulator complete
      }
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBN + cg * kColsPerEpiWarp;

#pragma unroll 1
      for (int c = 0; c < kColsPerEpiWarp / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        const int col0 = n0 + c * 32;
        float4 bq[8];
        const float4* bp = reinterpret_cast<const float4*>(bias_pad + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) bq[q] = __ldg(bp + q);
        tmem_ld_wait();
        if (c == kColsPerEpiWarp / 32 - 1) {
          // The tile's last columns are in registers: hand the TMEM accumulator back to the leader's MMA warp NOW,
          // before this chunk's soft-max / top-k work.  At the start of a run every logit enters the top-k lists
          // (~3x the steady-state instructions per chunk); held until the end of the tile, the accumulator kept the
          // MMA warp waiting ~6 us at the third tile of every CTA and again after every row-block boundary
          // (tools/head_fwd_timeline.py, rdy/acc/iss stamps).
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_acc_empty + 8 * acc);
        }
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq[q].x;
          v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq[q].y;
          v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq[q].z;
          v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq[q].w;
        }
        if (WRITE_LOGITS) {
          // One 2 KB staging buffer per warp: the previous chunk's TMA store was issued a whole chunk of
          // soft-max / top-k work ago, so waiting for it to have READ the buffer costs nothing.
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          // 16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3): the 64-byte TMA swizzle, and conflict-free
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t dst = my_out_row + ((q ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                         "r"(pack_bf16x2(v[8 * q + 0], v[8 * q + 1])), "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])),
                         "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7]))
                         : "memory");
          }
          fence_proxy_async_smem();  // staging writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_out, my_out, col0, m0 + quad * 32);  // rows >= M / columns >= ldc are clipped
            tma_store_commit();
          }
        }
        if (col0 + 32 > N) {  // tail tile: geocells >= C do not exist
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i >= N) v[i] = -INFINITY;
        }
        if (HAS_CEIL) {  // the softmax statistics are pass 0's; entries at or before the ceiling are out
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = (v[i] < ceil_v || (v[i] == ceil_v && col0 + i > ceil_i)) ? v[i] : -INFINITY;
        }
        float cmax = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) cmax = fmaxf(cmax, v[i]);
        if (!HAS_CEIL) {
          if (cmax > run_max) {
            run_sum *= ex2_approx((run_max - cmax) * kLog2e);  // run_max = -inf -> 0 * 0
            run_max = cmax;
          }
          if (run_max > -INFINITY) {
            const float ms = run_max * kLog2e;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              s0 += ex2_approx(fmaf(v[i + 0], kLog2e, -ms));
              s1 += ex2_approx(fmaf(v[i + 1], kLog2e, -ms));
              s2 += ex2_approx(fmaf(v[i + 2], kLog2e, -ms));
              s3 += ex2_approx(fmaf(v[i + 3], kLog2e, -ms));
            }
            run_sum += (s0 + s1) + (s2 + s3);
          }
        }
        // Top-k.  A logit can only matter if it beats a lower bound of the row's k-th best in this run:
        // this warp's own k-th best, or the one any of the three other warps of the row has published.
        // (Strict '>': a logit exactly equal to the bound may be dropped -- below the score-gap tolerance.)
        const float thr = fmaxf(tv[KTOP - 1], key_to_float(ld_volatile_s32_shared(thr_slot)));
        if (__any_sync(0xffffffffu, cmax > thr)) {
          uint32_t mask = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) mask |= (v[i] > thr) ? (1u << i) : 0u;
          const bool mine = mask != 0u;
          // Lanes are different rows, so WHICH columns qualify differs per lane.  Few per lane (steady
          // state): pop each lane's own bits, v[i] by a select tree.  Many (start of a run): walk the
          // union's columns in a warp-uniform, compile-time-indexed loop instead.
          if (__reduce_max_sync(0xffffffffu, __popc(mask)) >= 8u) {
            const uint32_t any_mask = __reduce_or_sync(0xffffffffu, mask);
#pragma unroll
            for (int i = 0; i < 32; ++i) {  // ascending column order: on ties the lower geocell index stays first
              if ((any_mask >> i) & 1u) {
                if ((mask >> i) & 1u) topk_insert<KTOP>(tv, ti, v[i], col0 + i);
              }
            }
          } else {
            while (mask) {
              const int i = __ffs(mask) - 1;
              mask &= mask - 1;
              topk_insert<KTOP>(tv, ti, select32(v, i), col0 + i);
            }
          }
          if (mine && tv[KTOP - 1] > -INFINITY) red_max_s32_shared(thr_slot, ordered_key(tv[KTOP - 1]));
        }
      }
      // end of this CTA's run over row block mb: flush the row state (one partial per epilogue warp)
      if (nb == sc.num_n - 1 || t == t_end - 1) {
        if (threadIdx.x == 64 && t == t_end - 1) stamp(timeline, 6);  // last tile's chunks done
        {
          const size_t p = (static_cast<size_t>(blockIdx.x) * sc.runs + (mb - mb_first)) * kColGroups + cg;
          if (!HAS_CEIL) {
            __stcg(&pmax[p * kBM + row_in_tile], run_max);
            __stcg(&psum[p * kBM + row_in_tile], run_sum);
          }
#pragma unroll
          for (int j = 0; j < KTOP; ++j) __stcg(&pkeys[(p * KTOP + j) * kBM + row_in_tile], make_key(tv[j], ti[j]));
        }
        if (threadIdx.x == 64 && t == t_end - 1) stamp(timeline, 7);  // partials stored
        run_max = -INFINITY;
        run_sum = 0.f;
#pragma unroll
        for (int j = 0; j < KTOP; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
        // Threshold slots: the warps of a CTA are never more than two tiles (hence two runs) apart --
        // the accumulators are double buffered -- so while this warp is in run r the others use slots
        // r-2 .. r+2 (mod 8); slot r+4 is free and is cleaned here, long before anyone enters run r+4.
        ++run_id;
        sm.thr[(run_id + 4) & (kThrSlots - 1)][row_in_tile] = kKeyMin;

        // Ticket of row block (mb, crank): one per contributing CTA.  Whoever draws the last one finds every
        // partial of these 128 rows in L2 and merges them (the other CTAs go on with their next run).
        // all 16 epilogue warps have flushed; ONE fence then publishes their stores (cumulative through the barrier)
        // before the ticket is drawn -- 512 membars per flush stalled the SM's memory pipeline for nothing
        named_bar_sync(kEpiBarrier, 32 * kEpiWarps);
        if (threadIdx.x == 64) {
          __threadfence();
          const int contrib = sc.owner((mb + 1) * sc.num_n - 1) - sc.owner(mb * sc.num_n) + 1;
          const unsigned int old = atomicAdd(&tickets[2 * mb + crank], 1u);
          const bool last = old + 1u == static_cast<unsigned int>(contrib);
          if (last) {
            tickets[2 * mb + crank] = 0u;  // left zeroed for the next launch
            __threadfence();               // (acquire side, once: the merge below reads the partials with ld.cg)
          }
          sm.is_last = last ? 1u : 0u;
        }
        named_bar_sync(kEpiBarrier, 32 * kEpiWarps);
        if (threadIdx.x == 64 && t == t_end - 1) stamp(timeline, 8);  // flushed, ticket drawn
        if (sm.is_last != 0u) {
          merge_block<KTOP, HAS_CEIL>(pmax, psum, pkeys, sc, mb, crank, M, out, timeline);
          if (threadIdx.x == 64) stamp(timeline, 9);  // merged a row block (last one wins)
        }
      }
    }
    if (WRITE_LOGITS) {
      if (lane == 0) tma_store_wait_all<0>();  // shared memory must outlive the last bulk store
    }
    if (threadIdx.x == 64) stamp(timeline, 10);  // epilogue done (bulk stores drained)
  }

  tc_fence_before();
  cluster_sync();  // the partner may still be signalling this CTA's barriers / the leader reading its operands
  if (threadIdx.x == 0) stamp(timeline, 11);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <bool WRITE_LOGITS>
static size_t fwd_smem_bytes() { return sizeof(FwdSmem<WRITE_LOGITS>) + 1024; }

static size_t fwd_partials(const FwdSched& sc) { return static_cast<size_t>(2 * sc.grid) * sc.runs * kColGroups; }

template <typename Kern, typename... Args>
static cudaError_t launch_pairs(Kern kern, int pairs, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, 1, 1);  // cluster shape (2,1,1) is compiled into the kernel
  cfg.blockDim = dim3(kFwdThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// workspace: [partials: pmax | psum | pkeys (64-bit candidates)][ceil_v (B) | ceil_i (B) | lse (B): passes beyond the first]
struct FwdWs {
  float *pmax, *psum;
  key_t64* pkeys;
  float* ceil_v[2];
  int* ceil_i[2];
  float* lse;
  size_t bytes;
};
static FwdWs carve_fwd_ws(void* base, const FwdSched& sc, int B, int ktop, bool multipass) {
  FwdWs w;
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t nbytes) {
    uint8_t* r = p;
    p += (nbytes + 255) & ~size_t(255);
    return r;
  };
  const size_t np = fwd_partials(sc) * kBM;
  w.pmax = reinterpret_cast<float*>(take(np * sizeof(float)));
  w.psum = reinterpret_cast<float*>(take(np * sizeof(float)));
  w.pkeys = reinterpret_cast<key_t64*>(take(np * ktop * sizeof(key_t64)));
  for (int i = 0; i < 2; ++i) {
    w.ceil_v[i] = reinterpret_cast<float*>(take(multipass ? sizeof(float) * B : 0));
    w.ceil_i[i] = reinterpret_cast<int*>(take(multipass ? sizeof(int) * B : 0));
  }
  w.lse = reinterpret_cast<float*>(take(multipass ? sizeof(float) * B : 0));
  w.bytes = static_cast<size_t>(p - static_cast<uint8_t*>(base));
  return w;
}

static long long* g_fwd_timeline = nullptr;  // gg_debug_head_fwd_timeline

template <int KTOP, bool WRITE_LOGITS, bool HAS_CEIL>
static int launch_one(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const CUtensorMap& tm_out, const float* bias_pad,
                      const FwdWs& w, unsigned int* tickets, const float* ceil_in_v, const int* ceil_in_i, int B, int C,
                      int D, const FwdSched& sc, const FwdOut& out, cudaStream_t stream) {
  auto kern = head_fwd_kernel<KTOP, WRITE_LOGITS, HAS_CEIL>;
  const size_t smem = fwd_smem_bytes<WRITE_LOGITS>();
  if (int e = set_max_dynamic_smem_once(kern, smem)) return e;
  GG_CUDA(launch_pairs(kern, sc.grid, smem, stream, tm_x, tm_w, tm_out, bias_pad, w.pmax, w.psum, w.pkeys, tickets,
                       ceil_in_v, ceil_in_i, B, C, D, sc, out, g_fwd_timeline));
  GG_LAUNCH_CHECK();
  return GG_OK;
}

}  // namespace gg

using namespace gg;

static int fwd_ktop(int k) { return k <= 5 ? 5 : 8; }

extern "C" size_t gg_head_fwd_workspace_bytes(int B, int C, int k) {
  const FwdSched sc = make_sched(B, C, device_sm_count());
  return carve_fwd_ws(nullptr, sc, B, fwd_ktop(k), k > 8).bytes;
}
extern "C" size_t gg_head_fwd_ticket_bytes(int B) { return sizeof(unsigned int) * 2 * ceil_div(B, 2 * kBM); }

extern "C" void gg_debug_head_fwd_timeline(long long* device_buf) { g_fwd_timeline = device_buf; }

extern "C" int gg_head_logits_ld(int C) { return ceil_div(C, 256) * 256; }
extern "C" int gg_head_bias_pad(int C) { return ceil_div(C, kBN) * kBN; }

extern "C" int gg_head_fwd(const void* x_bf16, const void* w_bf16, const float* bias_pad, int B, int C, int D,
                           void* logits_bf16, int ldc, int k, void* workspace, void* tickets, const float* centroids,
                           float* topk_val, long long* topk_idx, long long* pred_cell, float* pred_llh, float* lse,
                           gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && D > 0, GG_ERR_ARG, "gg_head_fwd: empty problem B=%d C=%d D=%d", B, C, D);
  GG_CHECK(D % 8 == 0, GG_ERR_ARG, "gg_head_fwd: embed dim D=%d must be a multiple of 8 (16-byte TMA pitch)", D);
  GG_CHECK(k >= 1 && k <= C, GG_ERR_ARG, "gg_head_fwd: num_candidates k=%d must be in [1, C=%d]", k, C);
  GG_CHECK(!logits_bf16 || (ldc >= C && ldc % 8 == 0), GG_ERR_ARG, "gg_head_fwd: ldc=%d must be >= C and a multiple of 8", ldc);
  GG_CHECK(x_bf16 && w_bf16 && bias_pad && workspace && tickets && topk_val && topk_idx, GG_ERR_ARG,
           "gg_head_fwd: null pointer");
  GG_CHECK(!pred_llh || centroids, GG_ERR_ARG, "gg_head_fwd: pred_llh needs the centroid table");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_x, tm_w;
  int rc = make_tmap_bf16_2d(&tm_x, x_bf16, D, B, static_cast<uint64_t>(D) * 2, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_w, w_bf16, D, C, static_cast<uint64_t>(D) * 2, kBK, kBN / 2);  // one CTA's half
  if (rc) return rc;
  CUtensorMap tm_out = tm_x;  // placeholder in serving (never dereferenced)
  if (logits_bf16) {
    rc = make_tmap_bf16_2d_sw64(&tm_out, logits_bf16, ldc, B, static_cast<uint64_t>(ldc) * 2, 32, 32);
    if (rc) return rc;
  }
  const FwdSched sc = make_sched(B, C, device_sm_count());
  const int ktop = fwd_ktop(k);
  const bool multipass = k > 8;
  const FwdWs w = carve_fwd_ws(workspace, sc, B, ktop, multipass);
  unsigned int* tk = static_cast<unsigned int*>(tickets);
  FwdOut o;
  o.centroids = centroids;
  o.topk_val = topk_val;
  o.topk_idx = topk_idx;
  o.pred_cell = pred_cell;
  o.pred_llh = pred_llh;
  o.lse = multipass && !lse ? w.lse : lse;
  o.k = k;
  o.k_off = 0;
  o.k_cnt = k < ktop ? k : ktop;
  o.ceil_v = multipass ? w.ceil_v[0] : nullptr;
  o.ceil_i = multipass ? w.ceil_i[0] : nullptr;
  o.lse_in = nullptr;
  if (ktop == 5) {
    rc = logits_bf16 ? launch_one<5, true, false>(tm_x, tm_w, tm_out, bias_pad, w, tk, nullptr, nullptr, B, C, D, sc, o, s)
                     : launch_one<5, false, false>(tm_x, tm_w, tm_out, bias_pad, w, tk, nullptr, nullptr, B, C, D, sc, o, s);
    return rc;
  }
  rc = logits_bf16 ? launch_one<8, true, false>(tm_x, tm_w, tm_out, bias_pad, w, tk, nullptr, nullptr, B, C, D, sc, o, s)
                   : launch_one<8, false, false>(tm_x, tm_w, tm_out, bias_pad, w, tk, nullptr, nullptr, B, C, D, sc, o, s);
  if (rc) return rc;
  // ranks 8, 9, ...: the GEMM again per 8 ranks, each pass below the previous pass's last entry
  const float* lse0 = o.lse;
  for (int off = 8, pass = 1; off < k; off += 8, ++pass) {
    FwdOut p = o;
    p.k_off = off;
    p.k_cnt = k - off < 8 ? k - off : 8;
    const int in = (pass - 1) & 1, outb = pass & 1;
    p.ceil_v = off + 8 < k ? w.ceil_v[outb] : nullptr;
    p.ceil_i = off + 8 < k ? w.ceil_i[outb] : nullptr;
    p.lse_in = lse0;
    rc = launch_one<8, false, true>(tm_x, tm_w, tm_out, bias_pad, w, tk, w.ceil_v[in], w.ceil_i[in], B, C, D, sc, p, s);
    if (rc) return rc;
  }
  return GG_OK;
}
