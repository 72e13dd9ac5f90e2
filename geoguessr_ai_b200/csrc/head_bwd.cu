// Geocell head backward: dW = scale * dlogits^T x  (C x D, fp32) and db = scale * sum_b dlogits.
//
// Replaces autograd's addmm backward for models/super_guessr.py:354 (reached from
// main_coordinator_idun_s3.py:423).  Both operands are consumed in the layout the forward pass
// left them in -- dlogits (B, ldc) and x (B, D), batch-major -- i.e. as MN-major UMMA operands:
// the contraction index (batch) is the slow index of both.  TMA loads 64(k) x 64(mn) boxes with
// the 128-byte swizzle.
//
// CTA pairs (tcgen05 cta_group::2): a cluster of two CTAs on the two SMs of a TPC executes one
// 256 (geocells) x 256 (embedding columns) x 16 MMA; each CTA holds its own 128 geocells of dlogits
// (2 boxes) and HALF of the x tile (2 of 4 boxes) per 64-row k-block -- 32 KB per CTA and stage, six
// stages -- and its 128 x 256 half of the accumulator in its own TMEM (double buffered).  The pair
// leader issues the MMAs; both CTAs' loads signal the leader's `full` barrier.
//
// Schedule: whole rounds of pair-tiles are handed out round-robin (pair p takes tiles p, p + P, ...: the
// pairs that run concurrently work on the few geocell blocks whose dlogits then sit in L2, so dlogits is
// streamed from HBM once).  The last, incomplete round is cut stream-K style: its (tile, k-block) units
// are split into one contiguous, equally long range per pair (at cfg2: 200 pair-tiles over 74 pairs = 2
// whole rounds + 52 tiles x 64 k-blocks = 45 units per pair instead of a third, 70 %-idle round).  A tile
// whose k-blocks are spread over two or three ranges is FINISHED by the pair that holds its last k-blocks; the
// others PARK their raw fp32 partial in the workspace and raise a flag.  A pair does its parking part FIRST,
// before the whole rounds, and its finishing part LAST: partials are parked ~10 us into the launch and nobody
// waits for anybody at the end (tools/head_bwd_timeline.py: with the tail round walked at the end and the
// finisher adding the parked partial in its epilogue, a middle pair parked at 72 us what the next pair needed to
// finish, and the launch ended at 81 us with the last MMA issued at 65).  The finisher does not add in its
// epilogue either: while the MMAs of its last whole tile run, its epilogue warps -- idle then -- sum the parked
// partials in pair order and write them INTO the TMEM accumulator its finishing part will use (tcgen05.st); the
// MMAs of that part accumulate on top and the epilogue is the ordinary one.  Waits only ever point at pairs
// that parked as their first piece of work, depending on nobody: no deadlock, and a fixed summation order
// (deterministic).
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace gg {

constexpr int kWM = 128;  // geocells per CTA (UMMA M = 256 over the pair)
constexpr int kWN = 256;  // embedding columns per pair tile (UMMA N; 128 per CTA in shared memory)
constexpr int kWK = 64;   // batch rows per stage
constexpr int kWStages = 6;
constexpr int kBwdThreads = 192;
constexpr uint32_t kAtomBytes = kWK * 128;                 // 64 k-rows x 128 B
constexpr uint32_t kWStageA = (kWM / 64) * kAtomBytes;     // 16 KB: my 128 geocells
constexpr uint32_t kWStageB = (kWN / 128) * kAtomBytes;    // 16 KB: my half of the embedding columns
constexpr size_t kPartialFloats = static_cast<size_t>(kWM) * kWN;  // one CTA's parked accumulator half
constexpr int kGradMaxWorld = 8;

struct BwdSmem {
  uint8_t a[kWStages][kWStageA];
  uint8_t b[kWStages][kWStageB];
  uint8_t out[4][2][32 * 128];  // per epilogue warp: two 32 x 32 fp32 staging tiles (128-byte swizzled TMA store boxes)
  uint64_t full[kWStages];
  uint64_t empty[kWStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  float dbrow[kWM];  // push mode: a block's 128 db entries on their way to the reducer (one bulk store)
  uint64_t preload;  // leader's: the finishing part's accumulator holds the parked partials (8 epilogue warps arrive)
  uint32_t tmem_base;
};

// A pair's work list: `rounds` whole tiles (pair, pair + P, ...), then its range of the tail round's units.
struct PairSchedule {
  int pair, npairs, num_k, rounds, tail_base;
  long long tail_total;
  long long u0, u1;  // tail units [u0, u1): unit = tail_tile * num_k + k-block
  // all_streamk: no whole rounds -- every pair gets one contiguous range of ALL (tile, k-block) units.  Same work and
  // the same number of parked partials (one per pair), but the pairs' tiles then complete at times spread evenly over
  // the launch instead of in three bursts: what the push mode needs (see gg_head_bwd)
  __device__ PairSchedule(int pair_, int npairs_, int num_tiles, int nk, bool all_streamk)
      : pair(pair_), npairs(npairs_), num_k(nk) {
    rounds = all_streamk ? 0 : num_tiles / npairs;
    tail_base = rounds * npairs;
    tail_total = static_cast<long long>(num_tiles - tail_base) * nk;
    u0 = tail_total * pair / npairs;
    u1 = tail_total * (pair + 1) / npairs;
  }
  // Does my range end inside a tile (its top part is parked for the pair that holds the tile's last k-blocks)?
  __device__ bool parks() const { return u1 > u0 && (u1 % num_k) != 0; }
  __device__ int tail_segments() const {
    return u1 > u0 ? static_cast<int>((u1 - 1) / num_k - u0 / num_k) + 1 : 0;
  }
  __device__ int segments() const { return rounds + tail_segments(); }
  // i-th segment in execution order: the parking part (if any), the whole rounds, the rest of the tail range from the
  // top down (whole tiles, then the finishing part); tile and k-block range [k0, k1)
  __device__ void get(int i, int& tile, int& k0, int& k1) const {
    const int pk = parks() ? 1 : 0;
    int j;  // tail segment, 0 = top of the range
    if (pk && i == 0) {
      j = 0;
    } else if (i - pk < rounds) {
      tile = (i - pk) * npairs + pair; k0 = 0; k1 = num_k;
      return;
    } else {
      j = i - rounds;
    }
    const int t = static_cast<int>((u1 - 1) / num_k) - j;
    const long long base = static_cast<long long>(t) * num_k;
    k0 = static_cast<int>((u0 > base ? u0 : base) - base);
    k1 = static_cast<int>((u1 < base + num_k ? u1 : base + num_k) - base);
    tile = tail_base + t;
  }
  // does pair q hold any of the tail units [a, b)?  (ranges of a short tail round can be empty)
  __device__ bool holds(int q, long long a, long long b) const {
    const long long q0 = tail_total * q / npairs, q1 = tail_total * (q + 1) / npairs;
    return q1 > q0 && q0 < b && q1 > a;
  }
  // owner of tail unit u
  __device__ int owner(long long u) const {
    return static_cast<int>(((u + 1) * npairs + tail_total - 1) / tail_total) - 1;  // ceil((u + 1) P / total) - 1
  }
};

// Data-parallel training (gg_grad_exchange): dW is cut into blocks of 128 geocells (one CTA's rows of a pair tile);
// block b is reduced by rank b % world.  In that mode this kernel does not write the gradient buffer at all: every
// finished tile -- and, with the first column tile, the block's 128 db entries -- goes straight into the REDUCER's
// staging slab for this source rank (posted TMA / plain stores over NVLink; the local slab for blocks this rank
// reduces itself), and every delivered tile is counted on the reducer's `ready` counter of the block with one relaxed
// add, issued after the tile's bulk stores have completed.  The transfer thus rides underneath the GEMM tile by tile,
// without any other kernel sharing the SMs with it and without a fence instruction in the epilogue.
struct GradPush {
  unsigned int* blk_count;             // (unused since every tile is counted on the reducer's own counter)
  unsigned int* ready[kGradMaxWorld];  // every rank's `ready` counters (peer-mapped)
  float* stage_b[kGradMaxWorld];       // rank r's db staging rows for THIS source rank (peer-mapped): [block / world][128]
  int world;                           // 0: plain local dW / db
};
struct StageMaps {
  CUtensorMap m[kGradMaxWorld];        // rank r's dW staging slab for this source rank: (rows = blocks-of-r x 128, D) fp32
};
__device__ __forceinline__ void red_release_sys_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Optional per-CTA timeline (tools/head_bwd_timeline.py, gg_debug_head_bwd_timeline): 64 globaltimer stamps per CTA.
// 0 entry, 1 set-up done, 2 exit, 3 last MMA issued, 4 number of segments, 5 parked-partial wait over; per segment i < 16:
// 8+i accumulator seen complete by the epilogue, 24+i epilogue done with it, 40+i its last MMA issued.
__device__ __forceinline__ void bwd_stamp(long long* tl, int slot) {
  if (tl != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[static_cast<size_t>(blockIdx.x) * 64 + slot] = t;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBwdThreads, 1)
head_bwd_kernel(const __grid_constant__ CUtensorMap tm_g,   // dlogits: inner C, rows B
                const __grid_constant__ CUtensorMap tm_x,   // x:       inner D, rows B
                const __grid_constant__ CUtensorMap tm_dw,  // dW (C, D) fp32: 32 x 32 store boxes
                const __grid_constant__ StageMaps stage, const __grid_constant__ GradPush sig, int C, int D, int Bk,
                float scale_in,
                const float* __restrict__ grad_scale, float* __restrict__ parked, int* __restrict__ flags,
                float* __restrict__ db, const float* __restrict__ db_partials, int db_parts, int db_ld,
                const float* __restrict__ db_ready, long long* __restrict__ timeline, int all_streamk) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (C + 2 * kWM - 1) / (2 * kWM);  // pairs of geocell blocks
  const int num_n = (D + kWN - 1) / kWN;
  const int num_tiles = num_m * num_n;              // pair-tiles
  const int crank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_k = (Bk + kWK - 1) / kWK;
  const float scale = grad_scale ? scale_in * __ldg(grad_scale) : scale_in;
  const PairSchedule ps(pair, npairs, num_tiles, num_k, all_streamk != 0);
  const int nseg = ps.segments();

  if (threadIdx.x == 0) {
    bwd_stamp(timeline, 0);
    if (timeline != nullptr) timeline[static_cast<size_t>(blockIdx.x) * 64 + 4] = nseg;
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dw);
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(&sm.full[s], 1);   // leader's: its own arrive.expect_tx, bytes from both CTAs' loads
      mbar_init(&sm.empty[s], 1);  // one multicast commit per round
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.acc_full[a], 1);
      mbar_init(&sm.acc_empty[a], 2 * 4);  // leader's: one arrive per epilogue warp of BOTH CTAs
    }
    mbar_init(&sm.preload, 2 * 4);
    fence_barrier_init();
  }
  if (warp == 1) {  // collective over the pair: one warp in each CTA
    tmem_alloc_pair(&sm.tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync();  // the partner's barriers exist before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  if (threadIdx.x == 0) bwd_stamp(timeline, 1);

  {
    if (warp == 0) {
      // ===== TMA producer: warp-uniform loop, one elected lane issues =====
      const uint32_t leader_full = mapa_u32(smem_u32(&sm.full[0]), 0);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nseg; ++i) {
        int t, k0, k1;
        ps.get(i, t, k0, k1);
        const int m0 = (2 * (t / num_n) + crank) * kWM, n0 = (t % num_n) * kWN + crank * (kWN / 2);
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&sm.empty[s], ph ^ 1);
          if (elect_one()) {
            if (crank == 0) mbar_arrive_expect_tx(&sm.full[s], 2 * (kWStageA + kWStageB));
#pragma unroll
            for (int i = 0; i < kWM / 64; ++i)  // my 128 geocells
              tma_load_2d_pair(sm.a[s] + i * kAtomBytes, &tm_g, leader_full + 8 * s, m0 + 64 * i, kb * kWK);
#pragma unroll
            for (int i = 0; i < kWN / 128; ++i)  // my half of the embedding columns
              tma_load_2d_pair(sm.b[s] + i * kAtomBytes, &tm_x, leader_full + 8 * s, n0 + 64 * i, kb * kWK);
          }
          __syncwarp();
          if (++s == kWStages) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer: pair leader only; warp-uniform loop, one elected lane issues =====
      if (crank == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(2 * kWM, kWN, 1, 1);
        // 16 k-rows of 128 B per MMA; atoms (64 mn-elements) kAtomBytes apart; 8-row groups 1024 B apart
        const uint64_t da_base = umma_desc_sw128(smem_u32(sm.a[0]), kAtomBytes, 1024);
        const uint64_t db_base = umma_desc_sw128(smem_u32(sm.b[0]), kAtomBytes, 1024);
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (; it < nseg; ++it) {
          const int acc = it & 1;
          const uint32_t acc_ph = (it >> 1) & 1;
          int t, k0, k1;
          ps.get(it, t, k0, k1);
          mbar_wait(&sm.acc_empty[acc], acc_ph ^ 1);  // both CTAs' epilogues have drained this accumulator
          // finishing part of a tile other pairs started: the epilogue warps have written the sum of the parked
          // partials into this accumulator; the MMAs go on top of it
          const bool finishing = k0 > 0 && k1 == num_k;
          if (finishing) mbar_wait(&sm.preload, 0);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kWN;
          for (int kb = k0; kb < k1; ++kb) {
            mbar_wait(&sm.full[s], ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = da_base + static_cast<uint64_t>(s * (kWStageA >> 4));
              const uint64_t db = db_base + static_cast<uint64_t>(s * (kWStageB >> 4));
              umma_pair_f16(d_tmem, da, db, idesc, kb != k0 || finishing);
#pragma unroll
              for (int k = 1; k < kWK / 16; ++k)
                umma_pair_f16_acc(d_tmem, da + k * (2048 >> 4), db + k * (2048 >> 4), idesc);
              umma_pair_commit(&sm.empty[s], 0x3);
              if (kb == k1 - 1) umma_pair_commit(&sm.acc_full[acc], 0x3);
            }
            __syncwarp();
            if (++s == kWStages) { s = 0; ph ^= 1; }
          }
          if (lane == 0 && it < 16) bwd_stamp(timeline, 40 + it);
        }
        if (lane == 0) bwd_stamp(timeline, 3);
      }
    } else {
      // ===== epilogue: 4 warps, thread = one geocell row of this CTA's accumulator half =====
      const int quad = warp & 3;
      const int rit = quad * 32 + lane;  // row in tile
      const uint32_t leader_acc_empty = mapa_u32(smem_u32(&sm.acc_empty[0]), 0);
      // parked partials: [pair][cta][column][row] so that a warp's accesses are contiguous
      float* const my_park = parked + (static_cast<size_t>(pair) * 2 + crank) * kPartialFloats + rit;
      // Finishing part (always the last segment): which pairs parked the tile's earlier k-blocks, and when to fetch them
      int fin_seg = -1, fin_q0 = 0;
      long long fin_a = 0, fin_b = 0;  // the tail units the other pairs parked: [fin_a, fin_b)
      if (nseg > 0) {
        int t, k0, k1;
        ps.get(nseg - 1, t, k0, k1);
        if (k0 > 0 && k1 == num_k) {
          fin_seg = nseg - 1;
          fin_a = static_cast<long long>(t - ps.tail_base) * num_k;
          fin_b = fin_a + k0;
          fin_q0 = ps.owner(fin_a);
        }
      }
      const uint32_t leader_preload = mapa_u32(smem_u32(&sm.preload), 0);
      // sum of the parked partials (pair order) -> the accumulator the finishing part will use
      auto preload = [&]() {
        if (threadIdx.x == 64) {
          for (int q = fin_q0; q < pair; ++q) {
            if (!ps.holds(q, fin_a, fin_b)) continue;
            const long long t0 = clock64();
            while (atomicAdd(&flags[q * 2 + crank], 0) == 0) {
              __nanosleep(64);
              if (clock64() - t0 > 8000000000LL) { printf("gg: head_bwd partial of pair %d never arrived\n", q); __trap(); }
            }
          }
          __threadfence();
          bwd_stamp(timeline, 5);
        }
        named_bar_sync(1, 128);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (fin_seg & 1) * kWN;
#pragma unroll 1
        for (int c = 0; c < kWN / 32; ++c) {
          uint32_t r[32];
          const float* src = parked + (static_cast<size_t>(fin_q0) * 2 + crank) * kPartialFloats + rit;
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__ldcg(src + static_cast<size_t>(c * 32 + j) * kWM));
          for (int q = fin_q0 + 1; q < pair; ++q) {
            if (!ps.holds(q, fin_a, fin_b)) continue;  // (warp-uniform)
            src = parked + (static_cast<size_t>(q) * 2 + crank) * kPartialFloats + rit;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldcg(src + static_cast<size_t>(c * 32 + j) * kWM));
          }
          tmem_st_32x32b_x32(taddr + c * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_preload);
      };
      if (fin_seg >= 0 && fin_seg < 2) preload();  // (no earlier segment uses that accumulator)
      // push mode: count a tile whose stores have been performed; the block's last column tile announces the block
      int pend_blk = -1, pend_dst = 0;
      auto announce = [&](int blk, int dst_rank) {
        named_bar_sync(1, 128);  // every warp's bulk stores of that tile have been performed at the reducer
        // One add per (tile, CTA) on the reducer's counter of the block (it waits for num_n per rank).  Relaxed: the
        // ordering IS the completion of the bulk stores (cp.async.bulk.wait_group) this thread and, through the
        // barrier, the other warps waited for before this instruction was issued.
        if (threadIdx.x == 64)
          asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(sig.ready[dst_rank] + blk), "r"(1u) : "memory");
      };
      int it = 0;
      int obuf = 0;  // staging buffer of the next dW store (alternates per store)
      for (; it < nseg; ++it) {
        int t, k0, k1;
        ps.get(it, t, k0, k1);
        const int m0 = (2 * (t / num_n) + crank) * kWM, n0 = (t % num_n) * kWN;
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        const int row = m0 + rit;
        // where this CTA's 128 geocells of the tile go: the gradient itself, or the reducer's staging slab
        const bool push = sig.world >= 1;
        const int blk = m0 / kWM;
        const int dst_rank = push ? blk % sig.world : 0;
        const CUtensorMap* const tm_dst = push ? &stage.m[dst_rank] : &tm_dw;
        const int dst_row0 = push ? (blk / sig.world) * kWM : m0;
        const bool park = k1 < num_k;      // another pair holds this tile's last k-blocks and finishes it
        mbar_wait(&sm.acc_full[acc], acc_ph);
        tc_fence_after();
        if (threadIdx.x == 64 && it < 16) bwd_stamp(timeline, 8 + it);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kWN;
        int ngroups = 0;  // bulk stores (one group each) this warp issues for this segment
#pragma unroll 1
        for (int c = 0; c < kWN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (park) {
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcg(my_park + static_cast<size_t>(c * 32 + j) * kWM, __uint_as_float(r[j]));
          } else if (col0 < D && m0 < C) {  // (warp-uniform)
            // 32 geocells x 32 columns through a swizzled staging tile and one TMA store: full 128-byte lines
            // instead of 32 row-strided 16-byte pieces per store instruction; rows >= C / columns >= D are clipped.
            uint8_t* const buf = sm.out[quad][obuf];
            obuf ^= 1;
            if (lane == 0) tma_store_wait_read<1>();  // the store issued two stores ago has read this buffer
            __syncwarp();
            const uint32_t dst_row = smem_u32(buf) + lane * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t dst = dst_row + ((q ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(__uint_as_float(r[4 * q + 0]) * scale),
                           "f"(__uint_as_float(r[4 * q + 1]) * scale), "f"(__uint_as_float(r[4 * q + 2]) * scale),
                           "f"(__uint_as_float(r[4 * q + 3]) * scale)
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tm_dst, buf, col0, dst_row0 + quad * 32);
              tma_store_commit();
            }
            ++ngroups;
          }
        }
        tc_fence_before();
        __syncwarp();
        // the accumulator this segment used is the finishing part's: fill it with the parked partials before the
        // MMA warp gets there (it is busy with the segment in between for another ~20 us)
        if (it == fin_seg - 2) preload();
        if (lane == 0) mbar_arrive_cluster(leader_acc_empty + 8 * acc);
        if (threadIdx.x == 64 && it < 16) bwd_stamp(timeline, 24 + it);
        // bias gradient: whoever finishes a geocell block's first column tile also sums that block's column-sum
        // partials from the loss kernel (fixed order; 128 consecutive geocells per warp-quartet: coalesced)
        const bool db_tile = (db != nullptr || push) && !park && n0 == 0 && m0 < C;  // (warp-uniform)
        if (db_tile) {
          float v = 0.f;
          if (row < C) {
          if (db_partials != nullptr) {
            float s0 = 0.f, s1 = 0.f;
            int i = 0;
            for (; i + 1 < db_parts; i += 2) {
              s0 += __ldg(db_partials + static_cast<size_t>(i) * db_ld + row);
              s1 += __ldg(db_partials + static_cast<size_t>(i + 1) * db_ld + row);
            }
            if (i < db_parts) s0 += __ldg(db_partials + static_cast<size_t>(i) * db_ld + row);
            v = (s0 + s1) * scale;
          } else {
            v = __ldg(db_ready + row);  // finished before this launch (column sums of dlogits)
          }
          }
          if (push) {
            // the block's 128 entries leave as ONE bulk store issued by thread 64, whose completion is awaited with the
            // tile's other bulk stores (below): no fence instruction in this epilogue -- every membar stalled the SM's
            // memory pipeline, operand loads included, for microseconds (tools/dp_overlap_probe.py: 14 us per launch)
            if (threadIdx.x == 64) tma_store_wait_read<0>();  // the previous block's entries have left shared memory
            named_bar_sync(1, 128);
            sm.dbrow[rit] = v;
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (threadIdx.x == 64) {
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                               sig.stage_b[dst_rank] + static_cast<size_t>(blk / sig.world) * kWM),
                           "r"(smem_u32(sm.dbrow)), "r"(static_cast<uint32_t>(kWM * sizeof(float)))
                           : "memory");
              tma_store_commit();
            }
            if (warp == 2) ++ngroups;  // (thread 64's warp)
          } else if (row < C) {
            db[row] = v;
          }
        }
        if (park) {  // publish: every epilogue thread's stores (one fence, cumulative through the barrier), then the flag
          named_bar_sync(1, 128);
          if (threadIdx.x == 64) {
            __threadfence();
            atomicExch(&flags[pair * 2 + crank], 1);
          }
        } else if (push && m0 < C) {
          // this CTA's 128 geocells x 256 columns of dW (and, with the first column tile, the block's db rows) are
          // on their way to the reducer.  Counting a tile needs its bulk stores PERFORMED (acknowledged by the reducer's memory), which under a busy
          // NVLink takes a while: the wait is for the PREVIOUS pushed tile -- everything but the groups just issued --
          // and that tile is counted now; this one when the next tile has been issued (or after the loop).
          if (lane == 0) {
            tma_store_wait_all_but(ngroups);
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (pend_blk >= 0) announce(pend_blk, pend_dst);
          pend_blk = blk;
          pend_dst = dst_rank;
        }
      }
      if (lane == 0) {
        tma_store_wait_all<0>();  // shared memory must outlive the last bulk store; the last pushed tile is performed
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      if (pend_blk >= 0) announce(pend_blk, pend_dst);
    }
  }
  tc_fence_before();
  cluster_sync();  // the partner may still be signalling this CTA's barriers / the leader reading its operands
  if (threadIdx.x == 0) bwd_stamp(timeline, 2);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// db: column sums of dlogits.  Block (32 x 8): each thread owns 8 adjacent columns (one 16-byte
// load per row), 8 row lanes; grid.y row slices write partial[slice][C], reduced by a second kernel
// (deterministic, no atomics).
constexpr int kDbSlices = 32;
__global__ void db_partial_kernel(const bf16* __restrict__ g, int ldc, int B, int C, float* __restrict__ partial) {
  __shared__ float red[8][32][8];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // columns bounded by C rounded up to 8 (<= ldc): with a geocell range (dlogits advanced by c0 columns) the rows'
  // 16-byte loads stay inside the allocation
  if (c0 < ((C + 7) & ~7)) {
    for (int r = blockIdx.y * 8 + threadIdx.y; r < B; r += 8 * gridDim.y) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + static_cast<size_t>(r) * ldc + c0));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x][i];
      if (c0 + i < C) partial[static_cast<size_t>(blockIdx.y) * C + c0 + i] = s;
    }
  }
}
// Block (32 columns x 8 slice lanes): lane y sums slices y, y+8, ... (coalesced 128-byte rows), then a
// fixed-order reduction over y in shared memory: deterministic, and wide enough to hide the latency.
__global__ void db_final_kernel(const float* __restrict__ partial, int C, int ld, int slices, float scale_in,
                                const float* __restrict__ grad_scale, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    int i = threadIdx.y;
    for (; i + 8 < slices; i += 16) {
      s0 += partial[static_cast<size_t>(i) * ld + c];
      s1 += partial[static_cast<size_t>(i + 8) * ld + c];
    }
    if (i < slices) s0 += partial[static_cast<size_t>(i) * ld + c];
  }
  red[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    const float scale = grad_scale ? scale_in * __ldg(grad_scale) : scale_in;
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x];
    db[c] = s * scale;
  }
}

}  // namespace gg

using namespace gg;

// workspace: [flags: one int per CTA][parked stream-K partials: 128 x 256 fp32 per CTA][db column-sum slices]
static size_t bwd_flag_bytes() { return ((static_cast<size_t>(device_sm_count()) * sizeof(int)) + 255) & ~size_t(255); }
static size_t bwd_park_bytes() { return static_cast<size_t>(device_sm_count()) * kPartialFloats * sizeof(float); }
extern "C" size_t gg_head_bwd_workspace_bytes(int C) {
  // + C floats: db finished ahead of the GEMM when it is pushed to a reducer without the loss kernel's partials
  return bwd_flag_bytes() + bwd_park_bytes() + static_cast<size_t>(kDbSlices + 1) * C * sizeof(float);
}

// staging region of one rank: `world` slabs (one per source rank), each = its blocks' dW rows then their db entries
static size_t grad_blocks_per_rank(int C, int world) { return (static_cast<size_t>(ceil_div(C, kWM)) + world - 1) / world; }
extern "C" size_t gg_grad_stage_floats(int C, int D, int world) {
  return static_cast<size_t>(world) * grad_blocks_per_rank(C, world) * kWM * (static_cast<size_t>(D) + 1);
}

static long long* g_bwd_timeline = nullptr;  // gg_debug_head_bwd_timeline
extern "C" void gg_debug_head_bwd_timeline(long long* device_buf) { g_bwd_timeline = device_buf; }

extern "C" int gg_head_bwd(const void* dlogits_bf16, int ldc, const void* x_bf16, int x_ld, int B, int C, int D,
                           float scale, const float* grad_scale, float* dW, float* db, const float* db_partials,
                           int db_parts, int db_ld, void* workspace, const unsigned long long* dp_ptrs, int dp_world,
                           int dp_rank, int schedule, gg_stream_t stream) {
  GG_CHECK(B > 0 && C > 0 && D > 0, GG_ERR_ARG, "gg_head_bwd: empty problem B=%d C=%d D=%d", B, C, D);
  GG_CHECK(dlogits_bf16 && x_bf16, GG_ERR_ARG, "gg_head_bwd: null pointer");
  GG_CHECK(ldc >= C && ldc % 8 == 0, GG_ERR_ARG, "gg_head_bwd: ldc=%d must be >= C and a multiple of 8", ldc);
  GG_CHECK(D % 8 == 0 && x_ld >= D && x_ld % 8 == 0, GG_ERR_ARG, "gg_head_bwd: D=%d / x_ld=%d must be multiples of 8", D, x_ld);
  GG_CHECK(workspace, GG_ERR_ARG, "gg_head_bwd: workspace (gg_head_bwd_workspace_bytes) is required");
  GG_CHECK(!db_partials || (db_parts > 0 && db_ld >= C), GG_ERR_ARG, "gg_head_bwd: bad db_partials shape");
  const bool push = dp_ptrs != nullptr && dp_world >= 1;  // (one rank: gg_grad_exchange_adamw's single-GPU form)
  GG_CHECK(dp_world >= 0 && dp_world <= kGradMaxWorld && (!push || (dp_ptrs && dp_rank >= 0 && dp_rank < dp_world)),
           GG_ERR_ARG, "gg_head_bwd: dp_world=%d dp_rank=%d (<= %d ranks) needs dp_ptrs", dp_world, dp_rank, kGradMaxWorld);
  GG_CHECK(push || dW, GG_ERR_ARG, "gg_head_bwd: dW is required");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_g, tm_x, tm_dw;
  int rc = make_tmap_bf16_2d(&tm_g, dlogits_bf16, C, B, static_cast<uint64_t>(ldc) * 2, 64, kWK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, x_bf16, D, B, static_cast<uint64_t>(x_ld) * 2, 64, kWK);
  if (rc) return rc;
  GradPush sig = {};
  StageMaps maps = {};
  if (push) {
    const size_t rows = grad_blocks_per_rank(C, dp_world) * kWM;        // staged geocell rows per (reducer, source)
    const size_t slab = rows * (static_cast<size_t>(D) + 1);            // floats
    sig.world = dp_world;
    sig.blk_count = reinterpret_cast<unsigned int*>(dp_ptrs[0]);
    GG_CHECK(sig.blk_count, GG_ERR_ARG, "gg_head_bwd: block counters missing");
    for (int r = 0; r < dp_world; ++r) {
      const unsigned long long ready = dp_ptrs[1 + r], stage_r = dp_ptrs[1 + dp_world + r];
      GG_CHECK(ready != 0 && stage_r != 0 && (stage_r & 15) == 0, GG_ERR_ARG, "gg_head_bwd: buffers of rank %d missing", r);
      sig.ready[r] = reinterpret_cast<unsigned int*>(ready);
      float* my_slab = reinterpret_cast<float*>(stage_r) + static_cast<size_t>(dp_rank) * slab;
      sig.stage_b[r] = my_slab + rows * D;
      rc = make_tmap_f32_2d(&maps.m[r], my_slab, D, rows, static_cast<uint64_t>(D) * 4, 32);
      if (rc) return rc;
    }
    tm_dw = maps.m[0];  // (not used in this mode)
  } else {
    rc = make_tmap_f32_2d(&tm_dw, dW, D, C, static_cast<uint64_t>(D) * 4, 32);
    if (rc) return rc;
    for (int r = 0; r < kGradMaxWorld; ++r) maps.m[r] = tm_dw;
  }
  const int pair_tiles = ceil_div(C, 2 * kWM) * ceil_div(D, kWN);
  const int pairs = std::max(1, std::min(pair_tiles, device_sm_count() / 2));
  const size_t smem = sizeof(BwdSmem) + 1024;
  if (int e = set_max_dynamic_smem_once(head_bwd_kernel, smem)) return e;
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  int* flags = reinterpret_cast<int*>(wsb);
  float* parked = reinterpret_cast<float*>(wsb + bwd_flag_bytes());
  float* db_slices = reinterpret_cast<float*>(wsb + bwd_flag_bytes() + bwd_park_bytes());
  float* db_tmp = db_slices + static_cast<size_t>(kDbSlices) * C;
  // with the loss kernel's column-sum partials at hand the GEMM's epilogue finishes db as well
  const bool db_fused = (db || push) && db_partials;
  auto db_from_dlogits = [&](float* out) -> int {
    dim3 blk(32, 8), grd(ceil_div(C, 256), kDbSlices);
    db_partial_kernel<<<grd, blk, 0, s>>>(static_cast<const bf16*>(dlogits_bf16), ldc, B, C, db_slices);
    GG_LAUNCH_CHECK();
    db_final_kernel<<<ceil_div(C, 32), dim3(32, 8), 0, s>>>(db_slices, C, C, kDbSlices, scale, grad_scale, out);
    GG_LAUNCH_CHECK();
    return GG_OK;
  };
  // pushed blocks carry their db entries: without the partials db is finished BEFORE the GEMM then
  if (push && !db_fused)
    if (int e = db_from_dlogits(db_tmp)) return e;
  GG_CUDA(cudaMemsetAsync(flags, 0, bwd_flag_bytes(), s));
  // Push mode: 7/8 of the fp32 tiles leave through NVLink (606 GB/s for SM-issued stores, tools/symm_probe.py).  With
  // whole rounds every CTA finishes its tiles at the same three moments, 19 MB each time, and the launch became a
  // sequence of bursts (137 us at 8 ranks); with one contiguous unit range per pair the completions -- and the pushes --
  // are spread over the whole launch.  On one GPU the rounds are 3-6 us faster (dlogits tiles shared in L2 by the
  // pairs that run the same geocell block at the same time).
  GG_CHECK(schedule == GG_BWD_SCHEDULE_AUTO || schedule == GG_BWD_ROUNDS || schedule == GG_BWD_STREAMK, GG_ERR_ARG,
           "gg_head_bwd: flags=%d", schedule);
  const int all_streamk = schedule == GG_BWD_SCHEDULE_AUTO ? (push ? 1 : 0) : (schedule == GG_BWD_STREAMK ? 1 : 0);
  // cluster shape (2,1,1) is compiled into the kernel
  head_bwd_kernel<<<2 * pairs, kBwdThreads, smem, s>>>(tm_g, tm_x, tm_dw, maps, sig, C, D, B, scale, grad_scale, parked,
                                                       flags, db_fused && !push ? db : nullptr, db_partials, db_parts, db_ld,
                                                       db_tmp, g_bwd_timeline, all_streamk);
  GG_LAUNCH_CHECK();
  if (db && !db_fused && !push)
    if (int e = db_from_dlogits(db)) return e;
  return GG_OK;
}
