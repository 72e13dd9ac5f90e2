/* geoguessr_b200.h -- C ABI of the B200-native post-encoder geolocation path.
 *
 * libgeoguessr_b200.so exports exactly these symbols.  All pointers are DEVICE pointers unless
 * noted; sizes are element counts; `stream` is a cudaStream_t passed as void*; every launcher
 * is asynchronous on that stream, re-entrant, allocates nothing and never synchronises.  Return
 * value: GG_OK or a GG_ERR_* code; gg_last_error() (host, thread-local) holds the message.  There
 * is no CPU fallback: a call on a machine without an sm_100 device fails with GG_ERR_CUDA.
 *
 * The reference (CogitoNTNU/geoguessr-ai) has no FFI: its boundary for this path is two
 * torch.nn.Module classes.  Each entry point below names the reference lines it replaces; the
 * Python modules geoguessr_ai_b200.SuperGuessr / ProtoRefiner (same constructor, forward and
 * state-dict contract) are built on top of these calls via ctypes.  See INTEGRATION.md.
 */
#ifndef GEOGUESSR_B200_H
#define GEOGUESSR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 4

#define GG_OK 0
#define GG_ERR_ARG 1         /* bad shape / alignment / null pointer */
#define GG_ERR_CUDA 2        /* CUDA runtime or driver error (no device, launch failure, ...) */
#define GG_ERR_UNSUPPORTED 3 /* valid request outside what the kernels cover */

typedef void* gg_stream_t; /* cudaStream_t */

int gg_abi_version(void);
const char* gg_last_error(void);

/* ---- layout helpers (host, pure) ---------------------------------------------------------- */
int gg_head_logits_ld(int C); /* row pitch (elements) of the bf16 logits / dlogits buffers: C rounded up to 256 */
int gg_head_bias_pad(int C);  /* length of the zero-padded fp32 bias vector: C rounded up to 256 */
int gg_hav_cpad(int C);       /* geocells rounded up to 256: column pitch of the centroid table and of db_partials */
size_t gg_centroid_table_floats(int C); /* 4-byte words in the table gg_centroid_unit_vectors fills: unit vectors in
                                           class order + a Morton-sorted copy with one bounding cap per 64 cells */
size_t gg_centroid_table_workspace_bytes(int C);
size_t gg_hav_row_stats_bytes(int B, int C);
size_t gg_head_fwd_workspace_bytes(int B, int C, int k);
size_t gg_head_fwd_ticket_bytes(int B); /* merge tickets of gg_head_fwd (one word per 128-row block) */
size_t gg_head_bwd_workspace_bytes(int C);
size_t gg_hav_ce_workspace_bytes(int B, int C);
int gg_hav_ce_db_parts(int B, int C); /* rows of the loss kernel's bias-gradient partial sums (one per row block) */
size_t gg_proto_retrieve_workspace_bytes(int B, int topk, int D, int ncell);

/* ---- a1: heading fusion -------------------------------------------------------------------
 * models/super_guessr.py:347  `output = layer_input.mean(dim=1)`  ((N,4,C) -> (N,C));
 * models/proto_refiner.py:150-151 (same mean for the refiner's queries).
 * emb (B,V,D) -> x (B, D) bf16 [split=0] or (B, 3D) = [hi|hi|lo] [split=1, fp32-faithful
 * "bf16x3" operands].  V=1 is the single-image pass-through (:351).  sqnorm (B) optional: ||x||^2.
 * in_dtype: element type of emb -- GG_IN_F32 (the reference's storage format, backend/s3bucket.py:848-859), or,
 * opt-in, GG_IN_BF16 / GG_IN_F16 embeddings read directly (half the bytes over PCIe; the mean is taken in fp32). */
#define GG_IN_F32 0
#define GG_IN_BF16 1
#define GG_IN_F16 2
int gg_fuse_headings(const void* emb, int in_dtype, void* x_bf16, int B, int V, int D, int split, float* sqnorm,
                     gg_stream_t stream);

/* Operand preparation for nn.Linear(D, C) (super_guessr.py:102): W (C,D) fp32 -> bf16 (C,D), or
 * (C,3D) = [hi|lo|hi] when split=1; b (C) -> zero-padded fp32 (gg_head_bias_pad(C)). */
int gg_prepare_head_weights(const float* w, const float* b, void* w_bf16, float* bias_pad, int C, int D, int split,
                            gg_stream_t stream);
/* Both of the above for one training step in a single launch (the weights move every step, so their bf16
 * operand is rebuilt next to the fusion of that step's batch): same arguments and results, no sqnorm. */
int gg_fuse_and_prepare(const void* emb, int in_dtype, void* x_bf16, int B, int V, int D, const float* w,
                        const float* b, void* w_bf16, float* bias_pad, int C, int split, gg_stream_t stream);
/* Prototype-bank operand: (rows, D) fp32 -> bf16 (rows, D), or the hi/lo split (rows, 3D) = [hi | lo | hi] that,
 * contracted against queries [hi | hi | lo], keeps fp32-faithful dot products ("bf16x3" retrieval). */
int gg_cast_bf16(const float* src, void* dst_bf16, long long rows, int D, int split, gg_stream_t stream);
/* ||row||^2 of a bf16 matrix (split: rows are [hi | lo | hi] and stand for hi + lo). */
int gg_row_sqnorm_bf16(const void* m_bf16, long long rows, int D, int split, float* out, gg_stream_t stream);

/* ---- a2-a4: geocell head ------------------------------------------------------------------
 * super_guessr.py:354 logits = cell_layer(output); :355 softmax; :358-361 argmax + centroid gather;
 * :365 torch.topk(probs, num_candidates).
 * x (B,D) bf16, w (C,D) bf16 (D here = the K extent, 3x the embed dim in split mode).
 * logits_bf16: (B, ldc) written only when non-null (training); serving never materialises them.
 * Outputs: topk_val (B,k) fp32 probabilities, sorted descending; topk_idx (B,k) int64; pred_cell (B)
 * int64 = argmax; pred_llh (B,2) fp32 = centroids[pred_cell] ((lng,lat)); lse (B) fp32 row
 * log-sum-exp (consumed by the loss).  pred_cell / pred_llh / lse may be null.  Any 1 <= k <= C, as
 * torch.topk: up to 8 candidates come out of the GEMM's epilogue in one pass; beyond that the GEMM runs once more
 * per further 8 ranks, each pass restricted to the logits ordered after the previous pass's last entry (exact:
 * every pass sees the same fp32 accumulators).
 * One launch: the CTA that flushes the last partial of a 128-row block merges that block's partials (softmax
 * denominators, top-k lists) -- `tickets` (gg_head_fwd_ticket_bytes(B) bytes) counts the flushes; the caller
 * zero-initialises it ONCE, every launch leaves it zeroed, and it serves one launch at a time (a stream). */
int gg_head_fwd(const void* x_bf16, const void* w_bf16, const float* bias_pad, int B, int C, int D, void* logits_bf16,
                int ldc, int k, void* workspace, void* tickets, const float* centroids, float* topk_val,
                long long* topk_idx, long long* pred_cell, float* pred_llh, float* lse, gg_stream_t stream);

/* Tuning aid (tools/head_fwd_timeline.py), not part of the path: while a device buffer of 64 int64 per CTA (2 per
 * SM) is set, every gg_head_fwd launch stamps %globaltimer at fixed points of each CTA's life (entry, set-up done,
 * first / last MMA issued, first / last accumulator complete, chunks done, fold, flush, merge, epilogue drained,
 * exit).  null switches it off (the default; the kernel then only tests a pointer). */
void gg_debug_head_fwd_timeline(long long* device_buf);
/* The same for gg_head_bwd (tools/head_bwd_timeline.py): 64 int64 per CTA -- entry, set-up done, exit, last MMA issued,
 * number of segments, end of the wait for a parked partial; per segment (tile or part of a tail tile) accumulator
 * complete / epilogue done / last MMA issued. */
void gg_debug_head_bwd_timeline(long long* device_buf);
/* Hardware probe (tools/symm_probe.py), not part of the path: stores `bytes` to dst -- e.g. a peer's symmetric-memory
 * mapping -- from `ctas` CTAs, mode 0 = coalesced 16-byte st.global, mode 1 = cp.async.bulk (TMA, 1-D) of `chunk` bytes
 * per instruction from shared memory. */
int gg_debug_nvlink_store_probe(void* dst, size_t bytes, int mode, int chunk, int ctas, gg_stream_t stream);

/* ---- a5-a8: haversine label-smoothed cross-entropy, forward + gradient ---------------------
 * models/utils.py:39-57 haversine_matrix; :20-32 smooth_labels (tau = config.py:52 = 65 km);
 * super_guessr.py:374-380 normalise + soft CE; autograd backward (main_coordinator_idun_s3.py:423).
 *
 * gg_centroid_unit_vectors: once per centroid table ((C,2) fp32 (lng,lat) degrees) -> cent_table
 * (gg_centroid_table_floats(C) words).  Geocell tables up to 16384 cells.
 *
 * gg_hav_row_stats: everything that depends on the labels only -- per row the label unit vector, the
 * nearest centroid (row minimum of the haversine matrix, utils.py:29), sum_c exp(-(d - dmin)/tau)
 * (super_guessr.py:377) and which 64-class groups hold cells with d < dmin + far_km.  labels (B,2) fp32
 * (lng,lat) degrees; row_stats: gg_hav_row_stats_bytes(B,C) bytes, consumed by gg_hav_ce_fwd_bwd.  It reads
 * no logits, so it can run before / beside the head GEMM.  By-products = the trainer's label derivation
 * (main_coordinator_idun_s3.py:390-391): nearest_cell (B) int64 = argmin_c d (first index on ties),
 * nearest_km (B) fp32; either may be null.
 * far_km: cells farther than dmin + far_km get t = 0 (65*ln(2^32) ~= 1442 km, the Python default,
 * keeps every target above 2^-32 of the nearest cell's, i.e. below the fp32 rounding of the reference's
 * own row sum; INFINITY evaluates every cell).
 *
 * gg_hav_ce_fwd_bwd: the B x C pass.  logits / dlogits (B, ldc) bf16 with ldc >= gg_hav_cpad(C);
 * dlogits = softmax(logits) - t, UNSCALED (gg_head_bwd applies 1/B).  loss_rows (B) fp32 =
 * -sum_c t log_softmax; loss_mean (optional, 1 float) = mean_scale * sum_b loss_rows, summed in a fixed
 * order (deterministic) -- the `.mean()` of super_guessr.py:380 with mean_scale = 1/B.  db_partials
 * (optional, (gg_hav_ce_db_parts(B,C), gg_hav_cpad(C)) fp32) = per-row-block column sums of dlogits,
 * finished by gg_head_bwd (saves a second pass over dlogits for the bias gradient).
 * row_stats also carries the kernel's finishing tickets and per-row fixed-point accumulators (zeroed by
 * gg_hav_row_stats, left zeroed by every launch): one gg_hav_ce_fwd_bwd at a time per row_stats buffer. */
int gg_centroid_unit_vectors(const float* centroids, float* cent_table, int C, void* workspace, gg_stream_t stream);
int gg_hav_row_stats(const float* labels, const float* cent_table, int B, int C, float tau, float far_km,
                     void* row_stats, long long* nearest_cell, float* nearest_km, gg_stream_t stream);
int gg_hav_ce_fwd_bwd(const void* logits_bf16, int ldc, const float* lse, const void* row_stats,
                      const float* cent_table, int B, int C, float tau, void* dlogits_bf16, float* loss_rows,
                      float* db_partials, void* workspace, float* loss_mean, float mean_scale, gg_stream_t stream);
/* super_guessr.py:383 nn.CrossEntropyLoss()(logits, labels_clf) and its gradient. */
int gg_hard_ce_fwd_bwd(const void* logits_bf16, int ldc, const float* lse, const long long* labels_clf, int B, int C,
                       void* dlogits_bf16, float* loss_rows, gg_stream_t stream);
/* `.mean()` of super_guessr.py:380: loss_out[0] = scale * sum_b loss_rows[b] (deterministic). */
int gg_loss_mean(const float* loss_rows, int B, float scale, float* loss_out, gg_stream_t stream);

/* ---- a8: head backward --------------------------------------------------------------------
 * autograd of super_guessr.py:354: dW (C,D) fp32 = scale * dlogits^T x, db (C) = scale * colsum.
 * x (B, x_ld) bf16 (first D columns used).  scale = 1/B (1/global batch under data parallelism);
 * grad_scale: optional DEVICE scalar (the upstream dL/dloss of autograd), multiplied in on the
 * device so that backward needs no host synchronisation.  db_partials (db_parts, db_ld): column
 * sums from gg_hav_ce_fwd_bwd; when null, db is computed from dlogits.  workspace
 * (gg_head_bwd_workspace_bytes(C) bytes, required): parked partial accumulators + flags of the stream-K
 * schedule, and the column-sum slices of the db pass.
 * Data parallelism (gg_grad_exchange below): with dp_ptrs given (dp_world >= 1) the kernel does not write dW / db at all (both may be
 * null): every finished tile of 128 geocells x 256 columns -- and the block's 128 db entries -- is stored straight
 * into the staging slab of the rank that reduces the block (block b -> rank b % dp_world), over NVLink for the other
 * ranks, and every delivered tile is counted on that rank's `ready` counter of the block (one relaxed add per tile and
 * CTA, issued after the tile's bulk stores have completed -- that completion is the ordering: no fence instruction).  dp_ptrs: HOST array of 1 + 2 * dp_world
 * device addresses as mapped on this device = {this rank's control region + GG_GRAD_CTRL_BLKCOUNT_OFF (reserved, unused), control region +
 * GG_GRAD_CTRL_READY_OFF of rank 0, 1, ..., staging region of rank 0, 1, ...}.  dp_ptrs null: plain local dW / db.
 * flags: GG_BWD_SCHEDULE_AUTO (0), or which work split to use -- GG_BWD_ROUNDS: whole rounds of tiles + a stream-K tail
 * round (fastest on one GPU), GG_BWD_STREAMK: one contiguous range of (tile, k-block) units per CTA pair, so that tiles
 * complete -- and, in push mode, leave over NVLink -- evenly over the launch (auto picks it when dp_ptrs is given).
 * The two splits add the same products in different orders: results agree to fp32 rounding, and bit for bit only
 * under the same split. */
#define GG_BWD_SCHEDULE_AUTO 0
#define GG_BWD_ROUNDS 1
#define GG_BWD_STREAMK 2
int gg_head_bwd(const void* dlogits_bf16, int ldc, const void* x_bf16, int x_ld, int B, int C, int D, float scale,
                const float* grad_scale, float* dW, float* db, const float* db_partials, int db_parts, int db_ld,
                void* workspace, const unsigned long long* dp_ptrs, int dp_world, int dp_rank, int flags,
                gg_stream_t stream);

/* dx of the same layer (autograd of super_guessr.py:354 w.r.t. its input; reached from
 * main_coordinator_idun_s3.py:423 whenever the encoder is trained -- TinyViT's last stage / CLIP's last layer,
 * super_guessr.py:127-153) with the heading mean's 1/V broadcast (:347) in the epilogue:
 * demb (B, V, D) fp32, every heading v: demb[b, v, :] = scale * grad_scale / V * sum_c dlogits[b, c] W[c, :].
 * w (C, w_ld >= D) bf16 = the forward's operand, consumed as it lies (no transposed copy); tcgen05. */
int gg_head_dx(const void* dlogits_bf16, int ldc, const void* w_bf16, int w_ld, int B, int C, int D, float scale,
               const float* grad_scale, int V, float* demb, gg_stream_t stream);

/* ---- f-1: the trainer's per-step accuracy metrics -----------------------------------------------
 * main_coordinator_idun_s3.py:399-408: top1 = mean(topk_idx[:, 0] == targets), topk = mean(any_j topk_idx[:, j] ==
 * targets), from the head's top-k indices and the nearest-centroid targets (gg_hav_row_stats' nearest_cell): two
 * device floats acc[0], acc[1] (deterministic single-CTA sum) instead of two `.item()` syncs per step. */
int gg_topk_accuracy(const long long* topk_idx, int k, const long long* targets, int B, float* acc, gg_stream_t stream);

/* ---- f-4: `hierarchical=True` heading fusion ----------------------------------------------------
 * models/super_guessr.py:89-99,340-345 + models/layers/positional_encoder.py:21-44, eval mode: z = x + PE[batch row],
 * 16-head self-attention over the V headings, token 0, output projection.  The fp32 module is reproduced to fp32
 * accuracy on the tensor cores by splitting operands into three bf16 terms laid side by side along K (six products):
 * gg_split3_bf16: src (rows, D) fp32 [+ pos_encoding[row / V] when given] -> (rows, 6D) bf16, role 0 = activation
 * [h|h|m|h|l|m], role 1 = weight [h|m|h|l|h|m];  gg_linear_bf16: out (M, N) fp32 = a (M, K) w(N, K)^T + bias, tcgen05;
 * gg_hier_attention: qkv (B*V, 3D) fp32 -> softmax_t(q[b,0,h] . k[b,t,h] / sqrt(D/heads)) weighted values of token 0,
 * written as the role-0 split operand (B, 6D) of the output projection. */
int gg_split3_bf16(const float* src, long long rows, int D, int role, const float* pos_encoding, int V, void* dst_bf16,
                   gg_stream_t stream);
int gg_linear_bf16(const void* a_bf16, int lda, const void* w_bf16, int ldw, const float* bias, int M, int N, int K,
                   float* out, int ldo, gg_stream_t stream);
int gg_hier_attention(const float* qkv, int B, int V, int D, int heads, void* ctx_split_bf16, gg_stream_t stream);

/* ---- a10-a15: ProtoRefiner ----------------------------------------------------------------
 * models/proto_refiner.py:165-203 (retrieval) -- for every (query i, candidate j < topk):
 * score = max_p -||proto_{c,p} - q_i||_2 (:190-193, _euclidean_distance :364-376), arg-best
 * prototype (:194) and its coordinates (:251-252); a cell without prototypes scores -100000 with
 * coordinates (0,0) (:181-187).  Bank: prototypes sorted by geocell, CSR cell_off (ncell+1, int32,
 * local to this rank's cell range [cell_lo, cell_hi)); bank (n_protos, D) bf16, bank_sqnorm
 * (n_protos) fp32 from gg_row_sqnorm_bf16, bank_coords (n_protos,2) fp32 (lng,lat).  q (B,D) bf16 and
 * q_sqnorm (B) from gg_fuse_headings.  cand (B, cand_ld) int64.  rec_out: (B*topk) 16-byte records
 * {score f32, lng f32, lat f32, prototype id i32 (proto_base + local row, -1 if none)}; pairs whose
 * cell lies outside [cell_lo, cell_hi) get score = -inf (another rank owns them).
 * group_off (ngroups+1 ints, DEVICE; local cell indices): geocells packed into accumulation groups of <= 256
 * prototypes, built once per bank by gg_proto_group_cells (HOST arrays in, returns ngroups): small neighbouring
 * cells share one 256-prototype unit, each pair masked to its own cell's columns.
 * metric: GG_METRIC_L2 = the executable reference (-cdist, :190, :364-376); GG_METRIC_COSINE = the reference's unused
 * _cosine_similarity (:347-362), opt-in: score = q.p / (|q| |p|), arg-max per cell.
 * flags: GG_RETRIEVE_NO_GATHER4 = copy the query rows into cell order first (workspace) instead of letting the
 * TMA engine gather them four rows at a time (tile::gather4). */
#define GG_METRIC_L2 0
#define GG_METRIC_COSINE 1
#define GG_RETRIEVE_NO_GATHER4 1
int gg_proto_group_cells(const int* cell_off_host, int ncell, int* group_off_host);
int gg_proto_retrieve(const void* q_bf16, const float* q_sqnorm, int B, int D, const long long* cand, int cand_ld,
                      int topk, const void* bank_bf16, const float* bank_sqnorm, const float* bank_coords,
                      long long n_protos, const int* cell_off, int cell_lo, int cell_hi, const int* group_off,
                      int ngroups, int proto_base, int metric, int flags, void* rec_out, void* workspace,
                      gg_stream_t stream);
/* models/proto_refiner.py:205-228: temperature softmax (:378-389), x candidate probs (:210), argmax
 * (:211), max-refinement guard with preprocessing/geo_utils.py:39-54 haversine (:216-223), outputs
 * (:225-228).  rec: nranks record arrays rank_stride records apart (all-gather layout); cand_probs
 * (B, cand_probs_ld) fp32 or null (= one-hot on candidate 0, :154-156).  out_llh (B,2) fp32,
 * out_cell (B) int64; optional out_guess (B) int32 (index of the chosen candidate, :226),
 * out_score (B,topk) fp32 and out_proto (B,topk) int32 (merged stage-1 result). */
int gg_proto_refine(const void* rec, int nranks, long long rank_stride, const float* cand_probs, int cand_probs_ld,
                    const long long* cand, int cand_ld, const float* initial_llh, int B, int topk, float temperature,
                    float max_refinement_km, float* out_llh, long long* out_cell, int* out_guess, float* out_score,
                    int* out_proto, gg_stream_t stream);

/* Within-cluster refinement (f-3, second half; the reference's _within_cluster_refinement, proto_refiner.py:239-269,
 * cannot run and would pick the FARTHEST member): a second gg_proto_retrieve over the member images of each pair's
 * best prototype (bank = image embeddings sorted by cluster, "cells" = clusters, cand = the stage-1 prototype ids)
 * finds the NEAREST image; this call then replaces the coordinates of every stage-1 record whose cluster has images:
 * rec[i].{lng, lat} = rec_img[i].{lng, lat} where rec_img[i].id >= 0.  Scores and prototype ids are kept. */
int gg_proto_take_image_coords(void* rec, const void* rec_img, long long n, gg_stream_t stream);
/* prototype ids of a record array as the int64 candidate list of that second retrieval: cand2[i] = rec[i].id */
int gg_proto_record_ids(const void* rec, long long n, long long* ids, gg_stream_t stream);

/* ---- next row f-3: prototype bank builder --------------------------------------------------------
 * models/proto_refiner.py:391-406 + :461-517 (Embeddings.generate_embeddings) with the encoder replaced by stored
 * embeddings: prototype p = mean over the members of cluster p (members[member_off[p] .. member_off[p+1]), location
 * indices into emb (L,V,D) fp32, walked in list order) of the member's mean over its V headings; members < 0, >= L
 * or with valid[idx] == 0 (non-finite coordinates, :470-472) are skipped; a cluster without valid members gives
 * the zero vector (:499-515).  bank_bf16 (P,D) is the retrieval operand; bank_f32 (optional) the unrounded mean;
 * count (optional, P) the members used. */
int gg_build_prototypes(const float* emb, long long L, int V, int D, const long long* member_off, const int* members,
                        const unsigned char* valid, long long P, void* bank_bf16, float* bank_f32, int* count,
                        gg_stream_t stream);

/* ---- a8 across GPUs: gradient averaging for data-parallel head training -------------------------
 * The reference leaves this to DDP / Accelerate after loss.backward() (main_coordinator_idun_s3.py:423-424;
 * SURVEY 8e).  peer_ptrs[r] = device address, valid on THIS device, of rank r's copy of the gradient buffer
 * (n_floats fp32, e.g. [dW | db | pad], n_floats % 4 == 0), all mapped over NVLink (symmetric memory).  Two-shot
 * all-reduce in place: this rank sums slice `rank` of every copy in rank order (deterministic, identical on all
 * ranks), scales by 1 / world and writes the slice into every copy.  world in {1, 2, 4, 8}.  The caller must
 * order the launch after every rank has written its buffer and must not read any copy before every rank's launch
 * has completed (a symmetric-memory barrier on either side).  gg_p2p_slice: the [lo, hi) range of rank's slice
 * in 16-byte units. */
void gg_p2p_slice(size_t n_floats, int world, int rank, size_t* lo4, size_t* hi4);
int gg_p2p_allreduce_avg(const unsigned long long* peer_ptrs, int world, int rank, size_t n_floats, gg_stream_t stream);
/* The same exchange through the NVSwitch multicast mapping of the buffer (NVLS): multimem.ld_reduce adds the
 * copies of slice `rank` inside the switch, multimem.st broadcasts the average to every copy; about one buffer
 * of traffic per GPU and direction for any number of ranks.  multicast_ptr: the symmetric-memory handle's
 * multicast address of the buffer on this device.  Same ordering contract as gg_p2p_allreduce_avg. */
int gg_nvls_allreduce_avg(void* multicast_ptr, int world, int rank, size_t n_floats, gg_stream_t stream);

/* The same average with the transfer fused into the dW GEMM (no host barrier; only the reduction of locally staged
 * copies and the broadcast of the averages are left for after the GEMM).  Every rank owns, in symmetric memory, a
 * control region of GG_GRAD_CTRL_BYTES (zeroed once, before the first step), its gradient buffer [dW (C,D) | db (C) |
 * pad to a multiple of 4 floats] and a staging region of gg_grad_stage_floats(C, D, world) floats: `world` slabs (one
 * per source rank), each = the dW rows of the blocks this rank reduces (block b of 128 geocells -> rank b % world)
 * followed by their db entries.  gg_head_bwd (dp_ptrs given) pushes every tile into the reducer's slab and counts it
 * on the block's `ready` counter there (ceil(D / 256) column tiles per rank make a block complete); gg_grad_exchange,
 * launched AFTER it on the same stream, waits per owned block for all ranks' tiles, adds the staged copies in rank order (deterministic, identical on all ranks), scales by 1 / world and
 * writes the average into every rank's gradient buffer -- multimem.st through the NVSwitch when the multicast
 * addresses are given, else posted peer stores -- and returns when every block of every reducer has landed in this
 * rank's gradient.  grad_ptrs / ctrl_ptrs: HOST arrays of `world` device addresses (rank order, as mapped on this
 * device); grad_mc / ctrl_mc: multicast addresses of the same buffers or both null; stage: THIS rank's staging
 * region.  world in {1, 2, 4, 8}.
 * flags: GG_GRAD_NO_WAIT = return once this rank's own blocks are exchanged, without waiting for the other reducers'
 * blocks to land (the caller orders the gradient's consumer behind every rank's exchange by other means; used by
 * the single-process emulation of the tests, where the "ranks" share one GPU's launch queues). */
#define GG_GRAD_CTRL_BYTES 16384
#define GG_GRAD_CTRL_BLKCOUNT_OFF 0
#define GG_GRAD_CTRL_READY_OFF 4096
#define GG_GRAD_NO_WAIT 1
size_t gg_grad_ctrl_bytes(void);
size_t gg_grad_stage_floats(int C, int D, int world);
int gg_grad_exchange(const unsigned long long* grad_ptrs, const unsigned long long* ctrl_ptrs, void* grad_mc,
                     void* ctrl_mc, const void* stage, int world, int rank, int C, int D, int flags, gg_stream_t stream);

/* The same exchange with the optimizer in it (sharded AdamW; replaces gg_grad_exchange AND the trainer's
 * torch.optim.AdamW step, main_coordinator_idun_s3.py:286-291,424, for the head's weight and bias).  After adding a
 * block's staged copies the reducer holds the averaged gradient in registers: it applies AdamW --
 *     p -= lr wd p;  m += (1 - b1)(g - m);  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 * (torch's fused kernel, fp32) -- to the rows of ITS blocks of the fp32 master weights master_w (C, D) / master_b (C)
 * with moments m_*, v_* (all local to this rank; rows of other ranks' blocks are not touched and go stale), and
 * writes what the next forward consumes into EVERY rank's operand buffers: the rows as bf16 (w16, (C, D): half the
 * bytes of the fp32 gradient the plain exchange broadcasts) and the bias as fp32 (bias, gg_head_bias_pad(C) entries,
 * pad kept at zero).  No gradient is materialised, no rank runs an optimizer pass over all of W, and the per-step
 * fp32 -> bf16 cast of W (gg_prepare_head_weights) is gone.  hyper: DEVICE floats {lr, beta1, beta2, eps,
 * weight_decay}; step: DEVICE counter of optimizer steps taken so far (t = *step + 1; incremented by the kernel), so a
 * captured CUDA graph replays correctly.  w16_ptrs / bias_ptrs / ctrl_ptrs: host arrays of `world` device addresses
 * in symmetric memory; *_mc: multicast addresses of the same three buffers or all null.  world in {1, 2, 4, 8}
 * (world = 1: gg_head_bwd with dp_world = 1 stages into the local slab; a single-GPU fused optimizer step).
 * A rank returns when every reducer's rows have landed in ITS operands.  Other ranks' operands are only overwritten
 * after that rank announced the block, i.e. after its dW GEMM: anything of the step that still reads w16 (gg_head_dx)
 * must be launched before gg_head_bwd. */
int gg_grad_exchange_adamw(const unsigned long long* w16_ptrs, const unsigned long long* bias_ptrs,
                           const unsigned long long* ctrl_ptrs, void* w16_mc, void* bias_mc, void* ctrl_mc,
                           const void* stage, int world, int rank, int C, int D, float* master_w, float* master_b,
                           float* m_w, float* v_w, float* m_b, float* v_b, const float* hyper, long long* step,
                           int flags, gg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOGUESSR_B200_H */
