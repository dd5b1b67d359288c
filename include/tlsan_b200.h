/* tlsan_b200.h -- C ABI of the B200-native TLSAN train / scoring hot path.
 *
 * The reference (TsingZ0/TLSAN) has no FFI: its boundary is the in-process Python surface
 * `class Model` (TLSAN/model.py:13-313) fed by the 9-tuple of TLSAN/input.py:54,107.  Every
 * entry point below names the reference call it replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (params, batch, workspace, outputs) unless a function says host; nothing is allocated or freed behind
 *     the ABI except the IPC arenas of tlsan_dp_arena_* (explicit create / release);
 *   - every call is asynchronous on the `cudaStream_t` passed as `void* stream`.  A train step also uses two
 *     internal streams per device (the occurrence sort and the table norms run beside the forward kernels, a
 *     pipelined step sorts the next batch behind the backward kernels); they fork from and join `stream` with
 *     events, so the caller sees ordinary stream semantics -- but buffers handed to a step must stay allocated
 *     until `stream` has passed it (and a batch announced as `next` until the step that consumes it);
 *   - return 0 on success, a negative TLSAN_E_* code otherwise; `tlsan_last_error()` returns
 *     a thread-local description; no C++ exception crosses the ABI;
 *   - indices are int32, values fp32; embedding rows are 128 B and must be 16-B aligned.
 * The internal streams / presort registry are per process and guarded by one lock: train-step entry points called from
 * several host threads enqueue one at a time (their kernels still overlap on the callers' streams).  The small
 * attention weights of the CUDA-core variant (TLSAN_FUSED_IMPL=ffma, not the default) are mirrored into one
 * __constant__ bank per process: two models using THAT variant concurrently on one device must be serialised by
 * the caller.
 */
#ifndef TLSAN_B200_H
#define TLSAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLSAN_ABI_VERSION 2   /* 2: tlsan_batch_t.hist_d, TLSAN_STAT_DP_ERR */

/* fixed architecture of the path (TLSAN/train.py:26-35 defaults; hidden_units must equal
 * item+cate embedding size, model.py:100-109) */
#define TLSAN_EMB 32        /* itemid/cateid/userid_embedding_size */
#define TLSAN_HID 64        /* hidden_units */
#define TLSAN_HEADS 8       /* num_heads */
#define TLSAN_DH 8          /* hidden_units / num_heads */
#define TLSAN_MAX_L 96      /* Ls <= 90 in the reference (build_dataset.py:7) */

/* packed small-parameter vector (`dense`), float offsets.  TF variable names in
 * tlsan_b200/model.py:DENSE_LAYOUT.  [k][j] row-major like the TF kernels (in x out). */
#define TLSAN_OFF_W1L 0      /* long  FWA map1 W [8][8]   model.py:380-381,447 */
#define TLSAN_OFF_B1L 64     /* long  FWA map1 bias [8] */
#define TLSAN_OFF_W2L 72     /* long  FWA map2 W [8][8]   model.py:382-383 */
#define TLSAN_OFF_B2L 136
#define TLSAN_OFF_W1S 144    /* short FWA map1 W */
#define TLSAN_OFF_B1S 208
#define TLSAN_OFF_W2S 216
#define TLSAN_OFF_B2S 280
#define TLSAN_OFF_WD 288     /* tf.layers.dense kernel [64][64]  model.py:347 */
#define TLSAN_OFF_BD 4384    /* tf.layers.dense bias [64] */
#define TLSAN_OFF_GAMMA 4448 /* gamma_parameter  model.py:58-60 */
#define TLSAN_DENSE_COUNT 4449
#define TLSAN_DENSE_PAD 4452

enum {
  TLSAN_OK = 0,
  TLSAN_E_DIMS = -1,      /* bad dimension (B<=0, L>TLSAN_MAX_L, ...) */
  TLSAN_E_ALIGN = -2,     /* pointer not 16-B aligned */
  TLSAN_E_NULL = -3,      /* required pointer is NULL */
  TLSAN_E_WORKSPACE = -4, /* workspace too small */
  TLSAN_E_CUDA = -5,      /* launch / runtime failure (cudaGetLastError) */
  TLSAN_E_UNSUPPORTED = -6
};

typedef struct {
  int32_t B;        /* rows in this (local) batch */
  int32_t L;        /* Ls: width of hist_i / hist_t and of usert_emb   (train.py:29) */
  int32_t S;        /* width of hist_i_new = max short length in batch (input.py:33,37) */
  int32_t NI, NU, NC;
  int32_t B_global; /* denominator of reduce_mean (model.py:171); = B on one GPU */
  int32_t reserved; /* flags; bit 0: tlsan_step_grads skips the table norms (row-sharded callers compute their own); bit 1: this batch was presorted into this workspace by the previous *_pipelined call */
} tlsan_dims_t;

/* Trainable state, model.py:56-81.  `emb` is ONE table: rows [0,NI) = item_emb,
 * [NI,NI+NC) = cate_emb, [NI+NC,NI+NC+NU) = user_emb (32 floats each). */
typedef struct {
  float* emb;
  float* usert;              /* usert_emb [NU][L] */
  float* item_b;             /* [NI] */
  float* dense;              /* [TLSAN_DENSE_PAD] */
  const int32_t* icl;        /* item_cate_list [NI]  (model.py:85) */
  const int32_t* cate_off;   /* [NC+1] CSR of items grouped by category (stable order) */
  const int32_t* cate_items; /* [NI] */
} tlsan_params_t;

/* The 9-tuple of TLSAN/input.py:54 / :107 after the int64->int32 feed cast
 * (model.py:210-222).  `i2` is batch[2] of the test tuple (negative item) or NULL. */
typedef struct {
  const int32_t* u;          /* batch[0] */
  const int32_t* i;          /* batch[1] */
  const int32_t* i2;         /* batch[2] as item (eval_auc) */
  const float* y;            /* batch[2] as label (train) */
  const int32_t* hist_i;     /* batch[3]  [B][L] */
  const int32_t* hist_i_new; /* batch[4]  [B][S] */
  const float* hist_t;       /* batch[5]  [B][L] */
  const int32_t* sl;         /* batch[6] */
  const int32_t* sl_new;     /* batch[7] */
  const int32_t* c;          /* batch[8]  u_cate */
  const int32_t* hist_d;     /* optional (may be NULL): RAW day gaps d[B][L] = cur_day - day + 1, 0 = padding.  When set,
                              * the long-term kernels bucket while they gather -- n = sum_j [d >= 2^j] = min(12,
                              * floor(log2 d)), weight float32(1/n), exactly proc_time_emb (build_dataset.py:16-21) +
                              * the float32 store of input.py:36,45 -- and hist_t is ignored (it may alias hist_d). */
} tlsan_batch_t;

/* scalars a train step leaves on the device (index into `stats`) */
enum {
  TLSAN_STAT_LOSS = 0,     /* self.loss, model.py:171-172 */
  TLSAN_STAT_BCE = 1,      /* reduce_mean(sigmoid_cross_entropy) */
  TLSAN_STAT_NORM = 2,     /* global norm used by clip_by_global_norm, model.py:201 */
  TLSAN_STAT_SCALE = 3,    /* clip multiplier */
  TLSAN_STAT_L2 = 4,       /* l2_norm, model.py:164-169 */
  TLSAN_STAT_DP_ERR = 5,   /* tlsan_dp_exchange: 1 = a wait on a peer rank timed out; that step's update was SKIPPED */
  TLSAN_STAT_COUNT = 8
};

int tlsan_abi_version(void);
const char* tlsan_last_error(void);

/* K2: n = sum_j [d >= 2^j], j=1..12, out = float32(1/n); d = 0 -> 0 (pad).
 * Replaces proc_time_emb, TLSAN/build_dataset.py:16-21 (+ cast at input.py:36,45).
 * `lut13` = device float[13], lut[n] = float32(1/n) built by the host. */
int tlsan_time_bucket(const int32_t* d, const float* lut13, float* out, int32_t* bucket_out,
                      int64_t n, void* stream);

/* K1: out[r] = [ item_emb[idx[r]] || cate_emb[icl[idx[r]]] ] * (tau ? tau[r] : 1).
 * Replaces the embedding_lookup/gather/concat/multiply group, model.py:84-86,105-113. */
int tlsan_gather_concat(const tlsan_dims_t* dims, const tlsan_params_t* p, const int32_t* idx,
                        const float* tau, float* out, int64_t n, void* stream);

/* Forward only: logits for 1 (i) or 2 (i, i2) candidates per row.
 * Replaces sess.run(self.logits, ...) in Model.eval_auc, model.py:237-263.
 * logits: [B][ncand]; ut (optional): [B][64] = u_t of model.py:135. */
int tlsan_score(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                int32_t ncand, float* logits, float* ut, void* stream);

/* Same logits as tlsan_score, faster for large batches: with a scratch workspace the 64x64 dense layer
 * (model.py:347) leaves the per-sample kernel and runs as ONE batched tensor-core GEMM between the
 * long-term and the short-term kernel. */
int tlsan_score_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes);
int tlsan_score_ws(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, int32_t ncand,
                   float* logits, float* ut, void* workspace, size_t workspace_bytes, void* stream);

/* bytes of workspace tlsan_train_step / tlsan_step_grads need for `dims`. */
int tlsan_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes);

/* One Model.train call (model.py:208-234): forward, loss, tf.gradients,
 * clip_by_global_norm, GradientDescentOptimizer.apply_gradients, all on `stream`.
 * `stats` = device float[TLSAN_STAT_COUNT]. */
int tlsan_train_step(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                     float lr, float reg, float clip_norm, void* workspace, size_t workspace_bytes,
                     float* stats, void* stream);

/* Data-parallel split of the same step.
 * tlsan_step_grads: forward + backward on the local rows and the local deterministic
 *   segmented reduce into ONE flat fp32 buffer `flat` of tlsan_flat_count() floats
 *   (sparse-part table gradients, dense gradients, sum-of-squares and loss partials),
 *   which the caller all-reduces (sum) across ranks;
 * tlsan_apply_flat: L2 term + clip + SGD from the reduced buffer (identical on every rank).  It must follow
 *   tlsan_step_grads on the SAME workspace with the weights unchanged in between: the sums of squares of the
 *   tables (l2_loss, model.py:164-169) are computed by tlsan_step_grads beside its forward kernels and read
 *   back from the workspace here. */
int tlsan_flat_count(const tlsan_dims_t* dims, int64_t* count);
int tlsan_step_grads(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                     void* workspace, size_t workspace_bytes, float* flat, void* stream);
int tlsan_apply_flat(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat,
                     float lr, float reg, float clip_norm, void* workspace, size_t workspace_bytes,
                     float* stats, void* stream);

/* The other optimisers of init_optimizer (model.py:188-193: adadelta / adam / rmsprop with the TF-1.8 defaults;
 * GradientDescentOptimizer is the `else` branch and the path above).  tlsan_apply_flat_opt = tlsan_apply_flat with the
 * update rule of `opt`: every element sees its aggregated, clipped gradient (sum of slices + reg * w) * scale, so TF's
 * sparse apply ops reduce to the dense formulas (the L2 term puts every row into the IndexedSlices).  slot1 / slot2
 * mirror the weight buffer, which must be ONE buffer emb | usert | item_b | dense (each padded to 16 B):
 *   adam      slot1 = m, slot2 = v (zeros);  rmsprop  slot1 = ms (ONES), slot2 = momentum (zeros);
 *   adadelta  slot1 = accum, slot2 = accum_update (zeros).  step = 1, 2, ... (adam's bias correction). */
enum { TLSAN_OPT_SGD = 0, TLSAN_OPT_ADAM = 1, TLSAN_OPT_RMSPROP = 2, TLSAN_OPT_ADADELTA = 3 };
typedef struct {
  int32_t kind;
  int32_t step;
  float beta1, beta2;   /* adam: 0.9, 0.999 */
  float rho;            /* rmsprop decay 0.9 ; adadelta rho 0.95 */
  float momentum;       /* rmsprop: 0.0 */
  float epsilon;        /* adam 1e-8, rmsprop 1e-10, adadelta 1e-8 */
  float* slot1;
  float* slot2;
} tlsan_opt_t;
int tlsan_apply_flat_opt(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat, float lr, float reg,
                         float clip_norm, const tlsan_opt_t* opt, void* workspace, size_t workspace_bytes, float* stats,
                         void* stream);

/* Data-parallel exchange over NVLink peer memory, in place of NCCL all-reduce + tlsan_apply_flat: a
 * fused reduce-scatter + optimiser step + all-gather.  Every rank owns an ARENA that its peers map through CUDA
 * IPC and passes it to tlsan_step_grads as the `flat` gradient buffer; rank r then sums slice r of the weight index
 * space out of all arenas (peer memory, fixed rank order: deterministic, identical on every rank), applies
 * L2 + clip + SGD to that slice, publishes it, and every rank pulls the other slices; ranks meet at three
 * release / acquire flags in the arenas (dense-gradient / norm partials ready -- that small exchange runs early, on
 * the library's side stream beside the segmented row reduce --, gradient rows ready, updated slice published).
 * The tables must be ONE device buffer: emb [(NI+NC+NU)*32] |
 * usert [NU*L, padded to 4] | item_b [NI, padded to 4], with params->usert / item_b pointing into it.
 *   tlsan_dp_arena_bytes / _create / _open / _release : arena size for (dims, world); cudaMalloc + IPC handle
 *       (64 bytes, to be exchanged by the host, e.g. torch.distributed.all_gather_object); map a peer's arena.
 *   tlsan_dp_exchange : arenas[world] = the mapped arenas in rank order; arenas[rank] = the own one = the buffer
 *       tlsan_step_grads just wrote; epoch = 1, 2, ... (one more every call, the same on every rank).
 *       A rank waits for its peers with a bounded spin (TLSAN_DP_TIMEOUT_S seconds, default 30): on a timeout the
 *       kernels of that step leave the weights untouched and stats[TLSAN_STAT_DP_ERR] is set (sticky) -- the
 *       caller must poll it (tlsan_b200.Model does, every dp_check_every steps) and stop: replicas have diverged. */
int tlsan_dp_arena_bytes(const tlsan_dims_t* dims, int32_t world, size_t* bytes);
int tlsan_dp_arena_create(size_t bytes, void** ptr, char* handle64);
int tlsan_dp_arena_open(const char* handle64, void** ptr);
int tlsan_dp_arena_release(void* ptr, int32_t owned);
int tlsan_dp_exchange(const tlsan_dims_t* dims, const tlsan_params_t* p, void* const* arenas, int32_t rank,
                      int32_t world, int32_t epoch, float lr, float reg, float clip_norm, void* workspace,
                      size_t workspace_bytes, float* stats, void* stream);

/* Pipelined variants: `next` (optional) names the batch of the FOLLOWING step and that step's own workspace.  Its
 * occurrence sort is enqueued behind the backward kernels of this step, where it runs beside the reduce, the
 * all-reduce and the update; the following call passes the same batch / workspace with dims->reserved bit 1 set
 * and starts without a sort.  If no valid presort exists for that workspace (side streams disabled with
 * TLSAN_SORT_OVERLAP=0, or the announcement was dropped) the step simply sorts in place.  Results are identical
 * to the plain entry points. */
typedef struct {
  const tlsan_dims_t* dims;
  const tlsan_batch_t* batch;
  void* workspace;
  size_t workspace_bytes;
  void* ready_event;   /* optional cudaEvent_t: the next batch's buffer is complete once it fires (its H2D copy on another
                        * stream); the presort waits for it, the current step does not.  NULL: the batch is ready */
} tlsan_next_t;
int tlsan_train_step_pipelined(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                               const tlsan_next_t* next, float lr, float reg, float clip_norm, void* workspace,
                               size_t workspace_bytes, float* stats, void* stream);
int tlsan_step_grads_pipelined(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                               const tlsan_next_t* next, void* workspace, size_t workspace_bytes, float* flat,
                               void* stream);

/* Full-catalogue ranking for Model.eval_prec / eval_recall (model.py:140-156,265-299):
 * rank[b] = #items scored above label[b] under u_t . all_emb^T + item_b with top_k tie order. */
int tlsan_label_rank(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* ut,
                     const int32_t* label, int32_t* rank, void* stream);

/* The same ranks on the tcgen05 tensor cores: a [B,64]x[64,NI] 3xTF32 GEMM whose accumulator tiles
 * stay in TMEM and are only counted, never stored (csrc/tlsan_rank_tc.cu).  The workspace holds the catalogue
 * re-laid out as UMMA operand tiles (73 728 B per 128 items), rebuilt on every call from the current weights. */
int tlsan_rank_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes);
int tlsan_label_rank_ws(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* ut, const int32_t* label,
                        int32_t* rank, void* workspace, size_t workspace_bytes, void* stream);

/* The same count against ONE row shard of the item tables (row-sharded configuration, SURVEY 8e last row): local
 * row j of the shard is global item j * gid_mul + gid_add; label_global[B] = global label ids, lab_rows[B][68] = the
 * labels' augmented rows [item_emb | cate_emb[icl] | item_b | 3 pad] (gathered by the rank that fetched them), so
 * every shard derives the label's score with the same instruction sequence and ties still break by global index.
 * rank_partial[b] = this shard's share; the caller sums over shards (integer all-reduce).
 * Workspace: 73 728 B per 128 shard rows + 256. */
int tlsan_label_rank_shard(int32_t B, int64_t n_local, const float* item_emb_shard, const float* item_b_shard,
                           const int32_t* icl_shard, const float* cate_emb, const float* ut,
                           const int32_t* label_global, const float* lab_rows, int32_t gid_mul, int32_t gid_add,
                           int32_t* rank_partial, void* workspace, size_t workspace_bytes, void* stream);

/* Host helper (no GPU work): pack the 9-tuple of TLSAN/input.py:54,107 (int64 ids, fp32 hist_t) into
 * ONE int32 staging buffer -- the int64->int32 feed cast of model.py:210-222 -- multi-threaded, with
 * the id range checks TF's CPU gather performs (nthreads = 0: min(8, host cores / LOCAL_WORLD_SIZE), so the ranks of a
 * node share the cores).  Segment order (each rounded up to 4 words):
 * u, i, second (i2 as int32, or y as fp32 bits), c, sl, sl_new, hist_i[B*L], hist_i_new[B*S],
 * hist_t[B*L].  Returns TLSAN_E_DIMS and names the field in tlsan_last_error() if an id is out of range. */
int tlsan_pack_batch_host(const tlsan_dims_t* dims, const int64_t* u, const int64_t* i, const int64_t* i2,
                          const float* y, const int64_t* hist_i, const int64_t* hist_i_new, const float* hist_t,
                          const int64_t* sl, const int64_t* sl_new, const int64_t* c, int32_t* out,
                          int64_t out_words, int32_t validate, int32_t nthreads);

/* The feed of Model.train / eval_auc (model.py:210-222,239-261) end to end: pack into `pinned` (page-locked host
 * memory) with a persistent thread pool, copy to `dev` on `stream` phase by phase while the rest is still being
 * packed, and leave the packed batch layout of tlsan_pack_batch_host at the front of `dev`.  The session matrix
 * hist_i_new [B][S] is almost all padding, so it travels ragged (offsets + valid items, in a tail region behind the
 * packed layout) and k_expand_sessions rebuilds the zero-padded matrix in HBM.  `pinned` and `dev` must hold
 * tlsan_stage_words(dims) int32 words; `pinned` may be reused once `stream` has passed the copies. */
int tlsan_stage_words(const tlsan_dims_t* dims, int64_t* words);
int tlsan_stage_batch_host(const tlsan_dims_t* dims, const int64_t* u, const int64_t* i, const int64_t* i2,
                           const float* y, const int64_t* hist_i, const int64_t* hist_i_new, const float* hist_t,
                           const int64_t* sl, const int64_t* sl_new, const int64_t* c, int32_t* pinned, int32_t* dev,
                           int64_t words, int32_t validate, int32_t nthreads, void* stream);

/* The packed feed: a batch that ALREADY sits in page-locked host memory in the staging layout of tlsan_stage_batch_host
 * (tlsan_b200/input.py, DataInput(..., packed=True), writes its arrays straight into such a buffer; tlsan_stage_batch_host
 * with dev == NULL packs a 9-tuple into one without copying).  Two host->device copies + the session expansion, no host
 * pass; n_new = number of valid session items (sum of sl_new clipped to S).  Ids are NOT range checked here: the
 * producer vouches for them (PackedBatch carries the id maxima of its dataset, Model compares them with its tables). */
int tlsan_stage_packed(const tlsan_dims_t* dims, const int32_t* pinned, int32_t* dev, int64_t words, int64_t n_new,
                       void* stream);

/* The same two helpers for batches whose integer fields are already int32 (no narrowing pass; tlsan_b200/input.py
 * emits such batches): half the host memory traffic of the feed, same range checks. */
int tlsan_pack_batch_host_i32(const tlsan_dims_t* dims, const int32_t* u, const int32_t* i, const int32_t* i2,
                              const float* y, const int32_t* hist_i, const int32_t* hist_i_new, const float* hist_t,
                              const int32_t* sl, const int32_t* sl_new, const int32_t* c, int32_t* out,
                              int64_t out_words, int32_t validate, int32_t nthreads);
int tlsan_stage_batch_host_i32(const tlsan_dims_t* dims, const int32_t* u, const int32_t* i, const int32_t* i2,
                               const float* y, const int32_t* hist_i, const int32_t* hist_i_new, const float* hist_t,
                               const int32_t* sl, const int32_t* sl_new, const int32_t* c, int32_t* pinned,
                               int32_t* dev, int64_t words, int32_t validate, int32_t nthreads, void* stream);

/* Device-resident dataset.  CSR image, in HBM, of the samples built by
 * TLSAN/build_dataset.py:58-59,71: per sample r its user, its long-term history
 * pre_items / pre_time [pre_off[r], pre_off[r+1]) (pre_time = float32(1/n), input.py:36,45), its session
 * new_items [new_off[r], new_off[r+1]), the candidate, the label (train: second_f) or negative item
 * (test: second_i) and u_cate. */
typedef struct {
  const int32_t* uid;
  const int64_t* pre_off;
  const int32_t* pre_items;
  const float* pre_time;
  const int64_t* new_off;
  const int32_t* new_items;
  const int32_t* cand;
  const int32_t* second_i;
  const float* second_f;
  const int32_t* ucate;
  int64_t n;
} tlsan_dataset_t;

/* Batch assembly on the GPU: rows idx[0..B) (device int32) in the layout of DataInput.__next__ /
 * DataInputTest.__next__ (TLSAN/input.py:17-54,70-107) -- last Ls history entries left-aligned and zero
 * padded, session zero padded to S columns -- written as the packed staging buffer of
 * tlsan_pack_batch_host.  S must be >= the longest session among the rows (input.py uses the batch max). */
int tlsan_collate(const tlsan_dataset_t* ds, const int32_t* idx, int32_t B, int32_t L, int32_t S, int32_t is_test,
                  int32_t* out, int64_t out_words, void* stream);

/* Dataset builder on the GPU (TLSAN/build_dataset.py:25-73).  Input: the review table sorted by user then
 * time (utils/2_remap_id.py:91) as three int32 columns, user_off[n_users + 1] = row range of every user.
 *   tlsan_ds_plan     per user u: counts4[u] = {train PAIRS, has a test sample, sum of their history lengths, sum of
 *                     their session lengths}, test2[u] = {first row (relative) and size of the test session}: session
 *                     segmentation (runs of equal day, :38-46) + the split rule i + count < min(len, 90) - 1 (:55-72)
 *   tlsan_ds_lengths  history / session length of every sample, unshuffled order (first_train[u] = index of user u's
 *                     first train sample = 2 * exclusive prefix of the pairs; first_test likewise)
 *   tlsan_ds_emit     writes every sample at its FINAL position (pos_train / pos_test: position after the reference's
 *                     random.shuffle) into two tlsan_dataset_t images whose offset arrays the caller has filled:
 *                     history, session, target, label / negative, u_cate (most frequent category, ties -> first seen,
 *                     :54) and the weights float32(1/n), n = sum_j [d >= 2^j] (:16-21; lut13[n]); train_gap / test_gap
 *                     (optional) receive the raw day gaps d.  neg[row] = sampled negative of every review row (:28-33),
 *                     pick[u] = index chosen by random.choice inside the test session (:66): the host draws both from
 *                     the Python random stream in the reference's call order. */
int tlsan_ds_plan(const int32_t* day, const int64_t* user_off, int32_t n_users, int32_t* counts4, int32_t* test2,
                  void* stream);
int tlsan_ds_lengths(const int32_t* day, const int64_t* user_off, int32_t n_users, const int64_t* first_train,
                     const int64_t* first_test, int32_t* len_pre_train, int32_t* len_new_train, int32_t* len_pre_test,
                     int32_t* len_new_test, void* stream);
int tlsan_ds_emit(const int32_t* reviewer, const int32_t* asin, const int32_t* day, const int32_t* item_cate,
                  const int64_t* user_off, int32_t n_users, const int64_t* first_train, const int64_t* first_test,
                  const int64_t* pos_train, const int64_t* pos_test, const int32_t* neg, const int32_t* pick,
                  const float* lut13, const tlsan_dataset_t* train, const tlsan_dataset_t* test, int32_t* train_gap,
                  int32_t* test_gap, void* stream);

/* Row-sharded item tables (SURVEY 8e / BASELINE config 5, NI = 10 M; no counterpart in the reference, whose
 * tables live in one TF process).  item_emb / item_b / icl are split by row over the ranks.  Per step a rank
 * asks the owners for the rows its batch touches (NCCL all-to-all of ids, then of rows), runs the unchanged
 * tlsan_step_grads on a COMPACT table (row r = r-th distinct id of its batch), sends the per-id gradient rows
 * back, and every owner applies L2 + clip + SGD to its shard.  Orchestration: tlsan_b200/sharded.py.
 * Exchange row = TLSAN_SHARD_ROW words: 32 floats item_emb row | item_b | icl (int bits) | 2 pad; gradient
 * rows: 32 floats d item_emb | d item_b | 3 pad.
 *   tlsan_shard_pack_rows     owner: out[k] = row local_ids[k] of its shard; *bad_flag = 1 on a foreign id
 *   tlsan_shard_unpack_rows   requester: compact table row dst_index[k] <- packed row k
 *   tlsan_shard_pack_grads    requester: out[k] = reduced gradient (item half + item_b) of compact row src_index[k]
 *   tlsan_shard_accum_grads   owner: g_emb/g_b[local_ids[k]] += packed[k]; ids of one call are distinct
 *   tlsan_reduce_cate         out[NC][32] = category gradient from the flat buffer of tlsan_step_grads
 *   tlsan_sumsq               partial[c] = sum of squares of a grid-strided slice of W (fixed order)
 *   tlsan_sgd_dense           W <- W - lr*((g + reg*W) * *scale) element-wise (g may be NULL: pure L2 decay)
 *   tlsan_shard_apply_replicated  norm / clip scale / loss statistics and the update of cate_emb, user_emb,
 *                             usert_emb and the small parameters from rank-summed gradients; item_sumsq =
 *                             rank-summed partial sums of squares of the whole sharded item_emb.
 *   tlsan_shard_apply_grads   owner: W[local_ids[k]] -= lr * *scale * packed[k] (ids < 0 = padding, skipped); the L2
 *                             decay of every shard row is a separate dense pass (tlsan_sgd_dense with g = NULL)
 *   tlsan_route_ids           device-side routing of the distinct item ids of a batch, no host round trip: position
 *                             p = owner * ceil(NI / world) + local of every id (cyclic: owner = id % world; block: p = id),
 *                             presence bitmap over the positions, compact row of an id = its rank among all requested
 *                             ids (owner-major order).  phase 0: bitmap + per-word popcounts into word_prefix (the
 *                             caller turns them into an INCLUSIVE prefix, e.g. torch.cumsum); phase 1:
 *                             send_ids[world][cap] = owner-local row ids requested from every owner (ascending, padded
 *                             with -1), slot_row[world][cap] = compact row of every request slot (-1 padding; the
 *                             dst_index / src_index of the pack / unpack helpers), counts[world], *overflow = 1 if a
 *                             group exceeds cap, and dst[q][e] = compact row of src[q][e] for the nfields id arrays of
 *                             the packed batch.  With fixed cap the all-to-alls run with equal splits; negative ids /
 *                             rows mark padding slots (zero rows, skipped by unpack / accum / apply).
 *   tlsan_route_bitmap_words  words of the presence bitmap (= of word_prefix) for (NI, world) */
#define TLSAN_SHARD_ROW 36
int tlsan_shard_apply_grads(const float* packed, const int32_t* local_ids, int64_t n, float* W_emb, float* W_b,
                            float lr, const float* scale, void* stream);
int tlsan_route_bitmap_words(int64_t NI, int32_t world, int64_t* words);
int tlsan_route_ids(const int32_t* const* src, int32_t* const* dst, const int64_t* n, int32_t nfields, int64_t NI,
                    int32_t world, int32_t cyclic, int32_t cap, uint32_t* bitmap, int32_t* word_prefix,
                    int32_t* send_ids, int32_t* slot_row, int32_t* counts, int32_t* overflow, int32_t phase,
                    void* stream);
int tlsan_shard_pack_rows(const float* emb_shard, const float* item_b_shard, const int32_t* icl_shard,
                          const int32_t* local_ids, int64_t n, int64_t n_local, float* out, int32_t* bad_flag,
                          void* stream);
int tlsan_shard_unpack_rows(const float* packed, const int32_t* dst_index, int64_t n, float* emb_c, float* item_b_c,
                            int32_t* icl_c, void* stream);
int tlsan_shard_pack_grads(const float* g_i, const float* g_b, const int32_t* src_index, int64_t n, float* out,
                           void* stream);
int tlsan_shard_accum_grads(const float* packed, const int32_t* local_ids, int64_t n, float* g_emb, float* g_b,
                            void* stream);
int tlsan_reduce_cate(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat, float* out, void* stream);
int tlsan_sgd_dense(float* W, const float* g, int64_t n, float lr, float reg, const float* scale, void* stream);
int tlsan_sumsq(const float* W, int64_t n, float* partial, int32_t npartial, void* stream);
int tlsan_shard_apply_replicated(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* gcate,
                                 const float* g_u, const float* dgrad, const float* item_sumsq,
                                 int32_t n_item_sumsq, float lr, float reg, float clip_norm, void* workspace,
                                 size_t workspace_bytes, float* stats, void* stream);

/* Instrumentation for bench.py (not on the product path).
 * tlsan_launch_count: kernels launched by this library since load (all threads).
 * tlsan_profile_begin(max_steps): from now on every tlsan_step_grads / tlsan_apply_flat records
 *   CUDA events on its stream at phase boundaries; tlsan_profile_end waits for them and fills
 *   ms[step][TLSAN_PHASE_COUNT] (elapsed per phase), returning the number of steps recorded. */
enum {
  TLSAN_PHASE_SORT = 0,      /* radix sort of the occurrence keys + segment bounds */
  TLSAN_PHASE_LONG_FWD = 1,  /* long-term FWA forward */
  TLSAN_PHASE_DENSE_FWD = 2, /* z = o_long Wd + bd (batched GEMM) */
  TLSAN_PHASE_SHORT = 3,     /* short-term FWA forward + logit + loss + its backward */
  TLSAN_PHASE_DENSE_BWD = 4, /* d o_long, dWd, dbd (batched GEMM) */
  TLSAN_PHASE_BWD_LONG = 5,  /* long-term FWA backward */
  TLSAN_PHASE_REDUCE = 6,    /* k_finalize1 + k_row_reduce */
  TLSAN_PHASE_APPLY = 7,     /* table sumsq, finalize2, row / cate updates */
  TLSAN_PHASE_COUNT = 8
};
long long tlsan_launch_count(void);
int tlsan_profile_begin(int32_t max_steps);
int tlsan_profile_end(float* ms, int32_t* steps);

#ifdef __cplusplus
}
#endif
#endif /* TLSAN_B200_H */
