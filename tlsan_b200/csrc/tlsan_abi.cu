// extern "C" entry points of include/tlsan_b200.h: argument validation + launch sequencing.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "tlsan_common.cuh"

static thread_local char g_err[512] = "";

void tlsan_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

long long g_tlsan_launches = 0;

// ---- optional phase timing (bench.py): events recorded on the caller's stream
static cudaEvent_t* g_ev = nullptr;      // [max_steps][TLSAN_PHASE_COUNT + 1]
static int g_ev_steps_cap = 0, g_ev_step = -1;
static bool g_prof_on = false;

void tlsan_profile_mark(int phase_done, cudaStream_t st) {
  if (!g_prof_on) return;
  if (phase_done < 0) {
    if (g_ev_step + 1 >= g_ev_steps_cap) { g_prof_on = false; return; }
    ++g_ev_step;
  }
  if (g_ev_step < 0) return;
  cudaEventRecord(g_ev[(size_t)g_ev_step * (TLSAN_PHASE_COUNT + 1) + (phase_done + 1)], st);
}

// All fused variants are sm_100a CUDA in this library.  TLSAN_FUSED_IMPL selects an older
// formulation for A/B measurements (the design asks for mma-vs-FFMA evidence):
//   ffma  = CUDA-core maps, 4 samples per warp      (tlsan_fwd_bwd.cu)
//   mma   = 3xTF32 mma tiles, synchronous gathers   (tlsan_fused_mma.cu)
//   async = mma tiles + cp.async sample pipeline     (tlsan_fused_async.cu)
//   hybrid = per kernel the faster of mma / async     (the round-1 default)
//   (unset) pf = long-term kernels with a metadata pre-pass + in-warp prefetch pipeline (tlsan_fused_pf.cu),
//           cp.async pipeline for the short-term kernel                                              [default]
static int fused_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TLSAN_FUSED_IMPL");
    v = (e && strcmp(e, "ffma") == 0) ? 0 : (e && strcmp(e, "mma") == 0) ? 1 : (e && strcmp(e, "async") == 0) ? 2
        : (e && strcmp(e, "hybrid") == 0) ? 3 : 4;
  }
  return v;
}
static bool use_mma() { return fused_impl() >= 1; }

// The radix sort of the occurrence keys only feeds the kernels that WRITE gradient rows (short-term kernel
// onwards), so it runs on a side stream beside the long-term forward and the dense GEMM (fork / join with
// events; works under stream capture too).  TLSAN_SORT_OVERLAP=0 keeps everything on the caller's stream.
struct SideStream { cudaStream_t st = nullptr, st2 = nullptr; cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, part = nullptr, dpx = nullptr; };
static SideStream* side_stream() {
  static SideStream per_dev[64];
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TLSAN_SORT_OVERLAP");
    enabled = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (!s.st) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s.st, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&s.st2, cudaStreamNonBlocking, lo) != cudaSuccess ||   // presort: fills gaps
        cudaEventCreateWithFlags(&s.fork2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.part, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.dpx, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      s.st = nullptr;
      return nullptr;
    }
  }
  return &s;
}
static bool g_prof_overlap = false;   // the recorded steps ran the sort on the side stream
static std::recursive_mutex g_api_mutex;   // guards the process-global state of the stateful entry points

// Pipelined steps: the occurrence sort of the NEXT batch is enqueued behind the backward kernels of the current
// one, into the next step's workspace; that step (dims->reserved bit 1) waits for the event instead of sorting.
// ev: ranks (inv) ready -- what the gradient-row writers wait for; ev_seg: segment bounds ready too (row reduce; the
// last kernel of the sort); ev_part: balanced partition of the short-term / backward kernels ready (end of the chain)
struct Presort { char* ws = nullptr; cudaEvent_t ev = nullptr, ev_seg = nullptr, ev_part = nullptr; bool valid = false; };
static Presort g_presort[32];
// `avoid`: the entry the calling step is consuming -- its events are still to be waited on by kernels that step
// enqueues later (row reduce -> ev_seg), so it must not be handed to another workspace and re-recorded
static Presort* presort_slot(char* ws, bool create, const Presort* avoid = nullptr) {
  for (auto& e : g_presort) if (e.ws == ws) return &e;
  if (!create) return nullptr;
  for (auto& e : g_presort)
    if (!e.valid && &e != avoid) {
      if (!e.ev && (cudaEventCreateWithFlags(&e.ev, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&e.ev_seg, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&e.ev_part, cudaEventDisableTiming) != cudaSuccess))
        return nullptr;
      e.ws = ws;
      return &e;
    }
  // every slot holds an announced-but-never-consumed presort (callers that dropped their model): one whose
  // kernels have finished can no longer race with anything and may be recycled
  for (auto& e : g_presort)
    if (&e != avoid && cudaEventQuery(e.ev_seg) == cudaSuccess && cudaEventQuery(e.ev_part) == cudaSuccess) {
      e.ws = ws;
      e.valid = false;
      return &e;
    }
  return nullptr;
}

#define REQUIRE(cond, code, ...)      \
  do {                                \
    if (!(cond)) {                    \
      tlsan_set_error(__VA_ARGS__);   \
      return code;                    \
    }                                 \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int check_dims(const tlsan_dims_t* d) {
  REQUIRE(d != nullptr, TLSAN_E_NULL, "dims is NULL");
  REQUIRE(d->B > 0, TLSAN_E_DIMS, "B must be > 0 (got %d)", d->B);
  REQUIRE(d->L >= 1 && d->L <= TLSAN_MAX_L, TLSAN_E_DIMS, "L must be in [1,%d] (got %d)", TLSAN_MAX_L, d->L);
  REQUIRE(d->S >= 1, TLSAN_E_DIMS, "S must be >= 1 (got %d)", d->S);
  REQUIRE(d->NI > 0 && d->NU > 0 && d->NC > 0, TLSAN_E_DIMS, "table sizes must be > 0");
  REQUIRE((long long)d->NI + d->NC + d->NU < (1ll << 30), TLSAN_E_DIMS, "row space too large");
  REQUIRE((long long)d->B * 2 * (d->L + d->S + 3) < (1ll << 31) - 1, TLSAN_E_DIMS, "B*(L+S+3) overflows int32");
  return TLSAN_OK;
}

static int check_params(const tlsan_params_t* p, bool train) {
  REQUIRE(p != nullptr, TLSAN_E_NULL, "params is NULL");
  REQUIRE(p->emb && p->usert && p->item_b && p->dense && p->icl, TLSAN_E_NULL, "params has a NULL table");
  REQUIRE(aligned16(p->emb) && aligned16(p->dense), TLSAN_E_ALIGN, "emb/dense must be 16-B aligned");
  if (train) REQUIRE(p->cate_off && p->cate_items, TLSAN_E_NULL, "cate_off/cate_items required for training");
  return TLSAN_OK;
}

static int check_batch(const tlsan_batch_t* b, bool train, int ncand) {
  REQUIRE(b != nullptr, TLSAN_E_NULL, "batch is NULL");
  REQUIRE(b->u && b->i && b->c && b->sl && b->sl_new && b->hist_i && b->hist_i_new && b->hist_t, TLSAN_E_NULL,
          "batch has a NULL field");
  if (train) REQUIRE(b->y != nullptr, TLSAN_E_NULL, "batch.y (labels) is NULL");
  if (ncand > 1) REQUIRE(b->i2 != nullptr, TLSAN_E_NULL, "batch.i2 is NULL but ncand == 2");
  REQUIRE(b->hist_d == nullptr || fused_impl() == 1 || fused_impl() >= 3, TLSAN_E_UNSUPPORTED,
          "raw day gaps (batch.hist_d) need the mma / hybrid / pf kernels (TLSAN_FUSED_IMPL)");
  return TLSAN_OK;
}

int tlsan_pdl_level() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TLSAN_PDL"); v = e ? atoi(e) : 1; }
  return v;
}

extern "C" {

int tlsan_abi_version(void) { return TLSAN_ABI_VERSION; }
const char* tlsan_last_error(void) { return g_err; }

int tlsan_time_bucket(const int32_t* d, const float* lut13, float* out, int32_t* bucket_out, int64_t n,
                      void* stream) {
  REQUIRE(n >= 0, TLSAN_E_DIMS, "n < 0");
  REQUIRE(d && lut13 && (out || bucket_out), TLSAN_E_NULL, "NULL argument");
  return tlsan_launch_bucket(d, lut13, out, bucket_out, n, (cudaStream_t)stream);
}

int tlsan_gather_concat(const tlsan_dims_t* dims, const tlsan_params_t* p, const int32_t* idx, const float* tau,
                        float* out, int64_t n, void* stream) {
  REQUIRE(dims && p && p->emb && p->icl, TLSAN_E_NULL, "NULL argument");
  REQUIRE(n >= 0, TLSAN_E_DIMS, "n < 0");
  if (n == 0) return TLSAN_OK;
  REQUIRE(idx && out, TLSAN_E_NULL, "idx/out is NULL");
  REQUIRE(aligned16(p->emb) && aligned16(out), TLSAN_E_ALIGN, "emb/out must be 16-B aligned");
  return tlsan_launch_gather(*dims, *p, idx, tau, out, n, (cudaStream_t)stream);
}

int tlsan_score(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, int32_t ncand,
                float* logits, float* ut, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, false))) return rc;
  REQUIRE(ncand == 1 || ncand == 2, TLSAN_E_DIMS, "ncand must be 1 or 2 (got %d)", ncand);
  if ((rc = check_batch(b, false, ncand))) return rc;
  REQUIRE(logits != nullptr, TLSAN_E_NULL, "logits is NULL");
  REQUIRE(ut == nullptr || aligned16(ut), TLSAN_E_ALIGN, "ut must be 16-B aligned");
  if (use_mma()) return tlsan_launch_score_mma(*dims, *p, *b, ncand, logits, ut, (cudaStream_t)stream);
  return tlsan_launch_score(*dims, *p, *b, ncand, logits, ut, (cudaStream_t)stream);
}

// scratch [B][TLSAN_SCR][64] + per-token metadata [B][L] x 16 B + the balanced partition, each 256-B aligned
static size_t score_ws_bytes(const tlsan_dims_t* d) {
  return tlsan_align_up((size_t)d->B * TLSAN_SCR * 64 * sizeof(float), 256) +
         tlsan_align_up((size_t)d->B * d->L * 16, 256) + tlsan_align_up(tlsan_partition_bytes(), 256) +
         tlsan_score_meta_bytes(d->B, d->S) + 512;
}

int tlsan_score_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(bytes != nullptr, TLSAN_E_NULL, "bytes is NULL");
  *bytes = score_ws_bytes(dims);
  return TLSAN_OK;
}

int tlsan_score_ws(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, int32_t ncand,
                   float* logits, float* ut, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, false))) return rc;
  REQUIRE(ncand == 1 || ncand == 2, TLSAN_E_DIMS, "ncand must be 1 or 2 (got %d)", ncand);
  if ((rc = check_batch(b, false, ncand))) return rc;
  REQUIRE(logits != nullptr && workspace != nullptr, TLSAN_E_NULL, "logits/workspace is NULL");
  REQUIRE(ut == nullptr || aligned16(ut), TLSAN_E_ALIGN, "ut must be 16-B aligned");
  REQUIRE(workspace_bytes >= score_ws_bytes(dims), TLSAN_E_WORKSPACE, "workspace too small");
  float* scratch = reinterpret_cast<float*>(tlsan_align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  if (!use_mma()) return tlsan_launch_score(*dims, *p, *b, ncand, logits, ut, (cudaStream_t)stream);
  return tlsan_launch_score_ws(*dims, *p, *b, ncand, logits, ut, scratch, (cudaStream_t)stream);
}

int tlsan_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(bytes != nullptr, TLSAN_E_NULL, "bytes is NULL");
  *bytes = tlsan_ws_layout(*dims).total + 256;
  return TLSAN_OK;
}

int tlsan_flat_count(const tlsan_dims_t* dims, int64_t* count) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(count != nullptr, TLSAN_E_NULL, "count is NULL");
  *count = (int64_t)tlsan_ws_layout(*dims).flat_count;
  return TLSAN_OK;
}

static char* ws_base(void* workspace) {
  return reinterpret_cast<char*>(tlsan_align_up(reinterpret_cast<uintptr_t>(workspace), 256));
}

static int step_grads_impl(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, void* workspace,
                           size_t workspace_bytes, float* flat, bool with_tsq, const tlsan_next_t* next, void* stream) {
  // the side streams, the presort registry and the phase recorder are per-process: host threads driving different
  // models enqueue their steps one at a time (the kernels themselves still overlap on their streams)
  std::lock_guard<std::recursive_mutex> guard(g_api_mutex);
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, true))) return rc;
  if ((rc = check_batch(b, true, 1))) return rc;
  REQUIRE(workspace && flat, TLSAN_E_NULL, "workspace/flat is NULL");
  REQUIRE(aligned16(flat), TLSAN_E_ALIGN, "flat must be 16-B aligned");
  const TlsanWs w = tlsan_ws_layout(*dims);
  REQUIRE(workspace_bytes >= w.total + 256, TLSAN_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes,
          w.total + 256);
  char* ws = ws_base(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* sorted_vals = nullptr;
  tlsan_profile_mark(-1, st);
  SideStream* side = use_mma() ? side_stream() : nullptr;
  cudaEvent_t sorted = nullptr, part_ready = nullptr, seg_ready = nullptr;
  const bool pf = fused_impl() == 4;                            // kernels that take the balanced partition
  int long_ctas = 3;
  // bit 1 = "the previous *_pipelined call announced this batch": honoured only if that call really enqueued the
  // presort (it does not when the side streams are off); otherwise the step sorts in place like a plain one
  Presort* ps = (dims->reserved & 2) && side ? presort_slot(ws, false) : nullptr;
  const bool presorted = ps && ps->valid;
  if (presorted) {
    ps->valid = false;
    sorted = ps->ev;
    seg_ready = ps->ev_seg;
    if (pf) part_ready = ps->ev_part;
    sorted_vals = tlsan_sorted_vals(w, ws);
    tlsan_profile_mark(TLSAN_PHASE_SORT, st);
    if (with_tsq) {
      TLSAN_CHECK_CUDA(cudaEventRecord(side->fork, st));
      TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
      if ((rc = tlsan_launch_table_sumsq(*dims, *p, w, ws, side->st))) return rc;
    }
    g_prof_overlap = false;
  } else if (side) {
    TLSAN_CHECK_CUDA(cudaEventRecord(side->fork, st));
    TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    if (Presort* stale = presort_slot(ws, false)) {
      // a presort announced for this workspace was not consumed (the caller trained on another batch): its
      // kernels may still be writing the sort buffers we are about to reuse
      // -- and, laid out for ANOTHER batch size, any other part of this workspace (the metadata pre-pass and the
      // forward's work counter are written by the main stream right away): both streams wait for it
      if (stale->valid) {
        TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st, stale->ev_seg, 0));
        TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st, stale->ev_part, 0));
        TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, stale->ev_seg, 0));
        TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, stale->ev_part, 0));
      }
      stale->valid = false;
    }
    if (!pf) long_ctas = tlsan_overlap_ctas();  // statically partitioned forward kernels leave room for the sort's CTAs
    if ((rc = tlsan_launch_sort(*dims, *p, *b, w, ws, &sorted_vals, nullptr, side->st))) return rc;
    tlsan_profile_mark(TLSAN_PHASE_SORT, side->st);
    TLSAN_CHECK_CUDA(cudaEventRecord(side->join, side->st));
    sorted = side->join;
    if (pf) {                    // balanced partition of the short-term / backward kernels: behind the sort, same stream
      if ((rc = tlsan_launch_partition_batch(*dims, *p, *b, long_ctas, ws + w.part, side->st))) return rc;
      TLSAN_CHECK_CUDA(cudaEventRecord(side->part, side->st));
      part_ready = side->part;
    }
    // ||W||^2 of the tables needs only the (still unchanged) weights: off the critical path too
    if (with_tsq && (rc = tlsan_launch_table_sumsq(*dims, *p, w, ws, side->st))) return rc;
    g_prof_overlap = true;
  } else {
    if ((rc = tlsan_launch_sort(*dims, *p, *b, w, ws, &sorted_vals, nullptr, st))) return rc;
    tlsan_profile_mark(TLSAN_PHASE_SORT, st);
    if (with_tsq && (rc = tlsan_launch_table_sumsq(*dims, *p, w, ws, st))) return rc;
    g_prof_overlap = false;
  }
  int grid_a = 0, grid_b = 0, grid_c = 0;
  if (fused_impl() >= 2)
    rc = tlsan_launch_fwd_bwd_async(*dims, *p, *b, w, ws, &grid_a, &grid_b, &grid_c, fused_impl() - 2, sorted,
                                    part_ready, presorted, long_ctas, st);
  else if (fused_impl() == 1)
    rc = tlsan_launch_fwd_bwd_mma(*dims, *p, *b, w, ws, &grid_a, &grid_b, &grid_c, sorted, long_ctas, st);
  else rc = tlsan_launch_fwd_bwd(*dims, *p, *b, w, ws, &grid_a, &grid_b, st);
  if (rc) return rc;
  if (next && side) {
    // sort of the next batch: behind the backward kernels of this step, beside its reduce / all-reduce / update
    REQUIRE(next->dims && next->batch && next->workspace, TLSAN_E_NULL, "next has a NULL field");
    if ((rc = check_dims(next->dims))) return rc;
    if ((rc = check_batch(next->batch, true, 1))) return rc;
    const TlsanWs wn = tlsan_ws_layout(*next->dims);
    REQUIRE(next->workspace_bytes >= wn.total + 256, TLSAN_E_WORKSPACE, "next workspace too small");
    char* wsn = ws_base(next->workspace);
    REQUIRE(wsn != ws, TLSAN_E_UNSUPPORTED, "the next batch needs its own workspace");
    Presort* consumed = ps;
    Presort* ps = presort_slot(wsn, true, consumed);
    REQUIRE(ps != nullptr, TLSAN_E_UNSUPPORTED, "too many presorted workspaces in flight");
    const int32_t* unused = nullptr;
    // released here, behind the long-term backward: measured best of the five possible points of the chain (an
    // earlier release only moves the sort's SM time from the reduce / update phase into the backward kernels)
    TLSAN_CHECK_CUDA(cudaEventRecord(side->fork2, st));
    TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st2, side->fork2, 0));
    if (next->ready_event) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st2, (cudaEvent_t)next->ready_event, 0));
    // ps->ev is recorded inside, before the segment-bounds kernel
    if ((rc = tlsan_launch_sort(*next->dims, *p, *next->batch, wn, wsn, &unused, ps->ev, side->st2))) return rc;
    TLSAN_CHECK_CUDA(cudaEventRecord(ps->ev_seg, side->st2));
    if (pf) {     // partition of the short-term / backward kernels: needed later than the ranks, so behind the sort
      if ((rc = tlsan_launch_partition_batch(*next->dims, *p, *next->batch, 3, wsn + wn.part, side->st2))) return rc;
      TLSAN_CHECK_CUDA(cudaEventRecord(ps->ev_part, side->st2));
    }
    ps->valid = true;
  }
  if (side) {
    // the fixed-order sum of the per-CTA partials and the segmented row reduce are independent: side by side
    TLSAN_CHECK_CUDA(cudaEventRecord(side->fork, st));
    TLSAN_CHECK_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    if ((rc = tlsan_launch_finalize1(w, ws, grid_a, grid_b, grid_c, flat + w.f_dgrad, side->st))) return rc;
    TLSAN_CHECK_CUDA(cudaEventRecord(side->join, side->st));
    if (seg_ready) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, seg_ready, 0));
    rc = tlsan_launch_row_reduce(*dims, w, ws, sorted_vals, flat + w.f_gi, flat + w.f_gb, flat + w.f_gu, st);
    TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  } else {
    if ((rc = tlsan_launch_finalize1(w, ws, grid_a, grid_b, grid_c, flat + w.f_dgrad, st))) return rc;
    rc = tlsan_launch_row_reduce(*dims, w, ws, sorted_vals, flat + w.f_gi, flat + w.f_gb, flat + w.f_gu, st);
  }
  tlsan_profile_mark(TLSAN_PHASE_REDUCE, st);
  return rc;
}

static int apply_flat_impl(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat, float lr, float reg,
                           float clip_norm, void* workspace, size_t workspace_bytes, float* stats, bool have_tsq,
                           void* stream, const tlsan_opt_t* opt = nullptr) {
  std::lock_guard<std::recursive_mutex> guard(g_api_mutex);
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, true))) return rc;
  REQUIRE(workspace && flat && stats, TLSAN_E_NULL, "workspace/flat/stats is NULL");
  REQUIRE(clip_norm > 0.f, TLSAN_E_DIMS, "clip_norm must be > 0");
  const TlsanWs w = tlsan_ws_layout(*dims);
  REQUIRE(workspace_bytes >= w.total + 256, TLSAN_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes,
          w.total + 256);
  char* ws = ws_base(workspace);
  if (opt && opt->kind != TLSAN_OPT_SGD) {
    REQUIRE(opt->kind >= TLSAN_OPT_ADAM && opt->kind <= TLSAN_OPT_ADADELTA, TLSAN_E_UNSUPPORTED, "unknown optimizer kind %d", opt->kind);
    REQUIRE(opt->slot1 && opt->slot2 && opt->step >= 1, TLSAN_E_NULL, "optimizer slots are NULL or step < 1");
    // the slots mirror ONE weight buffer: emb | usert | item_b | dense must be laid out like tlsan_dp_exchange needs
    const long long NR = (long long)dims->NI + dims->NC + dims->NU;
    const long long off_usert = NR * 32, off_itemb = off_usert + ((long long)dims->NU * dims->L + 3) / 4 * 4;
    const long long off_dense = off_itemb + ((long long)dims->NI + 3) / 4 * 4;
    REQUIRE(p->usert == p->emb + off_usert && p->item_b == p->emb + off_itemb && p->dense == p->emb + off_dense,
            TLSAN_E_UNSUPPORTED, "adam / rmsprop / adadelta need emb | usert | item_b | dense in one buffer (see header)");
  }
  rc = tlsan_launch_apply(*dims, *p, w, ws, flat + w.f_gi, flat + w.f_gb, flat + w.f_gu, flat + w.f_dgrad, lr,
                          reg, clip_norm, have_tsq, stats, opt, (cudaStream_t)stream);
  tlsan_profile_mark(TLSAN_PHASE_APPLY, (cudaStream_t)stream);
  return rc;
}

int tlsan_step_grads(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, void* workspace,
                     size_t workspace_bytes, float* flat, void* stream) {
  // the table norms go into the workspace here (beside the forward kernels): tlsan_apply_flat, which must
  // follow on the same workspace with the weights unchanged, reads them instead of recomputing
  return step_grads_impl(dims, p, b, workspace, workspace_bytes, flat, dims && !(dims->reserved & 1), nullptr, stream);
}

int tlsan_apply_flat(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat, float lr, float reg,
                     float clip_norm, void* workspace, size_t workspace_bytes, float* stats, void* stream) {
  return apply_flat_impl(dims, p, flat, lr, reg, clip_norm, workspace, workspace_bytes, stats, true, stream);
}

int tlsan_apply_flat_opt(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* flat, float lr, float reg,
                         float clip_norm, const tlsan_opt_t* opt, void* workspace, size_t workspace_bytes, float* stats,
                         void* stream) {
  return apply_flat_impl(dims, p, flat, lr, reg, clip_norm, workspace, workspace_bytes, stats, true, stream, opt);
}

int tlsan_step_grads_pipelined(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                               const tlsan_next_t* next, void* workspace, size_t workspace_bytes, float* flat,
                               void* stream) {
  return step_grads_impl(dims, p, b, workspace, workspace_bytes, flat, dims && !(dims->reserved & 1), next, stream);
}

int tlsan_train_step_pipelined(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b,
                               const tlsan_next_t* next, float lr, float reg, float clip_norm, void* workspace,
                               size_t workspace_bytes, float* stats, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(workspace != nullptr, TLSAN_E_NULL, "workspace is NULL");
  const TlsanWs w = tlsan_ws_layout(*dims);
  float* flat = reinterpret_cast<float*>(ws_base(workspace) + w.flat);
  if ((rc = step_grads_impl(dims, p, b, workspace, workspace_bytes, flat, true, next, stream))) return rc;
  return apply_flat_impl(dims, p, flat, lr, reg, clip_norm, workspace, workspace_bytes, stats, true, stream);
}

int tlsan_dp_arena_bytes(const tlsan_dims_t* dims, int32_t world, size_t* bytes) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(bytes != nullptr, TLSAN_E_NULL, "bytes is NULL");
  REQUIRE(world >= 1 && world <= 16, TLSAN_E_DIMS, "world must be in [1,16] (got %d)", world);
  *bytes = tlsan_dp_arena_bytes_impl(*dims, world);
  return TLSAN_OK;
}

int tlsan_dp_arena_create(size_t bytes, void** ptr, char* handle64) {
  REQUIRE(ptr && handle64 && bytes > 0, TLSAN_E_NULL, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  TLSAN_CHECK_CUDA(cudaMalloc(ptr, bytes));
  TLSAN_CHECK_CUDA(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  TLSAN_CHECK_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return TLSAN_OK;
}

int tlsan_dp_arena_open(const char* handle64, void** ptr) {
  REQUIRE(ptr && handle64, TLSAN_E_NULL, "NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  TLSAN_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TLSAN_OK;
}

int tlsan_dp_arena_release(void* ptr, int32_t owned) {
  if (!ptr) return TLSAN_OK;
  if (owned) TLSAN_CHECK_CUDA(cudaFree(ptr));
  else TLSAN_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return TLSAN_OK;
}

int tlsan_dp_exchange(const tlsan_dims_t* dims, const tlsan_params_t* p, void* const* arenas, int32_t rank,
                      int32_t world, int32_t epoch, float lr, float reg, float clip_norm, void* workspace,
                      size_t workspace_bytes, float* stats, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, true))) return rc;
  REQUIRE(arenas && workspace && stats, TLSAN_E_NULL, "NULL argument");
  REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, TLSAN_E_DIMS, "bad rank / world (%d / %d)", rank, world);
  REQUIRE(epoch >= 1 && clip_norm > 0.f, TLSAN_E_DIMS, "epoch must be >= 1 and clip_norm > 0");
  for (int i = 0; i < world; ++i) REQUIRE(arenas[i] != nullptr, TLSAN_E_NULL, "arena %d is NULL", i);
  // the tables must be ONE buffer: emb | usert | item_b (each padded to 16 B)
  const long long NR = (long long)dims->NI + dims->NC + dims->NU;
  const long long off_usert = NR * 32, off_itemb = off_usert + ((long long)dims->NU * dims->L + 3) / 4 * 4;
  REQUIRE(p->usert == p->emb + off_usert && p->item_b == p->emb + off_itemb, TLSAN_E_UNSUPPORTED,
          "tlsan_dp_exchange needs emb | usert | item_b in one buffer (see header)");
  const TlsanWs w = tlsan_ws_layout(*dims);
  REQUIRE(workspace_bytes >= w.total + 256, TLSAN_E_WORKSPACE, "workspace too small");
  std::lock_guard<std::recursive_mutex> guard(g_api_mutex);
  // the small exchange (dense gradients, norm partials -> clip scale) runs on the side stream, behind k_finalize1 of
  // the step and beside its segmented row reduce
  SideStream* side = use_mma() ? side_stream() : nullptr;
  rc = tlsan_launch_dp_exchange(*dims, *p, w, ws_base(workspace), reinterpret_cast<float* const*>(arenas), rank, world,
                                epoch, lr, reg, clip_norm, stats, side ? side->st : nullptr, side ? side->dpx : nullptr,
                                (cudaStream_t)stream);
  tlsan_profile_mark(TLSAN_PHASE_APPLY, (cudaStream_t)stream);
  return rc;
}

int tlsan_train_step(const tlsan_dims_t* dims, const tlsan_params_t* p, const tlsan_batch_t* b, float lr, float reg,
                     float clip_norm, void* workspace, size_t workspace_bytes, float* stats, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(workspace != nullptr, TLSAN_E_NULL, "workspace is NULL");
  const TlsanWs w = tlsan_ws_layout(*dims);
  float* flat = reinterpret_cast<float*>(ws_base(workspace) + w.flat);
  // one fused step: the table norms are computed beside the forward kernels (weights change only in apply)
  if ((rc = step_grads_impl(dims, p, b, workspace, workspace_bytes, flat, true, nullptr, stream))) return rc;
  return apply_flat_impl(dims, p, flat, lr, reg, clip_norm, workspace, workspace_bytes, stats, true, stream);
}

int tlsan_label_rank(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* ut, const int32_t* label,
                     int32_t* rank, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, false))) return rc;
  REQUIRE(ut && label && rank, TLSAN_E_NULL, "NULL argument");
  return tlsan_launch_label_rank(*dims, *p, ut, label, rank, (cudaStream_t)stream);
}

int tlsan_rank_workspace_bytes(const tlsan_dims_t* dims, size_t* bytes) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  REQUIRE(bytes != nullptr, TLSAN_E_NULL, "bytes is NULL");
  *bytes = tlsan_rank_ws_bytes(*dims);
  return TLSAN_OK;
}

int tlsan_label_rank_ws(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* ut, const int32_t* label,
                        int32_t* rank, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_dims(dims))) return rc;
  if ((rc = check_params(p, false))) return rc;
  REQUIRE(ut && label && rank && workspace, TLSAN_E_NULL, "NULL argument");
  REQUIRE(aligned16(ut), TLSAN_E_ALIGN, "ut must be 16-B aligned");
  REQUIRE(workspace_bytes >= tlsan_rank_ws_bytes(*dims), TLSAN_E_WORKSPACE, "workspace too small");
  return tlsan_launch_label_rank_tc(*dims, *p, ut, label, rank, ws_base(workspace), (cudaStream_t)stream);
}

long long tlsan_launch_count(void) { return g_tlsan_launches; }

int tlsan_profile_begin(int32_t max_steps) {
  REQUIRE(max_steps > 0 && max_steps <= 4096, TLSAN_E_DIMS, "max_steps must be in [1,4096]");
  if (g_ev_steps_cap < max_steps) {
    const int per = TLSAN_PHASE_COUNT + 1;
    cudaEvent_t* ev = new cudaEvent_t[(size_t)max_steps * per];
    for (int i = 0; i < max_steps * per; ++i) TLSAN_CHECK_CUDA(cudaEventCreate(&ev[i]));
    g_ev = ev;  // earlier (smaller) pool is intentionally kept alive: events may still be pending
    g_ev_steps_cap = max_steps;
  }
  g_ev_step = -1;
  g_prof_on = true;
  return TLSAN_OK;
}

int tlsan_profile_end(float* ms, int32_t* steps) {
  REQUIRE(ms && steps, TLSAN_E_NULL, "NULL argument");
  g_prof_on = false;
  const int per = TLSAN_PHASE_COUNT + 1;
  const int n = g_ev_step + 1;
  for (int s = 0; s < n; ++s) {
    TLSAN_CHECK_CUDA(cudaEventSynchronize(g_ev[(size_t)s * per + TLSAN_PHASE_COUNT]));
    for (int ph = 0; ph < TLSAN_PHASE_COUNT; ++ph) {
      // a phase runs from the previous mark to its own; with the sort on the side stream the long-term
      // forward starts at the step-start mark like the sort does (the two overlap)
      const int begin = (ph == TLSAN_PHASE_LONG_FWD && g_prof_overlap) ? 0 : ph;
      TLSAN_CHECK_CUDA(cudaEventElapsedTime(&ms[s * TLSAN_PHASE_COUNT + ph], g_ev[(size_t)s * per + begin],
                                            g_ev[(size_t)s * per + ph + 1]));
    }
  }
  *steps = n;
  g_ev_step = -1;
  return TLSAN_OK;
}

}  // extern "C"
