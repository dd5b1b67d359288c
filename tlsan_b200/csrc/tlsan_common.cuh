// Shared declarations for the TLSAN sm_100a kernels (internal; the public surface is
// include/tlsan_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tlsan_b200.h"

#define TLSAN_INVALID_KEY 0x7fffffff
#define TLSAN_TILE 32          // samples per CTA tile in the fused kernels (8 warps x 4)
#define TLSAN_THREADS 256
#define TLSAN_MAX_GRID 592     // upper bound on persistent grid (148 SMs x 4)
// per-CTA partial row: dense layout + [loss, sumsq, pad...]
#define TLSAN_PART_LOSS TLSAN_DENSE_PAD
#define TLSAN_PART_SUMSQ (TLSAN_DENSE_PAD + 1)
#define TLSAN_PART 4456
#define TLSAN_SCR 6            // 64-float slots per sample in scratch: do_long | o_long | max | 1/den | dz | z
#define TLSAN_SORT_CTA_KEYS 5120  // keys per CTA in one radix pass (16 warps x 320)

void tlsan_set_error(const char* fmt, ...);

#define TLSAN_CHECK_CUDA(expr)                                                      \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      tlsan_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TLSAN_E_CUDA;                                                          \
    }                                                                               \
  } while (0)

extern long long g_tlsan_launches;
void tlsan_profile_mark(int phase_done, cudaStream_t st);   // phase_done = -1: step start

#define TLSAN_CHECK_LAUNCH(name)                                                    \
  do {                                                                              \
    ++g_tlsan_launches;                                                             \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      tlsan_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return TLSAN_E_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline size_t tlsan_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch (PDL): the kernels of the step's main stream are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of kernel N+1 become resident while kernel N drains and
// run their prologue (weight images, partition bounds -- nothing kernel N writes) up to pdl_wait(), which returns
// once kernel N has completed and flushed.  Every such kernel calls pdl_trigger() only AFTER its own pdl_wait(): when
// kernel N+1 starts, kernel N-1 and everything before it are complete, so the pre-wait region may read whatever
// kernels <= N-1 wrote.  Launched without the attribute both instructions are no-ops.   TLSAN_PDL=0 switches it off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
int tlsan_pdl_level();     // TLSAN_PDL: 0 off, 1 = the update-phase kernels (reduce, fix-up, finalize, apply), 2 = every kernel of the step
template <typename... KArgs, typename... Args>
static inline cudaError_t tlsan_launch_kl(int level, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                          cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = tlsan_pdl_level() >= level ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t tlsan_launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args) {
  return tlsan_launch_kl(2, kern, grid, block, smem, st, args...);
}

// Workspace carve-up (byte offsets from a 256-B aligned base).
struct TlsanWs {
  int SLOTS;  // occurrence slots per sample: L long, S short, candidate, u_cate, user
  int SP, SPSH;  // slot stride = next power of two >= SLOTS (occurrence id = b * SP + j), log2(SP)
  int SI;     // slots with a 64-float payload (all but the user slot)
  int PU;     // floats per user payload: 32 (user_emb grad) + L (usert_emb grad), padded to 4
  int NR;     // unified row space NI + NC + NU
  int64_t nocc;
  int nchunks;
  size_t keys_a, keys_b, vals_a, vals_b, inv, hist, nvalid, seg_off;
  size_t rows_i, rows_u, gscal, scratch, meta, smeta, sscal, part;
  size_t part_a, part_b, part_c, tsq, rpart, flat;
  // flat gradient buffer (float offsets): [g_i (NI+NC)x64 | g_b NIpad | g_u NUxPU | dgrad PART]
  size_t f_gi, f_gb, f_gu, f_dgrad, flat_count;
  size_t total;
};

size_t tlsan_partition_bytes();
size_t tlsan_score_meta_bytes(int B, int S);   // scoring-layout short-term metadata of tlsan_score_ws (tlsan_fused_pf.cu)

static inline TlsanWs tlsan_ws_layout(const tlsan_dims_t& d) {
  TlsanWs w;
  w.SLOTS = d.L + d.S + 3;
  w.SI = d.L + d.S + 2;
  w.PU = (int)tlsan_align_up(32 + d.L, 4);
  w.NR = d.NI + d.NC + d.NU;
  w.SPSH = 0;
  while ((1 << w.SPSH) < w.SLOTS) ++w.SPSH;
  w.SP = 1 << w.SPSH;
  w.nocc = (int64_t)d.B * w.SP;
  w.nchunks = (int)((w.nocc + TLSAN_SORT_CTA_KEYS - 1) / TLSAN_SORT_CTA_KEYS);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = tlsan_align_up(o + bytes, 256); return r; };
  w.keys_a = take(w.nocc * 4);
  w.keys_b = take(w.nocc * 4);
  w.vals_a = take(w.nocc * 4);
  w.vals_b = take(w.nocc * 4);
  w.inv = take(w.nocc * 4);
  w.hist = take((size_t)256 * (w.nchunks + 1) * 4 + 1024);   // per-CTA digit counts + 256 digit totals
  w.nvalid = take(64);
  w.seg_off = take((size_t)(w.NR + 2) * 4);
  w.rows_i = take((size_t)d.B * w.SI * 64 * 4);
  w.rows_u = take((size_t)d.B * w.PU * 4);
  w.gscal = take((size_t)d.B * 4);
  w.scratch = take((size_t)d.B * TLSAN_SCR * 64 * 4);
  w.meta = take((size_t)d.B * d.L * 16);       // resolved per-token metadata of the long-term sequence (k_long_meta)
  w.smeta = take((size_t)d.B * (d.S + 2) * 8); // row pairs of candidate / user vector / session items
  w.sscal = take((size_t)d.B * 16);            // per-sample scalars of the short-term kernel
  w.part = take(tlsan_partition_bytes());      // balanced partitions of the three pipelined kernels + their scratch
  w.part_a = take((size_t)TLSAN_MAX_GRID * TLSAN_PART * 4);
  w.part_b = take((size_t)TLSAN_MAX_GRID * TLSAN_PART * 4);
  w.part_c = take((size_t)TLSAN_MAX_GRID * TLSAN_PART * 4);
  w.tsq = take((size_t)TLSAN_MAX_GRID * 4 * 4);
  w.rpart = take((size_t)2 * TLSAN_MAX_GRID * 8 * 72 * 4);   // head / tail partial rows of the balanced segmented reduce
  w.f_gi = 0;
  w.f_gb = (size_t)(d.NI + d.NC) * 64;
  w.f_gu = w.f_gb + tlsan_align_up(d.NI, 4);
  w.f_dgrad = w.f_gu + (size_t)d.NU * w.PU;
  w.flat_count = w.f_dgrad + TLSAN_PART;
  w.flat = take(w.flat_count * 4);
  w.total = o;
  return w;
}

// ---- launchers implemented in the .cu files (all asynchronous on `st`) ----
int tlsan_launch_upload_consts(const float* dense, cudaStream_t st);
int tlsan_launch_score(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                       float* logits, float* ut, cudaStream_t st);
int tlsan_launch_fwd_bwd(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b,
                         const TlsanWs& w, char* ws, int* grid_a, int* grid_b, cudaStream_t st);
int tlsan_launch_gather(const tlsan_dims_t& d, const tlsan_params_t& p, const int32_t* idx, const float* tau,
                        float* out, int64_t n, cudaStream_t st);
int tlsan_launch_bucket(const int32_t* dd, const float* lut, float* out, int32_t* bucket, int64_t n,
                        cudaStream_t st);
int tlsan_launch_sort(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, const TlsanWs& w,
                      char* ws, const int32_t** sorted_vals, cudaEvent_t ranks_ready, cudaStream_t st);
const int32_t* tlsan_sorted_vals(const TlsanWs& w, char* ws);
int tlsan_launch_row_reduce(const tlsan_dims_t& d, const TlsanWs& w, char* ws, const int32_t* sorted_vals,
                            float* g_i, float* g_b, float* g_u, cudaStream_t st);
int tlsan_launch_score_mma(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                           float* logits, float* ut, cudaStream_t st);
int tlsan_launch_fwd_bwd_mma(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b,
                             const TlsanWs& w, char* ws, int* grid_a, int* grid_b, int* grid_c, cudaEvent_t sorted,
                             int long_ctas, cudaStream_t st);
int tlsan_launch_fwd_bwd_async(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b,
                               const TlsanWs& w, char* ws, int* grid_a, int* grid_b, int* grid_c, int variant,
                               cudaEvent_t sorted, cudaEvent_t part_ready, bool part_early, int long_ctas,
                               cudaStream_t st);
int tlsan_launch_partition_batch(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int fwd_ctas,
                                 void* part, cudaStream_t st);
int tlsan_overlap_ctas();
int tlsan_launch_score_ws(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                          float* logits, float* ut, float* scratch, cudaStream_t st);
int tlsan_launch_finalize1(const TlsanWs& w, char* ws, int grid_a, int grid_b, int grid_c, float* dgrad,
                           cudaStream_t st);
int tlsan_launch_apply(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                       const float* g_i, const float* g_b, const float* g_u, const float* dgrad, float lr,
                       float reg, float clip, bool have_tsq, float* stats, const tlsan_opt_t* opt, cudaStream_t st);
int tlsan_launch_table_sumsq(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                             cudaStream_t st);
int tlsan_launch_apply_replicated(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                                  const float* gcate, const float* g_u, const float* dgrad, const float* item_sumsq,
                                  int n_item_sumsq, float lr, float reg, float clip, float* stats, cudaStream_t st);
size_t tlsan_dp_arena_bytes_impl(const tlsan_dims_t& d, int world);
int tlsan_launch_dp_exchange(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                             float* const* arenas, int rank, int world, int epoch, float lr, float reg, float clip,
                             float* stats, cudaStream_t side, cudaEvent_t side_done, cudaStream_t st);
int tlsan_launch_label_rank(const tlsan_dims_t& d, const tlsan_params_t& p, const float* ut, const int32_t* label,
                            int32_t* rank, cudaStream_t st);
size_t tlsan_rank_ws_bytes(const tlsan_dims_t& d);
int tlsan_launch_label_rank_tc(const tlsan_dims_t& d, const tlsan_params_t& p, const float* ut, const int32_t* label,
                               int32_t* rank, char* img, cudaStream_t st);
int tlsan_num_sms();
