// Full-catalogue ranking for Model.eval_prec / eval_recall (reference model.py:140-156, 265-299) on the
// 5th-generation tensor cores:
//
//     rank[b] = #{ j : s_bj > s_b,lab  or  (s_bj == s_b,lab and j < lab) },   s = u_t . all_emb^T + item_b
//
// is a [B,64] x [64,NI] GEMM whose [B,NI] result is never materialised: tcgen05.mma accumulates a
// 128 users x 128 items tile in TMEM, four epilogue warps read it back with tcgen05.ld and only count.
//
//   * fp32-level accuracy from TF32 tensor cores: 3xTF32, s = Uhi.Ehi + Ulo.Ehi + Uhi.Elo (hi = truncation to
//     tf32, lo = x - hi), 27 MMAs (3 terms x 9 k-steps of K = 8) per tile into one accumulator.
//   * item_b rides in the GEMM: K = 72 = 64 features + a (1 | item_b) column + 7 zero columns.
//   * k_build_catalogue writes the catalogue ONCE per call as ready-made UMMA operand tiles (K-major, no swizzle:
//     [k-chunk of 4 floats][128 rows][16 B], hi image then lo image, 73 728 B per 128 items), so the producer
//     of the GEMM kernel is ONE thread issuing one cp.async.bulk per tile onto an mbarrier.
//   * CTA = one work unit (128-user tile x a range of item tiles), 192 threads, warp-specialised:
//       warps 0-3  epilogue: build the user operand + the label operand, then per tile tcgen05.ld + count
//       warp 4     producer: cp.async.bulk of the next catalogue tile into a 2-stage shared-memory ring
//       warp 5     MMA issuer (one lane) + TMEM allocation; 2 accumulators of 128 columns (double buffered)
//   * the label's own score comes from the SAME instruction sequence (a first tile whose "items" are the 128
//     labels; user m reads the diagonal), so s_bj == s_b,lab holds exactly at j = lab and ties break by index
//     like tf.nn.top_k.
#include "tlsan_tc.cuh"

#define RK_M 128
#define RK_N 128
#define RK_K 72
#define RK_CHUNKS (RK_K / 4)                 // 16-byte k-chunks per row
#define RK_PLANE (RK_CHUNKS * RK_N * 16)     // one hi or lo image of a 128-row operand tile (36 864 B)
#define RK_TILE (2 * RK_PLANE)
#define RK_KSTEP_BYTES (2 * RK_N * 16)       // one MMA consumes K = 8 = two k-chunks
#define RK_THREADS 192
#define RK_TMEM_COLS 256
#define RK_SMEM (3 * RK_TILE + 256)

// k-chunk c (4 floats) of the augmented catalogue row of item j: [item_emb[j] | cate_emb[icl[j]] | item_b[j], 0...]
// (`cate` = first row of cate_emb: emb + NI * 32 in the replicated model, a separate buffer for a row shard)
__device__ __forceinline__ float4 catalogue_chunk(const float* __restrict__ cate, const float* __restrict__ emb,
                                                  const float* __restrict__ item_b, const int* __restrict__ icl, int j,
                                                  int c) {
  if (c < 8) return __ldg(reinterpret_cast<const float4*>(emb + (size_t)j * 32) + c);
  if (c < 16) return __ldg(reinterpret_cast<const float4*>(cate + (size_t)__ldg(icl + j) * 32) + (c - 8));
  if (c == 16) return make_float4(__ldg(item_b + j), 0.f, 0.f, 0.f);
  return make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void split_store(char* plane_hi, int c, int r, float4 v) {
  const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  *reinterpret_cast<float4*>(plane_hi + (size_t)c * (RK_N * 16) + r * 16) = h;
  *reinterpret_cast<float4*>(plane_hi + RK_PLANE + (size_t)c * (RK_N * 16) + r * 16) = l;
}

// catalogue -> UMMA operand tiles in global memory (one thread per (tile, k-chunk, row))
__global__ void __launch_bounds__(256) k_build_catalogue(int NI, const float* __restrict__ emb,
                                                         const float* __restrict__ cate,
                                                         const float* __restrict__ item_b, const int* __restrict__ icl,
                                                         char* __restrict__ img, int ntiles) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  const int r = (int)(g % RK_N), c = (int)((g / RK_N) % RK_CHUNKS);
  const long long tile = g / (RK_N * RK_CHUNKS);
  if (tile >= ntiles) return;
  const long long j = tile * RK_N + r;
  const float4 v = j < NI ? catalogue_chunk(cate, emb, item_b, icl, (int)j, c) : make_float4(0.f, 0.f, 0.f, 0.f);
  split_store(img + tile * RK_TILE, c, r, v);
}

// count the columns of one 32-column slice that rank above the label
template <bool MASKED>
__device__ __forceinline__ int count_above(const float (&s)[32], float sl, int dlab, int col0, int nvalid) {
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int col = col0 + i;
    bool above = (s[i] > sl) || (s[i] == sl && col < dlab);
    if (MASKED) above = above && col < nvalid;
    cnt += above ? 1 : 0;
  }
  return cnt;
}

// Row shards (tlsan_label_rank_shard): the catalogue holds the rows of ONE shard, local row j is global item
// j * gid_mul + gid_add; `label` holds GLOBAL ids and `lab_rows` [B][68] the labels' augmented rows
// (item_emb | cate_emb | item_b, pad) gathered by whoever owns them, so every shard derives the label's score with the
// same instruction sequence; rank[b] then counts this shard's share and the caller sums over the shards.
__global__ void __launch_bounds__(RK_THREADS, 1) k_label_rank_tc(int B, int NI, const float* __restrict__ emb,
                                                                 const float* __restrict__ cate,
                                                                 const float* __restrict__ item_b,
                                                                 const int* __restrict__ icl,
                                                                 const float* __restrict__ ut,
                                                                 const int* __restrict__ label,
                                                                 const float* __restrict__ lab_rows, int gid_mul,
                                                                 int gid_add, const char* __restrict__ img, int ntiles,
                                                                 int tiles_per_unit, int* __restrict__ rank) {
  extern __shared__ __align__(128) unsigned char smem[];
  char* sA = reinterpret_cast<char*>(smem);                  // user operand: hi plane, lo plane
  char* sB = sA + RK_TILE;                                   // 2 stages of catalogue tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 3 * RK_TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4),
                 bar_tempty = smem_u32(bars + 6), bar_diag = smem_u32(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * RK_M;
  const int t0 = blockIdx.y * tiles_per_unit, t1 = min(ntiles, t0 + tiles_per_unit);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);      // one arrival per epilogue warp
    }
    mbar_init(bar_diag, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(RK_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  int lab = 0;
  if (warp < 4) {
    // user operand row m = [u_t[b] | 1 | 0...], label operand row m = catalogue row of label[b]  (stage 0)
    const int m = threadIdx.x, b = m0 + m;
    const bool live = b < B;
    lab = live ? __ldg(label + b) : 0;
#pragma unroll 2
    for (int c = 0; c < RK_CHUNKS; ++c) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < 16) { if (live) v = __ldg(reinterpret_cast<const float4*>(ut + (size_t)b * 64) + c); }
      else if (c == 16) v.x = 1.f;
      split_store(sA, c, m, v);
      float4 lv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lab_rows) { if (live && c < 17) lv = __ldg(reinterpret_cast<const float4*>(lab_rows + (size_t)b * 68) + c); }
      else lv = catalogue_chunk(cate, emb, item_b, icl, lab, c);
      split_store(sB, c, m, lv);
    }
    // first local row whose global id is >= the label: ties against items before the label count as "above"
    if (lab_rows) lab = lab <= gid_add ? 0 : (lab - gid_add + gid_mul - 1) / gid_mul;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + RK_PLANE;

  auto issue_tile = [&](uint32_t b_hi, uint32_t d_tmem) {    // 3xTF32: hi.hi + lo.hi + hi.lo
    const uint32_t b_lo = b_hi + RK_PLANE;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t pa = term == 1 ? a_lo : a_hi, pb = term == 2 ? b_lo : b_hi;
#pragma unroll
      for (int ks = 0; ks < RK_K / 8; ++ks)
        umma_tf32(d_tmem, umma_desc(pa + ks * RK_KSTEP_BYTES, RK_N * 16, 128),
                  umma_desc(pb + ks * RK_KSTEP_BYTES, RK_N * 16, 128), umma_idesc_tf32(RK_M, RK_N, 0, 0),
                  (term | ks) ? 1u : 0u);
    }
  };

  // ---- label scores: tile of the 128 labels, user m reads D[m][m]
  if (warp == 5 && lane == 0) {
    issue_tile(smem_u32(sB), tmem);
    umma_commit(bar_diag);
  }
  float sl = 0.f;
  if (warp < 4) {
    mbar_wait(bar_diag, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + warp * 32, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) sl = i == lane ? v[i] : sl;
  }
  tc_fence_before();
  __syncthreads();            // stage 0 and accumulator 0 are free again
  tc_fence_after();

  if (warp == 4) {
    if (lane == 0) {
      for (int t = t0; t < t1; ++t) {
        const int k = t - t0, s = k & 1;
        mbar_wait(bar_empty + 8 * s, ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_full + 8 * s, RK_TILE);
        bulk_g2s(smem_u32(sB + (size_t)s * RK_TILE), img + (size_t)t * RK_TILE, RK_TILE, bar_full + 8 * s);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      for (int t = t0; t < t1; ++t) {
        const int k = t - t0, s = k & 1;
        mbar_wait(bar_full + 8 * s, (k >> 1) & 1);
        mbar_wait(bar_tempty + 8 * s, ((k >> 1) & 1) ^ 1);
        tc_fence_after();
        issue_tile(smem_u32(sB + (size_t)s * RK_TILE), tmem + s * RK_N);
        umma_commit(bar_empty + 8 * s);     // shared-memory stage may be refilled
        umma_commit(bar_tfull + 8 * s);     // accumulator is complete
      }
    }
  } else {
    int cnt = 0;
    for (int t = t0; t < t1; ++t) {
      const int k = t - t0, s = k & 1;
      mbar_wait(bar_tfull + 8 * s, (k >> 1) & 1);
      tc_fence_after();
      const int nbase = t * RK_N, dlab = lab - nbase, nvalid = NI - nbase;
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + s * RK_N;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        float v[32];
        tmem_ld32(taddr + q * 32, v);
        cnt += nvalid >= RK_N ? count_above<false>(v, sl, dlab, q * 32, nvalid)
                              : count_above<true>(v, sl, dlab, q * 32, nvalid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
    }
    const int b = m0 + threadIdx.x;
    if (b < B && cnt) atomicAdd(rank + b, cnt);    // integer: the sum over work units is order independent
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(RK_TMEM_COLS) : "memory");
  }
}

size_t tlsan_rank_ws_bytes(const tlsan_dims_t& d) {
  const size_t ntiles = ((size_t)d.NI + RK_N - 1) / RK_N;
  return ntiles * RK_TILE + 256;
}

static int launch_rank_tc(int B, int NI, const float* emb, const float* cate, const float* item_b, const int* icl,
                          const float* ut, const int32_t* label, const float* lab_rows, int gid_mul, int gid_add,
                          int32_t* rank, char* img, cudaStream_t st);

int tlsan_launch_label_rank_tc(const tlsan_dims_t& d, const tlsan_params_t& p, const float* ut, const int32_t* label,
                               int32_t* rank, char* img, cudaStream_t st) {
  return launch_rank_tc(d.B, d.NI, p.emb, p.emb + (size_t)d.NI * 32, p.item_b, p.icl, ut, label, nullptr, 1, 0, rank, img, st);
}

static int launch_rank_tc(int B, int NI, const float* emb, const float* cate, const float* item_b, const int* icl,
                          const float* ut, const int32_t* label, const float* lab_rows, int gid_mul, int gid_add,
                          int32_t* rank, char* img, cudaStream_t st) {
  struct { int B, NI; } d = {B, NI};
  struct { const float* emb; const float* item_b; const int* icl; } p = {emb, item_b, icl};
  static bool attr = false;
  if (!attr) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_label_rank_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM));
    attr = true;
  }
  const int ntiles = (d.NI + RK_N - 1) / RK_N, mtiles = (d.B + RK_M - 1) / RK_M;
  const long long nthreads = (long long)ntiles * RK_N * RK_CHUNKS;
  k_build_catalogue<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(d.NI, p.emb, cate, p.item_b, p.icl, img, ntiles);
  TLSAN_CHECK_LAUNCH("k_build_catalogue");
  TLSAN_CHECK_CUDA(cudaMemsetAsync(rank, 0, (size_t)d.B * sizeof(int32_t), st));
  // work units: enough (user tile, item range) pairs to fill the SMs about twice, each at least 8 item tiles
  int splits = (2 * tlsan_num_sms() + mtiles - 1) / mtiles;
  if (splits > (ntiles + 7) / 8) splits = (ntiles + 7) / 8;
  if (splits < 1) splits = 1;
  const int per = (ntiles + splits - 1) / splits;
  splits = (ntiles + per - 1) / per;
  k_label_rank_tc<<<dim3(mtiles, splits), RK_THREADS, RK_SMEM, st>>>(d.B, d.NI, p.emb, cate, p.item_b, p.icl, ut, label,
                                                                     lab_rows, gid_mul, gid_add, img, ntiles, per, rank);
  TLSAN_CHECK_LAUNCH("k_label_rank_tc");
  return TLSAN_OK;
}

// extern "C": full-catalogue label ranks against ONE row shard of the item tables (see k_label_rank_tc)
extern "C" int tlsan_label_rank_shard(int32_t B, int64_t n_local, const float* item_emb_shard, const float* item_b_shard,
                                      const int32_t* icl_shard, const float* cate_emb, const float* ut,
                                      const int32_t* label_global, const float* lab_rows, int32_t gid_mul,
                                      int32_t gid_add, int32_t* rank_partial, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  if (!item_emb_shard || !item_b_shard || !icl_shard || !cate_emb || !ut || !label_global || !lab_rows || !rank_partial ||
      !workspace) {
    tlsan_set_error("tlsan_label_rank_shard: NULL argument");
    return TLSAN_E_NULL;
  }
  if (B <= 0 || n_local <= 0 || n_local >= (1ll << 31) || gid_mul <= 0 || gid_add < 0) {
    tlsan_set_error("tlsan_label_rank_shard: bad dims");
    return TLSAN_E_DIMS;
  }
  const size_t ntiles = ((size_t)n_local + RK_N - 1) / RK_N;
  if (workspace_bytes < ntiles * RK_TILE + 256) {
    tlsan_set_error("tlsan_label_rank_shard: workspace too small");
    return TLSAN_E_WORKSPACE;
  }
  char* img = reinterpret_cast<char*>(tlsan_align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  return launch_rank_tc(B, (int)n_local, item_emb_shard, cate_emb, item_b_shard, icl_shard, ut, label_global, lab_rows,
                        gid_mul, gid_add, rank_partial, img, (cudaStream_t)stream);
}
