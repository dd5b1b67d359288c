// Row-sharded item tables (BASELINE config 5: NI = 10 M): item_emb / item_b / icl are split by row
// over the ranks; every step each rank asks the owners for the rows its batch touches (all-to-all of ids, then
// of rows), trains on a COMPACT table (row r = r-th distinct id of the local batch) with the unchanged fused
// kernels, and sends the per-id gradient rows back to the owners, which apply L2 + clip + SGD to their shard.
// This file holds the kernels either side of the collectives; orchestration in tlsan_b200/sharded.py.
//
// Exchange row = TLSAN_SHARD_ROW (36) words: 32 floats item_emb row | item_b | icl (int bits) | 2 pad  (144 B,
// 16-B aligned).  Gradient rows use the same shape: 32 floats d item_emb | d item_b | 3 pad.
#include "tlsan_common.cuh"

#define SHARD_THREADS 256

// out[k] = { emb[ids[k]][0..32), item_b[ids[k]], icl[ids[k]] }      9 lanes x 16 B per row
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_pack_rows(const float* __restrict__ emb,
                                                                   const float* __restrict__ item_b,
                                                                   const int* __restrict__ icl,
                                                                   const int* __restrict__ ids, long long n,
                                                                   long long n_local, float* __restrict__ out,
                                                                   int* __restrict__ bad) {
  const long long g = (long long)blockIdx.x * SHARD_THREADS + threadIdx.x;
  const long long k = g / 9;
  const int c = (int)(g - k * 9);
  if (k >= n) return;
  const long long id = ids[k];
  if (id < 0) {                           // padding slot of a fixed-capacity request list: zero row
    reinterpret_cast<float4*>(out + k * TLSAN_SHARD_ROW)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (id >= n_local) {                    // a foreign id reached this owner: flag, do not fault
    if (c == 0) atomicExch(bad, 1);
    return;
  }
  float4 v;
  if (c < 8) v = __ldg(reinterpret_cast<const float4*>(emb + id * 32) + c);
  else v = make_float4(__ldg(item_b + id), __int_as_float(__ldg(icl + id)), 0.f, 0.f);
  reinterpret_cast<float4*>(out + k * TLSAN_SHARD_ROW)[c] = v;
}

// emb_c[dst[k]] = packed row k; item_b_c / icl_c likewise
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_unpack_rows(const float* __restrict__ packed,
                                                                     const int* __restrict__ dst, long long n,
                                                                     float* __restrict__ emb_c,
                                                                     float* __restrict__ item_b_c,
                                                                     int* __restrict__ icl_c) {
  const long long g = (long long)blockIdx.x * SHARD_THREADS + threadIdx.x;
  const long long k = g / 9;
  const int c = (int)(g - k * 9);
  if (k >= n) return;
  const long long r = dst ? dst[k] : k;
  if (r < 0) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(packed + k * TLSAN_SHARD_ROW) + c);
  if (c < 8) reinterpret_cast<float4*>(emb_c + r * 32)[c] = v;
  else { item_b_c[r] = v.x; icl_c[r] = __float_as_int(v.y); }
}

// out[k] = { g_i[src[k]][0..32) (item half of the reduced row), g_b[src[k]] }
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_pack_grads(const float* __restrict__ g_i,
                                                                    const float* __restrict__ g_b,
                                                                    const int* __restrict__ src, long long n,
                                                                    float* __restrict__ out) {
  const long long g = (long long)blockIdx.x * SHARD_THREADS + threadIdx.x;
  const long long k = g / 9;
  const int c = (int)(g - k * 9);
  if (k >= n) return;
  const long long r = src ? src[k] : k;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= 0) {
    if (c < 8) v = __ldg(reinterpret_cast<const float4*>(g_i + r * 64) + c);
    else v = make_float4(__ldg(g_b + r), 0.f, 0.f, 0.f);
  }
  reinterpret_cast<float4*>(out + k * TLSAN_SHARD_ROW)[c] = v;
}

// g_emb[ids[k]] += packed[k][0..32), g_b[ids[k]] += packed[k][32].  The ids of ONE call are distinct (one
// source rank's list), so plain read-modify-write is race free; the caller issues the source ranks in rank
// order on one stream, which fixes the order of the additions.
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_accum_grads(const float* __restrict__ packed,
                                                                     const int* __restrict__ ids, long long n,
                                                                     float* __restrict__ g_emb,
                                                                     float* __restrict__ g_b) {
  const long long g = (long long)blockIdx.x * SHARD_THREADS + threadIdx.x;
  const long long k = g / 9;
  const int c = (int)(g - k * 9);
  if (k >= n) return;
  const long long id = ids[k];
  if (id < 0) return;                     // padding slot
  const float4 v = __ldg(reinterpret_cast<const float4*>(packed + k * TLSAN_SHARD_ROW) + c);
  if (c < 8) {
    float4* p = reinterpret_cast<float4*>(g_emb + id * 32) + c;
    float4 a = *p;
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    *p = a;
  } else {
    g_b[id] += v.x;
  }
}

// W[ids[k]] -= lr * scale * packed[k][0..32), b[ids[k]] -= lr * scale * packed[k][32]  (the L2 decay of every row is a
// separate dense pass, k_sgd_dense with g = NULL).  Same race rule as k_shard_accum_grads.
__global__ void __launch_bounds__(SHARD_THREADS) k_shard_apply_grads(const float* __restrict__ packed,
                                                                     const int* __restrict__ ids, long long n,
                                                                     float* __restrict__ W, float* __restrict__ b,
                                                                     float lr, const float* __restrict__ scale_p) {
  const long long g = (long long)blockIdx.x * SHARD_THREADS + threadIdx.x;
  const long long k = g / 9;
  const int c = (int)(g - k * 9);
  if (k >= n) return;
  const long long id = ids[k];
  if (id < 0) return;
  const float f = lr * *scale_p;
  const float4 v = __ldg(reinterpret_cast<const float4*>(packed + k * TLSAN_SHARD_ROW) + c);
  if (c < 8) {
    float4* p = reinterpret_cast<float4*>(W + id * 32) + c;
    float4 a = *p;
    a.x -= f * v.x; a.y -= f * v.y; a.z -= f * v.z; a.w -= f * v.w;
    *p = a;
  } else {
    b[id] -= f * v.x;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Device-side routing of the distinct item ids of a batch (no torch.unique, no host round trip).
// Every id has an owner-major POSITION p = owner * nloc + local (nloc = ceil(NI / W); cyclic partition: owner =
// id % W, local = id / W; block partition: p = id).  A presence bitmap over the positions is the set of distinct ids;
// a prefix of its word popcounts turns a position into its rank, and the COMPACT row of an id is its rank among all
// requested ids in owner-major order -- a pure function of the batch, so the compact table, the category CSR built on
// it and every summation order downstream are reproducible.  Request slot (owner o, k) = the k-th id of owner o's
// group; `cap` slots per owner are exchanged (equal-split all-to-alls, padding marked -1); a group that does not
// fit raises the overflow flag.
struct RouteGeo { int W, nloc, mod, cap; long long NI; };
__device__ __forceinline__ long long route_pos(const RouteGeo& g, int id) {
  return g.mod ? (long long)(id % g.W) * g.nloc + id / g.W : (long long)id;
}
struct RouteFields { const int* src[4]; int* dst[4]; long long n[4]; int nf; };

__global__ void __launch_bounds__(256) k_route_mark(const RouteFields f, const RouteGeo g, unsigned int* __restrict__ bits) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int q = 0; q < f.nf; ++q)
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < f.n[q]; e += stride) {
      const int id = __ldg(f.src[q] + e);
      if (id < 0 || id >= g.NI) continue;                      // validated on the host; never fault here
      const long long p = route_pos(g, id);
      const unsigned int m = 1u << (p & 31);
      if (!(*reinterpret_cast<volatile unsigned int*>(bits + (p >> 5)) & m)) atomicOr(bits + (p >> 5), m);   // padding zeros: one hot word
    }
}
__global__ void __launch_bounds__(256) k_route_popc(const unsigned int* __restrict__ bits, long long nwords,
                                                    int* __restrict__ cnt) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nwords) cnt[w] = __popc(bits[w]);
}
// set bits strictly before position p, given the INCLUSIVE prefix of the word popcounts
__device__ __forceinline__ int route_rank(const unsigned int* __restrict__ bits, const int* __restrict__ incl, long long p) {
  const long long w = p >> 5;
  const unsigned int word = __ldg(bits + w);
  return __ldg(incl + w) - __popc(word) + __popc(word & ((1u << (p & 31)) - 1u));
}
// per owner: its request list (owner-local row ids, ascending, padded with -1 to cap) and its count
__global__ void __launch_bounds__(256) k_route_emit(const unsigned int* __restrict__ bits, const int* __restrict__ incl,
                                                    long long nwords, const RouteGeo g, int* __restrict__ send_ids,
                                                    int* __restrict__ slot_row, int* __restrict__ counts,
                                                    int* __restrict__ overflow) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < g.W) {                                                // (first W threads also publish the group sizes)
    const int lo = route_rank(bits, incl, (long long)w * g.nloc);
    const long long endp = (long long)(w + 1) * g.nloc;
    const int hi = endp >= nwords * 32 ? __ldg(incl + nwords - 1) : route_rank(bits, incl, endp);
    counts[w] = hi - lo;
    if (hi - lo > g.cap) atomicExch(overflow, 1);
  }
  if (w >= nwords) return;
  unsigned int word = __ldg(bits + w);
  int before = __ldg(incl + w) - __popc(word);
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    const long long p = w * 32 + b;
    const int o = (int)(p / g.nloc);
    const int local = (int)(p - (long long)o * g.nloc);
    const int rank = before - route_rank(bits, incl, (long long)o * g.nloc);
    if (rank < g.cap) {
      send_ids[(long long)o * g.cap + rank] = local;
      slot_row[(long long)o * g.cap + rank] = before;           // dense compact row = rank among ALL requested ids
    }
    ++before;
  }
}
// id fields of the packed batch -> dense compact rows (rank of the id among all requested ids, owner-major order)
__global__ void __launch_bounds__(256) k_route_rewrite(const RouteFields f, const RouteGeo g,
                                                       const unsigned int* __restrict__ bits, const int* __restrict__ incl) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int q = 0; q < f.nf; ++q)
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < f.n[q]; e += stride) {
      const int id = __ldg(f.src[q] + e);
      int out = 0;
      if (id >= 0 && id < g.NI) out = route_rank(bits, incl, route_pos(g, id));
      f.dst[q][e] = out;
    }
}

// out[k][0..32) = sum of the cate halves of the reduced rows of category k (CSR order) + its direct row
__global__ void __launch_bounds__(256) k_reduce_cate(int NI, const float* __restrict__ g_i,
                                                     const int* __restrict__ cate_off,
                                                     const int* __restrict__ cate_items, float* __restrict__ out) {
  __shared__ float sh[8][32];
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lo = cate_off[k], hi = cate_off[k + 1];
  float acc = 0.f;
  for (int n = lo + warp; n < hi; n += 8) acc += __ldg(g_i + (size_t)__ldg(cate_items + n) * 64 + 32 + lane);
  sh[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float g = g_i[(size_t)(NI + k) * 64 + 32 + lane];
#pragma unroll
    for (int w = 0; w < 8; ++w) g += sh[w][lane];
    out[(size_t)k * 32 + lane] = g;
  }
}

// W <- W - lr * ((g + reg * W) * scale), g optional (NULL = pure L2 decay)
__global__ void __launch_bounds__(256) k_sgd_dense(float* __restrict__ W, const float* __restrict__ g, long long n,
                                                   float lr, float reg, const float* __restrict__ scale_p) {
  const float scale = *scale_p;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const float w = W[e];
    W[e] = w - lr * (((g ? g[e] : 0.f) + reg * w) * scale);
  }
}

__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ W, long long n, float* __restrict__ partial) {
  __shared__ float sh[8];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) s = fmaf(W[e], W[e], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += sh[w];
    partial[blockIdx.x] = r;
  }
}

static inline unsigned grid9(long long n) { return (unsigned)((n * 9 + SHARD_THREADS - 1) / SHARD_THREADS); }

#define SHARD_REQUIRE(cond, code, msg) \
  do {                                 \
    if (!(cond)) {                     \
      tlsan_set_error(msg);            \
      return code;                     \
    }                                  \
  } while (0)

extern "C" {

int tlsan_shard_pack_rows(const float* emb_shard, const float* item_b_shard, const int32_t* icl_shard,
                          const int32_t* local_ids, int64_t n, int64_t n_local, float* out, int32_t* bad_flag,
                          void* stream) {
  SHARD_REQUIRE(n >= 0 && n_local >= 0, TLSAN_E_DIMS, "tlsan_shard_pack_rows: negative count");
  if (n == 0) return TLSAN_OK;
  SHARD_REQUIRE(emb_shard && item_b_shard && icl_shard && local_ids && out && bad_flag, TLSAN_E_NULL,
                "tlsan_shard_pack_rows: NULL argument");
  SHARD_REQUIRE(((uintptr_t)emb_shard & 15) == 0 && ((uintptr_t)out & 15) == 0, TLSAN_E_ALIGN,
                "tlsan_shard_pack_rows: emb/out must be 16-B aligned");
  k_shard_pack_rows<<<grid9(n), SHARD_THREADS, 0, (cudaStream_t)stream>>>(emb_shard, item_b_shard, icl_shard,
                                                                         local_ids, n, n_local, out, bad_flag);
  TLSAN_CHECK_LAUNCH("k_shard_pack_rows");
  return TLSAN_OK;
}

int tlsan_shard_unpack_rows(const float* packed, const int32_t* dst_index, int64_t n, float* emb_c, float* item_b_c,
                            int32_t* icl_c, void* stream) {
  SHARD_REQUIRE(n >= 0, TLSAN_E_DIMS, "tlsan_shard_unpack_rows: n < 0");
  if (n == 0) return TLSAN_OK;
  SHARD_REQUIRE(packed && emb_c && item_b_c && icl_c, TLSAN_E_NULL, "tlsan_shard_unpack_rows: NULL argument");
  SHARD_REQUIRE(((uintptr_t)emb_c & 15) == 0 && ((uintptr_t)packed & 15) == 0, TLSAN_E_ALIGN,
                "tlsan_shard_unpack_rows: emb_c/packed must be 16-B aligned");
  k_shard_unpack_rows<<<grid9(n), SHARD_THREADS, 0, (cudaStream_t)stream>>>(packed, dst_index, n, emb_c, item_b_c,
                                                                           icl_c);
  TLSAN_CHECK_LAUNCH("k_shard_unpack_rows");
  return TLSAN_OK;
}

int tlsan_shard_pack_grads(const float* g_i, const float* g_b, const int32_t* src_index, int64_t n, float* out,
                           void* stream) {
  SHARD_REQUIRE(n >= 0, TLSAN_E_DIMS, "tlsan_shard_pack_grads: n < 0");
  if (n == 0) return TLSAN_OK;
  SHARD_REQUIRE(g_i && g_b && out, TLSAN_E_NULL, "tlsan_shard_pack_grads: NULL argument");
  SHARD_REQUIRE(((uintptr_t)g_i & 15) == 0 && ((uintptr_t)out & 15) == 0, TLSAN_E_ALIGN,
                "tlsan_shard_pack_grads: g_i/out must be 16-B aligned");
  k_shard_pack_grads<<<grid9(n), SHARD_THREADS, 0, (cudaStream_t)stream>>>(g_i, g_b, src_index, n, out);
  TLSAN_CHECK_LAUNCH("k_shard_pack_grads");
  return TLSAN_OK;
}

int tlsan_shard_accum_grads(const float* packed, const int32_t* local_ids, int64_t n, float* g_emb, float* g_b,
                            void* stream) {
  SHARD_REQUIRE(n >= 0, TLSAN_E_DIMS, "tlsan_shard_accum_grads: n < 0");
  if (n == 0) return TLSAN_OK;
  SHARD_REQUIRE(packed && local_ids && g_emb && g_b, TLSAN_E_NULL, "tlsan_shard_accum_grads: NULL argument");
  SHARD_REQUIRE(((uintptr_t)g_emb & 15) == 0 && ((uintptr_t)packed & 15) == 0, TLSAN_E_ALIGN,
                "tlsan_shard_accum_grads: g_emb/packed must be 16-B aligned");
  k_shard_accum_grads<<<grid9(n), SHARD_THREADS, 0, (cudaStream_t)stream>>>(packed, local_ids, n, g_emb, g_b);
  TLSAN_CHECK_LAUNCH("k_shard_accum_grads");
  return TLSAN_OK;
}

int tlsan_shard_apply_grads(const float* packed, const int32_t* local_ids, int64_t n, float* W_emb, float* W_b,
                            float lr, const float* scale, void* stream) {
  SHARD_REQUIRE(n >= 0, TLSAN_E_DIMS, "tlsan_shard_apply_grads: n < 0");
  if (n == 0) return TLSAN_OK;
  SHARD_REQUIRE(packed && local_ids && W_emb && W_b && scale, TLSAN_E_NULL, "tlsan_shard_apply_grads: NULL argument");
  SHARD_REQUIRE(((uintptr_t)W_emb & 15) == 0 && ((uintptr_t)packed & 15) == 0, TLSAN_E_ALIGN,
                "tlsan_shard_apply_grads: W_emb/packed must be 16-B aligned");
  k_shard_apply_grads<<<grid9(n), SHARD_THREADS, 0, (cudaStream_t)stream>>>(packed, local_ids, n, W_emb, W_b, lr, scale);
  TLSAN_CHECK_LAUNCH("k_shard_apply_grads");
  return TLSAN_OK;
}

int tlsan_route_bitmap_words(int64_t NI, int32_t world, int64_t* words) {
  SHARD_REQUIRE(NI > 0 && world > 0 && words, TLSAN_E_DIMS, "tlsan_route_bitmap_words: bad argument");
  const long long nloc = (NI + world - 1) / world;
  *words = (nloc * world + 31) / 32;
  return TLSAN_OK;
}

int tlsan_route_ids(const int32_t* const* src, int32_t* const* dst, const int64_t* n, int32_t nfields, int64_t NI,
                    int32_t world, int32_t cyclic, int32_t cap, uint32_t* bitmap, int32_t* word_prefix,
                    int32_t* send_ids, int32_t* slot_row, int32_t* counts, int32_t* overflow, int32_t phase,
                    void* stream) {
  SHARD_REQUIRE(src && dst && n && bitmap && word_prefix && send_ids && slot_row && counts && overflow, TLSAN_E_NULL,
                "tlsan_route_ids: NULL argument");
  SHARD_REQUIRE(nfields >= 1 && nfields <= 4 && NI > 0 && world > 0 && cap > 0, TLSAN_E_DIMS, "tlsan_route_ids: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  RouteFields f;
  f.nf = nfields;
  long long total = 0;
  for (int q = 0; q < 4; ++q) {
    f.src[q] = q < nfields ? src[q] : nullptr; f.dst[q] = q < nfields ? dst[q] : nullptr; f.n[q] = q < nfields ? n[q] : 0;
    total += f.n[q];
  }
  RouteGeo g;
  g.W = world; g.NI = NI; g.nloc = (int)((NI + world - 1) / world); g.mod = cyclic ? 1 : 0; g.cap = cap;
  const long long nwords = ((long long)g.nloc * world + 31) / 32;
  long long blocks = (total + 255) / 256;
  const long long capb = (long long)tlsan_num_sms() * 16;
  if (blocks > capb) blocks = capb;
  if (blocks < 1) blocks = 1;
  if (phase == 0) {            // presence bitmap + word popcounts; the caller turns word_prefix into an inclusive prefix
    TLSAN_CHECK_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)nwords * 4, st));
    k_route_mark<<<(unsigned)blocks, 256, 0, st>>>(f, g, bitmap);
    TLSAN_CHECK_LAUNCH("k_route_mark");
    k_route_popc<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(bitmap, nwords, word_prefix);
    TLSAN_CHECK_LAUNCH("k_route_popc");
    return TLSAN_OK;
  }
  // phase 1: request lists + compact batch
  TLSAN_CHECK_CUDA(cudaMemsetAsync(send_ids, 0xff, (size_t)world * cap * 4, st));
  TLSAN_CHECK_CUDA(cudaMemsetAsync(slot_row, 0xff, (size_t)world * cap * 4, st));
  k_route_emit<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(bitmap, word_prefix, nwords, g, send_ids, slot_row, counts,
                                                                 overflow);
  TLSAN_CHECK_LAUNCH("k_route_emit");
  k_route_rewrite<<<(unsigned)blocks, 256, 0, st>>>(f, g, bitmap, word_prefix);
  TLSAN_CHECK_LAUNCH("k_route_rewrite");
  return TLSAN_OK;
}

int tlsan_reduce_cate(const tlsan_dims_t* d, const tlsan_params_t* p, const float* flat, float* out, void* stream) {
  SHARD_REQUIRE(d && p && flat && out && p->cate_off && p->cate_items, TLSAN_E_NULL,
                "tlsan_reduce_cate: NULL argument");
  k_reduce_cate<<<d->NC, 256, 0, (cudaStream_t)stream>>>(d->NI, flat, p->cate_off, p->cate_items, out);
  TLSAN_CHECK_LAUNCH("k_reduce_cate");
  return TLSAN_OK;
}

int tlsan_sgd_dense(float* W, const float* g, int64_t n, float lr, float reg, const float* scale, void* stream) {
  SHARD_REQUIRE(W && scale && n >= 0, TLSAN_E_NULL, "tlsan_sgd_dense: bad argument");
  if (n == 0) return TLSAN_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)tlsan_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  k_sgd_dense<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, g, n, lr, reg, scale);
  TLSAN_CHECK_LAUNCH("k_sgd_dense");
  return TLSAN_OK;
}

int tlsan_sumsq(const float* W, int64_t n, float* partial, int32_t npartial, void* stream) {
  SHARD_REQUIRE(W && partial && npartial > 0 && n >= 0, TLSAN_E_NULL, "tlsan_sumsq: bad argument");
  k_sumsq<<<npartial, 256, 0, (cudaStream_t)stream>>>(W, n, partial);
  TLSAN_CHECK_LAUNCH("k_sumsq");
  return TLSAN_OK;
}

// L2 + clip + SGD of everything that is REPLICATED in the sharded configuration (cate_emb, user_emb, usert_emb,
// the 4449 small parameters) plus the step statistics.  `p` = the compact parameter block (p->emb = compact
// table with dims->NI item slots in front); gcate [NC][32], g_u [NU][PU] and dgrad [TLSAN_PART] are already
// summed over ranks; item_sumsq[n_item_sumsq] = partial sums of squares of the WHOLE sharded item_emb (summed
// over ranks), which enter the global norm and the l2 loss like the other tables (model.py:164-169,201).
int tlsan_shard_apply_replicated(const tlsan_dims_t* dims, const tlsan_params_t* p, const float* gcate,
                                 const float* g_u, const float* dgrad, const float* item_sumsq,
                                 int32_t n_item_sumsq, float lr, float reg, float clip_norm, void* workspace,
                                 size_t workspace_bytes, float* stats, void* stream) {
  SHARD_REQUIRE(dims && p && gcate && g_u && dgrad && item_sumsq && workspace && stats, TLSAN_E_NULL,
                "tlsan_shard_apply_replicated: NULL argument");
  SHARD_REQUIRE(p->emb && p->usert && p->dense, TLSAN_E_NULL, "tlsan_shard_apply_replicated: NULL table");
  SHARD_REQUIRE(clip_norm > 0.f && n_item_sumsq > 0, TLSAN_E_DIMS, "tlsan_shard_apply_replicated: bad scalar");
  const TlsanWs w = tlsan_ws_layout(*dims);
  SHARD_REQUIRE(workspace_bytes >= w.total + 256, TLSAN_E_WORKSPACE, "tlsan_shard_apply_replicated: workspace too small");
  char* ws = reinterpret_cast<char*>(tlsan_align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  int rc = tlsan_launch_apply_replicated(*dims, *p, w, ws, gcate, g_u, dgrad, item_sumsq, n_item_sumsq, lr, reg,
                                         clip_norm, stats, (cudaStream_t)stream);
  if (rc) return rc;
  // cate_emb <- cate_emb - lr * scale * (gcate + reg * cate_emb)
  k_sgd_dense<<<(dims->NC * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      p->emb + (size_t)dims->NI * 32, gcate, (long long)dims->NC * 32, lr, reg, stats + TLSAN_STAT_SCALE);
  TLSAN_CHECK_LAUNCH("k_sgd_dense");
  return TLSAN_OK;
}

}  // extern "C"
