// Device-resident dataset + batch assembly: the per-sample Python loops of the
// reference batcher (TLSAN/input.py:17-54, 70-107) as one kernel over a CSR image of the samples
// that lives in HBM.  Output = the packed int32 staging layout of tlsan_pack_batch_host, so the
// result feeds tlsan_train_step / tlsan_score directly; bit-exact vs input.py (tests/test_gpu_dataset.py).
#include "tlsan_common.cuh"

static inline long long up4ll(long long n) { return (n + 3) / 4 * 4; }

// one thread per (row, column): column t serves hist_i/hist_t (t < L), hist_i_new (t < S), scalars (t == 0)
__global__ void k_collate(tlsan_dataset_t ds, const int* __restrict__ idx, int B, int L, int S, int is_test,
                          int* __restrict__ out, long long o_u, long long o_i, long long o_2, long long o_c,
                          long long o_sl, long long o_sn, long long o_hi, long long o_hn, long long o_ht) {
  const int W = L > S ? L : S;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)B * W) return;
  const int b = (int)(g / W), t = (int)(g - (long long)b * W);
  const long long r = idx[b];
  const long long p0 = ds.pre_off[r], p1 = ds.pre_off[r + 1];
  const long long n0 = ds.new_off[r], n1 = ds.new_off[r + 1];
  const int len = (int)(p1 - p0);
  const int sl = len < L ? len : L;                 // input.py:30
  const int sn = (int)(n1 - n0);                    // input.py:31
  if (t < L) {                                      // keep the LAST k entries, left aligned (input.py:39-49)
    const bool ok = t < sl;
    const long long src = p1 - sl + t;
    out[o_hi + (long long)b * L + t] = ok ? ds.pre_items[src] : 0;
    reinterpret_cast<float*>(out)[o_ht + (long long)b * L + t] = ok ? ds.pre_time[src] : 0.f;
  }
  if (t < S) out[o_hn + (long long)b * S + t] = t < sn ? ds.new_items[n0 + t] : 0;   // input.py:50-51
  if (t == 0) {
    out[o_u + b] = ds.uid[r];
    out[o_i + b] = ds.cand[r];
    if (is_test) out[o_2 + b] = ds.second_i[r];
    else reinterpret_cast<float*>(out)[o_2 + b] = ds.second_f[r];
    out[o_c + b] = ds.ucate[r];
    out[o_sl + b] = sl;
    out[o_sn + b] = sn < S ? sn : S;
  }
}

extern "C" int tlsan_collate(const tlsan_dataset_t* ds, const int32_t* idx, int32_t B, int32_t L, int32_t S,
                             int32_t is_test, int32_t* out, int64_t out_words, void* stream) {
  if (!ds || !idx || !out || !ds->uid || !ds->pre_off || !ds->pre_items || !ds->pre_time || !ds->new_off ||
      !ds->new_items || !ds->cand || !ds->ucate || (is_test ? !ds->second_i : !ds->second_f)) {
    tlsan_set_error("tlsan_collate: NULL argument");
    return TLSAN_E_NULL;
  }
  if (B <= 0 || L < 1 || L > TLSAN_MAX_L || S < 1) {
    tlsan_set_error("tlsan_collate: bad dims B=%d L=%d S=%d", B, L, S);
    return TLSAN_E_DIMS;
  }
  const long long o_u = 0, o_i = o_u + up4ll(B), o_2 = o_i + up4ll(B), o_c = o_2 + up4ll(B), o_sl = o_c + up4ll(B),
                  o_sn = o_sl + up4ll(B), o_hi = o_sn + up4ll(B), o_hn = o_hi + up4ll((long long)B * L),
                  o_ht = o_hn + up4ll((long long)B * S), total = o_ht + up4ll((long long)B * L);
  if (out_words < total) {
    tlsan_set_error("tlsan_collate: output holds %lld words, need %lld", (long long)out_words, total);
    return TLSAN_E_WORKSPACE;
  }
  const int W = L > S ? L : S;
  const long long n = (long long)B * W;
  k_collate<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*ds, idx, B, L, S, is_test, out, o_u, o_i, o_2,
                                                                          o_c, o_sl, o_sn, o_hi, o_hn, o_ht);
  TLSAN_CHECK_LAUNCH("k_collate");
  return TLSAN_OK;
}

// ------------------------------------------------------------------ helpers for row-sharded tables
// (item_emb / item_b / icl sharded by row over the ranks, all-to-all exchange)

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

// W <- W - lr * ((g + reg * W) * scale), g optional (NULL = rows without a sparse gradient);
// also accumulates sum(W_old^2) partials when sq != NULL (fixed grid, fixed order)
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

extern "C" int tlsan_reduce_cate(const tlsan_dims_t* d, const tlsan_params_t* p, const float* flat, float* out,
                                 void* stream) {
  if (!d || !p || !flat || !out || !p->cate_off || !p->cate_items) {
    tlsan_set_error("tlsan_reduce_cate: NULL argument");
    return TLSAN_E_NULL;
  }
  k_reduce_cate<<<d->NC, 256, 0, (cudaStream_t)stream>>>(d->NI, flat, p->cate_off, p->cate_items, out);
  TLSAN_CHECK_LAUNCH("k_reduce_cate");
  return TLSAN_OK;
}

extern "C" int tlsan_sgd_dense(float* W, const float* g, int64_t n, float lr, float reg, const float* scale,
                               void* stream) {
  if (!W || !scale || n < 0) {
    tlsan_set_error("tlsan_sgd_dense: bad argument");
    return TLSAN_E_NULL;
  }
  if (n == 0) return TLSAN_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)tlsan_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  k_sgd_dense<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, g, n, lr, reg, scale);
  TLSAN_CHECK_LAUNCH("k_sgd_dense");
  return TLSAN_OK;
}

extern "C" int tlsan_sumsq(const float* W, int64_t n, float* partial, int32_t npartial, void* stream) {
  if (!W || !partial || npartial <= 0) {
    tlsan_set_error("tlsan_sumsq: bad argument");
    return TLSAN_E_NULL;
  }
  k_sumsq<<<npartial, 256, 0, (cudaStream_t)stream>>>(W, n, partial);
  TLSAN_CHECK_LAUNCH("k_sumsq");
  return TLSAN_OK;
}
