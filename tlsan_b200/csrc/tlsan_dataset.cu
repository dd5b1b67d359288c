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

// hist_i_new [B][S] from its ragged transfer form (tlsan_stage_batch_host): row b holds items[off[b] .. off[b]+sl_new[b])
// then zeros -- exactly the zero padding of DataInput.__next__ (TLSAN/input.py:50-51)
__global__ void k_expand_sessions(const int* __restrict__ sl_new, const int* __restrict__ off,
                                  const int* __restrict__ items, int* __restrict__ out, int B, int S) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)B * S) return;
  const int b = (int)(g / S), j = (int)(g - (long long)b * S);
  int sn = sl_new[b];
  sn = sn < 0 ? 0 : (sn > S ? S : sn);
  out[g] = j < sn ? items[off[b] + j] : 0;
}

int tlsan_launch_expand_sessions(const int32_t* sl_new, const int32_t* off, const int32_t* items, int32_t* hist_i_new,
                                 int B, int S, cudaStream_t st) {
  const long long n = (long long)B * S;
  k_expand_sessions<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sl_new, off, items, hist_i_new, B, S);
  TLSAN_CHECK_LAUNCH("k_expand_sessions");
  return TLSAN_OK;
}
