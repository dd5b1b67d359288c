// Dataset builder on the GPU: session segmentation, the train / test split rule, the dominant
// category and the time-gap weights of TLSAN/build_dataset.py:25-73 as kernels, emitting the samples straight into
// the CSR image a DeviceDataset trains from (tlsan_dataset_t) -- in the order the reference leaves them in.
//
// Per user (one thread; a user's reviews are contiguous and time-sorted, utils/2_remap_id.py:91):
//   sessions = runs of equal day (build_dataset.py:38-46); the first goes to the history; every later session of
//   `count` items yields, while i + count < min(len, 90) - 1, a positive and a negative train sample
//   (history pos[:i], session pos[i:i+count], target pos[i+count] / neg[i+count]) and i += count (:55-62);
//   the first session that fails the test becomes the user's ONE test sample (:63-72) and the user is done.
//   u_cate = the most frequent category of the history, ties -> first to appear (pd.value_counts, :54);
//   weight of history entry t = float32(1 / n), n = sum_j [d >= 2^j], d = day[i] - day[t] + 1 (:16-21).
// What stays on the host is exactly what consumes the Python `random` stream, in the reference's call order
// (negative sampling :28-33, the test-item choice :66, the two final shuffles :75-76); it reaches the kernels as
// three arrays: neg[row], pick[user] and the position of every sample after the shuffle.
//
//   k_ds_plan   per user: #train pairs, position / size of the test session, history + session lengths of every sample
//   k_ds_emit   per user: writes its samples at their shuffled positions
#include "tlsan_common.cuh"

#define DS_MAXLEN 90

struct DsIn {
  const int* asin; const int* day; const int* item_cate; const long long* user_off; int n_users;
};

// session walk shared by both kernels: calls on_train(i, count) / on_test(i, count) like the reference loop
template <typename FT, typename FE>
__device__ __forceinline__ void ds_walk(const int* __restrict__ day, long long s, long long e, FT on_train, FE on_test) {
  const int len = (int)(e - s);
  const int valid = min(len, DS_MAXLEN);
  int i = 1;
  while (i < len && day[s + i] == day[s]) ++i;                  // first session -> history
  while (i < len) {
    int count = 1;
    while (i + count < len && day[s + i + count] == day[s + i]) ++count;
    if (i + count < valid - 1) {
      on_train(i, count);
      i += count;
    } else {
      on_test(i, count);
      return;
    }
  }
}

// counts[u] = {train pairs, has test, sum of history lengths of its train PAIRS, sum of session lengths of its pairs}
// test[u] = {i, count} of the test session (count = 0: none)
__global__ void __launch_bounds__(128) k_ds_plan(const DsIn in, int4* __restrict__ counts, int2* __restrict__ test) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= in.n_users) return;
  const long long s = in.user_off[u], e = in.user_off[u + 1];
  int pairs = 0, pre = 0, nw = 0;
  int2 t = make_int2(0, 0);
  ds_walk(in.day, s, e, [&](int i, int c) { ++pairs; pre += i; nw += c; }, [&](int i, int c) { t = make_int2(i, c); });
  counts[u] = make_int4(pairs, t.y > 0 ? 1 : 0, pre, nw);
  test[u] = t;
}

// per-sample lengths in UNSHUFFLED order (train: pair k of user u -> samples first_tr[u] + 2k, +1)
__global__ void __launch_bounds__(128) k_ds_lengths(const DsIn in, const long long* __restrict__ first_tr,
                                                    const long long* __restrict__ first_te,
                                                    int* __restrict__ len_pre_tr, int* __restrict__ len_new_tr,
                                                    int* __restrict__ len_pre_te, int* __restrict__ len_new_te) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= in.n_users) return;
  const long long s = in.user_off[u], e = in.user_off[u + 1];
  long long j = first_tr[u];
  const long long jt = first_te[u];
  ds_walk(in.day, s, e,
          [&](int i, int c) { len_pre_tr[j] = i; len_pre_tr[j + 1] = i; len_new_tr[j] = c; len_new_tr[j + 1] = c; j += 2; },
          [&](int i, int c) { len_pre_te[jt] = i; len_new_te[jt] = c > 1 ? c - 1 : c; });
}

struct DsOut {          // CSR image of one split, rows in FINAL (shuffled) order
  int* uid; const long long* pre_off; int* pre_items; float* pre_time; int* pre_gap; const long long* new_off;
  int* new_items; int* cand; int* second_i; float* second_f; int* ucate;
};

// most frequent category of cates[0..i), ties -> first to appear
struct CateTab {
  int cat[DS_MAXLEN + 6], cnt[DS_MAXLEN + 6], n;
  __device__ __forceinline__ void add(int c) {
    for (int k = 0; k < n; ++k) if (cat[k] == c) { ++cnt[k]; return; }
    if (n < DS_MAXLEN + 6) { cat[n] = c; cnt[n] = 1; ++n; }
  }
  __device__ __forceinline__ int dominant() const {
    int best = 0, bn = 0;
    for (int k = 0; k < n; ++k) if (cnt[k] > bn) { best = cat[k]; bn = cnt[k]; }
    return best;
  }
};

__device__ __forceinline__ void ds_history(const DsIn& in, const DsOut& o, long long s, int i, long long p,
                                           const float* __restrict__ lut13) {
  const long long base = o.pre_off[p];
  const int cur = in.day[s + i];
  for (int t = 0; t < i; ++t) {
    const int d = cur - in.day[s + t] + 1;
    const int n = d >= 2 ? min(12, 31 - __clz(d)) : 0;
    o.pre_items[base + t] = in.asin[s + t];
    o.pre_time[base + t] = lut13[n];
    if (o.pre_gap) o.pre_gap[base + t] = d;
  }
}

__global__ void __launch_bounds__(128) k_ds_emit(const DsIn in, const int* __restrict__ reviewer,
                                                 const long long* __restrict__ first_tr,
                                                 const long long* __restrict__ first_te,
                                                 const long long* __restrict__ pos_tr, const long long* __restrict__ pos_te,
                                                 const int* __restrict__ neg, const int* __restrict__ pick,
                                                 const float* __restrict__ lut13, const DsOut tr, const DsOut te) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= in.n_users) return;
  const long long s = in.user_off[u], e = in.user_off[u + 1];
  const int uid = reviewer[s];
  long long j = first_tr[u];
  CateTab tab; tab.n = 0;
  int added = 0;                                                  // history entries already in the category table
  auto cate_upto = [&](int i) {
    for (; added < i; ++added) tab.add(in.item_cate[in.asin[s + added]]);
    return tab.dominant();
  };
  ds_walk(in.day, s, e,
          [&](int i, int c) {
            const int uc = cate_upto(i);
            for (int q = 0; q < 2; ++q) {                         // positive, then the sampled negative
              const long long p = pos_tr[j + q];
              tr.uid[p] = uid; tr.ucate[p] = uc;
              tr.cand[p] = q == 0 ? in.asin[s + i + c] : neg[s + i + c];
              tr.second_f[p] = q == 0 ? 1.f : 0.f;
              ds_history(in, tr, s, i, p, lut13);
              const long long nb = tr.new_off[p];
              for (int t = 0; t < c; ++t) tr.new_items[nb + t] = in.asin[s + i + t];
            }
            j += 2;
          },
          [&](int i, int c) {
            const long long p = pos_te[first_te[u]];
            const int uc = cate_upto(i);
            int pos_item = in.asin[s + i], drop = -1;
            if (c > 1) {                                          // rnd.choice(new_session); new_session.remove(pos_item)
              pos_item = in.asin[s + i + pick[u]];
              for (int t = 0; t < c; ++t) if (in.asin[s + i + t] == pos_item) { drop = t; break; }
            }
            int first = 0;                                        // neg_list[pos_list.index(pos_item)]
            while (in.asin[s + first] != pos_item) ++first;
            te.uid[p] = uid; te.ucate[p] = uc; te.cand[p] = pos_item; te.second_i[p] = neg[s + first];
            ds_history(in, te, s, i, p, lut13);
            long long nb = te.new_off[p];
            for (int t = 0; t < c; ++t) if (t != drop) te.new_items[nb++] = in.asin[s + i + t];
          });
}

extern "C" {

int tlsan_ds_plan(const int32_t* day, const int64_t* user_off, int32_t n_users, int32_t* counts4, int32_t* test2,
                  void* stream) {
  if (!day || !user_off || !counts4 || !test2 || n_users <= 0) {
    tlsan_set_error("tlsan_ds_plan: bad argument");
    return TLSAN_E_NULL;
  }
  DsIn in; in.asin = nullptr; in.day = day; in.item_cate = nullptr; in.user_off = (const long long*)user_off; in.n_users = n_users;
  k_ds_plan<<<(n_users + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<int4*>(counts4),
                                                                   reinterpret_cast<int2*>(test2));
  TLSAN_CHECK_LAUNCH("k_ds_plan");
  return TLSAN_OK;
}

int tlsan_ds_lengths(const int32_t* day, const int64_t* user_off, int32_t n_users, const int64_t* first_train,
                     const int64_t* first_test, int32_t* len_pre_train, int32_t* len_new_train, int32_t* len_pre_test,
                     int32_t* len_new_test, void* stream) {
  if (!day || !user_off || !first_train || !first_test || !len_pre_train || !len_new_train || !len_pre_test ||
      !len_new_test || n_users <= 0) {
    tlsan_set_error("tlsan_ds_lengths: bad argument");
    return TLSAN_E_NULL;
  }
  DsIn in; in.asin = nullptr; in.day = day; in.item_cate = nullptr; in.user_off = (const long long*)user_off; in.n_users = n_users;
  k_ds_lengths<<<(n_users + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      in, (const long long*)first_train, (const long long*)first_test, len_pre_train, len_new_train, len_pre_test,
      len_new_test);
  TLSAN_CHECK_LAUNCH("k_ds_lengths");
  return TLSAN_OK;
}

int tlsan_ds_emit(const int32_t* reviewer, const int32_t* asin, const int32_t* day, const int32_t* item_cate,
                  const int64_t* user_off, int32_t n_users, const int64_t* first_train, const int64_t* first_test,
                  const int64_t* pos_train, const int64_t* pos_test, const int32_t* neg, const int32_t* pick,
                  const float* lut13, const tlsan_dataset_t* train, const tlsan_dataset_t* test, int32_t* train_gap,
                  int32_t* test_gap, void* stream) {
  if (!reviewer || !asin || !day || !item_cate || !user_off || !first_train || !first_test || !pos_train || !pos_test ||
      !neg || !pick || !lut13 || !train || !test || n_users <= 0) {
    tlsan_set_error("tlsan_ds_emit: bad argument");
    return TLSAN_E_NULL;
  }
  DsIn in; in.asin = asin; in.day = day; in.item_cate = item_cate; in.user_off = (const long long*)user_off; in.n_users = n_users;
  auto out = [](const tlsan_dataset_t* d, int32_t* gap) {
    DsOut o;
    o.uid = const_cast<int*>(d->uid); o.pre_off = (const long long*)d->pre_off; o.pre_items = const_cast<int*>(d->pre_items);
    o.pre_time = const_cast<float*>(d->pre_time); o.pre_gap = gap; o.new_off = (const long long*)d->new_off;
    o.new_items = const_cast<int*>(d->new_items); o.cand = const_cast<int*>(d->cand);
    o.second_i = const_cast<int*>(d->second_i); o.second_f = const_cast<float*>(d->second_f); o.ucate = const_cast<int*>(d->ucate);
    return o;
  };
  k_ds_emit<<<(n_users + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      in, reviewer, (const long long*)first_train, (const long long*)first_test, (const long long*)pos_train,
      (const long long*)pos_test, neg, pick, lut13, out(train, train_gap), out(test, test_gap));
  TLSAN_CHECK_LAUNCH("k_ds_emit");
  return TLSAN_OK;
}

}  // extern "C"
