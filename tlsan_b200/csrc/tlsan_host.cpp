// Host-side packing of the input.py 9-tuple into the int32 staging layout of tlsan_batch_t
// (the int64 -> int32 feed cast of reference model.py:210-222), multi-threaded, with the range
// validation TF's CPU gather would do (InvalidArgumentError) folded into the same pass.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../../include/tlsan_b200.h"

void tlsan_set_error(const char* fmt, ...);

namespace {
// Range check without 64-bit min/max (which baseline x86-64 cannot vectorise): OR-accumulate
//   hi64 |= v >> 31            (non-zero  <=> v < 0 or v >= 2^31)
//   neg  |= (hi - 1) - (int32)v   (sign bit <=> low word >= hi, given hi64 == 0)
// Both are plain SSE2 shifts / subtracts / ORs, so the loop runs at memory speed.
struct Chk { uint64_t hi64 = 0; uint32_t neg = 0; bool used = false; };

void cvt(const int64_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n, int32_t hi, Chk& c) {
  uint64_t h = c.hi64; uint32_t g = c.neg;
  const int32_t hi1 = hi - 1;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = src[k];
    const int32_t w = (int32_t)v;
    h |= (uint64_t)v >> 31;
    g |= (uint32_t)(hi1 - w);
    dst[k] = w;
  }
  c.hi64 = h; c.neg = g; c.used = true;
}
inline bool bad(const Chk& c) { return c.used && (c.hi64 != 0 || (c.neg >> 31) != 0); }
inline int64_t up4(int64_t n) { return (n + 3) / 4 * 4; }
}  // namespace

extern "C" int tlsan_pack_batch_host(const tlsan_dims_t* d, const int64_t* u, const int64_t* i, const int64_t* i2,
                                     const float* y, const int64_t* hist_i, const int64_t* hist_i_new,
                                     const float* hist_t, const int64_t* sl, const int64_t* sl_new, const int64_t* c,
                                     int32_t* out, int64_t out_words, int32_t validate, int32_t nthreads) {
  if (!d || !u || !i || !hist_i || !hist_i_new || !hist_t || !sl || !sl_new || !c || !out || (!i2 && !y)) {
    tlsan_set_error("tlsan_pack_batch_host: NULL argument");
    return TLSAN_E_NULL;
  }
  const int64_t B = d->B, L = d->L, S = d->S;
  const int64_t o_u = 0, o_i = o_u + up4(B), o_2 = o_i + up4(B), o_c = o_2 + up4(B), o_sl = o_c + up4(B),
                o_sn = o_sl + up4(B), o_hi = o_sn + up4(B), o_hn = o_hi + up4(B * L), o_ht = o_hn + up4(B * S),
                total = o_ht + up4(B * L);
  if (out_words < total) {
    tlsan_set_error("tlsan_pack_batch_host: output holds %lld words, need %lld", (long long)out_words, (long long)total);
    return TLSAN_E_WORKSPACE;
  }
  int T = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > 8) T = 8;
  if (B * (2 * L + S) < 200000) T = 1;
  std::vector<Chk> k_hi(T), k_hn(T);
  Chk k_u, k_i, k_2, k_c, k_sl, k_sn;
  bool sl_low = false;                               // sl >= 1 (a first session always precedes a sample)
  auto work = [&](int t) {
    const int64_t n1 = B * L, n2 = B * S;
    const int64_t a1 = n1 * t / T, b1 = n1 * (t + 1) / T, a2 = n2 * t / T, b2 = n2 * (t + 1) / T;
    cvt(hist_i + a1, out + o_hi + a1, b1 - a1, d->NI, k_hi[t]);
    cvt(hist_i_new + a2, out + o_hn + a2, b2 - a2, d->NI, k_hn[t]);
    memcpy(out + o_ht + a1, hist_t + a1, (size_t)(b1 - a1) * 4);
    if (t == 0) {
      cvt(u, out + o_u, B, d->NU, k_u); cvt(i, out + o_i, B, d->NI, k_i); cvt(c, out + o_c, B, d->NC, k_c);
      cvt(sl, out + o_sl, B, (int32_t)L + 1, k_sl); cvt(sl_new, out + o_sn, B, (int32_t)S + 1, k_sn);
      for (int64_t k = 0; k < B; ++k) sl_low |= sl[k] < 1;
      if (i2) cvt(i2, out + o_2, B, d->NI, k_2);
      else memcpy(out + o_2, y, (size_t)B * 4);
    }
  };
  if (T == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  if (validate) {
    Chk hh, hn;
    for (int t = 0; t < T; ++t) {
      hh.hi64 |= k_hi[t].hi64; hh.neg |= k_hi[t].neg; hh.used = true;
      hn.hi64 |= k_hn[t].hi64; hn.neg |= k_hn[t].neg; hn.used = true;
    }
    struct { const char* name; bool bad; int64_t lo, hi; } chk[] = {
        {"u", bad(k_u), 0, d->NU}, {"i", bad(k_i), 0, d->NI}, {"c", bad(k_c), 0, d->NC},
        {"hist_i", bad(hh), 0, d->NI}, {"hist_i_new", bad(hn), 0, d->NI},
        {"sl", bad(k_sl) || sl_low, 1, L + 1}, {"sl_new", bad(k_sn), 0, S + 1}, {"second", bad(k_2), 0, d->NI}};
    for (auto& k : chk) {
      if (k.bad) {
        tlsan_set_error("batch field %s out of range [%lld, %lld)", k.name, (long long)k.lo, (long long)k.hi);
        return TLSAN_E_DIMS;
      }
    }
  }
  return TLSAN_OK;
}
