// Host-side packing of the input.py 9-tuple into the int32 staging layout of tlsan_batch_t
// (the int64 -> int32 feed cast of reference model.py:210-222), multi-threaded, with the range
// validation TF's CPU gather would do (InvalidArgumentError) folded into the same pass.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/tlsan_b200.h"

void tlsan_set_error(const char* fmt, ...);
int tlsan_launch_expand_sessions(const int32_t* sl_new, const int32_t* off, const int32_t* items, int32_t* hist_i_new,
                                 int B, int S, cudaStream_t st);

namespace {
// Range check without 64-bit min/max (which baseline x86-64 cannot vectorise): OR-accumulate
//   hi64 |= v >> 31            (non-zero  <=> v < 0 or v >= 2^31)
//   neg  |= (hi - 1) - (int32)v   (sign bit <=> low word >= hi, given hi64 == 0)
// Both are plain SSE2 shifts / subtracts / ORs, so the loop runs at memory speed.
struct Chk { uint64_t hi64 = 0; uint32_t neg = 0; bool used = false; };

template <typename IdT>
void cvt(const IdT* __restrict__ src, int32_t* __restrict__ dst, int64_t n, int32_t hi, Chk& c) {
  uint64_t h = c.hi64; uint32_t g = c.neg;
  const int32_t hi1 = hi - 1;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t v = (int64_t)src[k];
    const int32_t w = (int32_t)v;
    h |= (uint64_t)v >> 31;                 // int32 input: negative values sign-extend and are caught here too
    g |= (uint32_t)(hi1 - w);
    dst[k] = w;
  }
  c.hi64 = h; c.neg = g; c.used = true;
}
// worker threads per staging call: at most 8, and the host cores are shared by the ranks of the node (torchrun
// exports LOCAL_WORLD_SIZE): 8 ranks x 32 threads on 32 cores was the largest piece of the round-1 8-GPU e2e loss
int pool_threads(int32_t nthreads) {
  if (nthreads > 0) return nthreads < 8 ? nthreads : 8;
  static int cached = 0;
  if (!cached) {
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    const char* e = getenv("LOCAL_WORLD_SIZE");
    const int ranks = e && atoi(e) > 0 ? atoi(e) : 1;
    int t = hw / ranks;
    if (const char* o = getenv("TLSAN_PACK_THREADS")) t = atoi(o);
    cached = t < 1 ? 1 : (t > 8 ? 8 : t);
  }
  return cached;
}
inline bool bad(const Chk& c) { return c.used && (c.hi64 != 0 || (c.neg >> 31) != 0); }
inline int64_t up4(int64_t n) { return (n + 3) / 4 * 4; }
}  // namespace

// ---- persistent worker pool: one wake-up per call instead of one thread spawn per worker
// (spawning 7 threads costs ~0.3 ms, a third of the whole pack at B = 65 536)
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cuda_runtime.h>

namespace {
struct Pool {
  std::mutex call_mu;                 // one staging call at a time
  std::mutex mu;
  std::condition_variable cv;
  std::vector<std::thread> th;
  const std::function<void(int)>* job = nullptr;
  uint64_t gen = 0;
  int active = 0;                     // workers that take part in the current job
  std::atomic<int> done{0};
  void ensure(int nworkers) {
    while ((int)th.size() < nworkers) {
      const int id = (int)th.size() + 1;
      th.emplace_back([this, id] {
        uint64_t seen = 0;
        for (;;) {
          const std::function<void(int)>* j;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return gen != seen; });
            seen = gen;
            j = id <= active ? job : nullptr;
          }
          if (j) { (*j)(id); done.fetch_add(1, std::memory_order_release); }
        }
      });
      th.back().detach();
    }
  }
  // run fn(0..T-1): fn(0) on the caller, the rest on the pool; returns when all are done
  void run(int T, const std::function<void(int)>& fn) {
    if (T <= 1) { fn(0); return; }
    ensure(T - 1);
    done.store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu);
      job = &fn; active = T - 1; ++gen;
    }
    cv.notify_all();
    fn(0);
    while (done.load(std::memory_order_acquire) < T - 1) std::this_thread::yield();
  }
};
Pool& pool() { static Pool* p = new Pool; return *p; }   // leaked on purpose: workers outlive static destructors

struct Layout { int64_t B, L, S, o_u, o_i, o_2, o_c, o_sl, o_sn, o_hi, o_hn, o_ht, total; };
Layout layout_of(const tlsan_dims_t* d) {
  Layout y;
  y.B = d->B; y.L = d->L; y.S = d->S;
  y.o_u = 0; y.o_i = y.o_u + up4(y.B); y.o_2 = y.o_i + up4(y.B); y.o_c = y.o_2 + up4(y.B); y.o_sl = y.o_c + up4(y.B);
  y.o_sn = y.o_sl + up4(y.B); y.o_hi = y.o_sn + up4(y.B); y.o_hn = y.o_hi + up4(y.B * y.L);
  y.o_ht = y.o_hn + up4(y.B * y.S); y.total = y.o_ht + up4(y.B * y.L);
  return y;
}

// Pack in three phases (hist_i_new | hist_t + scalars | hist_i); after phase k is complete on every thread the
// caller's thread issues the host->device copy of that phase's words, which overlaps the packing of phase k+1.
template <typename IdT>
int pack_impl(const tlsan_dims_t* d, const IdT* u, const IdT* i, const IdT* i2, const float* y,
              const IdT* hist_i, const IdT* hist_i_new, const float* hist_t, const IdT* sl,
              const IdT* sl_new, const IdT* c, int32_t* out, int64_t out_words, int32_t validate,
              int32_t nthreads, int32_t* dev, cudaStream_t st) {
  if (!d || !u || !i || !hist_i || !hist_i_new || !hist_t || !sl || !sl_new || !c || !out || (!i2 && !y)) {
    tlsan_set_error("tlsan_pack_batch_host: NULL argument");
    return TLSAN_E_NULL;
  }
  const Layout Y = layout_of(d);
  const int64_t B = Y.B, L = Y.L, S = Y.S;
  if (out_words < Y.total) {
    tlsan_set_error("tlsan_pack_batch_host: output holds %lld words, need %lld", (long long)out_words, (long long)Y.total);
    return TLSAN_E_WORKSPACE;
  }
  int T = pool_threads(nthreads);
  if (B * (2 * L + S) < 200000) T = 1;
  std::vector<Chk> k_hi(T), k_hn(T);
  Chk k_u, k_i, k_2, k_c, k_sl, k_sn;
  bool sl_low = false;                               // sl >= 1 (a first session always precedes a sample)
  std::atomic<int> phase_done[3];
  for (auto& a : phase_done) a.store(0, std::memory_order_relaxed);
  cudaError_t cerr = cudaSuccess;
  auto copy_words = [&](int64_t lo, int64_t hi) {
    if (dev && cerr == cudaSuccess)
      cerr = cudaMemcpyAsync(dev + lo, out + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, st);
  };
  auto after = [&](int t, int ph) {
    phase_done[ph].fetch_add(1, std::memory_order_release);
    if (t != 0) return;
    while (phase_done[ph].load(std::memory_order_acquire) < T) std::this_thread::yield();
    if (ph == 0) copy_words(Y.o_hn, Y.o_ht);
    else if (ph == 1) { copy_words(0, Y.o_hi); copy_words(Y.o_ht, Y.total); }
    else copy_words(Y.o_hi, Y.o_hn);
  };
  const std::function<void(int)> work = [&](int t) {
    const int64_t n1 = B * L, n2 = B * S;
    const int64_t a1 = n1 * t / T, b1 = n1 * (t + 1) / T, a2 = n2 * t / T, b2 = n2 * (t + 1) / T;
    cvt(hist_i_new + a2, out + Y.o_hn + a2, b2 - a2, d->NI, k_hn[t]);
    after(t, 0);
    memcpy(out + Y.o_ht + a1, hist_t + a1, (size_t)(b1 - a1) * 4);
    if (t == T - 1) {
      cvt(u, out + Y.o_u, B, d->NU, k_u); cvt(i, out + Y.o_i, B, d->NI, k_i); cvt(c, out + Y.o_c, B, d->NC, k_c);
      cvt(sl, out + Y.o_sl, B, (int32_t)L + 1, k_sl); cvt(sl_new, out + Y.o_sn, B, (int32_t)S + 1, k_sn);
      for (int64_t k = 0; k < B; ++k) sl_low |= sl[k] < 1;
      if (i2) cvt(i2, out + Y.o_2, B, d->NI, k_2);
      else memcpy(out + Y.o_2, y, (size_t)B * 4);
    }
    after(t, 1);
    cvt(hist_i + a1, out + Y.o_hi + a1, b1 - a1, d->NI, k_hi[t]);
    after(t, 2);
  };
  {
    std::lock_guard<std::mutex> lk(pool().call_mu);
    pool().run(T, work);
  }
  if (cerr != cudaSuccess) {
    tlsan_set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(cerr));
    return TLSAN_E_CUDA;
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

// ---- compact staging: the session matrix hist_i_new [B][S] is almost all padding (S = longest session of the batch,
// 87 % of the sessions hold one item), so it travels ragged -- offsets [B] + the valid items -- and a kernel rebuilds
// the padded matrix in HBM.  Staging layout = the packed batch layout followed by [offsets up4(B) | items up4(B*S)].
template <typename IdT>
int stage_compact_impl(const tlsan_dims_t* d, const IdT* u, const IdT* i, const IdT* i2, const float* y,
                       const IdT* hist_i, const IdT* hist_i_new, const float* hist_t, const IdT* sl,
                       const IdT* sl_new, const IdT* c, int32_t* out, int32_t* dev, int64_t words,
                       int32_t validate, int32_t nthreads, cudaStream_t st) {
  if (!d || !u || !i || !hist_i || !hist_i_new || !hist_t || !sl || !sl_new || !c || !out || (!i2 && !y)) {
    tlsan_set_error("tlsan_stage_batch_host: NULL argument");
    return TLSAN_E_NULL;
  }
  const Layout Y = layout_of(d);
  const int64_t B = Y.B, L = Y.L, S = Y.S;
  const int64_t o_off = Y.total, o_rag = o_off + up4(B), need = o_rag + up4(B * S);
  if (words < need) {
    tlsan_set_error("tlsan_stage_batch_host: buffers hold %lld words, need %lld (tlsan_stage_words)", (long long)words,
                    (long long)need);
    return TLSAN_E_WORKSPACE;
  }
  int T = pool_threads(nthreads);
  if (B * (2 * L + S) < 200000) T = 1;
  struct PerThread { Chk u, i, s2, c, sl, sn, hi, hn; int64_t nnew = 0; bool sl_low = false; };
  std::vector<PerThread> pt(T);
  std::atomic<int> phase_done[4];
  for (auto& a : phase_done) a.store(0, std::memory_order_relaxed);
  cudaError_t cerr = cudaSuccess;
  int64_t total_new = 0;
  auto copy_words = [&](int64_t lo, int64_t hi) {
    if (dev && hi > lo && cerr == cudaSuccess)          // dev == NULL: pack only (tlsan_stage_packed copies later)
      cerr = cudaMemcpyAsync(dev + lo, out + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, st);
  };
  auto arrive = [&](int ph) { phase_done[ph].fetch_add(1, std::memory_order_release); };
  auto wait_all = [&](int ph) {
    while (phase_done[ph].load(std::memory_order_acquire) < T) std::this_thread::yield();
  };
  const std::function<void(int)> work = [&](int t) {
    PerThread& P = pt[t];
    const int64_t r0 = B * t / T, r1 = B * (t + 1) / T, n = r1 - r0;
    // phase 0: the per-row scalars of this thread's rows + the number of session items they hold
    cvt(u + r0, out + Y.o_u + r0, n, d->NU, P.u); cvt(i + r0, out + Y.o_i + r0, n, d->NI, P.i);
    cvt(c + r0, out + Y.o_c + r0, n, d->NC, P.c); cvt(sl + r0, out + Y.o_sl + r0, n, (int32_t)L + 1, P.sl);
    cvt(sl_new + r0, out + Y.o_sn + r0, n, (int32_t)S + 1, P.sn);
    if (i2) cvt(i2 + r0, out + Y.o_2 + r0, n, d->NI, P.s2);
    else memcpy(out + Y.o_2 + r0, y + r0, (size_t)n * 4);
    int64_t cnt = 0;
    for (int64_t b = r0; b < r1; ++b) {
      P.sl_low |= sl[b] < 1;
      const int64_t sn = sl_new[b] < 0 ? 0 : (sl_new[b] > S ? S : sl_new[b]);
      cnt += sn;
    }
    P.nnew = cnt;
    arrive(0);
    wait_all(0);
    int64_t off = 0;
    for (int q = 0; q < t; ++q) off += pt[q].nnew;
    // phase 1: offsets + the valid session items of this thread's rows
    for (int64_t b = r0; b < r1; ++b) {
      const int64_t sn = sl_new[b] < 0 ? 0 : (sl_new[b] > S ? S : sl_new[b]);
      out[o_off + b] = (int32_t)off;
      cvt(hist_i_new + b * S, out + o_rag + off, sn, d->NI, P.hn);
      off += sn;
    }
    arrive(1);
    if (t == 0) {
      wait_all(1);
      for (int q = 0; q < T; ++q) total_new += pt[q].nnew;
      copy_words(0, Y.o_hi);                      // scalars
      copy_words(o_off, o_off + B);               // session offsets
      copy_words(o_rag, o_rag + total_new);       // session items
    }
    const int64_t n1 = B * L, a1 = n1 * t / T, b1 = n1 * (t + 1) / T;
    memcpy(out + Y.o_ht + a1, hist_t + a1, (size_t)(b1 - a1) * 4);
    arrive(2);
    if (t == 0) { wait_all(2); copy_words(Y.o_ht, Y.total); }
    cvt(hist_i + a1, out + Y.o_hi + a1, b1 - a1, d->NI, P.hi);
    arrive(3);
    if (t == 0) { wait_all(3); copy_words(Y.o_hi, Y.o_hn); }
  };
  {
    std::lock_guard<std::mutex> lk(pool().call_mu);
    pool().run(T, work);
  }
  if (cerr != cudaSuccess) {
    tlsan_set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(cerr));
    return TLSAN_E_CUDA;
  }
  // hist_i_new [B][S] <- (sl_new, offsets, items), zero padded like input.py:50-51
  if (dev) {
    int rc = tlsan_launch_expand_sessions(dev + Y.o_sn, dev + o_off, dev + o_rag, dev + Y.o_hn, (int)B, (int)S, st);
    if (rc) return rc;
  }
  if (validate) {
    PerThread A;
    for (int t = 0; t < T; ++t) {
      Chk* dst[] = {&A.u, &A.i, &A.s2, &A.c, &A.sl, &A.sn, &A.hi, &A.hn};
      const Chk* src[] = {&pt[t].u, &pt[t].i, &pt[t].s2, &pt[t].c, &pt[t].sl, &pt[t].sn, &pt[t].hi, &pt[t].hn};
      for (int k = 0; k < 8; ++k) { dst[k]->hi64 |= src[k]->hi64; dst[k]->neg |= src[k]->neg; dst[k]->used |= src[k]->used; }
      A.sl_low |= pt[t].sl_low;
    }
    struct { const char* name; bool bad; int64_t lo, hi; } chk[] = {
        {"u", bad(A.u), 0, d->NU}, {"i", bad(A.i), 0, d->NI}, {"c", bad(A.c), 0, d->NC},
        {"hist_i", bad(A.hi), 0, d->NI}, {"hist_i_new", bad(A.hn), 0, d->NI},
        {"sl", bad(A.sl) || A.sl_low, 1, L + 1}, {"sl_new", bad(A.sn), 0, S + 1}, {"second", bad(A.s2), 0, d->NI}};
    for (auto& k : chk) {
      if (k.bad) {
        tlsan_set_error("batch field %s out of range [%lld, %lld)", k.name, (long long)k.lo, (long long)k.hi);
        return TLSAN_E_DIMS;
      }
    }
  }
  return TLSAN_OK;
}
}  // namespace

extern "C" int tlsan_stage_words(const tlsan_dims_t* d, int64_t* words) {
  if (!d || !words) {
    tlsan_set_error("tlsan_stage_words: NULL argument");
    return TLSAN_E_NULL;
  }
  const Layout Y = layout_of(d);
  *words = Y.total + up4(Y.B) + up4(Y.B * Y.S);
  return TLSAN_OK;
}

// A batch that already sits in page-locked memory in the staging layout (the packed feed of tlsan_b200.input: the
// batcher writes its arrays straight into that buffer): two host->device copies -- everything in front of the padded
// session matrix, everything behind it up to the last ragged item -- and the session expansion.  No host pass at all.
extern "C" int tlsan_stage_packed(const tlsan_dims_t* d, const int32_t* pinned, int32_t* dev, int64_t words,
                                  int64_t n_new, void* stream) {
  if (!d || !pinned || !dev) {
    tlsan_set_error("tlsan_stage_packed: NULL argument");
    return TLSAN_E_NULL;
  }
  const Layout Y = layout_of(d);
  const int64_t o_off = Y.total, o_rag = o_off + up4(Y.B), need = o_rag + up4(Y.B * Y.S);
  if (words < need || n_new < 0 || n_new > Y.B * Y.S) {
    tlsan_set_error("tlsan_stage_packed: buffers hold %lld words, need %lld; n_new %lld", (long long)words,
                    (long long)need, (long long)n_new);
    return TLSAN_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(dev, pinned, (size_t)Y.o_hn * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(dev + Y.o_ht, pinned + Y.o_ht, (size_t)(o_rag + n_new - Y.o_ht) * 4, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    tlsan_set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    return TLSAN_E_CUDA;
  }
  return tlsan_launch_expand_sessions(dev + Y.o_sn, dev + o_off, dev + o_rag, dev + Y.o_hn, (int)Y.B, (int)Y.S, st);
}

extern "C" int tlsan_pack_batch_host(const tlsan_dims_t* d, const int64_t* u, const int64_t* i, const int64_t* i2,
                                     const float* y, const int64_t* hist_i, const int64_t* hist_i_new,
                                     const float* hist_t, const int64_t* sl, const int64_t* sl_new, const int64_t* c,
                                     int32_t* out, int64_t out_words, int32_t validate, int32_t nthreads) {
  return pack_impl(d, u, i, i2, y, hist_i, hist_i_new, hist_t, sl, sl_new, c, out, out_words, validate, nthreads,
                   nullptr, nullptr);
}

extern "C" int tlsan_stage_batch_host(const tlsan_dims_t* d, const int64_t* u, const int64_t* i, const int64_t* i2,
                                      const float* y, const int64_t* hist_i, const int64_t* hist_i_new,
                                      const float* hist_t, const int64_t* sl, const int64_t* sl_new, const int64_t* c,
                                      int32_t* pinned, int32_t* dev, int64_t words, int32_t validate,
                                      int32_t nthreads, void* stream) {
  return stage_compact_impl(d, u, i, i2, y, hist_i, hist_i_new, hist_t, sl, sl_new, c, pinned, dev, words, validate,
                            nthreads, (cudaStream_t)stream);
}

// The same two entry points for batches whose integer fields are ALREADY int32 (tlsan_b200/input.py emits them):
// no 64 -> 32 bit narrowing, half the host memory traffic; the range checks are unchanged.
extern "C" int tlsan_pack_batch_host_i32(const tlsan_dims_t* d, const int32_t* u, const int32_t* i, const int32_t* i2,
                                         const float* y, const int32_t* hist_i, const int32_t* hist_i_new,
                                         const float* hist_t, const int32_t* sl, const int32_t* sl_new, const int32_t* c,
                                         int32_t* out, int64_t out_words, int32_t validate, int32_t nthreads) {
  return pack_impl(d, u, i, i2, y, hist_i, hist_i_new, hist_t, sl, sl_new, c, out, out_words, validate, nthreads,
                   nullptr, nullptr);
}

extern "C" int tlsan_stage_batch_host_i32(const tlsan_dims_t* d, const int32_t* u, const int32_t* i, const int32_t* i2,
                                          const float* y, const int32_t* hist_i, const int32_t* hist_i_new,
                                          const float* hist_t, const int32_t* sl, const int32_t* sl_new,
                                          const int32_t* c, int32_t* pinned, int32_t* dev, int64_t words,
                                          int32_t validate, int32_t nthreads, void* stream) {
  return stage_compact_impl(d, u, i, i2, y, hist_i, hist_i_new, hist_t, sl, sl_new, c, pinned, dev, words, validate,
                            nthreads, (cudaStream_t)stream);
}
