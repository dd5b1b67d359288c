// Deterministic grouping of gradient occurrences by table row (north-star item 5):
// occurrence keys -> stable LSD radix sort (8-bit digits) -> segment offsets per row.
// The sort is stable and keyed only by the row id, so inside a segment occurrences stay in
// (sample, slot) order and the later segmented reduce adds them in one fixed order.
//
// Occurrence slots of sample b (id = b*SP + j, SP = next power of two >= L+S+3):
//   j <  L        long-term token j        key = hist_i[b][j]        (valid iff j < sl[b])
//   j <  L+S      short-term token j-L     key = hist_i_new[b][j-L]  (valid iff j-L < sl_new[b])
//   j == L+S      candidate                key = i[b]
//   j == L+S+1    u_cate (cate row only)   key = NI + c[b]
//   j == L+S+2    user                     key = NI + NC + u[b]
// Keys live in the unified row space of tlsan_params_t::emb.
#include <stdlib.h>
#include <string.h>
#include "tlsan_common.cuh"

// key of occurrence id `g` straight from the batch (pass 1 never materialises the key array)
struct KeySrc {
  int B, L, S, spsh, NI, NC, NR;
  const int *u, *cand, *c, *sl, *sl_new, *hist_i, *hist_i_new;
};
// Branch-free on purpose (selects + one or two loads): the divergent five-way version was ~100 instructions per key
// and most of the instructions of both pass-0 kernels.  g < B * SP < 2^31 (check_dims).
__device__ __forceinline__ int occ_key(const KeySrc& k, int g) {
  const int b = g >> k.spsh, j = g & ((1 << k.spsh) - 1);
  const int LS = k.L + k.S, q = j - LS;               // q = 0 candidate | 1 u_cate | 2 user ; padding slots beyond
  const bool lng = j < k.L, hist = j < LS, scal = (unsigned)q <= 2u;
  const int jj = lng ? j : j - k.L;
  const int lim = hist ? __ldg((lng ? k.sl : k.sl_new) + b) : 0;
  const bool hv = hist && jj < lim;                   // a real history / session entry
  const int* hp = (lng ? k.hist_i : k.hist_i_new) + (size_t)b * (lng ? k.L : k.S) + jj;
  const int* sp = (q == 0 ? k.cand : q == 1 ? k.c : k.u) + b;
  const int add = hv || q == 0 ? 0 : q == 1 ? k.NI : k.NI + k.NC;
  int key = TLSAN_INVALID_KEY;
  if (hv || scal) key = __ldg(hv ? hp : sp) + add;
  // ids are range-checked when a batch is staged; a batch buffer that was freed and reused while a presort
  // announced for it was still queued must not turn into out-of-range segment writes either
  if ((unsigned)key >= (unsigned)k.NR) key = TLSAN_INVALID_KEY;
  return key;
}

// Radix pass geometry: a CTA of 16 warps owns TLSAN_SORT_CTA_KEYS = 5120 consecutive keys; a warp
// owns 320 of them, loaded up front as 10 coalesced loads per lane (so the serial ranking loop
// below runs from registers).  Histograms are kept per CTA: hist[digit][cta].
#define SORT_WARPS 16
#define SORT_KPL 10    // keys per lane: 2 M occurrence slots (B 65 536 x 32) = 410 CTAs, one resident wave at 3 CTAs per SM
#define SORT_KPW (32 * SORT_KPL)
static_assert(SORT_WARPS * SORT_KPW == TLSAN_SORT_CTA_KEYS, "workspace layout and radix pass disagree on the CTA tile");

template <bool FROM_BATCH>
__device__ __forceinline__ void load_keys(const int* __restrict__ keys, const KeySrc& src, long long n,
                                          long long wbase, int lane, int (&k)[SORT_KPL]) {
#pragma unroll
  for (int it = 0; it < SORT_KPL; ++it) {
    const long long idx = wbase + it * 32 + lane;
    k[it] = idx < n ? (FROM_BATCH ? occ_key(src, (int)idx) : keys[idx]) : TLSAN_INVALID_KEY;
  }
}

template <bool FROM_BATCH>
__global__ void __launch_bounds__(32 * SORT_WARPS) k_radix_hist(const int* __restrict__ keys, const KeySrc src,
                                                                 long long ncap, const int* __restrict__ nvalid,
                                                                 int shift, int nblk, int* __restrict__ hist) {
  __shared__ int cnt[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long n = nvalid ? (long long)*nvalid : ncap;
  int k[SORT_KPL];
  load_keys<FROM_BATCH>(keys, src, n, ((long long)blockIdx.x * SORT_WARPS + warp) * SORT_KPW, lane, k);
#pragma unroll
  for (int it = 0; it < SORT_KPL; ++it)
    if (k[it] != TLSAN_INVALID_KEY) atomicAdd(&cnt[(k[it] >> shift) & 255], 1);
  __syncthreads();
  if (threadIdx.x < 256) hist[(size_t)threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}

// Exclusive scan of every digit's row hist[d][0..nblk) in place (one warp per digit, coalesced
// 32-wide steps) and the digit totals -> tot[d].  The scatter kernel adds the scan over digits.
__global__ void __launch_bounds__(256) k_radix_scan_rows(int* __restrict__ hist, int nblk, int* __restrict__ tot) {
  const int lane = threadIdx.x & 31;
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5);
  int* row = hist + (size_t)d * nblk;
  int run = 0;
  for (int base = 0; base < nblk; base += 32) {
    const int i = base + lane;
    const int v = i < nblk ? row[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (i < nblk) row[i] = run + x - v;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) tot[d] = run;
}

// One pass = count + rank in ONE sweep over the warp's keys: in index order a key's rank among the warp's keys of the
// same digit is (matching keys seen in earlier iterations) + (matching lower lanes of this iteration); the running
// count per digit lives in the warp's shared-memory row and ends up as the warp's digit histogram.  The CTA turns the
// 16 histograms into offsets, the keys are re-ordered by digit INSIDE the CTA through shared memory, and thread i then
// writes the i-th key of that order: keys of one digit go to consecutive output positions, so the global stores are
// runs of ~6 (pass 0) / ~17 (pass 1) keys instead of single 4-byte writes to as many sectors (the scattered writes,
// 2.2 M sectors per pass, bounded the first versions: 34 / 24 us per pass).
template <bool FROM_BATCH>
__global__ void __launch_bounds__(32 * SORT_WARPS, FROM_BATCH ? 3 : 2) k_radix_scatter(
    const int* __restrict__ keys_in, const KeySrc src, const int* __restrict__ vals_in, long long ncap,
    const int* __restrict__ nvalid, int* __restrict__ nvalid_out, int shift, int nblk, const int* __restrict__ hist,
    const int* __restrict__ tot, int* __restrict__ keys_out, int* __restrict__ vals_out,
    int* __restrict__ inv_out /* last pass only: occurrence id -> sorted rank */) {
  constexpr int TILE = SORT_WARPS * SORT_KPW;
  __shared__ int offkey[TILE > SORT_WARPS * 256 ? TILE : SORT_WARPS * 256];
  int (*off)[256] = reinterpret_cast<int (*)[256]>(offkey);   // per-warp digit counts -> offsets; then the staged keys
  __shared__ int sval[TILE];
  __shared__ int gstart[256];                 // output position of the CTA's first key of digit d
  __shared__ int cstart[257];                 // position of that key in the CTA's digit-ordered tile; [256] = #keys
  __shared__ int wtot[8], wtot2[8];
  int* skey = offkey;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n = nvalid ? (long long)*nvalid : ncap;
  // a CTA past the valid keys (later passes run on the compacted list) has nothing to rank or write
  if (blockIdx.x > 0 && (long long)blockIdx.x * SORT_WARPS * SORT_KPW >= n) return;
  for (int d = lane; d < 256; d += 32) off[warp][d] = 0;
  const long long wbase = ((long long)blockIdx.x * SORT_WARPS + warp) * SORT_KPW;
  int k[SORT_KPL], v[FROM_BATCH ? 1 : SORT_KPL], rk[SORT_KPL];   // pass 0 (from the batch): value = occurrence id = index
  load_keys<FROM_BATCH>(keys_in, src, n, wbase, lane, k);
  if (!FROM_BATCH) {
#pragma unroll
    for (int it = 0; it < SORT_KPL; ++it) {
      const long long idx = wbase + it * 32 + lane;
      v[it] = (vals_in && k[it] != TLSAN_INVALID_KEY) ? vals_in[idx] : (int)idx;
    }
  }
  // digit totals: loaded early, scanned below
  const int my_tot = threadIdx.x < 256 ? tot[threadIdx.x] : 0;
  const int my_hist = threadIdx.x < 256 ? hist[(size_t)threadIdx.x * nblk + blockIdx.x] : 0;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < SORT_KPL; ++it) {      // in index order: the pass is stable
    const bool act = k[it] != TLSAN_INVALID_KEY;
    const int digit = (k[it] >> shift) & 255;
    // lanes holding the same digit: eight ballots, one per digit bit (match.any is far slower than that on sm_100a:
    // the pass was bound by it -- ~150 cycles each, one at a time per SM)
    unsigned m = __ballot_sync(0xffffffffu, act);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
      const bool one = (digit >> bit) & 1;
      const unsigned bal = __ballot_sync(0xffffffffu, one);
      m &= one ? bal : ~bal;
    }
    const int below = __popc(m & lt);
    const int seen = act ? off[warp][digit] : 0;
    rk[it] = seen + below;
    __syncwarp();
    if (act && below == 0) off[warp][digit] = seen + __popc(m);
    __syncwarp();
  }
  __syncthreads();
  int ctot = 0;
  if (threadIdx.x < 256) {  // thread d: the warps' counts of digit d -> offsets inside the CTA's run of d
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const int c = off[w][threadIdx.x];
      off[w][threadIdx.x] = ctot;
      ctot += c;
    }
    // two exclusive scans over the 256 digits (8 warps x 32): the global digit totals and this CTA's
    int x = my_tot, y = ctot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int xs = __shfl_up_sync(0xffffffffu, x, o), ys = __shfl_up_sync(0xffffffffu, y, o);
      if (lane >= o) { x += xs; y += ys; }
    }
    if (lane == 31) { wtot[warp] = x; wtot2[warp] = y; }
    gstart[threadIdx.x] = x - my_tot;
    cstart[threadIdx.x] = y - ctot;
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    int pre = 0, pre2 = 0;
    for (int w = 0; w < warp; ++w) { pre += wtot[w]; pre2 += wtot2[w]; }
    if (blockIdx.x == 0 && threadIdx.x == 255) *nvalid_out = pre + gstart[255] + my_tot;
    gstart[threadIdx.x] += pre + my_hist;      // digit base + this CTA's base inside the digit
    cstart[threadIdx.x] += pre2;
    if (threadIdx.x == 255) cstart[256] = cstart[255] + ctot;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < SORT_KPL; ++it) {      // position inside the CTA's digit-ordered tile
    const int digit = (k[it] >> shift) & 255;
    if (k[it] != TLSAN_INVALID_KEY) rk[it] += cstart[digit] + off[warp][digit];
  }
  __syncthreads();                             // the offsets are consumed: their storage becomes the key stage
#pragma unroll
  for (int it = 0; it < SORT_KPL; ++it) {
    if (k[it] != TLSAN_INVALID_KEY) {
      skey[rk[it]] = k[it];
      sval[rk[it]] = FROM_BATCH ? (int)(wbase + it * 32 + lane) : v[it];
    }
  }
  __syncthreads();
  const int total = cstart[256];
  for (int i = threadIdx.x; i < total; i += 32 * SORT_WARPS) {
    const int key = skey[i], val = sval[i];
    const int digit = (key >> shift) & 255;
    const int pos = gstart[digit] + (i - cstart[digit]);
    keys_out[pos] = key;
    vals_out[pos] = val;
    if (inv_out) inv_out[val] = pos;
  }
}

// seg_off[r] = first sorted position with key >= r, r in [0, NR]: one thread per ROW, binary search over the sorted
// keys (a per-key thread that fills the rows between two keys serialises on long runs of untouched rows -- the compact
// table of the row-sharded configuration has runs of 10^5)
__global__ void k_seg_bounds(const int* __restrict__ keys, const int* __restrict__ nvalid, int NR,
                             int* __restrict__ seg_off) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > NR) return;
  int lo = 0, hi = *nvalid;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < r) lo = mid + 1; else hi = mid;
  }
  seg_off[r] = lo;
}

// which ping-pong buffer holds the sorted occurrence ids after the last pass (pass k writes b, a, b, ...)
const int32_t* tlsan_sorted_vals(const TlsanWs& w, char* ws) {
  int bits = 1;
  while ((1ll << bits) < (long long)w.NR) ++bits;
  const int passes = (bits + 7) / 8;
  return reinterpret_cast<const int32_t*>(ws + ((passes & 1) ? w.vals_b : w.vals_a));
}

int tlsan_launch_sort(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, const TlsanWs& w,
                      char* ws, const int32_t** sorted_vals, cudaEvent_t ranks_ready, cudaStream_t st) {
  int* keys_a = reinterpret_cast<int*>(ws + w.keys_a);
  int* keys_b = reinterpret_cast<int*>(ws + w.keys_b);
  int* vals_a = reinterpret_cast<int*>(ws + w.vals_a);
  int* vals_b = reinterpret_cast<int*>(ws + w.vals_b);
  int* hist = reinterpret_cast<int*>(ws + w.hist);
  int* nvalid = reinterpret_cast<int*>(ws + w.nvalid);
  int* seg_off = reinterpret_cast<int*>(ws + w.seg_off);
  const long long nocc = w.nocc;
  KeySrc src;
  src.B = d.B; src.L = d.L; src.S = d.S; src.spsh = w.SPSH; src.NI = d.NI; src.NC = d.NC; src.NR = w.NR;
  src.u = b.u; src.cand = b.i; src.c = b.c; src.sl = b.sl; src.sl_new = b.sl_new;
  src.hist_i = b.hist_i; src.hist_i_new = b.hist_i_new;
  int bits = 1;
  while ((1ll << bits) < (long long)w.NR) ++bits;
  const int passes = (bits + 7) / 8;
  const int nblk = (int)((nocc + SORT_WARPS * SORT_KPW - 1) / (SORT_WARPS * SORT_KPW));
  int* tot = hist + (size_t)256 * nblk;
  const int* kin = nullptr; const int* vin = nullptr;
  int* kout = keys_b; int* vout = vals_b;
  for (int pass = 0; pass < passes; ++pass) {
    const int* nv = pass == 0 ? nullptr : nvalid;
    if (pass == 0) k_radix_hist<true><<<nblk, 32 * SORT_WARPS, 0, st>>>(kin, src, nocc, nv, 8 * pass, nblk, hist);
    else k_radix_hist<false><<<nblk, 32 * SORT_WARPS, 0, st>>>(kin, src, nocc, nv, 8 * pass, nblk, hist);
    TLSAN_CHECK_LAUNCH("k_radix_hist");
    k_radix_scan_rows<<<32, 256, 0, st>>>(hist, nblk, tot);
    TLSAN_CHECK_LAUNCH("k_radix_scan_rows");
    int* inv = pass == passes - 1 ? reinterpret_cast<int*>(ws + w.inv) : nullptr;
    if (pass == 0)
      k_radix_scatter<true><<<nblk, 32 * SORT_WARPS, 0, st>>>(kin, src, vin, nocc, nv, nvalid, 8 * pass, nblk, hist,
                                                             tot, kout, vout, inv);
    else
      k_radix_scatter<false><<<nblk, 32 * SORT_WARPS, 0, st>>>(kin, src, vin, nocc, nv, nvalid, 8 * pass, nblk, hist,
                                                              tot, kout, vout, inv);
    TLSAN_CHECK_LAUNCH("k_radix_scatter");
    kin = kout; vin = vout;
    if (kout == keys_b) { kout = keys_a; vout = vals_a; } else { kout = keys_b; vout = vals_b; }
  }
  // the sorted ranks (inv) are complete here: the gradient-row writers need nothing more, only the row reduce reads
  // the segment bounds
  if (ranks_ready) TLSAN_CHECK_CUDA(cudaEventRecord(ranks_ready, st));
  k_seg_bounds<<<(unsigned)((w.NR + 1 + 255) / 256), 256, 0, st>>>(kin, nvalid, w.NR, seg_off);
  TLSAN_CHECK_LAUNCH("k_seg_bounds");
  *sorted_vals = vin;
  return TLSAN_OK;
}
