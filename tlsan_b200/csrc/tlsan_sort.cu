// Deterministic grouping of gradient occurrences by table row (north-star item 5):
// occurrence keys -> stable LSD radix sort (8-bit digits) -> segment offsets per row.
// The sort is stable and keyed only by the row id, so inside a segment occurrences stay in
// (sample, slot) order and the later segmented reduce adds them in one fixed order.
//
// Occurrence slots of sample b (id = b*SLOTS + j):
//   j <  L        long-term token j        key = hist_i[b][j]        (valid iff j < sl[b])
//   j <  L+S      short-term token j-L     key = hist_i_new[b][j-L]  (valid iff j-L < sl_new[b])
//   j == L+S      candidate                key = i[b]
//   j == L+S+1    u_cate (cate row only)   key = NI + c[b]
//   j == L+S+2    user                     key = NI + NC + u[b]
// Keys live in the unified row space of tlsan_params_t::emb.
#include "tlsan_common.cuh"

__global__ void k_build_keys(int B, int L, int S, int NI, int NC, const int* __restrict__ u,
                             const int* __restrict__ cand, const int* __restrict__ c, const int* __restrict__ sl,
                             const int* __restrict__ sl_new, const int* __restrict__ hist_i,
                             const int* __restrict__ hist_i_new, int* __restrict__ keys) {
  const int SLOTS = L + S + 3;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)B * SLOTS) return;
  const int b = (int)(g / SLOTS), j = (int)(g - (long long)b * SLOTS);
  int key;
  if (j < L) key = j < sl[b] ? hist_i[(size_t)b * L + j] : TLSAN_INVALID_KEY;
  else if (j < L + S) key = (j - L) < sl_new[b] ? hist_i_new[(size_t)b * S + (j - L)] : TLSAN_INVALID_KEY;
  else if (j == L + S) key = cand[b];
  else if (j == L + S + 1) key = NI + c[b];
  else key = NI + NC + u[b];
  keys[g] = key;
}

// One warp = one chunk of TLSAN_SORT_CHUNK consecutive keys, 8 warps (chunks) per CTA.
// Histograms are kept per CTA: hist[digit][cta]; the scatter kernel recounts its 8 chunks
// to split the CTA's range among its warps (keys are L2-resident, the recount is cheap).
__device__ __forceinline__ void count_chunk(const int* __restrict__ keys, long long n, int chunk, int shift,
                                            int lane, int* cnt /* [256] smem, zeroed */) {
  const long long base = (long long)chunk * TLSAN_SORT_CHUNK;
#pragma unroll 4
  for (int it = 0; it < TLSAN_SORT_CHUNK / 32; ++it) {
    const long long idx = base + it * 32 + lane;
    const int key = idx < n ? keys[idx] : TLSAN_INVALID_KEY;
    if (key != TLSAN_INVALID_KEY) atomicAdd(&cnt[(key >> shift) & 255], 1);
  }
}

__global__ void __launch_bounds__(256) k_radix_hist(const int* __restrict__ keys, long long ncap,
                                                    const int* __restrict__ nvalid, int shift, int nchunks,
                                                    int nblk, int* __restrict__ hist) {
  __shared__ int cnt[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = blockIdx.x * 8 + warp;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  const long long n = nvalid ? (long long)*nvalid : ncap;
  if (chunk < nchunks) count_chunk(keys, n, chunk, shift, lane, cnt[warp]);
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += cnt[w][threadIdx.x];
  hist[(size_t)threadIdx.x * nblk + blockIdx.x] = t;
}

// exclusive scan of hist[n] in place (single CTA of 1024 threads, contiguous span per thread)
__global__ void __launch_bounds__(1024) k_radix_scan(int* __restrict__ hist, int n, int* __restrict__ total_out) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += hist[i];
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  int run = (warp > 0 ? wsum[warp - 1] : 0) + x - s;   // exclusive prefix of this thread's span
  for (int i = lo; i < hi; ++i) {
    const int v = hist[i];
    hist[i] = run;
    run += v;
  }
  if (threadIdx.x == 1023) *total_out = wsum[31];
}

__global__ void __launch_bounds__(256) k_radix_scatter(const int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                       long long ncap, const int* __restrict__ nvalid, int shift,
                                                       int nchunks, int nblk, const int* __restrict__ hist,
                                                       int* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int off[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = blockIdx.x * 8 + warp;
  for (int d = lane; d < 256; d += 32) off[warp][d] = 0;
  __syncwarp();
  const long long n = nvalid ? (long long)*nvalid : ncap;
  if (chunk < nchunks) count_chunk(keys_in, n, chunk, shift, lane, off[warp]);
  __syncthreads();
  {  // counts -> start offsets: CTA base of the digit + counts of the lower warps
    int run = hist[(size_t)threadIdx.x * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = off[w][threadIdx.x];
      off[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
  if (chunk >= nchunks) return;
  const long long base = (long long)chunk * TLSAN_SORT_CHUNK;
  const unsigned lt = (1u << lane) - 1u;
  for (int it = 0; it < TLSAN_SORT_CHUNK / 32; ++it) {
    const long long idx = base + it * 32 + lane;
    const int key = idx < n ? keys_in[idx] : TLSAN_INVALID_KEY;
    const bool act = key != TLSAN_INVALID_KEY;
    const int digit = act ? (key >> shift) & 255 : 256 + lane;  // inactive lanes match only themselves
    const unsigned m = __match_any_sync(0xffffffffu, digit);
    const int rank = __popc(m & lt);
    int pos = 0;
    if (act) pos = off[warp][digit] + rank;
    __syncwarp();
    if (act && rank == 0) off[warp][digit] += __popc(m);
    __syncwarp();
    if (act) {
      keys_out[pos] = key;
      vals_out[pos] = vals_in ? vals_in[idx] : (int)idx;
    }
  }
}

// seg_off[r] = first sorted position with key >= r, r in [0, NR]
__global__ void k_seg_bounds(const int* __restrict__ keys, const int* __restrict__ nvalid, int NR,
                             int* __restrict__ seg_off) {
  const int n = *nvalid;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int prev = i > 0 ? keys[i - 1] : -1;
  const int cur = i < n ? keys[i] : NR;
  for (int r = prev + 1; r <= cur; ++r) seg_off[r] = i;
}

int tlsan_launch_sort(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, const TlsanWs& w,
                      char* ws, const int32_t** sorted_vals, cudaStream_t st) {
  int* keys_a = reinterpret_cast<int*>(ws + w.keys_a);
  int* keys_b = reinterpret_cast<int*>(ws + w.keys_b);
  int* vals_a = reinterpret_cast<int*>(ws + w.vals_a);
  int* vals_b = reinterpret_cast<int*>(ws + w.vals_b);
  int* hist = reinterpret_cast<int*>(ws + w.hist);
  int* nvalid = reinterpret_cast<int*>(ws + w.nvalid);
  int* seg_off = reinterpret_cast<int*>(ws + w.seg_off);
  const long long nocc = w.nocc;
  k_build_keys<<<(unsigned)((nocc + 255) / 256), 256, 0, st>>>(d.B, d.L, d.S, d.NI, d.NC, b.u, b.i, b.c, b.sl,
                                                               b.sl_new, b.hist_i, b.hist_i_new, keys_a);
  TLSAN_CHECK_LAUNCH("k_build_keys");
  int bits = 1;
  while ((1ll << bits) < (long long)w.NR) ++bits;
  const int passes = (bits + 7) / 8;
  const int nblk = (w.nchunks + 7) / 8;
  const int* kin = keys_a; const int* vin = nullptr;
  int* kout = keys_b; int* vout = vals_b;
  for (int pass = 0; pass < passes; ++pass) {
    const int* nv = pass == 0 ? nullptr : nvalid;
    k_radix_hist<<<nblk, 256, 0, st>>>(kin, nocc, nv, 8 * pass, w.nchunks, nblk, hist);
    TLSAN_CHECK_LAUNCH("k_radix_hist");
    k_radix_scan<<<1, 1024, 0, st>>>(hist, 256 * nblk, nvalid);
    TLSAN_CHECK_LAUNCH("k_radix_scan");
    k_radix_scatter<<<nblk, 256, 0, st>>>(kin, vin, nocc, nv, 8 * pass, w.nchunks, nblk, hist, kout, vout);
    TLSAN_CHECK_LAUNCH("k_radix_scatter");
    kin = kout; vin = vout;
    if (kout == keys_b) { kout = keys_a; vout = vals_a; } else { kout = keys_b; vout = vals_b; }
  }
  k_seg_bounds<<<(unsigned)((nocc + 1 + 255) / 256), 256, 0, st>>>(kin, nvalid, w.NR, seg_off);
  TLSAN_CHECK_LAUNCH("k_seg_bounds");
  *sorted_vals = vin;
  return TLSAN_OK;
}
