// Long-term kernels with a resolved-metadata pre-pass and an in-warp prefetch pipeline (default path of the train
// step and of large-batch scoring since round 2).
//
// Same math and lane layout as tlsan_fused_mma.cu (one warp = one sample, 16-row 3xTF32 mma tiles).  What changed is
// how a sample's data reaches the warp.  The round-1 kernels walked, per sample, a chain of dependent loads
//        u, sl  ->  hist_i / hist_t / usert[u]  ->  icl[id]  ->  token rows (item row | cate row)
// and a quarter to a third of their stall samples sat on it (profiles/r01_ncu_full_summary.txt: long_scoreboard).
//   * k_long_meta resolves the chain ONCE per step for every token: meta[b][t] = {item row, category row,
//     P[u,t] * hist_t, hist_t} (16 B, coalesced) -- the time-gap bucketing of raw day gaps (build_dataset.py:16-21) happens
//     here too, while gathering;
//   * k_pf_long keeps, per warp, two shared-memory buffers of one ROUND (up to 16 tokens of one sample) each and a
//     three-slot ring of round metadata, and runs a two-deep pipeline entirely on cp.async (no prefetch registers):
//     while round n is computed from buffer n % 2, the rows of round n+1 are in flight (16 lanes x 16 B per token,
//     addresses from metadata that landed a round ago) and the metadata of round n+2 is being copied.  No load a
//     tile depends on is issued less than one round of compute earlier.
// Work assignment: the backward and short-term kernels take a balanced STATIC partition of the samples (k_part_*:
// contiguous ranges of equal cost per warp), so every per-warp partial sum is accumulated in a fixed order and results
// are bit-identical from run to run; the forward, whose outputs are all per sample, claims samples dynamically.
//
// Tile-level changes against tlsan_mma_common.cuh:
//   * log2(e) is folded into W2 / b2 once per kernel: the softmax works in the log2 domain (no multiply per exp);
//   * forward: up to FOUR tokens (two independent mma chains) per iteration and ONE running-max update for all of
//     them (5 ex2 per 4 tokens and feature instead of 8), which also halves the dependent FMNMX / FFMA chain;
//   * backward: the tile of tlsan_mma_common.cuh with the log2-domain exponent.  (A variant that fed the weight-
//     gradient outer products from LDS.128 broadcasts instead of quad shuffles and reduced d tau through shared
//     memory was measured and dropped: +25 % instructions after register allocation, profiles/experiments/.)
//
//   k_long_meta   per-token metadata of the long-term sequence and of the session / candidate / user rows
//                                                                                  (model.py:84-86,98-99,109)
//   k_pf_long<1>  long-term FWA forward -> o_long + softmax statistics (scratch)  (model.py:98-109,334-345)
//   k_pf_long<3>  backward of the long-term FWA and of the time-aware position term
//   k_pf_short    short-term FWA forward + logit + loss + backward of logit / short FWA, same pipeline, one
//                 sample per round                                                 (model.py:135-137,164-172,350-364)
#include <stdlib.h>
#include "tlsan_mma_common.cuh"

#define PF_R 16                          // tokens per round
#define PF_WARPS 8
#define PF_THREADS (32 * PF_WARPS)
#define LOG2E 1.4426950408889634f

// ---------------------------------------------------------------------------------------------------------------
// per-warp shared-memory geometry (bytes); RW = rows per buffer = min(16, L rounded up to 2)
struct PfGeo {
  int rw, o_tau, o_stats, buf, o_ring, ring, per_warp, total;
};
static PfGeo pf_geo(int L, bool bwd) {
  PfGeo g;
  g.rw = L >= PF_R ? PF_R : (L + 1) / 2 * 2;
  int o = g.rw * 256;
  g.o_tau = o; o += 64;                                         // tau of the round's tokens
  g.o_stats = o; if (bwd) o += 1024;                            // do_long | o_long | max | 1/den of the sample
  g.buf = o;
  o = 2 * g.buf;
  g.o_ring = o;                                                 // 3 x { int4 meta[16] ; int pos[16] }
  g.ring = 16 * 16 + 16 * 4;
  o += 3 * g.ring;
  g.per_warp = (o + 127) / 128 * 128;
  g.total = g.per_warp * PF_WARPS;
  if (bwd && g.total < (int)(PF_WARPS * 160 * 4)) g.total = PF_WARPS * 160 * 4;   // end-of-kernel reduction staging
  return g;
}

struct PfArgs {
  FArgs a;
  const int4* meta;         // [B][L] {item row, category row, P*hist_t, hist_t}
  const int* starts;        // [#warps + 1] balanced partition (k_partition): backward only
  int* counter;             // forward: next unclaimed sample (k_long_meta resets it to #warps * chunk)
  int chunk;                // forward: samples per claim
  PfGeo g;
};

__device__ __forceinline__ void cp16_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp4_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp8_s(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// forward weights with log2(e) folded into the second map: m2' = m2 * log2(e)
__device__ __forceinline__ FwaW load_fwa_log2(const float* __restrict__ dense, int base, int g, int t) {
  FwaW w;
  w.W1 = load_b(dense + base, g, t);
  w.W2 = make_b(dense[base + 72 + (2 * t) * 8 + g] * LOG2E, dense[base + 72 + (2 * t + 1) * 8 + g] * LOG2E);
  w.b1[0] = dense[base + 64 + 2 * t]; w.b1[1] = dense[base + 64 + 2 * t + 1];
  w.b2[0] = dense[base + 136 + 2 * t] * LOG2E; w.b2[1] = dense[base + 136 + 2 * t + 1] * LOG2E;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// pre-pass: B x L threads resolve the long-term history (one token each), then B threads the short-term kernel's
// rows (one sample each; smeta == NULL: scoring, long-term part only).  smeta[b][0] = candidate, [1] = user vector,
// [2 + j] = session item j, each {row of the item / user half, row of the category half};
// sscal[b] = {sl_new, candidate, y, item_b[candidate]}
// score_ncand > 0: SCORING layout of the short-term metadata (k_pf_score): smeta[b][0] = candidate 1, [1] = user vector,
// [2] = candidate 2 (= candidate 1 when there is only one), [3 + j] = session item j; sscal[b] = {sl_new, 0,
// item_b[candidate 1], item_b[candidate 2]}
__global__ void __launch_bounds__(256) k_long_meta(const FArgs a, int4* __restrict__ meta, int2* __restrict__ smeta,
                                                   int4* __restrict__ sscal, int* __restrict__ counter, int counter0,
                                                   int score_ncand) {
  // the per-sample threads of the short-term part have the longest dependent chain (candidate -> icl / item_b, session
  // ids -> icl, a loop for long sessions): they get the FIRST blocks so that the per-token threads run beside them
  // instead of the kernel ending on them
  const long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nshort = smeta ? a.B : 0;
  const long long nlong = (long long)a.B * a.L;
  const long long gidx = g0 < nshort ? nlong + g0 : g0 - nshort;   // index in the original (tokens, then samples) order
  pdl_wait();                                                   // usert / item_b: the previous step's update
  pdl_trigger();
  if (g0 == 0 && counter) { counter[0] = counter0; counter[1] = counter0; }   // work counters of k_pf_long<1> / k_pf_score
  if (g0 >= nlong + nshort) return;
  if (gidx < nlong) {
    const int b = (int)(gidx / a.L), t = (int)(gidx - (long long)b * a.L);
    int4 m = make_int4(0, 0, 0, 0);
    if (t < __ldg(a.sl + b)) {
      const int id = __ldg(a.hist_i + gidx);
      const float ht = a.hist_d ? bucket_weight(__ldg(a.hist_d + gidx)) : __ldg(a.hist_t + gidx);   // bucketing fused
      const float pt = __ldg(a.usert + (size_t)__ldg(a.u + b) * a.L + t) * ht;                     // model.py:99
      m = make_int4(id, a.NI + __ldg(a.icl + id), __float_as_int(pt), __float_as_int(ht));
    }
    meta[gidx] = m;
    return;
  }
  const long long bb = gidx - nlong;
  if (!smeta || bb >= a.B) return;
  const int b = (int)bb;
  const int cand = __ldg(a.i + b), u = __ldg(a.u + b), uc = __ldg(a.c + b), s = min(__ldg(a.sl_new + b), a.S);
  if (score_ncand > 0) {
    int2* out = smeta + (size_t)b * (a.S + 3);
    const int* hn = a.hist_i_new + (size_t)b * a.S;
    const int cand2 = score_ncand > 1 ? __ldg(a.i2 + b) : cand;
    const int id0 = s > 0 ? __ldg(hn) : 0, id1 = s > 1 ? __ldg(hn + 1) : 0;
    const int c1 = __ldg(a.icl + cand), c2 = __ldg(a.icl + cand2), c0i = __ldg(a.icl + id0), c1i = __ldg(a.icl + id1);
    const float ib1 = __ldg(a.item_b + cand), ib2 = __ldg(a.item_b + cand2);
    out[0] = make_int2(cand, a.NI + c1);
    out[1] = make_int2(a.NI + a.NC + u, a.NI + uc);
    out[2] = make_int2(cand2, a.NI + c2);
    if (s > 0) out[3] = make_int2(id0, a.NI + c0i);
    if (s > 1) out[4] = make_int2(id1, a.NI + c1i);
    for (int j = 2; j < s; ++j) {
      const int id = __ldg(hn + j);
      out[3 + j] = make_int2(id, a.NI + __ldg(a.icl + id));
    }
    sscal[b] = make_int4(s, 0, __float_as_int(ib1), __float_as_int(ib2));
    return;
  }
  const float y = __ldg(a.y + b);
  int2* out = smeta + (size_t)b * (a.S + 2);
  const int* hn = a.hist_i_new + (size_t)b * a.S;
  int id0 = s > 0 ? __ldg(hn) : 0, id1 = s > 1 ? __ldg(hn + 1) : 0;          // 96 % of the sessions hold <= 2 items
  const int ccand = __ldg(a.icl + cand);
  const float ib = __ldg(a.item_b + cand);
  const int c0 = __ldg(a.icl + id0), c1 = __ldg(a.icl + id1);
  out[0] = make_int2(cand, a.NI + ccand);
  out[1] = make_int2(a.NI + a.NC + u, a.NI + uc);
  if (s > 0) out[2] = make_int2(id0, a.NI + c0);
  if (s > 1) out[3] = make_int2(id1, a.NI + c1);
  for (int j = 2; j < s; ++j) {
    const int id = __ldg(hn + j);
    out[2 + j] = make_int2(id, a.NI + __ldg(a.icl + id));
  }
  sscal[b] = make_int4(s, cand, __float_as_int(y), __float_as_int(ib));
}

// ---------------------------------------------------------------------------------------------------------------
// Balanced static partition.  Samples differ in length (1..L history entries, 1..S session items), so dealing them
// round-robin leaves the slowest of a few thousand warps ~30 % above the mean and every CTA waiting for its slowest
// warp (ncu: 9 % of the backward's stall samples sat on the final barrier).  Instead warp w of a kernel with NW warps
// takes the CONTIGUOUS range [starts[w], starts[w+1]) whose cost prefix sums are equal shares of the total:
// starts[w] = min { b : P(b) >= floor(w * total / NW) }, P = exclusive prefix of cost.  The partition is a pure
// function of the batch and the grid, so the per-warp summation order -- and every result bit -- stays reproducible.
//   cost_long(b) = 2 ceil(sl / 2) + 2          cost_short(b) = 2 ceil((sl_new + 1) / 2) + 3        (tiles + overhead)
struct PartArgs { const int* sl; const int* sl_new; int B; int nw[3]; int* starts[3]; };   // [0] unused (the forward
                                                    // claims samples dynamically, see k_pf_long), [1] bwd, [2] short

// Two small kernels.  k_part_scan (PART_CTAS CTAs): CTA c owns samples [c * per, (c+1) * per) -- per-chunk (8 samples)
// exclusive prefixes LOCAL to the CTA and the CTA's total, for both cost kinds.  k_part_bounds: every CTA scans the 64
// CTA totals in shared memory, each thread places one boundary by binary search -- over the CTA bases first, then over
// one CTA's chunk prefixes in global memory -- and a walk of at most 8 samples; comparisons are
// w * total <= P * nw + nw - 1 (no division).
#define PART_CTAS 64
#define PART_MAXB (200 * 1024)                                   // sizes the chunk-prefix scratch (tlsan_partition_bytes)
__device__ __forceinline__ int part_cost_of(int kind, int v) {
  return kind == 0 ? 2 * ((min(max(v, 0), 120) + 1) / 2) + 2 : 2 * ((min(max(v, 0), 120) + 2) / 2) + 3;
}
// scan[kind][chunk] = CTA-local exclusive prefix ; tot[kind][cta]
__global__ void __launch_bounds__(256) k_part_scan(const PartArgs p, unsigned int* __restrict__ scan,
                                                   unsigned int* __restrict__ tot, int nchunk, int cpc) {
  __shared__ unsigned int wsum[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c_lo = blockIdx.x * cpc, c_hi = min(c_lo + cpc, nchunk);       // this CTA's chunks
  for (int kind = 0; kind < 2; ++kind) {
    const int* src = kind == 0 ? p.sl : p.sl_new;
    unsigned int carry = 0;
    for (int c0 = c_lo; c0 < c_hi; c0 += 256) {                  // 256 chunks (2048 samples) per sweep, one per thread
      const int c = c0 + tid;
      unsigned int mine = 0;
      if (c < c_hi) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int b = c * 8 + i; if (b < p.B) mine += part_cost_of(kind, __ldg(src + b)); }
      }
      unsigned int inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      __syncthreads();
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      unsigned int base = carry;
      for (int k = 0; k < wid; ++k) base += wsum[k];
      if (c < c_hi) scan[(size_t)kind * nchunk + c] = base + inc - mine;
      unsigned int all = 0;
      for (int k = 0; k < 8; ++k) all += wsum[k];
      carry += all;
    }
    if (tid == 0) tot[kind * PART_CTAS + blockIdx.x] = carry;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_part_bounds(const PartArgs p, const unsigned int* __restrict__ scan,
                                                     const unsigned int* __restrict__ tot, int nchunk, int cpc) {
  __shared__ unsigned int cbase[2][PART_CTAS + 1];               // global prefix at the first chunk of every scan CTA
  __shared__ unsigned int ctot[2 * PART_CTAS];
  const int tid = threadIdx.x;
  // exclusive scan of the 64 CTA totals of each kind: one load per thread, then every thread sums its predecessors
  // out of shared memory
  if (tid < 2 * PART_CTAS) ctot[tid] = __ldg(tot + tid);
  __syncthreads();
  if (tid < 2 * PART_CTAS) {
    const int kind = tid / PART_CTAS, c = tid - kind * PART_CTAS;
    unsigned int run = 0;
    for (int q = 0; q < c; ++q) run += ctot[kind * PART_CTAS + q];
    cbase[kind][c] = run;
    if (c == PART_CTAS - 1) cbase[kind][PART_CTAS] = run + ctot[tid];
  }
  __syncthreads();
  const int stride = TLSAN_MAX_GRID * PF_WARPS + 1;
  const int idx = blockIdx.x * 256 + tid;
  const int t = idx / stride, w = idx - t * stride;
  if (t >= 3 || p.nw[t] <= 0 || w > p.nw[t]) return;
  const int nw = p.nw[t], kind = t == 2 ? 1 : 0;
  const int* src = kind == 0 ? p.sl : p.sl_new;
  // global exclusive prefix of chunk c (c == nchunk: the total); the CTA bases sit in shared memory, the CTA-local
  // prefixes stay in global memory (an earlier version staged all of them in every CTA's shared memory: 20 us)
  auto P = [&](int c) -> unsigned int {
    return c >= nchunk ? cbase[kind][PART_CTAS] : cbase[kind][c / cpc] + __ldg(scan + (size_t)kind * nchunk + c);
  };
  int out = p.B;
  if (w < nw) {
    const unsigned long long total = cbase[kind][PART_CTAS] > 0 ? cbase[kind][PART_CTAS] : 1;
    const unsigned long long lhs = (unsigned long long)w * total;
    auto ge = [&](unsigned int x) { return lhs <= (unsigned long long)x * nw + (nw - 1); };
    // first chunk whose prefix satisfies ge (nchunk: none) -- first among the 65 CTA bases (shared memory), then
    // inside one scan CTA's chunks (<= 8 dependent global loads)
    int qlo = 0, qhi = PART_CTAS + 1;
    while (qlo < qhi) { const int mid = (qlo + qhi) >> 1; if (ge(cbase[kind][mid])) qhi = mid; else qlo = mid + 1; }
    int lo = qlo == 0 ? 0 : (int)min((long long)(qlo - 1) * cpc + 1, (long long)nchunk);
    int hi = qlo == 0 ? 0 : (int)min((long long)qlo * cpc, (long long)nchunk);
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ge(P(mid))) hi = mid; else lo = mid + 1; }
    int b = lo * 8;                                              // inside chunk lo - 1 (after its first sample) or at chunk lo
    if (lo > 0) {
      const int base = (lo - 1) * 8;
      int v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = base + i < p.B ? __ldg(src + base + i) : 0;
      unsigned int x = P(lo - 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x += base + i < p.B ? part_cost_of(kind, v[i]) : 0;
        if (b == lo * 8 && ge(x)) b = base + i + 1;
      }
    }
    out = min(b, p.B);
  }
  p.starts[t][w] = out;
}

__global__ void k_partition_uniform(const PartArgs p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = idx / (TLSAN_MAX_GRID * PF_WARPS + 1), w = idx - t * (TLSAN_MAX_GRID * PF_WARPS + 1);
  if (t < 3 && p.nw[t] > 0 && w <= p.nw[t]) p.starts[t][w] = (int)((long long)w * p.B / p.nw[t]);
}

// ---------------------------------------------------------------------------------------------------------------
// online softmax over the sequence axis in the log2 domain, for the lane's two features
struct SoftL2 {
  float mx[2], den[2], acc[2];
  __device__ __forceinline__ void init() {
    mx[0] = mx[1] = -INFINITY; den[0] = den[1] = 0.f; acc[0] = acc[1] = 0.f;
  }
  // two tokens (second may be masked: m = -inf, x = 0)
  __device__ __forceinline__ void push2(const float (&m)[4], const float (&x)[4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float nm = fmax3(mx[j], m[j], m[2 + j]);
      const float sc = ex2f(mx[j] - nm), e0 = ex2f(m[j] - nm), e1 = ex2f(m[2 + j] - nm);
      den[j] = fmaf(den[j], sc, e0 + e1);
      acc[j] = fmaf(acc[j], sc, fmaf(e0, x[j], e1 * x[2 + j]));
      mx[j] = nm;
    }
  }
  // four tokens
  __device__ __forceinline__ void push4(const float (&ma)[4], const float (&xa)[4], const float (&mb)[4],
                                        const float (&xb)[4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float nm = fmaxf(fmax3(mx[j], ma[j], ma[2 + j]), fmaxf(mb[j], mb[2 + j]));
      const float sc = ex2f(mx[j] - nm);
      const float e0 = ex2f(ma[j] - nm), e1 = ex2f(ma[2 + j] - nm), e2 = ex2f(mb[j] - nm), e3 = ex2f(mb[2 + j] - nm);
      den[j] = fmaf(den[j], sc, (e0 + e1) + (e2 + e3));
      acc[j] = fmaf(acc[j], sc, fmaf(e0, xa[j], e1 * xa[2 + j]) + fmaf(e2, xb[j], e3 * xb[2 + j]));
      mx[j] = nm;
    }
  }
};

// backward of one tile: tile_bwd of tlsan_mma_common.cuh with the softmax exponent in the log2 domain
// (w.W2 / w.b2 carry log2(e); wt holds the UNscaled transposed maps for d pre / d x)
// core: the maps m1 / m2 of the tile are given (recomputed by tile_bwd_l2, or kept from the forward when the whole
// sequence was one tile)
__device__ __forceinline__ void tile_bwd_core(const float (&x)[4], const float (&m1)[4], const float (&m2)[4], bool okB,
                                              const float (&o)[2], const float (&kf)[2], const float (&nmx)[2],
                                              const FwaWT& wt, int lane, float (&dx)[4], FwaGrad& G) {
  float ado[4], dm2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = i & 1;
    const float aw = ex2f(m2[i] + nmx[j]) * kf[j];             // softmax weight * d out
    ado[i] = (i < 2 || okB) ? aw : 0.f;
    dm2[i] = ado[i] * (x[i] - o[j]);
  }
  G.b2[0] += dm2[0] + dm2[2]; G.b2[1] += dm2[1] + dm2[3];
  float dpre[4] = {0.f, 0.f, 0.f, 0.f};
  mma3(dpre, dm2, wt.W2T);
#pragma unroll
  for (int i = 0; i < 4; ++i) dpre[i] = m1[i] > 0.f ? dpre[i] : 0.f;
  G.b1[0] += dpre[0] + dpre[2]; G.b1[1] += dpre[1] + dpre[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) dx[i] = ado[i];
  mma3(dx, dpre, wt.W1T);
  const int qbase = lane & ~3;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float2 dmb[2] = {make_float2(dm2[2 * r], dm2[2 * r]), make_float2(dm2[2 * r + 1], dm2[2 * r + 1])};
    const float2 dpb[2] = {make_float2(dpre[2 * r], dpre[2 * r]), make_float2(dpre[2 * r + 1], dpre[2 * r + 1])};
#pragma unroll
    for (int tq = 0; tq < 4; ++tq) {
      const float2 mk = make_float2(__shfl_sync(0xffffffffu, m1[2 * r], qbase + tq),
                                    __shfl_sync(0xffffffffu, m1[2 * r + 1], qbase + tq));
      const float2 xk = make_float2(__shfl_sync(0xffffffffu, x[2 * r], qbase + tq),
                                    __shfl_sync(0xffffffffu, x[2 * r + 1], qbase + tq));
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        G.W2p[tq][jj] = __ffma2_rn(mk, dmb[jj], G.W2p[tq][jj]);
        G.W1p[tq][jj] = __ffma2_rn(xk, dpb[jj], G.W1p[tq][jj]);
      }
    }
  }
}

__device__ __forceinline__ void tile_bwd_l2(const float (&x)[4], bool okB, const float (&o)[2], const float (&kf)[2],
                                            const float (&nmx)[2], const FwaW& w, const FwaWT& wt, int lane,
                                            float (&dx)[4], FwaGrad& G) {
  float m1[4], m2[4];
  tile_maps(x, w, m1, m2);
  tile_bwd_core(x, m1, m2, okB, o, kf, nmx, wt, lane, dx, G);
}

// fixed-order reduction of a warp's weight-gradient accumulators into red[0..143]
__device__ __forceinline__ void pf_reduce_grads(const FwaGrad& G, const LaneGeo& L, float* __restrict__ red) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      float r1 = G.w1(k, jj), r2 = G.w2(k, jj);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      }
      if (L.g == 0) { red[k * 8 + 2 * L.t + jj] = r1; red[72 + k * 8 + 2 * L.t + jj] = r2; }
    }
  }
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    float r1 = G.b1[jj], r2 = G.b2[jj];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      r1 += __shfl_xor_sync(0xffffffffu, r1, o);
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
    }
    if (L.g == 0) { red[64 + 2 * L.t + jj] = r1; red[136 + 2 * L.t + jj] = r2; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the warp's sequence of rounds: (sample b, first token r0, length ell); b >= B marks the end
struct PfIter { int b, r0, ell; };

template <int KIND>
__global__ void __launch_bounds__(PF_THREADS, KIND == 1 ? 3 : 2) k_pf_long(const PfArgs A) {
  constexpr bool BWD = KIND == 3;
  // Work assignment.  Backward: the balanced static partition (its per-warp partial sums must be accumulated in a
  // fixed order).  Forward: every output is per sample, so the order samples are processed in cannot change a bit of
  // the result -- warps CLAIM chunks of A.chunk samples from a global counter instead.  A CTA that becomes resident
  // late (the occurrence sort of the next batch shares the SMs: a third CTA per SM often has to wait for a sort CTA
  // to retire) then simply claims less, where the static partition made the whole kernel wait for it
  // (long-term forward 52 us alone, 100 us beside the sort).
  constexpr bool DYN = KIND == 1;
  extern __shared__ __align__(128) unsigned char smem[];
  const FArgs& a = A.a;
  const PfGeo& g = A.g;
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PF_WARPS + warp;
  int bBeg, bEnd;                                               // static: this warp's contiguous samples
  if (DYN) { bBeg = min(gw * A.chunk, a.B); bEnd = min(bBeg + A.chunk, a.B); }   // first chunk: no claim needed
  else { bBeg = __ldg(A.starts + gw); bEnd = __ldg(A.starts + gw + 1); }
  const int h = L.lane >> 4, c16 = L.lane & 15;
  unsigned char* mine = smem + (size_t)warp * g.per_warp;
  const FwaW wl = load_fwa_log2(a.dense, TLSAN_OFF_W1L, L.g, L.t);

  const float gamma = a.dense[TLSAN_OFF_GAMMA];
  int curEnd = bEnd, nxt = a.B;                                 // dynamic: end of the newest chunk, first sample of the claimed one
  auto claim = [&]() -> int {
    int v = 0;
    if (L.lane == 0) v = atomicAdd(A.counter, A.chunk);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  auto ell_of = [&](int b) { return b < a.B ? __ldg(a.sl + b) : -1; };
  // the sample after b in this warp's sequence; a.B: end marker.  Called exactly once per sample (stateful when DYN).
  auto succ = [&](int b) -> int {
    if (b >= a.B) return a.B;
    if (b + 1 < curEnd) return b + 1;
    if (!DYN) return a.B;
    const int nb = nxt;
    if (nb >= a.B) return a.B;
    curEnd = min(nb + A.chunk, a.B);
    nxt = claim();                                              // needed a whole chunk from now
    return nb;
  };
  auto ring = [&](int n) { return mine + g.o_ring + (n % 3) * g.ring; };
  // copy the metadata of a round into ring slot `dst`: lane t < 16 serves token r0 + t
  auto issue_meta = [&](const PfIter& it, unsigned char* dst) {
    const int tok = it.r0 + c16;
    if (h == 0 && tok < it.ell) {
      cp16_s(smem_addr(dst) + c16 * 16, A.meta + (size_t)it.b * a.L + tok);
      if (BWD) cp4_s(smem_addr(dst) + 256 + c16 * 4, a.inv + ((size_t)it.b << a.spsh) + tok);
    }
  };
  // start the copies of a round into `buf` (its metadata has landed in ring slot `src`): tau, the token rows,
  // (backward) the sample's record
  auto issue_round = [&](const PfIter& it, const unsigned char* src, unsigned char* buf) {
    const int cnt = min(PF_R, max(it.ell - it.r0, 0));
    int4 m = make_int4(0, 0, 0, 0);
    if (c16 < cnt) m = reinterpret_cast<const int4*>(src)[c16];
    if (h == 0) reinterpret_cast<float*>(buf + g.o_tau)[c16] = gamma * __int_as_float(m.z);   // tau   (model.py:109)
    // 16 lanes x 16 B per token (chunks 0-7 item row, 8-15 category row); two tokens per instruction
    const uint32_t dst = smem_addr(buf) + h * 256 + c16 * 16;
    const int col = (c16 & 7) * 4;
    for (int tt = 0; tt < cnt; tt += 2) {
      const int idt = __shfl_sync(0xffffffffu, m.x, tt + h), crt = __shfl_sync(0xffffffffu, m.y, tt + h);
      if (tt + h < cnt) cp16_s(dst + tt * 256, a.emb + (size_t)(c16 < 8 ? idt : crt) * 32 + col);
    }
    if (BWD && it.r0 == 0 && it.ell >= 0) {                    // do_long | o_long | max | 1/den of the sample
      const float* sc = a.scratch + (size_t)it.b * (TLSAN_SCR * 64);
      cp16_s(smem_addr(buf + g.o_stats) + L.lane * 16, sc + L.lane * 4);
      cp16_s(smem_addr(buf + g.o_stats) + 512 + L.lane * 16, sc + 128 + L.lane * 4);
    }
  };
  auto advance = [&](const PfIter& it, int b_next, int ell_next) {
    PfIter n;
    if (it.r0 + PF_R < it.ell) { n.b = it.b; n.r0 = it.r0 + PF_R; n.ell = it.ell; }
    else { n.b = b_next; n.r0 = 0; n.ell = ell_next; }
    return n;
  };

  // ---- pipeline prologue: metadata of rounds 0 and 1, then the rows of round 0
  const FwaWT wlt = BWD ? load_fwa_t(a.dense, TLSAN_OFF_W1L, L.g, L.t) : FwaWT();
  PfIter it0, it1, it2;
  int bN = a.B, eN = -1;                                        // the sample after the newest iterator's, its length
  auto first_rounds = [&]() {
    it0.b = bBeg < bEnd ? bBeg : a.B; it0.r0 = 0; it0.ell = ell_of(it0.b);
    bN = succ(it0.b); eN = ell_of(bN);
    it1 = advance(it0, bN, eN);
    if (it1.r0 == 0) { bN = succ(it1.b); eN = ell_of(bN); }
  };
  if (!DYN) first_rounds();                                     // batch data only: overlaps the previous kernel's tail
  pdl_wait();                                                   // meta, work counter (k_long_meta) / scratch, ranks (backward)
  pdl_trigger();
  if (DYN) {
    if (bBeg < a.B) nxt = claim();
    first_rounds();
  }
  issue_meta(it0, ring(0));
  issue_meta(it1, ring(1));
  cp_commit();
  cp_wait_group<0>();
  __syncwarp();
  issue_round(it0, ring(0), mine);
  cp_commit();

  // per-kernel state
  SoftL2 st; st.init();
  FwaGrad G;
  if (BWD) G.init();
  float ggamma = 0.f, sq_acc = 0.f;
  float o[2] = {0.f, 0.f}, kf[2] = {0.f, 0.f}, nmx[2] = {0.f, 0.f};

  for (int n = 0; it0.b < a.B; ++n) {
    unsigned char* buf = mine + (size_t)(n & 1) * g.buf;
    cp_wait_group<0>();                                         // rows of round n + metadata of round n+1 (a round old)
    __syncwarp();
    // ---- next round's rows, the metadata of the round after it
    issue_round(it1, ring(n + 1), mine + (size_t)((n + 1) & 1) * g.buf);
    it2 = advance(it1, bN, eN);
    if (it2.r0 == 0) { bN = succ(it2.b); eN = ell_of(bN); }
    issue_meta(it2, ring(n + 2));
    cp_commit();

    // ---- compute round n
    const int cnt = min(PF_R, max(it0.ell - it0.r0, 0));
    const float (*rows)[64] = reinterpret_cast<const float (*)[64]>(buf);
    const float* tau = reinterpret_cast<const float*>(buf + g.o_tau);
    if (!BWD) {
      int j = 0;
      for (; j + 2 < cnt; j += 4) {                            // 3 or 4 tokens: two independent tiles
        const float4 t4 = *reinterpret_cast<const float4*>(tau + j);
        const bool ok3 = j + 3 < cnt;
        const float2 e0 = *reinterpret_cast<const float2*>(&rows[j][L.f0]);
        const float2 e1 = *reinterpret_cast<const float2*>(&rows[j + 1][L.f0]);
        const float2 e2 = *reinterpret_cast<const float2*>(&rows[j + 2][L.f0]);
        const float2 e3 = ok3 ? *reinterpret_cast<const float2*>(&rows[j + 3][L.f0]) : make_float2(0.f, 0.f);
        const float xa[4] = {e0.x * t4.x, e0.y * t4.x, e1.x * t4.y, e1.y * t4.y};
        const float xb[4] = {e2.x * t4.z, e2.y * t4.z, ok3 ? e3.x * t4.w : 0.f, ok3 ? e3.y * t4.w : 0.f};
        float m1a[4], m2a[4], m1b[4], m2b[4];
        tile_maps(xa, wl, m1a, m2a);
        tile_maps(xb, wl, m1b, m2b);
        if (!ok3) { m2b[2] = -INFINITY; m2b[3] = -INFINITY; }
        st.push4(m2a, xa, m2b, xb);
      }
      if (j < cnt) {                                           // 1 or 2 tokens
        const bool okB = j + 1 < cnt;
        const float tA = tau[j], tB = okB ? tau[j + 1] : 0.f;
        const float2 eA = *reinterpret_cast<const float2*>(&rows[j][L.f0]);
        const float2 eB = okB ? *reinterpret_cast<const float2*>(&rows[j + 1][L.f0]) : make_float2(0.f, 0.f);
        const float x[4] = {eA.x * tA, eA.y * tA, eB.x * tB, eB.y * tB};
        float m1t[4], m2t[4];
        tile_maps(x, wl, m1t, m2t);
        if (!okB) { m2t[2] = -INFINITY; m2t[3] = -INFINITY; }
        st.push2(m2t, x);
      }
      if (it0.r0 + PF_R >= it0.ell) {                          // last round of the sample
        float* sc = a.scratch + (size_t)it0.b * (TLSAN_SCR * 64) + L.f0;
        const float i0 = st.den[0] > 0.f ? 1.f / st.den[0] : 0.f, i1 = st.den[1] > 0.f ? 1.f / st.den[1] : 0.f;
        st2(sc + 64, st.acc[0] * i0, st.acc[1] * i1);          // o_long
        st2(sc + 128, st.mx[0], st.mx[1]);                      // running max (log2 domain)
        st2(sc + 192, i0, i1);                                  // 1 / denominator
        st.init();
      }
    } else {
      if (it0.r0 == 0) {
        const float* sc = reinterpret_cast<const float*>(buf + g.o_stats) + L.f0;
        const float2 dol2 = *reinterpret_cast<const float2*>(sc);
        const float2 o2 = *reinterpret_cast<const float2*>(sc + 64);
        const float2 mx2 = *reinterpret_cast<const float2*>(sc + 128);
        const float2 inv2 = *reinterpret_cast<const float2*>(sc + 192);
        o[0] = o2.x; o[1] = o2.y;
        kf[0] = inv2.x * dol2.x; kf[1] = inv2.y * dol2.y;       // (1/den) * d out
        nmx[0] = -mx2.x; nmx[1] = -mx2.y;
      }
      const unsigned char* mr = ring(n);                        // this round's metadata: {id, crow, P*hist_t, hist_t}, rank
      const int* pos = reinterpret_cast<const int*>(mr + 256);
      float* ru = a.rows_u + (size_t)it0.b * a.PU + 32;
      // d tau of the round's tokens: lane q < 16 collects token 2q (first of tile q), lane 16 + q token 2q + 1
      float dtau_l = 0.f;
      const bool upper = L.lane >= 16;
      for (int j = 0; j < cnt; j += 2) {
        const bool okB = j + 1 < cnt;
        const float tA = tau[j], tB = okB ? tau[j + 1] : 0.f;
        const float2 eA = *reinterpret_cast<const float2*>(&rows[j][L.f0]);
        const float2 eB = okB ? *reinterpret_cast<const float2*>(&rows[j + 1][L.f0]) : make_float2(0.f, 0.f);
        const float x[4] = {eA.x * tA, eA.y * tA, eB.x * tB, eB.y * tB};
        float dx[4];
        tile_bwd_l2(x, okB, o, kf, nmx, wl, wlt, L.lane, dx, G);
        // gradient of the gathered slices (tau * dX) and of tau (<dX, e>)
        const float rA0 = dx[0] * tA, rA1 = dx[1] * tA;
        sq_acc = fmaf(rA0, rA0, sq_acc); sq_acc = fmaf(rA1, rA1, sq_acc);
        st2(a.rows_i + (size_t)pos[j] * 64 + L.f0, rA0, rA1);
        if (okB) {
          const float rB0 = dx[2] * tB, rB1 = dx[3] * tB;
          sq_acc = fmaf(rB0, rB0, sq_acc); sq_acc = fmaf(rB1, rB1, sq_acc);
          st2(a.rows_i + (size_t)pos[j + 1] * 64 + L.f0, rB0, rB1);
        }
        // both dot products in ONE butterfly: the halves of the warp swap the partial they do not keep, then the
        // lower half sums token A's 32 partials and the upper half token B's (5 shuffles instead of 10; eB = 0 and
        // dx[2..3] = 0 without a second token)
        const float pA = fmaf(dx[0], eA.x, dx[1] * eA.y), pB = fmaf(dx[2], eB.x, dx[3] * eB.y);
        float v = (upper ? pB : pA) + __shfl_xor_sync(0xffffffffu, upper ? pA : pB, 16);
#pragma unroll
        for (int o2 = 8; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
        if ((L.lane & 15) == (j >> 1)) dtau_l = v;
      }
      {
        const int tok = 2 * (L.lane & 15) + (upper ? 1 : 0);
        if (tok < cnt) {
          const int4 m = reinterpret_cast<const int4*>(mr)[tok];
          ggamma = fmaf(dtau_l, __int_as_float(m.z), ggamma);  // d gamma += d tau * P[u,t] hist_t
          const float dp = dtau_l * gamma * __int_as_float(m.w); // d usert_emb[u, t] = d tau * gamma * hist_t
          sq_acc = fmaf(dp, dp, sq_acc);
          ru[it0.r0 + tok] = dp;
        }
      }
      if (it0.r0 + PF_R >= it0.ell)
        for (int tt = max(it0.ell, 0) + L.lane; tt < a.PU - 32; tt += 32) ru[tt] = 0.f;
    }
    __syncwarp();                                               // buffer n % 2 / ring slot n % 3 may be overwritten
    it0 = it1; it1 = it2;
  }
  cp_wait_group<0>();
  if (!BWD) return;

  // ---- per-CTA partial sums, fixed order: butterfly over g -> warps 0..7 -> global
  __syncthreads();                                              // the buffers are dead: reuse them
  float (*red)[160] = reinterpret_cast<float (*)[160]>(smem);
  pf_reduce_grads(G, L, red[warp]);
  {
    const float r1 = warp_sum_f(ggamma), r2 = warp_sum_f(sq_acc);
    if (L.lane == 0) { red[warp][144] = r1; red[warp][145] = r2; }
  }
  __syncthreads();
  if (threadIdx.x < 146) {
    float r = 0.f;
#pragma unroll
    for (int wv = 0; wv < PF_WARPS; ++wv) r += red[wv][threadIdx.x];
    const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1L + threadIdx.x
                                      : (threadIdx.x == 144 ? TLSAN_OFF_GAMMA : TLSAN_PART_SUMSQ);
    a.part[(size_t)blockIdx.x * TLSAN_PART + dst] = r;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// short-term kernel: one sample per pipeline step.  Buffer = { rows[PS_RS + 2][64] : staged session items, candidate,
// user vector ; z[64] } ; ring slot = { int4 scalars ; int2 rows[2 + 32] ; int rank[2 + 32] }.
#define PS_RS 6        // session items whose rows are staged (99.5 % of the sessions; longer ones gather the rest directly)
struct PsArgs {
  FArgs a;
  const int2* smeta;        // [B][S + 2]
  const int4* sscal;        // [B]
  const int* starts;        // [#warps + 1] balanced partition (k_partition)
};
#define PS_BUF ((PS_RS + 2) * 256 + 256)
#define PS_RING (16 + 34 * 8 + 34 * 4 + 8)
#define PS_PER_WARP ((2 * PS_BUF + 3 * PS_RING + 127) / 128 * 128)

__global__ void __launch_bounds__(PF_THREADS, 2) k_pf_short(const PsArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FArgs& a = A.a;
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PF_WARPS + warp;
  const int bBeg = __ldg(A.starts + gw), bEnd = __ldg(A.starts + gw + 1);     // this warp's contiguous samples
  const int h = L.lane >> 4, c16 = L.lane & 15;
  unsigned char* mine = smem + (size_t)warp * PS_PER_WARP;
  const FwaW w = load_fwa_log2(a.dense, TLSAN_OFF_W1S, L.g, L.t);
  const FwaWT wt = load_fwa_t(a.dense, TLSAN_OFF_W1S, L.g, L.t);
  FwaGrad G; G.init();
  float loss_acc = 0.f, sq_acc = 0.f;
  auto ring = [&](int n) { return mine + 2 * PS_BUF + (n % 3) * PS_RING; };
  const int SW = a.S + 2;
  // metadata of sample b -> ring slot: scalars, the row pairs and the sorted ranks of its first 34 slots
  auto issue_meta = [&](int b, unsigned char* dst) {
    if (b >= bEnd) return;
    if (L.lane == 0) cp16_s(smem_addr(dst), A.sscal + b);
    const uint32_t rp = smem_addr(dst) + 16, pp = rp + 34 * 8;
    for (int k = L.lane; k < min(SW, 34); k += 32) {
      cp8_s(rp + k * 8, A.smeta + (size_t)b * SW + k);
      // slot order of the occurrence sort: session item j -> L + j, candidate -> L + S, virtual (u_cate) -> L + S + 1
      const int slot = k == 0 ? a.L + a.S : k == 1 ? a.L + a.S + 1 : a.L + (k - 2);
      cp4_s(pp + k * 4, a.inv + ((size_t)b << a.spsh) + slot);
    }
  };
  // rows of sample b (metadata in ring slot `src`) -> buffer
  auto issue_rows = [&](int b, const unsigned char* src, unsigned char* buf) {
    if (b >= bEnd) return;
    const int s = reinterpret_cast<const int*>(src)[0];
    const int2* rp = reinterpret_cast<const int2*>(src + 16);
    // staged row r: 0..PS_RS-1 = session items, PS_RS = candidate, PS_RS + 1 = user vector; 16 lanes x 16 B per row
    const int n = min(s, PS_RS);
    const uint32_t dst = smem_addr(buf) + c16 * 16;
    const int col = (c16 & 7) * 4;
    for (int r = h; r < n; r += 2) {
      const int2 p = rp[2 + r];
      cp16_s(dst + r * 256, a.emb + (size_t)(c16 < 8 ? p.x : p.y) * 32 + col);
    }
    {
      const int2 p = rp[h];                                     // half 0: candidate, half 1: user vector
      cp16_s(dst + (PS_RS + h) * 256, a.emb + (size_t)(c16 < 8 ? p.x : p.y) * 32 + col);
    }
    if (h == 0) cp16_s(dst + (PS_RS + 2) * 256, a.scratch + (size_t)b * (TLSAN_SCR * 64) + 320 + c16 * 4);   // z
  };

  int b0 = bBeg;
  pdl_wait();                                                   // z (k_dense_fwd_mma), smeta / sscal, ranks
  pdl_trigger();
  issue_meta(b0, ring(0));
  issue_meta(b0 + 1, ring(1));
  cp_commit();
  cp_wait_group<0>();
  __syncwarp();
  issue_rows(b0, ring(0), mine);
  cp_commit();

  for (int n = 0; b0 < bEnd; ++n, ++b0) {
    const int b = b0;
    unsigned char* buf = mine + (size_t)(n & 1) * PS_BUF;
    cp_wait_group<0>();
    __syncwarp();
    issue_rows(b + 1, ring(n + 1), mine + (size_t)((n + 1) & 1) * PS_BUF);
    issue_meta(b + 2, ring(n + 2));
    cp_commit();

    const unsigned char* mr = ring(n);
    const int4 sc4 = *reinterpret_cast<const int4*>(mr);
    const int s = sc4.x;
    const float yb = __int_as_float(sc4.z), ib = __int_as_float(sc4.w);
    const int2* rp = reinterpret_cast<const int2*>(mr + 16);
    const int* pos = reinterpret_cast<const int*>(mr + 16 + 34 * 8);
    const float (*rows)[64] = reinterpret_cast<const float (*)[64]>(buf);
    const int ntok = s + 1;
    const float2 zz = *reinterpret_cast<const float2*>(&rows[PS_RS + 2][L.f0]);
    const float z[2] = {zz.x, zz.y};
    // token n >= 1 is session item n-1: staged row, else direct gather
    auto item_x = [&](int it) -> float2 {
      if (it < PS_RS) return *reinterpret_cast<const float2*>(&rows[it][L.f0]);
      int id, cr;
      if (it < 32) { const int2 p = rp[2 + it]; id = p.x; cr = p.y; }
      else { id = __ldg(a.hist_i_new + (size_t)b * a.S + it); cr = a.NI + __ldg(a.icl + id); }
      return ldg2(a.emb + (size_t)(L.half ? cr : id) * 32 + L.col);
    };
    // ================= short-term FWA forward over [z ; e(hist_i_new)] (model.py:350-364)
    SoftL2 ss; ss.init();
    // the first tile ([z ; first session item]: the WHOLE sequence of 87 % of the samples) keeps its inputs and maps
    // in registers for the backward pass below; later tiles are recomputed there
    float x0[4], m10[4], m20[4];
    {
      const bool okB = 1 < ntok;
      x0[0] = z[0]; x0[1] = z[1];
      if (okB) { const float2 e = item_x(0); x0[2] = e.x; x0[3] = e.y; } else { x0[2] = 0.f; x0[3] = 0.f; }
      tile_maps(x0, w, m10, m20);
      if (!okB) { m20[2] = -INFINITY; m20[3] = -INFINITY; }
      ss.push2(m20, x0);
    }
    for (int k = 2; k < ntok; k += 2) {
      const bool okB = k + 1 < ntok;
      float x[4];
      { const float2 e = item_x(k - 1); x[0] = e.x; x[1] = e.y; }
      if (okB) { const float2 e = item_x(k); x[2] = e.x; x[3] = e.y; } else { x[2] = 0.f; x[3] = 0.f; }
      float m1[4], m2[4];
      tile_maps(x, w, m1, m2);
      if (!okB) { m2[2] = -INFINITY; m2[3] = -INFINITY; }
      ss.push2(m2, x);
    }
    const float inv_s[2] = {1.f / ss.den[0], 1.f / ss.den[1]};
    const float v[2] = {ss.acc[0] * inv_s[0], ss.acc[1] * inv_s[1]};
    // ---- user vector, candidate, logit (model.py:84-95,135-137)
    const float2 q = *reinterpret_cast<const float2*>(&rows[PS_RS][L.f0]);
    const float2 p = *reinterpret_cast<const float2*>(&rows[PS_RS + 1][L.f0]);
    const float ut[2] = {v[0] + p.x, v[1] + p.y};
    const float logit = warp_sum_f(fmaf(ut[0], q.x, ut[1] * q.y)) + ib;
    // ---- sigmoid cross entropy (model.py:171) and its gradient through reduce_mean
    const float ex = expf(-fabsf(logit));
    const float bce = fmaxf(logit, 0.f) - logit * yb + log1pf(ex);
    const float sig = logit >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
    const float gl = (sig - yb) * a.invB;
    if (L.lane == 0) { loss_acc += bce; sq_acc = fmaf(gl, gl, sq_acc); a.gscal[b] = gl; }
    // gradient-row destinations: ring rank k = 0 candidate, 1 virtual (u_cate), 2 + j session item j
    auto slot_row = [&](int k) -> float* {
      const int ps = k < 34 ? pos[k] : __ldg(a.inv + ((size_t)b << a.spsh) + a.L + (k - 2));
      return a.rows_i + (size_t)ps * 64 + L.f0;
    };
    const float dq[2] = {gl * ut[0], gl * ut[1]};
    const float du[2] = {gl * q.x, gl * q.y};
    sq_acc = fmaf(dq[0], dq[0], sq_acc); sq_acc = fmaf(dq[1], dq[1], sq_acc);
    sq_acc = fmaf(du[0], du[0], sq_acc); sq_acc = fmaf(du[1], du[1], sq_acc);
    st2(slot_row(0), dq[0], dq[1]);                                       // -> item_emb[i] | cate_emb[icl[i]]
    float* rvirt = slot_row(1);
    if (L.half) st2(rvirt, du[0], du[1]);                                 // -> cate_emb[u_cate]
    else { st2(rvirt, 0.f, 0.f); st2(a.rows_u + (size_t)b * a.PU + L.f0, du[0], du[1]); }  // -> user_emb[u]
    // ---- short-term FWA backward, d v = du
    const float kf[2] = {inv_s[0] * du[0], inv_s[1] * du[1]};
    const float nmx[2] = {-ss.mx[0], -ss.mx[1]};
    float dz[2];
    {
      const bool okB = 1 < ntok;
      float dx[4];
      tile_bwd_core(x0, m10, m20, okB, v, kf, nmx, wt, L.lane, dx, G);
      dz[0] = dx[0]; dz[1] = dx[1];
      if (okB) {
        sq_acc = fmaf(dx[2], dx[2], sq_acc); sq_acc = fmaf(dx[3], dx[3], sq_acc);
        st2(slot_row(2), dx[2], dx[3]);
      }
    }
    for (int k = 2; k < ntok; k += 2) {
      const bool okB = k + 1 < ntok;
      float x[4], dx[4];
      { const float2 e = item_x(k - 1); x[0] = e.x; x[1] = e.y; }
      if (okB) { const float2 e = item_x(k); x[2] = e.x; x[3] = e.y; } else { x[2] = 0.f; x[3] = 0.f; }
      tile_bwd_l2(x, okB, v, kf, nmx, w, wt, L.lane, dx, G);
      sq_acc = fmaf(dx[0], dx[0], sq_acc); sq_acc = fmaf(dx[1], dx[1], sq_acc);
      st2(slot_row(2 + k - 1), dx[0], dx[1]);
      if (okB) {
        sq_acc = fmaf(dx[2], dx[2], sq_acc); sq_acc = fmaf(dx[3], dx[3], sq_acc);
        st2(slot_row(2 + k), dx[2], dx[3]);
      }
    }
    st2(a.scratch + (size_t)b * (TLSAN_SCR * 64) + 256 + L.f0, dz[0], dz[1]);   // -> k_dense_bwd_mma
    __syncwarp();
  }
  cp_wait_group<0>();

  // ---- per-CTA partial sums, fixed order: butterfly over g -> warps 0..7 -> global
  __syncthreads();
  float (*red)[160] = reinterpret_cast<float (*)[160]>(smem);
  pf_reduce_grads(G, L, red[warp]);
  {
    const float r1 = warp_sum_f(loss_acc), r2 = warp_sum_f(sq_acc);
    if (L.lane == 0) { red[warp][144] = r1; red[warp][145] = r2; }
  }
  __syncthreads();
  if (threadIdx.x < 146) {
    float r = 0.f;
#pragma unroll
    for (int wv = 0; wv < PF_WARPS; ++wv) r += red[wv][threadIdx.x];
    const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1S + threadIdx.x
                                      : (threadIdx.x == 144 ? TLSAN_PART_LOSS : TLSAN_PART_SUMSQ);
    a.part[(size_t)blockIdx.x * TLSAN_PART + dst] = r;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// short-term part of SCORING (tlsan_score_ws, Model.eval_auc): short FWA forward over [z ; e(hist_i_new)], u_t and the
// logits of one or two candidates from the same u_t (model.py:135-137,251-261,350-364).  Same pipeline as k_pf_short
// (metadata two samples ahead, rows one ahead), samples claimed dynamically like the long-term forward -- nothing is
// accumulated across samples.  Replaces the round-robin k_fwd_mma<3>, whose per-sample chain
// (u, sl_new, candidates -> icl -> rows) was exposed: 61 -> ~35 us at B 65 536.
struct PscArgs {
  FArgs a;
  const int2* smeta;        // [B][S + 3]  (scoring layout of k_long_meta)
  const int4* sscal;        // [B]
  int* counter;             // next unclaimed sample (k_long_meta resets it to #warps * chunk)
  int chunk, ncand;
};
#define PSC_ROWS (PS_RS + 3)                     // staged rows: PS_RS session items, candidate 1, user vector, candidate 2
#define PSC_BUF (PSC_ROWS * 256 + 256)           // + z
#define PSC_RING (16 + 35 * 8 + 8)               // int4 scalars ; int2 rows[3 + 32]
#define PSC_PER_WARP ((2 * PSC_BUF + 3 * PSC_RING + 127) / 128 * 128)

__global__ void __launch_bounds__(PF_THREADS, 3) k_pf_score(const PscArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FArgs& a = A.a;
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PF_WARPS + warp;
  const int h = L.lane >> 4, c16 = L.lane & 15;
  unsigned char* mine = smem + (size_t)warp * PSC_PER_WARP;
  const FwaW w = load_fwa_log2(a.dense, TLSAN_OFF_W1S, L.g, L.t);
  const int SW = a.S + 3;
  auto ring = [&](int n) { return mine + 2 * PSC_BUF + (n % 3) * PSC_RING; };
  int curEnd, nxt = a.B;
  auto claim = [&]() -> int {
    int v = 0;
    if (L.lane == 0) v = atomicAdd(A.counter, A.chunk);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  auto succ = [&](int b) -> int {                 // the sample after b in this warp's sequence (a.B: none); once per sample
    if (b >= a.B) return a.B;
    if (b + 1 < curEnd) return b + 1;
    const int nb = nxt;
    if (nb >= a.B) return a.B;
    curEnd = min(nb + A.chunk, a.B);
    nxt = claim();
    return nb;
  };
  auto issue_meta = [&](int b, unsigned char* dst) {
    if (b >= a.B) return;
    if (L.lane == 0) cp16_s(smem_addr(dst), A.sscal + b);
    const uint32_t rp = smem_addr(dst) + 16;
    for (int k = L.lane; k < min(SW, 35); k += 32) cp8_s(rp + k * 8, A.smeta + (size_t)b * SW + k);
  };
  auto issue_rows = [&](int b, const unsigned char* src, unsigned char* buf) {
    if (b >= a.B) return;
    const int s = reinterpret_cast<const int*>(src)[0];
    const int2* rp = reinterpret_cast<const int2*>(src + 16);
    const int n = min(s, PS_RS);
    const uint32_t dst = smem_addr(buf) + c16 * 16;
    const int col = (c16 & 7) * 4;
    for (int r = h; r < n; r += 2) {
      const int2 p = rp[3 + r];
      cp16_s(dst + r * 256, a.emb + (size_t)(c16 < 8 ? p.x : p.y) * 32 + col);
    }
    {
      const int2 p = rp[h];                                     // half 0: candidate 1, half 1: user vector
      cp16_s(dst + (PS_RS + h) * 256, a.emb + (size_t)(c16 < 8 ? p.x : p.y) * 32 + col);
    }
    if (h == 0) {
      const int2 p = rp[2];                                     // candidate 2
      cp16_s(dst + (PS_RS + 2) * 256, a.emb + (size_t)(c16 < 8 ? p.x : p.y) * 32 + col);
    } else {
      cp16_s(dst + (PS_RS + 3) * 256, a.scratch + (size_t)b * (TLSAN_SCR * 64) + 320 + c16 * 4);   // z
    }
  };

  pdl_wait();                                                   // z (k_dense_fwd_mma), smeta / sscal, the work counter
  pdl_trigger();
  int b0 = min(gw * A.chunk, a.B);
  curEnd = min(b0 + A.chunk, a.B);
  if (b0 < a.B) nxt = claim();
  int b1 = succ(b0), b2 = succ(b1);
  issue_meta(b0, ring(0));
  issue_meta(b1, ring(1));
  cp_commit();
  cp_wait_group<0>();
  __syncwarp();
  issue_rows(b0, ring(0), mine);
  cp_commit();

  for (int n = 0; b0 < a.B; ++n) {
    const int b = b0;
    unsigned char* buf = mine + (size_t)(n & 1) * PSC_BUF;
    cp_wait_group<0>();
    __syncwarp();
    issue_rows(b1, ring(n + 1), mine + (size_t)((n + 1) & 1) * PSC_BUF);
    issue_meta(b2, ring(n + 2));
    cp_commit();

    const unsigned char* mr = ring(n);
    const int4 sc4 = *reinterpret_cast<const int4*>(mr);
    const int s = sc4.x;
    const int2* rp = reinterpret_cast<const int2*>(mr + 16);
    const float (*rows)[64] = reinterpret_cast<const float (*)[64]>(buf);
    const int ntok = s + 1;
    const float2 zz = *reinterpret_cast<const float2*>(&rows[PS_RS + 3][L.f0]);
    auto item_x = [&](int it) -> float2 {        // session item `it`: staged row, else direct gather
      if (it < PS_RS) return *reinterpret_cast<const float2*>(&rows[it][L.f0]);
      int id, cr;
      if (it < 32) { const int2 p = rp[3 + it]; id = p.x; cr = p.y; }
      else { id = __ldg(a.hist_i_new + (size_t)b * a.S + it); cr = a.NI + __ldg(a.icl + id); }
      return ldg2(a.emb + (size_t)(L.half ? cr : id) * 32 + L.col);
    };
    SoftL2 ss; ss.init();
    for (int k = 0; k < ntok; k += 2) {
      const bool okB = k + 1 < ntok;
      float x[4];
      if (k == 0) { x[0] = zz.x; x[1] = zz.y; } else { const float2 e = item_x(k - 1); x[0] = e.x; x[1] = e.y; }
      if (okB) { const float2 e = item_x(k); x[2] = e.x; x[3] = e.y; } else { x[2] = 0.f; x[3] = 0.f; }
      float m1[4], m2[4];
      tile_maps(x, w, m1, m2);
      if (!okB) { m2[2] = -INFINITY; m2[3] = -INFINITY; }
      ss.push2(m2, x);
    }
    const float v[2] = {ss.acc[0] / ss.den[0], ss.acc[1] / ss.den[1]};
    const float2 q1 = *reinterpret_cast<const float2*>(&rows[PS_RS][L.f0]);
    const float2 p = *reinterpret_cast<const float2*>(&rows[PS_RS + 1][L.f0]);
    const float ut[2] = {v[0] + p.x, v[1] + p.y};
    const float l1 = warp_sum_f(fmaf(ut[0], q1.x, ut[1] * q1.y)) + __int_as_float(sc4.z);
    if (L.lane == 0) a.logits[(size_t)b * A.ncand] = l1;
    if (a.ut) st2(a.ut + (size_t)b * 64 + L.f0, ut[0], ut[1]);
    if (A.ncand > 1) {                            // Model.eval_auc second run (model.py:251-261): same u_t, other item
      const float2 q2 = *reinterpret_cast<const float2*>(&rows[PS_RS + 2][L.f0]);
      const float l2 = warp_sum_f(fmaf(ut[0], q2.x, ut[1] * q2.y)) + __int_as_float(sc4.w);
      if (L.lane == 0) a.logits[(size_t)b * A.ncand + 1] = l2;
    }
    __syncwarp();
    b0 = b1; b1 = b2; b2 = succ(b2);
  }
  cp_wait_group<0>();
}

// ------------------------------------------------------------------ launchers
size_t tlsan_long_meta_bytes(int B, int L) { return (size_t)B * L * sizeof(int4); }

// smeta / sscal may be NULL (scoring: long-term part only)
// `part`: the partition block of the workspace; its last 16 bytes hold the work counter of the long-term forward, reset
// here for the kernel launched with `fwd_ctas` CTAs per SM
static int* pf_counter(const void* part);
static int pf_chunk();
static int pf_grid(int B, int ctas_per_sm);
int tlsan_launch_long_meta(const FArgs& a, void* meta, void* smeta, void* sscal, void* part, int fwd_ctas,
                           int score_ncand, cudaStream_t st) {
  const long long n = (long long)a.B * a.L + (smeta ? a.B : 0);
  tlsan_launch_k(k_long_meta, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, a, reinterpret_cast<int4*>(meta),
                 reinterpret_cast<int2*>(smeta), reinterpret_cast<int4*>(sscal), pf_counter(part),
                 pf_grid(a.B, fwd_ctas) * PF_WARPS * pf_chunk(), score_ncand);
  TLSAN_CHECK_LAUNCH("k_long_meta");
  return TLSAN_OK;
}

// samples a forward warp claims at a time: small enough that the last claims end together (a sample is ~2.8 us of one
// warp), large enough that the claims (one atomic on one address each) stay far below the L2's same-address rate
static int pf_chunk() {
  static int v = 0;
  if (!v) { const char* e = getenv("TLSAN_PF_CHUNK"); v = e ? atoi(e) : 4; if (v < 1 || v > 1024) v = 4; }
  return v;
}
static int* pf_counter(const void* part) {
  return reinterpret_cast<int*>(reinterpret_cast<char*>(const_cast<void*>(part)) + tlsan_partition_bytes() - 16);
}
// grids of the three pipelined kernels (the partition is computed for exactly these)
static int pf_grid(int B, int ctas_per_sm) {
  const int need = (B + PF_WARPS - 1) / PF_WARPS, cap = tlsan_num_sms() * ctas_per_sm;
  return need < cap ? need : cap;
}
size_t tlsan_partition_bytes() {
  return ((size_t)3 * (TLSAN_MAX_GRID * PF_WARPS + 4) + (size_t)2 * ((PART_MAXB + 7) / 8) + 2 * PART_CTAS + 64) * sizeof(int);
}

// balanced partitions for the forward (fwd_ctas CTAs per SM), backward and short-term kernels of one batch
int tlsan_launch_partition(const FArgs& a, int fwd_ctas, bool train, void* part, cudaStream_t st) {
  PartArgs p;
  p.sl = a.sl; p.sl_new = a.sl_new; p.B = a.B;
  int* base = reinterpret_cast<int*>(part);
  for (int t = 0; t < 3; ++t) p.starts[t] = base + (size_t)t * (TLSAN_MAX_GRID * PF_WARPS + 4);
  (void)fwd_ctas;                                                // the forward claims its samples dynamically
  if (!train) return TLSAN_OK;                                   // scoring: nothing to partition
  p.nw[0] = 0;
  p.nw[1] = train ? pf_grid(a.B, 2) * PF_WARPS : 0;
  p.nw[2] = train ? pf_grid(a.B, 2) * PF_WARPS : 0;
  const int nchunk = (a.B + 7) / 8;
  if (a.B > PART_MAXB) {                                         // chunk prefixes would not fit shared memory: equal counts
    k_partition_uniform<<<(3 * (TLSAN_MAX_GRID * PF_WARPS + 1) + 255) / 256, 256, 0, st>>>(p);
    TLSAN_CHECK_LAUNCH("k_partition_uniform");
    return TLSAN_OK;
  }
  // scratch of the two kernels lives behind the three boundary arrays
  unsigned int* scan = reinterpret_cast<unsigned int*>(base + (size_t)3 * (TLSAN_MAX_GRID * PF_WARPS + 4));
  unsigned int* tot = scan + (size_t)2 * ((PART_MAXB + 7) / 8);
  const int cpc = (nchunk + PART_CTAS - 1) / PART_CTAS;          // chunks per CTA
  k_part_scan<<<PART_CTAS, 256, 0, st>>>(p, scan, tot, nchunk, cpc);
  TLSAN_CHECK_LAUNCH("k_part_scan");
  k_part_bounds<<<(3 * (TLSAN_MAX_GRID * PF_WARPS + 1) + 255) / 256, 256, 0, st>>>(p, scan, tot, nchunk, cpc);
  TLSAN_CHECK_LAUNCH("k_part_bounds");
  return TLSAN_OK;
}
FArgs tlsan_make_fargs(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b);
int tlsan_launch_partition_batch(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int fwd_ctas,
                                 void* part, cudaStream_t st) {
  const FArgs a = tlsan_make_fargs(d, p, b);
  return tlsan_launch_partition(a, fwd_ctas, true, part, st);
}
static const int* part_starts(const void* part, int which) {
  return reinterpret_cast<const int*>(part) + (size_t)which * (TLSAN_MAX_GRID * PF_WARPS + 4);
}

template <int KIND>
static int launch_pf_long(const FArgs& a, const void* meta, const void* part, int ctas_per_sm, int* grid_out,
                          cudaStream_t st) {
  PfArgs A;
  A.a = a; A.meta = reinterpret_cast<const int4*>(meta); A.g = pf_geo(a.L, KIND == 3);
  A.starts = part_starts(part, 1);
  A.counter = pf_counter(part); A.chunk = pf_chunk();
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_pf_long<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set = true;
  }
  const int gr = pf_grid(a.B, ctas_per_sm);
  if (grid_out) *grid_out = gr;
  tlsan_launch_k(k_pf_long<KIND>, dim3(gr), dim3(PF_THREADS), (size_t)A.g.total, st, A);
  TLSAN_CHECK_LAUNCH(KIND == 1 ? "k_pf_long<fwd>" : "k_pf_long<bwd>");
  return TLSAN_OK;
}

// ctas_per_sm (forward): 3 fills the SM.  The samples are claimed dynamically, so CTAs that become resident late
// (the sort kernels share the SMs) cost nothing; tlsan_launch_long_meta must have been given the same value (it
// seeds the work counter with grid x warps x chunk).
int tlsan_launch_long_fwd_pf(const FArgs& a, const void* meta, const void* part, int ctas_per_sm, cudaStream_t st) {
  return launch_pf_long<1>(a, meta, part, ctas_per_sm, nullptr, st);
}
int tlsan_launch_bwd_long_pf(const FArgs& a, const void* meta, const void* part, int* grid_b, cudaStream_t st) {
  return launch_pf_long<3>(a, meta, part, 2, grid_b, st);
}

int tlsan_launch_short_pf(const FArgs& a, const void* smeta, const void* sscal, const void* part, int* grid_a,
                          cudaStream_t st) {
  PsArgs A;
  A.a = a; A.smeta = reinterpret_cast<const int2*>(smeta); A.sscal = reinterpret_cast<const int4*>(sscal);
  A.starts = part_starts(part, 2);
  const int smem = PS_PER_WARP * PF_WARPS;
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_pf_short, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int gr = pf_grid(a.B, 2);
  if (grid_a) *grid_a = gr;
  tlsan_launch_k(k_pf_short, dim3(gr), dim3(PF_THREADS), (size_t)smem, st, A);
  TLSAN_CHECK_LAUNCH("k_pf_short");
  return TLSAN_OK;
}

size_t tlsan_score_meta_bytes(int B, int S) {
  return tlsan_align_up((size_t)B * (S + 3) * sizeof(int2), 256) + tlsan_align_up((size_t)B * sizeof(int4), 256);
}
int tlsan_launch_score_pf(const FArgs& a, const void* smeta, const void* sscal, const void* part, int ncand,
                          cudaStream_t st) {
  PscArgs A;
  A.a = a; A.smeta = reinterpret_cast<const int2*>(smeta); A.sscal = reinterpret_cast<const int4*>(sscal);
  A.counter = pf_counter(part) + 1; A.chunk = pf_chunk(); A.ncand = ncand;
  const int smem = PSC_PER_WARP * PF_WARPS;
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_pf_score, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  tlsan_launch_k(k_pf_score, dim3(pf_grid(a.B, 3)), dim3(PF_THREADS), (size_t)smem, st, A);
  TLSAN_CHECK_LAUNCH("k_pf_score");
  return TLSAN_OK;
}
