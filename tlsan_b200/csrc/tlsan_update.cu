// Gradient aggregation and the optimiser step (reference model.py:164-172,185-205):
//   k_row_reduce   deterministic segmented reduce of the per-occurrence gradient rows
//   k_finalize1    fixed-order sum of the per-CTA dense partials
//   k_table_sumsq  sum of squares of the four L2-regularised tables (l2_loss, model.py:164-169)
//   k_finalize2    global norm, clip scale, loss, SGD on the 4449 small parameters
//   k_apply_rows / k_apply_cate   W <- W - lr * scale * (g_sparse + reg * W) for every table row
//   k_label_rank   full-catalogue rank of the label item (model.py:140-156)
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tlsan_common.cuh"

// ------------------------------------------------------------------ balanced segmented reduce
// Giving one warp one ROW (round 1) fails on Zipf item popularity (and 15 categories on the Movies-TV shape)
// a few warps own segments of thousands of occurrences while the rest idle -- ncu: long_scoreboard 74 % at 41 % of the
// warps active, 3.1 TB/s.  Here the SORTED OCCURRENCE LIST is what is divided: warp w sums positions
// [w R, (w+1) R) of the item / category part (R = ceil(T / #warps) rounded up to 16, T = seg_off[NI + NC]), 16 rows
// (4 KB) in flight per warp, walking the segment ends it meets:
//   * a segment that lies inside the range is written straight to g_i / g_b;
//   * the range's first segment, if it began in an earlier range, goes to head[w]; its last one, if it continues in
//     the next range, to tail[w] (a range inside ONE long segment is a head only);
//   * k_row_fix (one warp per row) then writes  g[r] = tail[w_a] + head[w_a + 1] + ... + head[w_b]  for the rows whose
//     segment spans ranges w_a < w_b, and zeros for the rows without occurrences.
// Ranges, walk order and the order of the fix-up sum are functions of the batch and the grid only: bit-reproducible.
// The user rows (short segments, payload indexed by sample) keep the row-per-warp loop.
#define RR_PSTRIDE 72          // floats per partial: 64 + item_b gradient + pad
__device__ __forceinline__ int rr_range(int T, int NW) {
  const int r = ((T + NW - 1) / NW + 15) & ~15;
  return r < 16 ? 16 : r;
}
__global__ void __launch_bounds__(256, 4) k_row_reduce_bal(int NI, int NC, int NU, int L, int S, int spsh, int PU,
                                                        const int* __restrict__ seg_off, const int* __restrict__ keys,
                                                        const int* __restrict__ vals,
                                                        const float* __restrict__ rows_i,
                                                        const float* __restrict__ rows_u,
                                                        const float* __restrict__ gscal, float* __restrict__ g_i,
                                                        float* __restrict__ g_b, float* __restrict__ g_u,
                                                        float* __restrict__ head, float* __restrict__ tail) {
  const int lane = threadIdx.x & 31;
  const int NW = gridDim.x * 8;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int smask = (1 << spsh) - 1;
  const int T = seg_off[NI + NC];
  const int R = rr_range(T, NW);
  const int p0 = min(gw * R, T), p1 = min(p0 + R, T);
  pdl_wait();                                       // gradient rows of the backward kernels (the sort finished long ago)
  pdl_trigger();
  if (p0 < p1) {
    // lane j of a batch at position p holds the key of position p - 1 + j (j = 0..17): the neighbours decide whether
    // the first / last segment of the range is shared with the adjacent ranges
    auto key_at = [&](int p, int j) { const int q = p - 1 + j; return (j < 18 && q >= 0 && q < T) ? __ldg(keys + q) : -1; };
    auto occ_at = [&](int p, int j) { const int q = p - 1 + j; return (j >= 1 && j < 17 && q < p1) ? __ldg(vals + q) : 0; };
    int key_n = key_at(p0, lane), occ_n = occ_at(p0, lane);
    float2 acc = make_float2(0.f, 0.f);
    float accb = 0.f;
    bool first = true;                              // still inside the first segment of the range
    const bool head_shared = __shfl_sync(0xffffffffu, key_n, 0) == __shfl_sync(0xffffffffu, key_n, 1);   // key[p0 - 1] == key[p0]
    for (int p = p0; p < p1; p += 16) {
      const int cnt = min(16, p1 - p);
      const int key = key_n, occ = occ_n;
      float2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        v[i] = i < cnt ? __ldcs(reinterpret_cast<const float2*>(rows_i + (size_t)(p + i) * 64) + lane) : make_float2(0.f, 0.f);
      if (p + 16 < p1) { key_n = key_at(p + 16, lane); occ_n = occ_at(p + 16, lane); }
      // item_b gradient: the candidate slot (j == L + S) of sample b contributes gscal[b]
      const bool is_cand = lane >= 1 && lane <= cnt && (occ & smask) == L + S;
      const float gb = is_cand ? __ldg(gscal + (occ >> spsh)) : 0.f;
      const unsigned candmask = __ballot_sync(0xffffffffu, is_cand) >> 1;
      const int knext = __shfl_down_sync(0xffffffffu, key, 1);
      // bit i: position p + i is the last of its segment as far as this range can tell
      const bool last_batch = p + 16 >= p1;             // the range's last position always closes a (partial) segment
      const unsigned endmask = (__ballot_sync(0xffffffffu, lane >= 1 && (key != knext || (last_batch && lane == cnt))) >> 1) & 0xffffu;
      const bool tail_shared = last_batch && __shfl_sync(0xffffffffu, key, cnt) == __shfl_sync(0xffffffffu, key, cnt + 1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < cnt) {
          acc.x += v[i].x; acc.y += v[i].y;
          if (candmask & (1u << i)) accb += __shfl_sync(0xffffffffu, gb, i + 1);
          if (endmask & (1u << i)) {
            const int k = __shfl_sync(0xffffffffu, key, i + 1);
            const bool to_head = first && head_shared;
            const bool to_tail = !to_head && tail_shared && i == cnt - 1;
            float* dst = to_head ? head + (size_t)gw * RR_PSTRIDE : to_tail ? tail + (size_t)gw * RR_PSTRIDE : g_i + (size_t)k * 64;
            reinterpret_cast<float2*>(dst)[lane] = acc;
            if (lane == 0) {
              if (to_head || to_tail) dst[64] = accb;
              else if (k < NI) g_b[k] = accb;
            }
            acc = make_float2(0.f, 0.f); accb = 0.f; first = false;
          }
        }
      }
    }
  }
  // ---- user rows: payload rows_u[sample], one warp per row
  const int NR = NI + NC + NU;
  for (int r = NI + NC + gw; r < NR; r += NW) {
    const int lo = __ldg(seg_off + r), hi = __ldg(seg_off + r + 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = lo; base < hi; base += 32) {
      const int mine = base + lane < hi ? __ldg(vals + base + lane) : 0;
      const int cnt = min(32, hi - base);
      for (int k = 0; k < cnt; ++k) {
        const int b = __shfl_sync(0xffffffffu, mine, k) >> spsh;
        const float* src = rows_u + (size_t)b * PU;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          if (lane + 32 * qq < PU) acc[qq] += __ldg(src + lane + 32 * qq);
      }
    }
#pragma unroll
    for (int qq = 0; qq < 4; ++qq)
      if (lane + 32 * qq < PU) g_u[(size_t)(r - NI - NC) * PU + lane + 32 * qq] = acc[qq];
  }
}

// rows whose segment spans several ranges: tail of the first range + heads of the others, in range order; rows
// without occurrences: zeros.  NW = the warp count of k_row_reduce_bal.
// One THREAD classifies one row (coalesced segment bounds, 256 rows per CTA); the rows that need work -- a few
// hundred of 22 721 on the Electronics shape, none of the ~600 000 of a row-sharded compact table -- are then taken
// by the CTA's warps in row order.
__global__ void __launch_bounds__(256) k_row_fix(int NI, int NC, int NW, const int* __restrict__ seg_off,
                                                 const float* __restrict__ head, const float* __restrict__ tail,
                                                 float* __restrict__ g_i, float* __restrict__ g_b) {
  __shared__ int todo[256];
  __shared__ int ntodo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // rows are dealt to the CTAs in groups of 8 (32-B sectors of seg_off): the rows that need a fix-up are the hot ones
  // and cluster at neighbouring ids -- in contiguous blocks of 256 a few CTAs got all of them (20 us, 7 us this way)
  const int r0 = ((threadIdx.x >> 3) * gridDim.x + blockIdx.x) * 8 + (threadIdx.x & 7);
  const int R = rr_range(__ldg(seg_off + NI + NC), NW);
  if (threadIdx.x == 0) ntodo = 0;
  __syncthreads();
  bool need = false;
  if (r0 < NI + NC) {
    const int lo = __ldg(seg_off + r0), hi = __ldg(seg_off + r0 + 1);
    need = lo >= hi || lo / R != (hi - 1) / R;      // else: written directly by its range
  }
  // compact the flagged rows (any fixed order: every row's sum is computed on its own)
  const unsigned m = __ballot_sync(0xffffffffu, need);
  __shared__ int wcnt[8];
  if (lane == 0) wcnt[warp] = __popc(m);
  __syncthreads();
  int base = 0;
  for (int q = 0; q < warp; ++q) base += wcnt[q];
  if (need) todo[base + __popc(m & ((1u << lane) - 1))] = r0;
  if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; ++q) t += wcnt[q]; ntodo = t; }
  __syncthreads();
  // EVERY thread waits, also those with nothing to fix: a grid that finished without waiting would let the next
  // kernel of the chain start while k_row_reduce_bal is still writing
  pdl_wait();
  pdl_trigger();
  for (int k = warp; k < ntodo; k += 8) {
    const int r = todo[k];
    const int lo = __ldg(seg_off + r), hi = __ldg(seg_off + r + 1);
    float2 acc = make_float2(0.f, 0.f);
    float accb = 0.f;
    if (lo < hi) {
      const int wa = lo / R, wb = (hi - 1) / R;
      acc = reinterpret_cast<const float2*>(tail + (size_t)wa * RR_PSTRIDE)[lane];
      accb = tail[(size_t)wa * RR_PSTRIDE + 64];
      for (int w = wa + 1; w <= wb; ++w) {
        const float2 h = reinterpret_cast<const float2*>(head + (size_t)w * RR_PSTRIDE)[lane];
        acc.x += h.x; acc.y += h.y;
        accb += head[(size_t)w * RR_PSTRIDE + 64];
      }
    }
    reinterpret_cast<float2*>(g_i + (size_t)r * 64)[lane] = acc;
    if (r < NI && lane == 0) g_b[r] = accb;
  }
}

// ------------------------------------------------------------------ dense partials
// part_a: fused kernel A (short FWA, loss, sumsq; + dense grads in the FFMA variant)
// part_b: long backward (long FWA, gamma, sumsq) ; part_c: k_dense_grad (dense kernel/bias), grid_c = 0 if unused
// CTA = 32 entries x 8 row groups: group q adds partial rows q, q+8, ... (coalesced 128-B reads),
// then the 8 group sums are added in order.  Fixed order, no atomics.
__global__ void __launch_bounds__(256) k_finalize1(const float* __restrict__ part_a, int grid_a,
                                                   const float* __restrict__ part_b, int grid_b,
                                                   const float* __restrict__ part_c, int grid_c,
                                                   float* __restrict__ dgrad) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (e < TLSAN_PART) {
    const bool is_dense = e >= TLSAN_OFF_WD && e < TLSAN_OFF_BD + 64;
    const bool from_a = (e >= TLSAN_OFF_W1S && e < TLSAN_OFF_WD) || (is_dense && grid_c == 0) ||
                        e == TLSAN_PART_LOSS || e == TLSAN_PART_SUMSQ;
    const bool from_b = e < TLSAN_OFF_W1S || e == TLSAN_OFF_GAMMA || e == TLSAN_PART_SUMSQ;
    if (is_dense)
      for (int c = q; c < grid_c; c += 8) s += part_c[(size_t)c * TLSAN_PART + e];
    if (from_a)
      for (int c = q; c < grid_a; c += 8) s += part_a[(size_t)c * TLSAN_PART + e];
    if (from_b)
      for (int c = q; c < grid_b; c += 8) s += part_b[(size_t)c * TLSAN_PART + e];
  }
  sh[q][lane] = s;
  __syncthreads();
  if (q == 0 && e < TLSAN_PART) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][lane];
    dgrad[e] = t;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) r += sh[w];
  return r;
}

// tsq[cta][4] = partial sum of squares of (item_emb, cate_emb, user_emb, usert_emb)
__global__ void __launch_bounds__(256) k_table_sumsq(const float* __restrict__ emb, const float* __restrict__ usert,
                                                     long long n_item, long long n_cate, long long n_user,
                                                     long long n_usert, float* __restrict__ tsq) {
  __shared__ float sh[8];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_emb = n_item + n_cate + n_user;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_emb; e += stride) {
    const float v = emb[e];
    const int w = e < n_item ? 0 : (e < n_item + n_cate ? 1 : 2);
    s[w] = fmaf(v, v, s[w]);
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_usert; e += stride) {
    const float v = usert[e];
    s[3] = fmaf(v, v, s[3]);
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float r = block_sum_256(s[w], sh);
    if (threadIdx.x == 0) tsq[blockIdx.x * 4 + w] = r;
  }
}

// ---- optimisers of init_optimizer (model.py:188-195).  Every trainable element sees its AGGREGATED, clipped gradient
// g = (sum of its slices + reg * w) * scale each step: the L2 term makes the IndexedSlices of a table cover all of its
// rows, so TF's sparse apply ops reduce to the dense formulas (tf.train.*Optimizer defaults of TF 1.8):
//   sgd      w -= lr g
//   adam     m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; w -= lr c1 m / (sqrt(v) + eps), c1 = sqrt(1-b2^t)/(1-b1^t)
//   rmsprop  ms = rho ms + (1-rho) g^2 ; mom = momentum mom + lr g / sqrt(ms + eps) ; w -= mom      (ms starts at 1)
//   adadelta acc = rho acc + (1-rho) g^2 ; u = sqrt(acc_u + eps) / sqrt(acc + eps) g ; acc_u = rho acc_u + (1-rho) u^2 ;
//            w -= lr u
// The two slot arrays mirror the weight buffer (element of w at wbase + i <-> s1[i], s2[i]).
struct OptArgs { int kind; float a, b, eps, c1; float* s1; float* s2; const float* wbase; };
__device__ __forceinline__ float opt_step(float w, float g, float lr, const OptArgs& o, const float* wp) {
  if (o.kind == TLSAN_OPT_SGD) return w - lr * g;
  const size_t i = (size_t)(wp - o.wbase);
  if (o.kind == TLSAN_OPT_ADAM) {
    const float m = o.a * o.s1[i] + (1.f - o.a) * g, v = o.b * o.s2[i] + (1.f - o.b) * g * g;
    o.s1[i] = m; o.s2[i] = v;
    return w - lr * o.c1 * m / (sqrtf(v) + o.eps);
  }
  if (o.kind == TLSAN_OPT_RMSPROP) {
    const float ms = o.a * o.s1[i] + (1.f - o.a) * g * g;
    const float mom = o.b * o.s2[i] + lr * g / sqrtf(ms + o.eps);
    o.s1[i] = ms; o.s2[i] = mom;
    return w - mom;
  }
  const float acc = o.a * o.s1[i] + (1.f - o.a) * g * g;             // adadelta
  const float u = sqrtf(o.s2[i] + o.eps) / sqrtf(acc + o.eps) * g;
  o.s1[i] = acc; o.s2[i] = o.a * o.s2[i] + (1.f - o.a) * u * u;
  return w - lr * u;
}
static OptArgs opt_args(const tlsan_opt_t* opt, const float* wbase) {
  OptArgs o;
  o.kind = opt ? opt->kind : TLSAN_OPT_SGD; o.a = o.b = o.eps = o.c1 = 0.f; o.s1 = o.s2 = nullptr; o.wbase = wbase;
  if (!opt || o.kind == TLSAN_OPT_SGD) return o;
  o.s1 = opt->slot1; o.s2 = opt->slot2;
  if (o.kind == TLSAN_OPT_ADAM) {
    o.a = opt->beta1; o.b = opt->beta2; o.eps = opt->epsilon;
    o.c1 = (float)(sqrt(1.0 - pow((double)opt->beta2, (double)opt->step)) / (1.0 - pow((double)opt->beta1, (double)opt->step)));
  } else if (o.kind == TLSAN_OPT_RMSPROP) {
    o.a = opt->rho; o.b = opt->momentum; o.eps = opt->epsilon;
  } else {
    o.a = opt->rho; o.eps = opt->epsilon;
  }
  return o;
}

// global norm (TF style: un-aggregated slices, see oracle header), clip scale, loss, dense SGD
__device__ __forceinline__ double block_sum_1024d(double v, double* sh) {   // fixed order: lanes, then warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < 32; ++w) r += sh[w];
  return r;
}

// extra_sq[n_extra] (optional): partial sums of squares of table rows that live elsewhere (row-sharded item_emb)
__global__ void __launch_bounds__(1024) k_finalize2(const float* __restrict__ dgrad, const float* __restrict__ tsq,
                                                    int ntsq, const float* __restrict__ extra_sq, int n_extra,
                                                    float invB, float lr, float reg, float clip,
                                                    float* __restrict__ dense, float* __restrict__ stats,
                                                    const OptArgs opt) {
  __shared__ double sh[32];
  float g[5];
  double s = 0.0;
  pdl_wait();
  pdl_trigger();
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int e = threadIdx.x + 1024 * q;
    g[q] = e < TLSAN_DENSE_COUNT ? dgrad[e] : 0.f;
    s += (double)g[q] * (double)g[q];
  }
  const double dense_sq = block_sum_1024d(s, sh);
  double t = 0.0;
  for (int c = threadIdx.x; c < ntsq; c += 1024)
    t += (double)tsq[c * 4] + (double)tsq[c * 4 + 1] + (double)tsq[c * 4 + 2] + (double)tsq[c * 4 + 3];
  for (int c = threadIdx.x; c < n_extra; c += 1024) t += (double)extra_sq[c];
  t = block_sum_1024d(t, sh);
  const double sq = (double)dgrad[TLSAN_PART_SUMSQ] + dense_sq + (double)reg * (double)reg * t;
  const float norm = (float)sqrt(sq);
  const float scale = clip * fminf(1.f / norm, 1.f / clip);   // tf.clip_by_global_norm
  if (threadIdx.x == 0) {
    const float l2 = (float)(0.5 * t);
    const float bce = dgrad[TLSAN_PART_LOSS] * invB;
    stats[TLSAN_STAT_LOSS] = bce + reg * l2;
    stats[TLSAN_STAT_BCE] = bce;
    stats[TLSAN_STAT_NORM] = norm;
    stats[TLSAN_STAT_SCALE] = scale;
    stats[TLSAN_STAT_L2] = l2;
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int e = threadIdx.x + 1024 * q;
    if (e < TLSAN_DENSE_COUNT) {
      if (opt.kind == TLSAN_OPT_SGD) dense[e] -= lr * (g[q] * scale);
      else dense[e] = opt_step(dense[e], g[q] * scale, lr, opt, dense + e);
    }
  }
}

// element-wise update of item_emb, user_emb, usert_emb and item_b from the reduced buffers
__global__ void __launch_bounds__(256) k_apply_rows(int NI, int NC, int NU, int L, int PU, float* __restrict__ emb,
                                                    float* __restrict__ usert, float* __restrict__ item_b,
                                                    const float* __restrict__ g_i, const float* __restrict__ g_b,
                                                    const float* __restrict__ g_u, float lr, float reg,
                                                    const float* __restrict__ stats, const OptArgs opt) {
  pdl_wait();                                       // the clip scale of k_finalize2
  pdl_trigger();
  const float scale = stats[TLSAN_STAT_SCALE];
  // float4 units: item rows (8 per row), user rows (8 per row); then scalars: usert, item_b
  const long long v1 = (long long)NI * 8, v2 = v1 + (long long)NU * 8;
  const long long n3 = (long long)NU * L, n4 = n3 + NI;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < v2 + n4; e += stride) {
    if (e < v2) {
      float4* wp; float4 g;
      if (e < v1) {
        wp = reinterpret_cast<float4*>(emb) + e;
        g = *reinterpret_cast<const float4*>(g_i + (e >> 3) * 64 + (e & 7) * 4);
      } else {
        const long long x = e - v1;
        wp = reinterpret_cast<float4*>(emb + (size_t)(NI + NC) * 32) + x;
        g = *reinterpret_cast<const float4*>(g_u + (x >> 3) * PU + (x & 7) * 4);
      }
      float4 w = *wp;
      if (opt.kind == TLSAN_OPT_SGD) {
        w.x -= lr * ((g.x + reg * w.x) * scale); w.y -= lr * ((g.y + reg * w.y) * scale);
        w.z -= lr * ((g.z + reg * w.z) * scale); w.w -= lr * ((g.w + reg * w.w) * scale);
      } else {
        const float* f = reinterpret_cast<const float*>(wp);
        w.x = opt_step(w.x, (g.x + reg * w.x) * scale, lr, opt, f); w.y = opt_step(w.y, (g.y + reg * w.y) * scale, lr, opt, f + 1);
        w.z = opt_step(w.z, (g.z + reg * w.z) * scale, lr, opt, f + 2); w.w = opt_step(w.w, (g.w + reg * w.w) * scale, lr, opt, f + 3);
      }
      *wp = w;
    } else {
      const long long x = e - v2;
      if (x < n3) {
        const long long u = x / L; const int t = (int)(x - u * L);
        const float w = usert[x];
        const float g = (g_u[u * PU + 32 + t] + reg * w) * scale;
        usert[x] = opt.kind == TLSAN_OPT_SGD ? w - lr * g : opt_step(w, g, lr, opt, usert + x);
      } else {
        const long long y = x - n3;
        const float g = g_b[y] * scale;
        // item_b has no L2 term: its gradient is a truly sparse IndexedSlices.  TF's sparse RMSProp / Adadelta kernels
        // leave the rows outside the slices (weights AND slots) untouched; sparse Adam (non-lazy) decays every row,
        // which is the dense formula with g = 0.  A row is "outside" when its summed gradient is exactly 0.
        if (opt.kind == TLSAN_OPT_SGD) item_b[y] = item_b[y] - lr * g;
        else if (opt.kind == TLSAN_OPT_ADAM || g != 0.f) item_b[y] = opt_step(item_b[y], g, lr, opt, item_b + y);
      }
    }
  }
}

// one CTA per category: sum the cate halves of its items' reduced rows in CSR order
// blockDim = 256, or 1024 when categories are few and large (Movies-TV: 15 categories of ~1 900 items): warp w adds
// items lo + w, lo + w + nw, ... in that order, the warp sums are added in warp order -- fixed for a given shape
__global__ void __launch_bounds__(1024) k_apply_cate(int NI, float* __restrict__ emb, const float* __restrict__ g_i,
                                                     const int* __restrict__ cate_off,
                                                     const int* __restrict__ cate_items, float lr, float reg,
                                                     const float* __restrict__ stats, const OptArgs opt) {
  __shared__ float sh[32][32];
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int lo = cate_off[k], hi = cate_off[k + 1];
  float acc = 0.f;
  int n = lo + warp;
  for (; n + 7 * nw < hi; n += 8 * nw) {            // 8 independent row loads in flight per lane
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(g_i + (size_t)__ldg(cate_items + n + q * nw) * 64 + 32 + lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += v[q];
  }
  for (; n < hi; n += nw) acc += __ldg(g_i + (size_t)__ldg(cate_items + n) * 64 + 32 + lane);
  sh[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float g = g_i[(size_t)(NI + k) * 64 + 32 + lane];   // direct u_cate occurrences
    for (int w = 0; w < nw; ++w) g += sh[w][lane];
    const float scale = stats[TLSAN_STAT_SCALE];
    float* wp = emb + (size_t)(NI + k) * 32 + lane;
    const float w = *wp;
    const float gg = (g + reg * w) * scale;
    *wp = opt.kind == TLSAN_OPT_SGD ? w - lr * gg : opt_step(w, gg, lr, opt, wp);
  }
  // launched as the programmatic dependent of k_apply_rows: everything above needs only what k_finalize2 and the
  // kernels before it wrote (category rows are nobody else's), so it ran BESIDE k_apply_rows; waiting for it here
  // keeps the chain's invariant (a kernel finishes only after its predecessor)
  pdl_wait();
  pdl_trigger();
}

// rank[b] = #{ j : s_j > s_label  or  (s_j == s_label and j < label) },  s = u_t . all_emb^T + item_b
__global__ void __launch_bounds__(256) k_label_rank(int NI, const float* __restrict__ emb,
                                                    const float* __restrict__ item_b, const int* __restrict__ icl,
                                                    const float* __restrict__ ut, const int* __restrict__ label,
                                                    int* __restrict__ rank) {
  __shared__ float su[64];
  __shared__ int scnt[8];
  const int b = blockIdx.x;
  if (threadIdx.x < 64) su[threadIdx.x] = ut[(size_t)b * 64 + threadIdx.x];
  __syncthreads();
  auto score = [&](int j) {
    const float4* pi = reinterpret_cast<const float4*>(emb + (size_t)j * 32);
    const float4* pc = reinterpret_cast<const float4*>(emb + (size_t)(NI + __ldg(icl + j)) * 32);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(pi + q);
      s = fmaf(su[4 * q], v.x, s); s = fmaf(su[4 * q + 1], v.y, s);
      s = fmaf(su[4 * q + 2], v.z, s); s = fmaf(su[4 * q + 3], v.w, s);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(pc + q);
      s = fmaf(su[32 + 4 * q], v.x, s); s = fmaf(su[32 + 4 * q + 1], v.y, s);
      s = fmaf(su[32 + 4 * q + 2], v.z, s); s = fmaf(su[32 + 4 * q + 3], v.w, s);
    }
    return s + __ldg(item_b + j);
  };
  const int lab = label[b];
  const float sl = score(lab);
  int cnt = 0;
  for (int j = threadIdx.x; j < NI; j += 256) {
    const float s = score(j);
    cnt += (s > sl) || (s == sl && j < lab);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) scnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += scnt[w];
    rank[b] = t;
  }
}

// ------------------------------------------------------------------ launchers
int tlsan_launch_row_reduce(const tlsan_dims_t& d, const TlsanWs& w, char* ws, const int32_t* sorted_vals,
                            float* g_i, float* g_b, float* g_u, cudaStream_t st) {
  const int* seg_off = reinterpret_cast<const int*>(ws + w.seg_off);
  const int rgrid = tlsan_num_sms() * 4;
  // the sorted keys sit in the ping-pong buffer of the same parity as the sorted occurrence ids
  const int* keys = reinterpret_cast<const int*>(
      ws + (sorted_vals == reinterpret_cast<const int32_t*>(ws + w.vals_b) ? w.keys_b : w.keys_a));
  float* head = reinterpret_cast<float*>(ws + w.rpart);
  float* tail = head + (size_t)TLSAN_MAX_GRID * 8 * RR_PSTRIDE;
  tlsan_launch_kl(1, k_row_reduce_bal, dim3(rgrid), dim3(256), 0, st, d.NI, d.NC, d.NU, d.L, d.S, w.SPSH, w.PU, seg_off,
                  keys, (const int*)sorted_vals, reinterpret_cast<const float*>(ws + w.rows_i),
                  reinterpret_cast<const float*>(ws + w.rows_u), reinterpret_cast<const float*>(ws + w.gscal), g_i, g_b,
                  g_u, head, tail);
  TLSAN_CHECK_LAUNCH("k_row_reduce_bal");
  tlsan_launch_kl(1, k_row_fix, dim3((d.NI + d.NC + 255) / 256), dim3(256), 0, st, d.NI, d.NC, rgrid * 8, seg_off,
                  (const float*)head, (const float*)tail, g_i, g_b);
  TLSAN_CHECK_LAUNCH("k_row_fix");
  return TLSAN_OK;
}

int tlsan_launch_finalize1(const TlsanWs& w, char* ws, int grid_a, int grid_b, int grid_c, float* dgrad,
                           cudaStream_t st) {
  k_finalize1<<<(TLSAN_PART + 31) / 32, 256, 0, st>>>(reinterpret_cast<const float*>(ws + w.part_a), grid_a,
                                                        reinterpret_cast<const float*>(ws + w.part_b), grid_b,
                                                        reinterpret_cast<const float*>(ws + w.part_c), grid_c,
                                                        dgrad);
  TLSAN_CHECK_LAUNCH("k_finalize1");
  return TLSAN_OK;
}

static int tsq_grid() {
  int ntsq = tlsan_num_sms() * 2;
  return ntsq > TLSAN_MAX_GRID ? TLSAN_MAX_GRID : ntsq;
}

// ||W||^2 partials of the four regularised tables; depends on the weights only, so a fused train step
// runs it on the side stream while the forward kernels run
int tlsan_launch_table_sumsq(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                             cudaStream_t st) {
  k_table_sumsq<<<tsq_grid(), 256, 0, st>>>(p.emb, p.usert, (long long)d.NI * 32, (long long)d.NC * 32,
                                            (long long)d.NU * 32, (long long)d.NU * d.L,
                                            reinterpret_cast<float*>(ws + w.tsq));
  TLSAN_CHECK_LAUNCH("k_table_sumsq");
  return TLSAN_OK;
}

int tlsan_launch_apply(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                       const float* g_i, const float* g_b, const float* g_u, const float* dgrad, float lr,
                       float reg, float clip, bool have_tsq, float* stats, const tlsan_opt_t* optd, cudaStream_t st) {
  const OptArgs opt = opt_args(optd, p.emb);
  float* tsq = reinterpret_cast<float*>(ws + w.tsq);
  const int ntsq = tsq_grid();
  if (!have_tsq) {
    int rc = tlsan_launch_table_sumsq(d, p, w, ws, st);
    if (rc) return rc;
  }
  const float invB = 1.0f / (float)(d.B_global > 0 ? d.B_global : d.B);
  tlsan_launch_kl(1, k_finalize2, dim3(1), dim3(1024), 0, st, dgrad, (const float*)tsq, ntsq, (const float*)nullptr, 0, invB,
                 lr, reg, clip, p.dense, stats, opt);
  TLSAN_CHECK_LAUNCH("k_finalize2");
  const long long n4 = (long long)d.NI * 8 + (long long)d.NU * 8 + (long long)d.NU * d.L + d.NI;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)tlsan_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  tlsan_launch_kl(1, k_apply_rows, dim3((unsigned)blocks), dim3(256), 0, st, d.NI, d.NC, d.NU, d.L, w.PU, p.emb, p.usert,
                 p.item_b, g_i, g_b, g_u, lr, reg, (const float*)stats, opt);
  TLSAN_CHECK_LAUNCH("k_apply_rows");
  tlsan_launch_kl(1, k_apply_cate, dim3(d.NC), dim3(d.NI / d.NC > 256 ? 1024 : 256), 0, st, d.NI, p.emb, g_i, (const int*)p.cate_off,
                 (const int*)p.cate_items, lr, reg, (const float*)stats, opt);
  TLSAN_CHECK_LAUNCH("k_apply_cate");
  return TLSAN_OK;
}

// ------------------------------------------------------------------ data-parallel exchange over NVLink peer memory
// One fused reduce-scatter + optimiser step + all-gather in place of NCCL all-reduce + apply.
// Every rank lets tlsan_step_grads write its flat gradient buffer INTO an IPC-shared arena; after folding the
// category gradient into the (otherwise unused) item half of the category rows it raises flag 1.  Rank r then
// sums, for slice r of the WEIGHT index space (emb | usert | item_b), the matching gradient words of all ranks
// straight out of peer memory (fixed rank order: deterministic and identical everywhere), applies L2 + clip + SGD
// to that slice of its own weights, publishes the new values in its arena and raises flag 2; finally every rank
// pulls the other slices.  The 4 449 small parameters and the statistics are reduced redundantly by every rank
// (18 KB per peer), so the clip scale is known before the slices are touched.
// arena = [flat: flat_count floats | wnew: chunk floats | 64 ints: flag1, flag2, ..., err at [8], counters at [16..]]
struct DpLayout { long long off_usert, off_itemb, n_tab, chunk, flat_count; };
static DpLayout dp_layout(const tlsan_dims_t& d, int world) {
  DpLayout y;
  const long long NR = (long long)d.NI + d.NC + d.NU;
  y.off_usert = NR * 32;
  y.off_itemb = y.off_usert + ((long long)d.NU * d.L + 3) / 4 * 4;
  y.n_tab = y.off_itemb + ((long long)d.NI + 3) / 4 * 4;
  const long long q = (y.n_tab / 4 + world - 1) / world;        // float4 units per rank
  y.chunk = q * 4;
  y.flat_count = (long long)tlsan_ws_layout(d).flat_count;
  return y;
}
size_t tlsan_dp_arena_bytes_impl(const tlsan_dims_t& d, int world) {
  const DpLayout y = dp_layout(d, world);
  return (size_t)(y.flat_count + y.chunk) * 4 + 256;
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {      // peer memory: never through the (incoherent) L1
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// the last CTA of a grid to get here publishes `epoch` in `flag` (release at system scope)
__device__ __forceinline__ void dp_signal_when_grid_done(int* counter, int* flag, int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(counter, 1) == (int)gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
  }
}

struct DpPeers { float* arena[16]; };

// spin until every peer's flag `which` has reached `epoch`.  Bounded (a dead peer must not hang the GPU): after
// `timeout_ns` the sticky error word is raised and the caller's kernel SKIPS its work -- the weights of this rank
// stay as they were and stats[TLSAN_STAT_DP_ERR] tells the host.  Returns false on (any earlier) error.
__device__ __forceinline__ bool dp_wait(const DpPeers& peers, long long flag_off, int which, int world, int epoch,
                                        int* err, long long timeout_ns) {
  if (threadIdx.x < world) {
    const int* f = reinterpret_cast<const int*>(peers.arena[threadIdx.x] + flag_off) + which;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_flag(f) < epoch) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if ((long long)(t1 - t0) > timeout_ns || *reinterpret_cast<volatile int*>(err)) { atomicExch(err, 1); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return *reinterpret_cast<volatile int*>(err) == 0;
}
static long long dp_timeout_ns() {
  static long long v = 0;
  if (!v) { const char* e = getenv("TLSAN_DP_TIMEOUT_S"); const double s = e ? atof(e) : 30.0; v = (long long)((s > 0 ? s : 30.0) * 1e9); }
  return v;
}

// category gradient of category k (item-row cate halves in CSR order + the direct u_cate row) -> item half of flat
// row NI + k, i.e. where the weight-indexed exchange expects the gradient of emb row NI + k; then flag 1
__global__ void __launch_bounds__(256) k_dp_reduce_cate(int NI, float* __restrict__ g_i, const int* __restrict__ cate_off,
                                                        const int* __restrict__ cate_items, int* __restrict__ flags,
                                                        int epoch) {
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
    g_i[(size_t)(NI + k) * 64 + lane] = g;
  }
  dp_signal_when_grid_done(flags + 16, flags + 0, epoch);
}

// dense gradients + loss / norm partials summed over ranks (fixed order) into the partial-row layout k_finalize2 reads.
// Runs EARLY, on the library's side stream right behind k_finalize1 (which produced this rank's partials) and beside the
// segmented row reduce: block 0 publishes "partials ready" (flag 2), every block waits for the peers' flag 2, then one
// element per thread with every peer's word in flight together (small CTAs: they must find room beside the row reduce).  The clip scale is therefore known before the
// gradient rows are complete, and the two single-CTA kernels (this + k_finalize2) leave the critical path.
__global__ void __launch_bounds__(256) k_dp_dense_sum(DpPeers peers, DpLayout y, long long f_dgrad, int rank, int world,
                                                       int epoch, float* __restrict__ dtot, int* __restrict__ err,
                                                       long long timeout_ns, float* __restrict__ stats) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int* mine = reinterpret_cast<int*>(peers.arena[rank] + y.flat_count + y.chunk) + 2;
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(mine), "r"(epoch) : "memory");
  }
  if (!dp_wait(peers, y.flat_count + y.chunk, 2, world, epoch, err, timeout_ns)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[TLSAN_STAT_DP_ERR] = 1.f;
    return;
  }
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < TLSAN_PART && (e < TLSAN_DENSE_COUNT || e == TLSAN_PART_LOSS || e == TLSAN_PART_SUMSQ)) {
    float v[16];
#pragma unroll
    for (int p = 0; p < 16; ++p)
      if (p < world) v[p] = ld_peer1(peers.arena[p] + f_dgrad + e);
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p)
      if (p < world) s += v[p];
    dtot[e] = s;
  }
}

// gradient words (one float4 of the weight index space) of one rank, read from its flat buffer in peer memory
__device__ __forceinline__ float4 dp_grad4(const float* __restrict__ flat, long long e, int NIC, int NU, int L, int PU,
                                           long long f_gb, long long f_gu, const DpLayout& y) {
  if (e < (long long)NIC * 32) return ld_peer4(flat + (e >> 5) * 64 + (e & 31));            // item and category rows
  if (e < y.off_usert) {                                                                       // user rows
    const long long x = e - (long long)NIC * 32;
    return ld_peer4(flat + f_gu + (x >> 5) * PU + (x & 31));
  }
  if (e < y.off_itemb) {                                                                       // usert_emb [NU][L]
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long x = e - y.off_usert + i;
      const long long u = x / L;
      v[i] = u < NU ? ld_peer1(flat + f_gu + u * PU + 32 + (x - u * L)) : 0.f;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
  }
  return ld_peer4(flat + f_gb + (e - y.off_itemb));                                            // item_b
}

// slice `rank` of the table part of wflat: sum over ranks, update, publish; then flag 2.
// W = compiled rank count (2, 4, 8, 16 >= world): U = 8 / W elements per thread and sweep, i.e. always ~8 peer reads
// (float4) in flight per thread before the first use -- a peer read is a 1-2 us round trip, and the volatile loads are
// not pipelined across loop iterations by the compiler (one element per sweep: 35 us for the whole table on one GPU).
template <int W>
__global__ void __launch_bounds__(256) k_dp_apply_slice(DpPeers peers, DpLayout y, int rank, int world, int NIC, int NU,
                                                        int L, int PU, long long f_gb, long long f_gu,
                                                        float* __restrict__ wflat, float lr, float reg,
                                                        float* __restrict__ stats, int epoch,
                                                        int* __restrict__ err, long long timeout_ns) {
  constexpr int U = W >= 8 ? 1 : 8 / W;
  // every peer's gradient rows complete (flag 1 of k_dp_reduce_cate)?  A timed-out wait leaves the weights untouched.
  if (!dp_wait(peers, y.flat_count + y.chunk, 0, world, epoch, err, timeout_ns)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[TLSAN_STAT_DP_ERR] = 1.f;
    return;
  }
  const float scale = stats[TLSAN_STAT_SCALE];
  const long long lo = (long long)rank * y.chunk, hi = min(lo + y.chunk, y.n_tab);
  float* wnew = peers.arena[rank] + y.flat_count;
  const long long nthr4 = (long long)gridDim.x * blockDim.x * 4;
  for (long long e0 = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e0 < hi; e0 += U * nthr4) {
    float4 v[U][W], w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * nthr4;
      if (e < hi) {
#pragma unroll
        for (int p = 0; p < W; ++p)
          if (p < world) v[u][p] = dp_grad4(peers.arena[p], e, NIC, NU, L, PU, f_gb, f_gu, y);
        w[u] = *reinterpret_cast<const float4*>(wflat + e);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long e = e0 + u * nthr4;
      if (e < hi) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);         // fixed rank order: identical on every rank
#pragma unroll
        for (int p = 0; p < W; ++p)
          if (p < world) { g.x += v[u][p].x; g.y += v[u][p].y; g.z += v[u][p].z; g.w += v[u][p].w; }
        const float r = e < y.off_itemb ? reg : 0.f;         // item_b carries no L2 term (model.py:164-169)
        float4 x = w[u];
        x.x -= lr * ((g.x + r * x.x) * scale); x.y -= lr * ((g.y + r * x.y) * scale);
        x.z -= lr * ((g.z + r * x.z) * scale); x.w -= lr * ((g.w + r * x.w) * scale);
        *reinterpret_cast<float4*>(wflat + e) = x;
        *reinterpret_cast<float4*>(wnew + (e - lo)) = x;
      }
    }
  }
  int* flags = reinterpret_cast<int*>(peers.arena[rank] + y.flat_count + y.chunk);
  dp_signal_when_grid_done(flags + 17, flags + 1, epoch);
}

// the other ranks' updated slices -> local weights
__global__ void __launch_bounds__(256) k_dp_gather(DpPeers peers, DpLayout y, int rank, int world, int epoch,
                                                   float* __restrict__ wflat, int* __restrict__ err,
                                                   long long timeout_ns, float* __restrict__ stats) {
  if (!dp_wait(peers, y.flat_count + y.chunk, 1, world, epoch, err, timeout_ns)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[TLSAN_STAT_DP_ERR] = 1.f;
    return;
  }
  // one pass over the (world - 1) foreign slices, 4 float4 loads in flight per thread
  const long long c4 = y.chunk / 4, total = c4 * (world - 1);
  const long long nthr = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 4 * nthr) {
    float4 v[4];
    long long dst[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long j = i + q * nthr;
      dst[q] = -1;
      if (j < total) {
        const int pi = (int)(j / c4);
        const int p = pi < rank ? pi : pi + 1;
        const long long off = (j - (long long)pi * c4) * 4, e = (long long)p * y.chunk + off;
        if (e < y.n_tab) { v[q] = ld_peer4(peers.arena[p] + y.flat_count + off); dst[q] = e; }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (dst[q] >= 0) *reinterpret_cast<float4*>(wflat + dst[q]) = v[q];
  }
}

// `arenas[rank]` must be the buffer tlsan_step_grads wrote its flat gradients to
// `side` / `side_done` (optional): the library's side stream -- ordered behind k_finalize1 of this step, NOT behind the
// segmented row reduce -- and an event to record on it
int tlsan_launch_dp_exchange(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                             float* const* arenas, int rank, int world, int epoch, float lr, float reg, float clip,
                             float* stats, cudaStream_t side, cudaEvent_t side_done, cudaStream_t st) {
  const DpLayout y = dp_layout(d, world);
  DpPeers peers;
  for (int i = 0; i < 16; ++i) peers.arena[i] = i < world ? arenas[i] : nullptr;
  float* flat = arenas[rank];
  int* flags = reinterpret_cast<int*>(flat + y.flat_count + y.chunk);
  int* err = flags + 8;
  // (1) small exchange: dense gradients, loss and norm partials -> clip scale, dense parameters updated
  cudaStream_t s1 = side ? side : st;
  float* dtot = reinterpret_cast<float*>(ws + w.part_a);          // the per-CTA partials are consumed by now
  k_dp_dense_sum<<<(TLSAN_PART + 255) / 256, 256, 0, s1>>>(peers, y, (long long)w.f_dgrad, rank, world, epoch, dtot,
                                                             err, dp_timeout_ns(), stats);
  TLSAN_CHECK_LAUNCH("k_dp_dense_sum");
  const float invB = 1.0f / (float)(d.B_global > 0 ? d.B_global : d.B);
  k_finalize2<<<1, 1024, 0, s1>>>(dtot, reinterpret_cast<float*>(ws + w.tsq), tsq_grid(), nullptr, 0, invB, lr, reg,
                                  clip, p.dense, stats, opt_args(nullptr, p.emb));
  TLSAN_CHECK_LAUNCH("k_finalize2");
  if (side) TLSAN_CHECK_CUDA(cudaEventRecord(side_done, side));
  // (2) gradient rows: fold the category halves, publish (flag 1)
  k_dp_reduce_cate<<<d.NC, 256, 0, st>>>(d.NI, flat + w.f_gi, p.cate_off, p.cate_items, flags, epoch);
  TLSAN_CHECK_LAUNCH("k_dp_reduce_cate");
  if (side) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, side_done, 0));
  // (3) reduce-scatter + update of this rank's slice (waits for every peer's flag 1), publish (flag 2); (4) all-gather
#define DP_APPLY(Wc)                                                                                                   \
  k_dp_apply_slice<Wc><<<tlsan_num_sms() * 2, 256, 0, st>>>(peers, y, rank, world, d.NI + d.NC, d.NU, d.L, w.PU,       \
                                                            (long long)w.f_gb, (long long)w.f_gu, p.emb, lr, reg,      \
                                                            stats, epoch, err, dp_timeout_ns())
  if (world <= 2) DP_APPLY(2); else if (world <= 4) DP_APPLY(4); else if (world <= 8) DP_APPLY(8); else DP_APPLY(16);
#undef DP_APPLY
  TLSAN_CHECK_LAUNCH("k_dp_apply_slice");
  k_dp_gather<<<tlsan_num_sms() * 2, 256, 0, st>>>(peers, y, rank, world, epoch, p.emb, err, dp_timeout_ns(), stats);
  TLSAN_CHECK_LAUNCH("k_dp_gather");
  return TLSAN_OK;
}

// Replicated part of the row-sharded configuration (tlsan_shard.cu): the item slots of the compact table are
// scratch, so the table kernels run with an item count of 0 on the tail of `p.emb` (cate rows first).
int tlsan_launch_apply_replicated(const tlsan_dims_t& d, const tlsan_params_t& p, const TlsanWs& w, char* ws,
                                  const float* gcate, const float* g_u, const float* dgrad, const float* item_sumsq,
                                  int n_item_sumsq, float lr, float reg, float clip, float* stats, cudaStream_t st) {
  (void)gcate;
  float* tsq = reinterpret_cast<float*>(ws + w.tsq);
  int ntsq = tlsan_num_sms() * 2;
  if (ntsq > TLSAN_MAX_GRID) ntsq = TLSAN_MAX_GRID;
  float* tail = p.emb + (size_t)d.NI * 32;          // cate rows, then user rows
  k_table_sumsq<<<ntsq, 256, 0, st>>>(tail, p.usert, 0, (long long)d.NC * 32, (long long)d.NU * 32,
                                      (long long)d.NU * d.L, tsq);
  TLSAN_CHECK_LAUNCH("k_table_sumsq");
  const float invB = 1.0f / (float)(d.B_global > 0 ? d.B_global : d.B);
  k_finalize2<<<1, 1024, 0, st>>>(dgrad, tsq, ntsq, item_sumsq, n_item_sumsq, invB, lr, reg, clip, p.dense, stats,
                                  opt_args(nullptr, p.emb));
  TLSAN_CHECK_LAUNCH("k_finalize2");
  const long long n4 = (long long)d.NU * 8 + (long long)d.NU * d.L;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)tlsan_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  k_apply_rows<<<(unsigned)blocks, 256, 0, st>>>(0, d.NC, d.NU, d.L, w.PU, tail, p.usert, nullptr, nullptr, nullptr,
                                                 g_u, lr, reg, stats, opt_args(nullptr, p.emb));
  TLSAN_CHECK_LAUNCH("k_apply_rows");
  return TLSAN_OK;
}

int tlsan_launch_label_rank(const tlsan_dims_t& d, const tlsan_params_t& p, const float* ut, const int32_t* label,
                            int32_t* rank, cudaStream_t st) {
  k_label_rank<<<d.B, 256, 0, st>>>(d.NI, p.emb, p.item_b, p.icl, ut, label, rank);
  TLSAN_CHECK_LAUNCH("k_label_rank");
  return TLSAN_OK;
}
