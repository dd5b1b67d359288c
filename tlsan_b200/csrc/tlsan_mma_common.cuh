// Device-side building blocks shared by the tensor-core TLSAN kernels (internal header):
// 3xTF32 mma tiles, the feature-wise-attention forward / backward of one 16-row tile, the
// online softmax, lane geometry and token gathers.  See tlsan_fused_mma.cu for the layout story.
#pragma once
#include "tlsan_fused.cuh"

#define MMA_THREADS 256
#define MMA_WARPS 8

struct BMat { uint32_t h0, h1, l0, l1; };          // B fragment (b0, b1) split into tf32 hi / lo
struct FwaW { BMat W1, W2; float b1[2], b2[2]; };  // forward weights of one FWA
struct FwaWT { BMat W2T, W1T; };                   // transposed fragments for the backward

// fp32 -> tf32 by truncation (one LOP3).  `cvt.rna.tf32.f32` is emulated with ~5 integer
// instructions on sm_100a (ncu: it was a quarter of the backward tile); with the hi/lo split
// truncation loses nothing: lo = x - hi is exact and its own truncation error is ~2^-21 relative.
__device__ __forceinline__ uint32_t to_tf32(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ BMat make_b(float b0, float b1) {
  BMat m;
  m.h0 = to_tf32(b0); m.h1 = to_tf32(b1);
  m.l0 = to_tf32(b0 - __uint_as_float(m.h0)); m.l1 = to_tf32(b1 - __uint_as_float(m.h1));
  return m;
}
// out[n] = sum_f in[f] W[f][n]  : B[slot t] = W[2t][g], B[slot t+4] = W[2t+1][g]
__device__ __forceinline__ BMat load_b(const float* __restrict__ W, int g, int t) {
  return make_b(W[(2 * t) * 8 + g], W[(2 * t + 1) * 8 + g]);
}
// out[n] = sum_f in[f] W[n][f]  (transposed product of the backward)
__device__ __forceinline__ BMat load_bt(const float* __restrict__ W, int g, int t) {
  return make_b(W[g * 8 + 2 * t], W[g * 8 + 2 * t + 1]);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// d += x B with 3xTF32.  x in D order {A:2t, A:2t+1, B:2t, B:2t+1}; as an A fragment:
// a0 = (row g, slot t) = x[0], a1 = (row g+8, slot t) = x[2], a2 = (row g, slot t+4) = x[1], a3 = x[3].
__device__ __forceinline__ void mma3(float (&d)[4], const float (&x)[4], const BMat& B) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = to_tf32(x[i]);
    l[i] = __float_as_uint(x[i] - __uint_as_float(h[i]));   // low 13 bits are dropped by the tensor core: no LOP
  }
  // the two small terms chain on one accumulator, the big term runs beside them; packed final add
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  mma_tf32(c, l[0], l[2], l[1], l[3], B.h0, B.h1);
  mma_tf32(d, h[0], h[2], h[1], h[3], B.h0, B.h1);
  mma_tf32(c, h[0], h[2], h[1], h[3], B.l0, B.l1);
  const float2 s0 = __fadd2_rn(make_float2(d[0], d[1]), make_float2(c[0], c[1]));
  const float2 s1 = __fadd2_rn(make_float2(d[2], d[3]), make_float2(c[2], c[3]));
  d[0] = s0.x; d[1] = s0.y; d[2] = s1.x; d[3] = s1.y;
}
// exp(x) for x <= 0 (softmax weights): one FMUL + MUFU.EX2, no denormal fix-up code
__device__ __forceinline__ float exp_neg(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

__device__ __forceinline__ FwaW load_fwa(const float* __restrict__ dense, int base, int g, int t) {
  FwaW w;
  w.W1 = load_b(dense + base, g, t);
  w.W2 = load_b(dense + base + 72, g, t);
  w.b1[0] = dense[base + 64 + 2 * t]; w.b1[1] = dense[base + 64 + 2 * t + 1];
  w.b2[0] = dense[base + 136 + 2 * t]; w.b2[1] = dense[base + 136 + 2 * t + 1];
  return w;
}
__device__ __forceinline__ FwaWT load_fwa_t(const float* __restrict__ dense, int base, int g, int t) {
  FwaWT w;
  w.W2T = load_bt(dense + base + 72, g, t);
  w.W1T = load_bt(dense + base, g, t);
  return w;
}

// m1 = relu(x W1 + b1), m2 = m1 W2 + b2   (model.py:380-383)
__device__ __forceinline__ void tile_maps(const float (&x)[4], const FwaW& w, float (&m1)[4], float (&m2)[4]) {
  m1[0] = w.b1[0]; m1[1] = w.b1[1]; m1[2] = w.b1[0]; m1[3] = w.b1[1];
  mma3(m1, x, w.W1);
#pragma unroll
  for (int i = 0; i < 4; ++i) m1[i] = fmaxf(m1[i], 0.f);
  m2[0] = w.b2[0]; m2[1] = w.b2[1]; m2[2] = w.b2[0]; m2[3] = w.b2[1];
  mma3(m2, m1, w.W2);
}

// online softmax over the sequence axis for the lane's two features
struct Soft2 {
  float mx[2], den[2], acc[2];
  __device__ __forceinline__ void init() {
    mx[0] = mx[1] = -INFINITY; den[0] = den[1] = 0.f; acc[0] = acc[1] = 0.f;
  }
  __device__ __forceinline__ void push(float m0, float m1, float x0, float x1) {
    const float mm[2] = {m0, m1}, xx[2] = {x0, x1};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float d = mm[j] - mx[j];
      const float e = exp_neg(-fabsf(d));
      const bool up = d > 0.f;
      const float c = up ? e : 1.f, n = up ? 1.f : e;
      den[j] = fmaf(den[j], c, n);
      acc[j] = fmaf(acc[j], c, n * xx[j]);
      mx[j] = up ? mm[j] : mx[j];
    }
  }
};

// per-lane gradient accumulators of one FWA weight set: rows k = 0..7, the lane's 2 columns
// W?p[tq][jj] = (dW[2tq][col jj], dW[2tq+1][col jj]): row pairs packed for FFMA2 (fma.rn.f32x2)
struct FwaGrad {
  float2 W1p[4][2], W2p[4][2];
  float b1[2], b2[2];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) { W1p[q][jj] = make_float2(0.f, 0.f); W2p[q][jj] = make_float2(0.f, 0.f); }
    b1[0] = b1[1] = b2[0] = b2[1] = 0.f;
  }
  __device__ __forceinline__ float w1(int k, int jj) const { return (k & 1) ? W1p[k >> 1][jj].y : W1p[k >> 1][jj].x; }
  __device__ __forceinline__ float w2(int k, int jj) const { return (k & 1) ? W2p[k >> 1][jj].y : W2p[k >> 1][jj].x; }
};

// backward of one tile.  okB = second token of the tile is real.
__device__ __forceinline__ void tile_bwd(const float (&x)[4], bool okB, const float (&o)[2], const float (&dout)[2],
                                         const float (&mx)[2], const float (&inv)[2], const FwaW& w,
                                         const FwaWT& wt, int lane, float (&dx)[4], FwaGrad& G) {
  float m1[4], m2[4];
  tile_maps(x, w, m1, m2);
  // softmax weight times d out: a * do, with (1/den) * do folded into one per-feature factor
  const float kf[2] = {inv[0] * dout[0], inv[1] * dout[1]};
  // exp(m2 - mx) = ex2(m2 * log2e - mx * log2e): the second product is per sample (hoisted out of the tile loop)
  const float nmx[2] = {-mx[0] * 1.4426950408889634f, -mx[1] * 1.4426950408889634f};
  float ado[4], dm2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = i & 1;
    float ee;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ee) : "f"(fmaf(m2[i], 1.4426950408889634f, nmx[j])));
    const float aw = ee * kf[j];
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
  // dW2[k][j] += m1[k] dm2[j], dW1[k][j] += x[k] dpre[j]: the 8 k-values of a row live in the
  // 4 lanes of the row's quad (2 each) -> quad shuffles give the pair (k = 2tq, 2tq+1), then one
  // FFMA2 per (pair, column) on the lane's 2 columns.
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

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// lane geometry: features f0, f0+1 of the 64-float token; which table half; offset inside the row
struct LaneGeo {
  int lane, g, t, f0, half, col;
  __device__ __forceinline__ void init() {
    lane = threadIdx.x & 31; g = lane >> 2; t = lane & 3;
    f0 = 8 * g + 2 * t; half = g >> 2; col = (g & 3) * 8 + 2 * t;
  }
};

// float32(1/n), n = 0..12, exactly as the reference computes it: float64 1/n (build_dataset.py:20) rounded to
// float32 by the store into the batch array (input.py:36,45); n = 0 (d < 2, unreachable) -> 0
static __constant__ float c_bucket_w[13] = {0.f, (float)(1.0 / 1.0), (float)(1.0 / 2.0), (float)(1.0 / 3.0),
                                            (float)(1.0 / 4.0), (float)(1.0 / 5.0), (float)(1.0 / 6.0),
                                            (float)(1.0 / 7.0), (float)(1.0 / 8.0), (float)(1.0 / 9.0),
                                            (float)(1.0 / 10.0), (float)(1.0 / 11.0), (float)(1.0 / 12.0)};
// n = sum_j [d >= 2^j], j = 1..12  (build_dataset.py:16-21)  =  min(12, floor(log2 d)) for d >= 2
__device__ __forceinline__ float bucket_weight(int d) {
  const int n = d >= 2 ? min(12, 31 - __clz(d)) : 0;
  return c_bucket_w[n];
}

// ---- token meta of up to 32 tokens, one per lane (coalesced), broadcast by shuffle
struct LongMeta { int id, crow; float tau, pt, ht; };
__device__ __forceinline__ LongMeta load_long_meta(const FArgs& a, int b, int u, int tt, int ell, float gamma) {
  LongMeta m;
  const bool ok = tt < ell;
  m.id = ok ? __ldg(a.hist_i + (size_t)b * a.L + tt) : 0;
  if (a.hist_d) m.ht = ok ? bucket_weight(__ldg(a.hist_d + (size_t)b * a.L + tt)) : 0.f;   // bucketing fused into the gather
  else m.ht = ok ? __ldg(a.hist_t + (size_t)b * a.L + tt) : 0.f;
  const float pu = ok ? __ldg(a.usert + (size_t)u * a.L + tt) : 0.f;
  m.pt = pu * m.ht;                  // P[u,t] * hist_t    (model.py:99)
  m.tau = gamma * m.pt;              // gamma * (...)      (model.py:109)
  m.crow = a.NI + __ldg(a.icl + m.id);
  return m;
}
__device__ __forceinline__ const float* row_ptr(const FArgs& a, const LaneGeo& L, int id, int crow) {
  return a.emb + (size_t)(L.half ? crow : id) * 32 + L.col;
}

// one tile's inputs: two gathered token slices (float2 each) and their tau
struct Pair { float2 eA, eB; float tA, tB; bool okB; };
__device__ __forceinline__ Pair fetch_pair(const FArgs& a, const LaneGeo& L, const LongMeta& me, int j, int cnt) {
  Pair p;
  p.okB = j + 1 < cnt;
  const int idA = __shfl_sync(0xffffffffu, me.id, j), crA = __shfl_sync(0xffffffffu, me.crow, j);
  const int idB = __shfl_sync(0xffffffffu, me.id, (j + 1) & 31), crB = __shfl_sync(0xffffffffu, me.crow, (j + 1) & 31);
  p.tA = __shfl_sync(0xffffffffu, me.tau, j);
  p.tB = __shfl_sync(0xffffffffu, me.tau, (j + 1) & 31);
  p.eA = ldg2(row_ptr(a, L, idA, crA));
  p.eB = p.okB ? ldg2(row_ptr(a, L, idB, crB)) : make_float2(0.f, 0.f);
  return p;
}


// ---- per-round row staging: ALL token rows of a (<= 32 token) round are copied global -> shared
// with cp.async right after the round's ids are known (16 lanes x 16 B per token, two tokens per
// LDGSTS), then one wait: the tiles read their slices with conflict-free LDS.64.  Without it the
// first use of every tile's gather stalls on an L2 round trip (16-17 % of all samples in ncu).
__device__ __forceinline__ void cp16_async(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void stage_round_rows(const FArgs& a, const LongMeta& me, int cnt, int lane,
                                                 float (*buf)[64]) {
  const int hs = lane >> 4, c = lane & 15;
  __syncwarp();                                   // every lane is done with the previous round's rows
  for (int k = 0; k < cnt; k += 2) {
    const int t = k + hs;
    const int idt = __shfl_sync(0xffffffffu, me.id, t & 31), crt = __shfl_sync(0xffffffffu, me.crow, t & 31);
    if (t < cnt) cp16_async(&buf[t][c * 4], a.emb + (size_t)(c < 8 ? idt : crt) * 32 + (c & 7) * 4);
  }
  cp_async_wait_all();
  __syncwarp();
}
