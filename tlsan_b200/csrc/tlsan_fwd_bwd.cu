// Fused TLSAN forward / backward kernels for sm_100a.
//
// Mapping (all fused kernels): one thread = one (sample, head) pair; the 8 heads of a sample
// sit in 8 adjacent lanes, a warp carries 4 samples, a 256-thread CTA a tile of 32 samples,
// and a persistent grid strides over tiles.  A head is an 8-feature slice (the 8x8 maps are
// shared by all heads, reference model.py:374,447), so the feature-wise softmax over the
// sequence axis (model.py:386) runs as an online softmax entirely in registers, one token
// at a time, with no cross-thread traffic.  The 8x8 weights are FFMA constant-bank operands.
//
//   k_score     forward only (Model.eval_auc, model.py:237-263), 1 or 2 candidates per row
//   k_fwd_bwd_a forward + loss + backward of logit / short-term FWA / dense (model.py:84-137,
//               164-172, 198) ; leaves d(o_long) and the long softmax statistics in scratch
//   k_bwd_b     backward of the long-term FWA and of the time-aware position term
//   k_gather    K1 standalone (model.py:84-86,105-113) ; k_bucket  K2 (build_dataset.py:16-21)
#include "tlsan_common.cuh"

__constant__ float c_small[292];  // dense[0..288) (both FWA weight sets) ; [288] = gamma

#include "tlsan_fused.cuh"

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8_plain(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 8 lanes (heads) of one sample.  Only those 8 lanes are named in the mask: the
// call sites sit inside per-sample loops whose trip count differs between the 4 samples of a warp.
__device__ __forceinline__ float oct_sum(float v) {
  const unsigned m = 0xffu << (threadIdx.x & 24);
  v += __shfl_xor_sync(m, v, 1);
  v += __shfl_xor_sync(m, v, 2);
  v += __shfl_xor_sync(m, v, 4);
  return v;
}

// m1 = relu(x W1 + b1), m2 = m1 W2 + b2      (model.py:380-383 via :397-454)
template <int BASE>
__device__ __forceinline__ void fwa_maps(const float (&x)[8], float (&m1)[8], float (&m2)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a = fmaf(x[k], c_small[BASE + k * 8 + j], a);
    m1[j] = fmaxf(a + c_small[BASE + 64 + j], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a = fmaf(m1[k], c_small[BASE + 72 + k * 8 + j], a);
    m2[j] = a + c_small[BASE + 136 + j];
  }
}

// online softmax over the sequence axis, independently per feature (model.py:386-387).
// One exp per feature and token: exactly one of the two rescale factors is 1.
struct Soft {
  float mx[8], den[8], acc[8];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mx[j] = -INFINITY; den[j] = 0.f; acc[j] = 0.f; }
  }
  __device__ __forceinline__ void push(const float (&m2)[8], const float (&x)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = m2[j] - mx[j];
      const float t = __expf(-fabsf(d));
      const bool up = d > 0.f;
      const float c = up ? t : 1.f;   // rescale of the running sums
      const float e = up ? 1.f : t;   // weight of the new token
      den[j] = fmaf(den[j], c, e);
      acc[j] = fmaf(acc[j], c, e * x[j]);
      mx[j] = up ? m2[j] : mx[j];
    }
  }
};

// backward of one FWA token: given x, o, do, softmax stats -> dx, weight grads
template <int BASE>
__device__ __forceinline__ void fwa_bwd_token(const float (&x)[8], const float (&o)[8], const float (&dout)[8],
                                              const float (&mx)[8], const float (&inv)[8], float (&dx)[8],
                                              float (&gW1)[64], float (&gb1)[8], float (&gW2)[64],
                                              float (&gb2)[8]) {
  float m1[8], m2[8];
  fwa_maps<BASE>(x, m1, m2);
  float dm2[8], ado[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = __expf(m2[j] - mx[j]) * inv[j];
    ado[j] = a * dout[j];
    dm2[j] = ado[j] * (x[j] - o[j]);
    gb2[j] += dm2[j];
  }
  float dpre[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(dm2[j], c_small[BASE + 72 + k * 8 + j], s);
    dpre[k] = m1[k] > 0.f ? s : 0.f;
    gb1[k] += dpre[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = ado[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(dpre[j], c_small[BASE + k * 8 + j], s);
    dx[k] = s;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gW2[k * 8 + j] = fmaf(m1[k], dm2[j], gW2[k * 8 + j]);
      gW1[k * 8 + j] = fmaf(x[k], dpre[j], gW1[k * 8 + j]);
    }
  }
}

// shared-memory image of the 64x64 dense layer, column-permuted so that head h's two
// float4 (features 8h..8h+3 and 8h+4..8h+7) sit at [k][4h] and [k][32+4h]: conflict-free.
__device__ __forceinline__ int perm_col(int f) { return ((f & 7) >> 2) * 32 + (f >> 3) * 4 + (f & 3); }

struct SmemDense {
  float wd[64 * 64];   // wd[k][perm(f)]  = Wd[k][f]
  float wdt[64 * 64];  // wdt[j][perm(f)] = Wd[f][j]
  float bd[64];
  float o[TLSAN_TILE * 64];   // o_long of the tile
  float dz[TLSAN_TILE * 64];  // d loss / d z of the tile
  float red[8 * 160];         // end-of-kernel reduction staging
};

__device__ __forceinline__ void load_dense_smem(SmemDense& sm, const float* __restrict__ dense, bool need_t) {
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int k = e >> 6, f = e & 63;
    const float w = dense[TLSAN_OFF_WD + e];
    sm.wd[k * 64 + perm_col(f)] = w;
    if (need_t) sm.wdt[f * 64 + perm_col(k)] = w;
  }
  if (threadIdx.x < 64) sm.bd[threadIdx.x] = dense[TLSAN_OFF_BD + threadIdx.x];
}

// out[8h+q] = sum_k vec[k] * W[k][perm(8h+q)]   (vec broadcast from smem, W from smem)
__device__ __forceinline__ void dense_apply(const float* __restrict__ vec, const float* __restrict__ w, int h,
                                            float (&out)[8]) {
#pragma unroll 4
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 v4 = *reinterpret_cast<const float4*>(vec + 4 * k4);
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * k4 + kk;
      const float4 w0 = *reinterpret_cast<const float4*>(w + k * 64 + 4 * h);
      const float4 w1 = *reinterpret_cast<const float4*>(w + k * 64 + 32 + 4 * h);
      out[0] = fmaf(vv[kk], w0.x, out[0]); out[1] = fmaf(vv[kk], w0.y, out[1]);
      out[2] = fmaf(vv[kk], w0.z, out[2]); out[3] = fmaf(vv[kk], w0.w, out[3]);
      out[4] = fmaf(vv[kk], w1.x, out[4]); out[5] = fmaf(vv[kk], w1.y, out[5]);
      out[6] = fmaf(vv[kk], w1.z, out[6]); out[7] = fmaf(vv[kk], w1.w, out[7]);
    }
  }
}

// row of the unified table holding this head's slice of e(item): heads 0-3 read item_emb[id],
// heads 4-7 read cate_emb[icl[id]]  (model.py:84-86)
__device__ __forceinline__ const float* tok_ptr(const FArgs& a, int id, int half, int sub) {
  const int row = half ? a.NI + __ldg(a.icl + id) : id;
  return a.emb + (size_t)row * 32 + sub * 8;
}

// ------------------------------------------------------------------ fused forward (+ backward A)
template <bool TRAIN>
__global__ void __launch_bounds__(TLSAN_THREADS, 1) k_fused(const FArgs a, const int ncand) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemDense& sm = *reinterpret_cast<SmemDense*>(smem_raw);
  load_dense_smem(sm, a.dense, TRAIN);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane & 7, half = h >> 2, sub = h & 3;
  const int sidx = warp * 4 + (lane >> 3);
  const float gamma = c_small[288];
  const int ntiles = (a.B + TLSAN_TILE - 1) / TLSAN_TILE;

  // persistent accumulators (TRAIN only)
  float gW1[64], gW2[64], gb1[8], gb2[8], gWd[16], gbd[8];
  float loss_acc = 0.f, sq_acc = 0.f;
  if (TRAIN) {
#pragma unroll
    for (int e = 0; e < 64; ++e) { gW1[e] = 0.f; gW2[e] = 0.f; }
#pragma unroll
    for (int e = 0; e < 8; ++e) { gb1[e] = 0.f; gb2[e] = 0.f; gbd[e] = 0.f; }
#pragma unroll
    for (int e = 0; e < 16; ++e) gWd[e] = 0.f;
  }

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * TLSAN_TILE + sidx;
    const bool valid = b < a.B;
    const int bb = valid ? b : 0;
    const int u = __ldg(a.u + bb);
    const int ell = valid ? __ldg(a.sl + bb) : 0;
    const int s = valid ? __ldg(a.sl_new + bb) : -1;  // short tokens 0..s (token 0 = z)

    // ---- long-term FWA forward (model.py:98-109, 334-345)
    Soft st;
    st.init();
    for (int t = 0; t < ell; ++t) {
      const int id = __ldg(a.hist_i + (size_t)bb * a.L + t);
      float e[8], x[8], m1[8], m2[8];
      ld8(tok_ptr(a, id, half, sub), e);
      const float tau = gamma * (__ldg(a.usert + (size_t)u * a.L + t) * __ldg(a.hist_t + (size_t)bb * a.L + t));
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = e[q] * tau;
      fwa_maps<0>(x, m1, m2);
      st.push(m2, x);
    }
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = ell > 0 ? st.acc[q] / st.den[q] : 0.f;
    st8(sm.o + sidx * 64 + 8 * h, o);
    if (TRAIN && valid) {
      float inv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) inv[q] = 1.f / st.den[q];
      float* sc = a.scratch + (size_t)b * (TLSAN_SCR * 64) + 8 * h;
      st8(sc + 64, o);
      st8(sc + 128, st.mx);
      st8(sc + 192, inv);
    }
    __syncwarp();

    // ---- dense: z = o_long Wd + bd  (model.py:347), token 0 of the short sequence (:350)
    float z[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) z[q] = 0.f;
    dense_apply(sm.o + sidx * 64, sm.wd, h, z);
#pragma unroll
    for (int q = 0; q < 8; ++q) z[q] += sm.bd[8 * h + q];

    // ---- short-term FWA forward over [z ; e(hist_i_new)]  (model.py:350-364)
    st.init();
    for (int t = 0; t <= s; ++t) {
      float x[8], m1[8], m2[8];
      if (t == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = z[q];
      } else {
        const int id = __ldg(a.hist_i_new + (size_t)bb * a.S + (t - 1));
        ld8(tok_ptr(a, id, half, sub), x);
      }
      fwa_maps<144>(x, m1, m2);
      st.push(m2, x);
    }
    float v[8], inv_s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      inv_s[q] = s >= 0 ? 1.f / st.den[q] : 0.f;
      v[q] = st.acc[q] * inv_s[q];
    }

    // ---- user vector, candidate, logit  (model.py:84-95,135-137)
    float p[8], ut[8];
    {
      const int urow = half ? a.NI + __ldg(a.c + bb) : a.NI + a.NC + u;
      ld8(a.emb + (size_t)urow * 32 + sub * 8, p);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) ut[q] = v[q] + p[q];
    const int cand = __ldg(a.i + bb);
    float qv[8];
    ld8(tok_ptr(a, cand, half, sub), qv);
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) dot = fmaf(ut[q], qv[q], dot);
    const float logit = oct_sum(dot) + __ldg(a.item_b + cand);

    if (!TRAIN) {
      if (valid) {
        if (h == 0) a.logits[(size_t)b * ncand] = logit;
        if (a.ut) st8(a.ut + (size_t)b * 64 + 8 * h, ut);
      }
      if (ncand > 1) {  // Model.eval_auc second run, model.py:251-261: same u_t, other item
        const int cand2 = __ldg(a.i2 + bb);
        float q2[8];
        ld8(tok_ptr(a, cand2, half, sub), q2);
        float d2 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) d2 = fmaf(ut[q], q2[q], d2);
        const float l2 = oct_sum(d2) + __ldg(a.item_b + cand2);
        if (valid && h == 0) a.logits[(size_t)b * ncand + 1] = l2;
      }
      __syncwarp();
      continue;
    }

    // =========================== backward (TRAIN) ===========================
    float dz[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) dz[q] = 0.f;
    if (valid) {
      // sigmoid cross entropy (model.py:171) and its gradient through reduce_mean
      const float yb = __ldg(a.y + b);
      const float ex = expf(-fabsf(logit));
      const float bce = fmaxf(logit, 0.f) - logit * yb + log1pf(ex);
      const float sig = logit >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      const float g = (sig - yb) * a.invB;
      if (h == 0) {
        loss_acc += bce;
        sq_acc = fmaf(g, g, sq_acc);  // item_b gather slice
        a.gscal[b] = g;
      }
      float* rcand = grad_row(a, b, a.L + a.S) + 8 * h;
      float* rvirt = grad_row(a, b, a.L + a.S + 1) + 8 * h;
      float dq[8], du[8], zero[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        dq[q] = g * ut[q];
        du[q] = g * qv[q];
        zero[q] = 0.f;
        sq_acc = fmaf(dq[q], dq[q], sq_acc);
        sq_acc = fmaf(du[q], du[q], sq_acc);
      }
      st8(rcand, dq);                                        // -> item_emb[i] | cate_emb[icl[i]]
      if (half) st8(rvirt, du);                              // -> cate_emb[u_cate]
      else {
        st8(rvirt, zero);
        st8(a.rows_u + (size_t)b * a.PU + 8 * sub, du);      // -> user_emb[u]
      }
      // short-term FWA backward (dv = du)
      for (int t = 0; t <= s; ++t) {
        float x[8], dx[8];
        if (t == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = z[q];
        } else {
          const int id = __ldg(a.hist_i_new + (size_t)b * a.S + (t - 1));
          ld8(tok_ptr(a, id, half, sub), x);
        }
        fwa_bwd_token<144>(x, v, du, st.mx, inv_s, dx, gW1, gb1, gW2, gb2);
        if (t == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) dz[q] = dx[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) sq_acc = fmaf(dx[q], dx[q], sq_acc);
          st8(grad_row(a, b, a.L + (t - 1)) + 8 * h, dx);
        }
      }
    }
    st8(sm.dz + sidx * 64 + 8 * h, dz);
#pragma unroll
    for (int q = 0; q < 8; ++q) gbd[q] += dz[q];
    __syncthreads();
    // dense kernel gradient: gWd[k][j] += o_long[k] * dz[j] over the tile (fixed order)
    {
      const int r4 = threadIdx.x >> 4, c4 = threadIdx.x & 15;
#pragma unroll 4
      for (int ss = 0; ss < TLSAN_TILE; ++ss) {
        const float4 o4 = *reinterpret_cast<const float4*>(sm.o + ss * 64 + 4 * r4);
        const float4 z4 = *reinterpret_cast<const float4*>(sm.dz + ss * 64 + 4 * c4);
        const float oo[4] = {o4.x, o4.y, o4.z, o4.w};
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) gWd[i * 4 + j] = fmaf(oo[i], zz[j], gWd[i * 4 + j]);
      }
    }
    // d o_long = dz Wd^T
    float dol[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) dol[q] = 0.f;
    dense_apply(sm.dz + sidx * 64, sm.wdt, h, dol);
    if (valid) st8(a.scratch + (size_t)b * (TLSAN_SCR * 64) + 8 * h, dol);
    __syncthreads();
  }

  if (TRAIN) {
    // ---- per-CTA partial sums, fixed order: lanes (butterfly) -> warps 0..7 -> global
    float* part = a.part + (size_t)blockIdx.x * TLSAN_PART;
#pragma unroll
    for (int e = 0; e < 64; ++e) {
      const float r1 = warp_sum(gW1[e]);
      const float r2 = warp_sum(gW2[e]);
      if (lane == 0) { sm.red[warp * 160 + e] = r1; sm.red[warp * 160 + 72 + e] = r2; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float r1 = warp_sum(gb1[e]);
      const float r2 = warp_sum(gb2[e]);
      if (lane == 0) { sm.red[warp * 160 + 64 + e] = r1; sm.red[warp * 160 + 136 + e] = r2; }
    }
    {
      const float r1 = warp_sum(loss_acc), r2 = warp_sum(sq_acc);
      if (lane == 0) { sm.red[warp * 160 + 144] = r1; sm.red[warp * 160 + 145] = r2; }
    }
    // bd gradient: sum over the 4 samples of the warp (lanes with equal h), then warps
    __syncthreads();  // sm.o / sm.dz free again: reuse sm.o as [8 warps][64] staging for gbd
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float r = gbd[q];
      r += __shfl_xor_sync(0xffffffffu, r, 8);
      r += __shfl_xor_sync(0xffffffffu, r, 16);
      if (lane < 8) sm.o[warp * 64 + 8 * h + q] = r;
    }
    __syncthreads();
    if (threadIdx.x < 146) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) r += sm.red[w * 160 + threadIdx.x];
      const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1S + threadIdx.x
                                        : (threadIdx.x == 144 ? TLSAN_PART_LOSS : TLSAN_PART_SUMSQ);
      part[dst] = r;
    }
    if (threadIdx.x < 64) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) r += sm.o[w * 64 + threadIdx.x];
      part[TLSAN_OFF_BD + threadIdx.x] = r;
    }
    {
      const int r4 = threadIdx.x >> 4, c4 = threadIdx.x & 15;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + TLSAN_OFF_WD + (4 * r4 + i) * 64 + 4 * c4) =
            make_float4(gWd[i * 4], gWd[i * 4 + 1], gWd[i * 4 + 2], gWd[i * 4 + 3]);
    }
  }
}

// ------------------------------------------------------------------ backward B: long-term FWA
__global__ void __launch_bounds__(TLSAN_THREADS, 1) k_bwd_long(const FArgs a) {
  __shared__ float red[8 * 160];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane & 7, half = h >> 2, sub = h & 3;
  const int sidx = warp * 4 + (lane >> 3);
  const float gamma = c_small[288];
  const int ntiles = (a.B + TLSAN_TILE - 1) / TLSAN_TILE;

  float gW1[64], gW2[64], gb1[8], gb2[8];
  float ggamma = 0.f, sq_acc = 0.f;
#pragma unroll
  for (int e = 0; e < 64; ++e) { gW1[e] = 0.f; gW2[e] = 0.f; }
#pragma unroll
  for (int e = 0; e < 8; ++e) { gb1[e] = 0.f; gb2[e] = 0.f; }

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile * TLSAN_TILE + sidx;
    const bool valid = b < a.B;
    const int bb = valid ? b : 0;
    const int u = __ldg(a.u + bb);
    const int ell = valid ? __ldg(a.sl + bb) : 0;
    float dol[8], o[8], mx[8], inv[8];
    {
      const float* sc = a.scratch + (size_t)bb * (TLSAN_SCR * 64) + 8 * h;
      ld8_plain(sc, dol); ld8_plain(sc + 64, o); ld8_plain(sc + 128, mx); ld8_plain(sc + 192, inv);
    }
    float* ru = a.rows_u + (size_t)bb * a.PU + 32;
    for (int t = 0; t < ell; ++t) {
      const int id = __ldg(a.hist_i + (size_t)bb * a.L + t);
      float e[8], x[8], dx[8];
      ld8(tok_ptr(a, id, half, sub), e);
      const float ht = __ldg(a.hist_t + (size_t)bb * a.L + t);
      const float pt = __ldg(a.usert + (size_t)u * a.L + t) * ht;   // P[u,t] * hist_t  (model.py:99)
      const float tau = gamma * pt;                                  // (:109)
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = e[q] * tau;
      fwa_bwd_token<0>(x, o, dol, mx, inv, dx, gW1, gb1, gW2, gb2);
      float dtp = 0.f, row[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        dtp = fmaf(dx[q], e[q], dtp);
        row[q] = dx[q] * tau;            // gradient of the gathered embedding slice
        sq_acc = fmaf(row[q], row[q], sq_acc);
      }
      st8(grad_row(a, b, t) + 8 * h, row);
      const float dtau = oct_sum(dtp);
      if (h == 0) {
        ggamma = fmaf(dtau, pt, ggamma);
        const float dp = dtau * gamma * ht;   // d usert_emb[u,t]
        sq_acc = fmaf(dp, dp, sq_acc);
        ru[t] = dp;
      }
    }
    if (valid && h == 0)
      for (int t = ell; t < a.PU - 32; ++t) ru[t] = 0.f;
  }

  float* part = a.part + (size_t)blockIdx.x * TLSAN_PART;
#pragma unroll
  for (int e = 0; e < 64; ++e) {
    const float r1 = warp_sum(gW1[e]);
    const float r2 = warp_sum(gW2[e]);
    if (lane == 0) { red[warp * 160 + e] = r1; red[warp * 160 + 72 + e] = r2; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float r1 = warp_sum(gb1[e]);
    const float r2 = warp_sum(gb2[e]);
    if (lane == 0) { red[warp * 160 + 64 + e] = r1; red[warp * 160 + 136 + e] = r2; }
  }
  {
    const float r1 = warp_sum(ggamma), r2 = warp_sum(sq_acc);
    if (lane == 0) { red[warp * 160 + 144] = r1; red[warp * 160 + 145] = r2; }
  }
  __syncthreads();
  if (threadIdx.x < 146) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += red[w * 160 + threadIdx.x];
    const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1L + threadIdx.x
                                      : (threadIdx.x == 144 ? TLSAN_OFF_GAMMA : TLSAN_PART_SUMSQ);
    part[dst] = r;
  }
}

// ------------------------------------------------------------------ K1 / K2 standalone
// 16 lanes x float4 per 64-float output row; lanes 0-7 item_emb row, lanes 8-15 cate_emb row.
__global__ void k_gather(const float* __restrict__ emb, const int* __restrict__ icl, const int* __restrict__ idx,
                         const float* __restrict__ tau, float* __restrict__ out, long long n, int NI) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r = g >> 4;
  const int q = (int)(g & 15);
  if (r >= n) return;
  const int id = __ldg(idx + r);
  const int row = q < 8 ? id : NI + __ldg(icl + id);
  float4 v = __ldg(reinterpret_cast<const float4*>(emb + (size_t)row * 32) + (q & 7));
  if (tau) {
    const float t = __ldg(tau + r);
    v.x *= t; v.y *= t; v.z *= t; v.w *= t;
  }
  reinterpret_cast<float4*>(out + (size_t)r * 64)[q] = v;
}

// n = sum_j [d >= 2^j] (j = 1..12) = min(12, floor(log2 d)) for d >= 2 ; d < 2 -> 0
__global__ void k_bucket(const int* __restrict__ d, const float* __restrict__ lut, float* __restrict__ out,
                         int* __restrict__ bucket, long long n) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int v = d[g];
  const int nb = v >= 2 ? min(12, 31 - __clz(v)) : 0;
  if (out) out[g] = __ldg(lut + nb);
  if (bucket) bucket[g] = nb;
}

// ------------------------------------------------------------------ host launchers
static int g_num_sms = 0;
int tlsan_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms = 148;
  }
  return g_num_sms;
}

int tlsan_launch_upload_consts(const float* dense, cudaStream_t st) {
  TLSAN_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_small, dense, 288 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
  TLSAN_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_small, dense + TLSAN_OFF_GAMMA, sizeof(float), 288 * sizeof(float),
                                           cudaMemcpyDeviceToDevice, st));
  return TLSAN_OK;
}

FArgs tlsan_make_fargs(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b) {
  FArgs a;
  a.B = d.B; a.L = d.L; a.S = d.S; a.NI = d.NI; a.NC = d.NC; a.NU = d.NU;
  a.SI = d.L + d.S + 2; a.PU = (int)tlsan_align_up(32 + d.L, 4);
  a.invB = 1.0f / (float)(d.B_global > 0 ? d.B_global : d.B);
  a.emb = p.emb; a.usert = p.usert; a.item_b = p.item_b; a.dense = p.dense; a.icl = p.icl;
  a.u = b.u; a.i = b.i; a.i2 = b.i2; a.c = b.c; a.sl = b.sl; a.sl_new = b.sl_new;
  a.hist_i = b.hist_i; a.hist_i_new = b.hist_i_new; a.y = b.y; a.hist_t = b.hist_t; a.hist_d = b.hist_d;
  a.logits = nullptr; a.ut = nullptr; a.rows_i = nullptr; a.rows_u = nullptr; a.gscal = nullptr;
  a.scratch = nullptr; a.part = nullptr; a.inv = nullptr; a.spsh = 0;
  return a;
}

static int fused_grid(int B) {
  const int ntiles = (B + TLSAN_TILE - 1) / TLSAN_TILE;
  const int g = tlsan_num_sms();
  return ntiles < g ? ntiles : g;
}

int tlsan_launch_score(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                       float* logits, float* ut, cudaStream_t st) {
  int rc = tlsan_launch_upload_consts(p.dense, st);
  if (rc) return rc;
  FArgs a = tlsan_make_fargs(d, p, b);
  a.logits = logits; a.ut = ut;
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SmemDense)));
    attr_set = true;
  }
  k_fused<false><<<fused_grid(d.B), TLSAN_THREADS, sizeof(SmemDense), st>>>(a, ncand);
  TLSAN_CHECK_LAUNCH("k_fused<score>");
  return TLSAN_OK;
}

int tlsan_launch_fwd_bwd(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, const TlsanWs& w,
                         char* ws, int* grid_a, int* grid_b, cudaStream_t st) {
  int rc = tlsan_launch_upload_consts(p.dense, st);
  if (rc) return rc;
  FArgs a = tlsan_make_fargs(d, p, b);
  a.rows_i = reinterpret_cast<float*>(ws + w.rows_i);
  a.inv = reinterpret_cast<const int*>(ws + w.inv); a.spsh = w.SPSH;
  a.rows_u = reinterpret_cast<float*>(ws + w.rows_u);
  a.gscal = reinterpret_cast<float*>(ws + w.gscal);
  a.scratch = reinterpret_cast<float*>(ws + w.scratch);
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SmemDense)));
    attr_set = true;
  }
  const int g = fused_grid(d.B);
  *grid_a = g; *grid_b = g;
  a.part = reinterpret_cast<float*>(ws + w.part_a);
  k_fused<true><<<g, TLSAN_THREADS, sizeof(SmemDense), st>>>(a, 1);
  TLSAN_CHECK_LAUNCH("k_fused<train>");
  // this variant fuses long forward, dense and short into one kernel: report it all as SHORT
  tlsan_profile_mark(TLSAN_PHASE_LONG_FWD, st); tlsan_profile_mark(TLSAN_PHASE_DENSE_FWD, st);
  tlsan_profile_mark(TLSAN_PHASE_SHORT, st); tlsan_profile_mark(TLSAN_PHASE_DENSE_BWD, st);
  a.part = reinterpret_cast<float*>(ws + w.part_b);
  k_bwd_long<<<g, TLSAN_THREADS, 0, st>>>(a);
  TLSAN_CHECK_LAUNCH("k_bwd_long");
  tlsan_profile_mark(TLSAN_PHASE_BWD_LONG, st);
  return TLSAN_OK;
}

int tlsan_launch_gather(const tlsan_dims_t& d, const tlsan_params_t& p, const int32_t* idx, const float* tau,
                        float* out, int64_t n, cudaStream_t st) {
  if (n == 0) return TLSAN_OK;
  const long long threads = n * 16;
  k_gather<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p.emb, p.icl, idx, tau, out, n, d.NI);
  TLSAN_CHECK_LAUNCH("k_gather");
  return TLSAN_OK;
}

int tlsan_launch_bucket(const int32_t* dd, const float* lut, float* out, int32_t* bucket, int64_t n,
                        cudaStream_t st) {
  if (n == 0) return TLSAN_OK;
  k_bucket<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dd, lut, out, bucket, n);
  TLSAN_CHECK_LAUNCH("k_bucket");
  return TLSAN_OK;
}
