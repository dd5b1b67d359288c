// Fused TLSAN forward / backward kernels, tensor-core formulation (default path).
//
// One WARP works on one sample at a time.  A work tile is 16 (token, head) rows = 2 tokens x 8
// heads of that sample; the four 8x8 maps of the feature-wise attention (x W1, m1 W2 and the two
// transposed products of the backward) are `mma.sync.m16n8k8` TF32 tiles with the 3xTF32 hi/lo
// split (fp32-level accuracy: the 1e-4 logit tolerance does not survive plain TF32).
//
// Register layout of every per-row 8-vector (x, m1, m2, dm2, dpre, dx): lane (g = lane/4,
// t = lane%4) holds features {2t, 2t+1} of head g for token A (v[0], v[1]) and token B
// (v[2], v[3]) -- exactly the accumulator (D) fragment of the mma.  Permuting the K index of
// the next product (slot t <-> feature 2t, slot t+4 <-> feature 2t+1, applied to the rows of
// the weight fragment) makes the same registers a valid A fragment, so the chain
// x -> m1 -> m2 and dm2 -> dm1 -> dx never leaves registers and needs no shuffles.
// The per-feature softmax over the sequence (model.py:386) is an online softmax per lane.
// A warp touches one token as one 256-B coalesced float2 access (item row | cate row).
//
//   k_fwd_mma<0>     scoring, fused forward, 1 or 2 candidates (Model.eval_auc, model.py:237-263)
//   k_fwd_mma<1>     long-term FWA forward (model.py:98-109,334-345)
//   k_dense_fwd_mma  z = o_long Wd + bd as one batched GEMM (model.py:347)
//   k_fwd_mma<2>     short FWA forward + logit + loss + backward of logit / short FWA (model.py:135-137,164-172,350-364)
//   k_dense_bwd_mma  d o_long = dz Wd^T ; dWd = O^T dZ ; dbd
//   k_bwd_long_mma   backward of the long FWA and of the time-aware position term (model.py:98-109)
#include <stdlib.h>
#include <string.h>
#include "tlsan_mma_common.cuh"

// long-term FWA forward of one sample (model.py:98-109, 334-345) -> softmax state.
// buf = per-warp [32][64] row staging area.
__device__ __forceinline__ void long_forward(const FArgs& a, const LaneGeo& L, int b, int u, int ell, float gamma,
                                             const FwaW& w, float (*buf)[64], Soft2& st) {
  st.init();
  for (int r0 = 0; r0 < ell; r0 += 32) {
    const LongMeta me = load_long_meta(a, b, u, r0 + L.lane, ell, gamma);
    const int cnt = min(32, ell - r0);
    stage_round_rows(a, me, cnt, L.lane, buf);
    for (int j = 0; j < cnt; j += 2) {
      const bool okB = j + 1 < cnt;
      const float tA = __shfl_sync(0xffffffffu, me.tau, j), tB = __shfl_sync(0xffffffffu, me.tau, (j + 1) & 31);
      const float2 eA = *reinterpret_cast<const float2*>(&buf[j][L.f0]);
      const float2 eB = okB ? *reinterpret_cast<const float2*>(&buf[j + 1][L.f0]) : make_float2(0.f, 0.f);
      const float x[4] = {eA.x * tA, eA.y * tA, okB ? eB.x * tB : 0.f, okB ? eB.y * tB : 0.f};
      float m1[4], m2[4];
      tile_maps(x, w, m1, m2);
      st.push(m2[0], m2[1], x[0], x[1]);
      if (okB) st.push(m2[2], m2[3], x[2], x[3]);
    }
  }
}

// shared-memory image of the dense layer (natural layouts: lane reads float2 at [k][f0])
struct SmemMma {
  float red[MMA_WARPS][160];     // MODE 2: end-of-kernel reduction staging
  float rows[MMA_WARPS][32][64]; // MODE 0/1: per-warp staging of one round of long-term token rows
  float vec[MMA_WARPS][64];      // MODE 0: per-warp staging of o_long
  float bd[64];
  float wd[64 * 64];             // MODE 0: Wd[k][f]
};

// out[jj] += sum_k vec[k] * W[k][f0 + jj]
__device__ __forceinline__ void dense2(const float* __restrict__ vec, const float* __restrict__ W, int f0,
                                       float (&out)[2]) {
#pragma unroll 4
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 v = *reinterpret_cast<const float4*>(vec + 4 * k4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float2 wv = *reinterpret_cast<const float2*>(W + (4 * k4 + kk) * 64 + f0);
      out[0] = fmaf(vv[kk], wv.x, out[0]);
      out[1] = fmaf(vv[kk], wv.y, out[1]);
    }
  }
}

// MODE 0: scoring, fully fused (long FWA -> dense -> short FWA -> logits), dense as FFMA from smem
// MODE 1: training, long-term FWA forward only: o_long and its softmax statistics -> scratch
// MODE 2: training, short-term FWA forward (z read from scratch) + logit + loss + backward of
//         logit / short FWA -> gradient rows, dz -> scratch.  The 64x64 dense layer between
//         MODE 1 and MODE 2 (and its backward) runs as batched tensor-core GEMMs (k_dense_*).
// MODE 3: scoring with a workspace: short-term FWA forward (z from scratch) + logits; used after
//         MODE 1 + k_dense_fwd_mma by tlsan_score_ws (the dense layer leaves the per-sample kernel).
template <int MODE>
__global__ void __launch_bounds__(MMA_THREADS, (MODE == 1 || MODE == 3) ? 3 : 2) k_fwd_mma(const FArgs a, const int ncand) {
  constexpr bool TRAIN = MODE == 2;
  constexpr bool LONG = MODE == 0 || MODE == 1;     // runs the long-term FWA itself
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemMma& sm = *reinterpret_cast<SmemMma*>(smem_raw);
  if (MODE == 0) {
    for (int e = threadIdx.x; e < 64 * 64; e += MMA_THREADS) sm.wd[e] = a.dense[TLSAN_OFF_WD + e];
    if (threadIdx.x < 64) sm.bd[threadIdx.x] = a.dense[TLSAN_OFF_BD + threadIdx.x];
    __syncthreads();
  }

  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  const float gamma = a.dense[TLSAN_OFF_GAMMA];
  FwaW wl, ws;
  FwaWT wst;
  FwaGrad G;
  float loss_acc = 0.f, sq_acc = 0.f;
  if (LONG) wl = load_fwa(a.dense, TLSAN_OFF_W1L, L.g, L.t);
  if (MODE != 1) ws = load_fwa(a.dense, TLSAN_OFF_W1S, L.g, L.t);
  if (TRAIN) { wst = load_fwa_t(a.dense, TLSAN_OFF_W1S, L.g, L.t); G.init(); }
  float* vec = sm.vec[warp];
  const int nwarps = gridDim.x * MMA_WARPS;
  // per-warp staging of one round of long-term token rows.  MODE 1 (training) owns the whole dynamic smem and
  // sizes it by L (a round holds min(32, ell) tokens): 16-row slots when L <= 16 leave shared memory for the
  // radix-sort kernels that run beside it on the side stream.
  float (*rowbuf)[64] = MODE == 1 ? reinterpret_cast<float (*)[64]>(smem_raw) + (size_t)warp * (a.L <= 16 ? 16 : 32)
                                  : sm.rows[warp];

  for (int b = blockIdx.x * MMA_WARPS + warp; b < a.B; b += nwarps) {
    const int u = __ldg(a.u + b);
    float z[2] = {0.f, 0.f};
    if (LONG) {
      // ---- long-term FWA forward
      const int ell = __ldg(a.sl + b);
      Soft2 st;
      long_forward(a, L, b, u, ell, gamma, wl, rowbuf, st);
      const float o[2] = {st.den[0] > 0.f ? st.acc[0] / st.den[0] : 0.f, st.den[1] > 0.f ? st.acc[1] / st.den[1] : 0.f};
      if (MODE == 1) {
        float* sc = a.scratch + (size_t)b * (TLSAN_SCR * 64) + L.f0;
        st2(sc + 64, o[0], o[1]);
        st2(sc + 128, st.mx[0], st.mx[1]);
        st2(sc + 192, 1.f / st.den[0], 1.f / st.den[1]);
        continue;
      }
      __syncwarp();
      st2(vec + L.f0, o[0], o[1]);
      __syncwarp();
      // ---- z = o_long Wd + bd   (model.py:347)
      dense2(vec, sm.wd, L.f0, z);
      z[0] += sm.bd[L.f0]; z[1] += sm.bd[L.f0 + 1];
    } else {
      const float2 zz = *reinterpret_cast<const float2*>(a.scratch + (size_t)b * (TLSAN_SCR * 64) + 320 + L.f0);
      z[0] = zz.x; z[1] = zz.y;
    }
    const int s = __ldg(a.sl_new + b);
    const int cand = __ldg(a.i + b), uc = __ldg(a.c + b);
    // ---- short-term FWA forward over s+1 tokens (model.py:350-364)
    const int ntok = s + 1;
    Soft2 ss; ss.init();
    for (int r0 = 0; r0 < ntok; r0 += 32) {       // round covers tokens r0 .. r0+31 (token n >= 1 is item n-1)
      // lane l holds the meta of item r0 + l, i.e. of token r0 + l + 1
      const int item = r0 + L.lane;
      const int id_l = item < s ? __ldg(a.hist_i_new + (size_t)b * a.S + item) : 0;
      const int crow_l = a.NI + __ldg(a.icl + id_l);
      const int cnt = min(32, ntok - r0);          // tokens in this round
      for (int j = 0; j < cnt; j += 2) {
        float x[4]; bool okB;
        // token index n = r0 + j (A) and n+1 (B); item index = n - 1 -> lane (n - 1 - r0) = j - 1 (A), j (B)
        const int nA = r0 + j, nB = nA + 1;
        okB = nB < ntok;
        const int laneA = (j - 1) & 31, laneB = j & 31;
        int idA = __shfl_sync(0xffffffffu, id_l, laneA), crA = __shfl_sync(0xffffffffu, crow_l, laneA);
        const int idB = __shfl_sync(0xffffffffu, id_l, laneB), crB = __shfl_sync(0xffffffffu, crow_l, laneB);
        if (nA == 0) { x[0] = z[0]; x[1] = z[1]; }
        else {
          if (j == 0) {   // first token of a later round: item r0-1 was not loaded in this round
            idA = __ldg(a.hist_i_new + (size_t)b * a.S + (nA - 1));
            crA = a.NI + __ldg(a.icl + idA);
          }
          const float2 e = ldg2(row_ptr(a, L, idA, crA)); x[0] = e.x; x[1] = e.y;
        }
        if (okB) { const float2 e = ldg2(row_ptr(a, L, idB, crB)); x[2] = e.x; x[3] = e.y; }
        else { x[2] = 0.f; x[3] = 0.f; }
        float m1[4], m2[4];
        tile_maps(x, ws, m1, m2);
        ss.push(m2[0], m2[1], x[0], x[1]);
        if (okB) ss.push(m2[2], m2[3], x[2], x[3]);
      }
    }
    const float inv_s[2] = {1.f / ss.den[0], 1.f / ss.den[1]};
    const float v[2] = {ss.acc[0] * inv_s[0], ss.acc[1] * inv_s[1]};
    // ---- user vector, candidate, logit (model.py:84-95,135-137)
    const float2 p = ldg2(a.emb + (size_t)(L.half ? a.NI + uc : a.NI + a.NC + u) * 32 + L.col);
    const float ut[2] = {v[0] + p.x, v[1] + p.y};
    const int ccrow = a.NI + __ldg(a.icl + cand);
    const float2 q = ldg2(row_ptr(a, L, cand, ccrow));
    const float logit = warp_sum_f(fmaf(ut[0], q.x, ut[1] * q.y)) + __ldg(a.item_b + cand);

    if (!TRAIN) {
      if (L.lane == 0) a.logits[(size_t)b * ncand] = logit;
      if (a.ut) st2(a.ut + (size_t)b * 64 + L.f0, ut[0], ut[1]);
      if (ncand > 1) {   // Model.eval_auc second run (model.py:251-261): same u_t, other item
        const int c2 = __ldg(a.i2 + b);
        const int c2row = a.NI + __ldg(a.icl + c2);
        const float2 q2 = ldg2(row_ptr(a, L, c2, c2row));
        const float l2 = warp_sum_f(fmaf(ut[0], q2.x, ut[1] * q2.y)) + __ldg(a.item_b + c2);
        if (L.lane == 0) a.logits[(size_t)b * ncand + 1] = l2;
      }
      continue;
    }

    // =========================== backward ===========================
    const float yb = __ldg(a.y + b);
    const float ex = expf(-fabsf(logit));
    const float bce = fmaxf(logit, 0.f) - logit * yb + log1pf(ex);     // model.py:171
    const float sig = logit >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
    const float gl = (sig - yb) * a.invB;                               // d loss / d logit
    if (L.lane == 0) { loss_acc += bce; sq_acc = fmaf(gl, gl, sq_acc); a.gscal[b] = gl; }
    float* rcand = grad_row(a, b, a.L + a.S) + L.f0;
    float* rvirt = grad_row(a, b, a.L + a.S + 1) + L.f0;
    const float dq[2] = {gl * ut[0], gl * ut[1]};
    const float du[2] = {gl * q.x, gl * q.y};
    sq_acc = fmaf(dq[0], dq[0], sq_acc); sq_acc = fmaf(dq[1], dq[1], sq_acc);
    sq_acc = fmaf(du[0], du[0], sq_acc); sq_acc = fmaf(du[1], du[1], sq_acc);
    st2(rcand, dq[0], dq[1]);                                           // -> item_emb[i] | cate_emb[icl[i]]
    if (L.half) st2(rvirt, du[0], du[1]);                               // -> cate_emb[u_cate]
    else { st2(rvirt, 0.f, 0.f); st2(a.rows_u + (size_t)b * a.PU + L.f0, du[0], du[1]); }  // -> user_emb[u]
    // short-term FWA backward, d v = du
    float dz[2] = {0.f, 0.f};
    for (int r0 = 0; r0 < ntok; r0 += 32) {
      const int item = r0 + L.lane;
      const int id_l = item < s ? __ldg(a.hist_i_new + (size_t)b * a.S + item) : 0;
      const int crow_l = a.NI + __ldg(a.icl + id_l);
      const int cnt = min(32, ntok - r0);
      for (int j = 0; j < cnt; j += 2) {
        float x[4], dx[4];
        const int nA = r0 + j, nB = nA + 1;
        const bool okB = nB < ntok;
        const int laneA = (j - 1) & 31, laneB = j & 31;
        int idA = __shfl_sync(0xffffffffu, id_l, laneA), crA = __shfl_sync(0xffffffffu, crow_l, laneA);
        const int idB = __shfl_sync(0xffffffffu, id_l, laneB), crB = __shfl_sync(0xffffffffu, crow_l, laneB);
        if (nA == 0) { x[0] = z[0]; x[1] = z[1]; }
        else {
          if (j == 0) {
            idA = __ldg(a.hist_i_new + (size_t)b * a.S + (nA - 1));
            crA = a.NI + __ldg(a.icl + idA);
          }
          const float2 e = ldg2(row_ptr(a, L, idA, crA)); x[0] = e.x; x[1] = e.y;
        }
        if (okB) { const float2 e = ldg2(row_ptr(a, L, idB, crB)); x[2] = e.x; x[3] = e.y; }
        else { x[2] = 0.f; x[3] = 0.f; }
        tile_bwd(x, okB, v, du, ss.mx, inv_s, ws, wst, L.lane, dx, G);
        if (nA == 0) { dz[0] = dx[0]; dz[1] = dx[1]; }
        else {
          sq_acc = fmaf(dx[0], dx[0], sq_acc); sq_acc = fmaf(dx[1], dx[1], sq_acc);
          st2(grad_row(a, b, a.L + (nA - 1)) + L.f0, dx[0], dx[1]);
        }
        if (okB) {
          sq_acc = fmaf(dx[2], dx[2], sq_acc); sq_acc = fmaf(dx[3], dx[3], sq_acc);
          st2(grad_row(a, b, a.L + (nB - 1)) + L.f0, dx[2], dx[3]);
        }
      }
    }
    // ---- dz -> scratch: k_dense_bwd_mma turns it into d o_long, dWd and dbd
    st2(a.scratch + (size_t)b * (TLSAN_SCR * 64) + 256 + L.f0, dz[0], dz[1]);
  }

  if (TRAIN) {
    // lanes with equal t hold the same (k, column) slots: butterfly over g, then warps in order
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
        if (L.g == 0) { sm.red[warp][k * 8 + 2 * L.t + jj] = r1; sm.red[warp][72 + k * 8 + 2 * L.t + jj] = r2; }
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
      if (L.g == 0) { sm.red[warp][64 + 2 * L.t + jj] = r1; sm.red[warp][136 + 2 * L.t + jj] = r2; }
    }
    {
      const float r1 = warp_sum_f(loss_acc), r2 = warp_sum_f(sq_acc);
      if (L.lane == 0) { sm.red[warp][144] = r1; sm.red[warp][145] = r2; }
    }
    __syncthreads();
    if (threadIdx.x < 146) {
      float r = 0.f;
#pragma unroll
      for (int wv = 0; wv < MMA_WARPS; ++wv) r += sm.red[wv][threadIdx.x];
      const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1S + threadIdx.x
                                        : (threadIdx.x == 144 ? TLSAN_PART_LOSS : TLSAN_PART_SUMSQ);
      a.part[(size_t)blockIdx.x * TLSAN_PART + dst] = r;
    }
  }
}

// ------------------------------------------------------------------ backward of the long-term FWA
__global__ void __launch_bounds__(MMA_THREADS, 2) k_bwd_long_mma(const FArgs a) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  float (*red)[160] = reinterpret_cast<float (*)[160]>(smem_b);                   // reused after the loop
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  float (*rowsw)[64] = reinterpret_cast<float (*)[64]>(smem_b) + (size_t)warp * (a.L <= 16 ? 16 : 32);   // [warp][16|32][64]
  const float gamma = a.dense[TLSAN_OFF_GAMMA];
  const FwaW wl = load_fwa(a.dense, TLSAN_OFF_W1L, L.g, L.t);
  const FwaWT wlt = load_fwa_t(a.dense, TLSAN_OFF_W1L, L.g, L.t);
  FwaGrad G; G.init();
  float ggamma = 0.f, sq_acc = 0.f;
  const int nwarps = gridDim.x * MMA_WARPS;

  for (int b = blockIdx.x * MMA_WARPS + warp; b < a.B; b += nwarps) {
    const int u = __ldg(a.u + b), ell = __ldg(a.sl + b);
    const float* sc = a.scratch + (size_t)b * (TLSAN_SCR * 64) + L.f0;
    const float2 dol2 = *reinterpret_cast<const float2*>(sc);
    const float2 o2 = *reinterpret_cast<const float2*>(sc + 64);
    const float2 mx2 = *reinterpret_cast<const float2*>(sc + 128);
    const float2 inv2 = *reinterpret_cast<const float2*>(sc + 192);
    const float dol[2] = {dol2.x, dol2.y}, o[2] = {o2.x, o2.y}, mx[2] = {mx2.x, mx2.y}, inv[2] = {inv2.x, inv2.y};
    float* ru = a.rows_u + (size_t)b * a.PU + 32;
    for (int r0 = 0; r0 < ell; r0 += 32) {
      const LongMeta me = load_long_meta(a, b, u, r0 + L.lane, ell, gamma);
      const int cnt = min(32, ell - r0);
      float dtau_l = 0.f;                          // lane j collects d tau of token r0 + j
      const int inv_l = L.lane < cnt ? __ldg(a.inv + ((size_t)b << a.spsh) + r0 + L.lane) : 0;   // sorted ranks
      stage_round_rows(a, me, cnt, L.lane, rowsw);
      for (int j = 0; j < cnt; j += 2) {
        Pair cur;
        cur.okB = j + 1 < cnt;
        cur.tA = __shfl_sync(0xffffffffu, me.tau, j); cur.tB = __shfl_sync(0xffffffffu, me.tau, (j + 1) & 31);
        cur.eA = *reinterpret_cast<const float2*>(&rowsw[j][L.f0]);
        cur.eB = cur.okB ? *reinterpret_cast<const float2*>(&rowsw[j + 1][L.f0]) : make_float2(0.f, 0.f);
        const int posA = __shfl_sync(0xffffffffu, inv_l, j), posB = __shfl_sync(0xffffffffu, inv_l, (j + 1) & 31);
        const bool okB = cur.okB;
        const float2 eA = cur.eA, eB = cur.eB;
        const float tA = cur.tA, tB = cur.tB;
        const float x[4] = {eA.x * tA, eA.y * tA, okB ? eB.x * tB : 0.f, okB ? eB.y * tB : 0.f};
        float dx[4];
        tile_bwd(x, okB, o, dol, mx, inv, wl, wlt, L.lane, dx, G);
        // gradient of the gathered slices (tau * dX) and of tau (<dX, e>)
        const float rA0 = dx[0] * tA, rA1 = dx[1] * tA;
        sq_acc = fmaf(rA0, rA0, sq_acc); sq_acc = fmaf(rA1, rA1, sq_acc);
        st2(a.rows_i + (size_t)posA * 64 + L.f0, rA0, rA1);
        const float dtA = warp_sum_f(fmaf(dx[0], eA.x, dx[1] * eA.y));
        if (L.lane == j) dtau_l = dtA;
        if (okB) {
          const float rB0 = dx[2] * tB, rB1 = dx[3] * tB;
          sq_acc = fmaf(rB0, rB0, sq_acc); sq_acc = fmaf(rB1, rB1, sq_acc);
          st2(a.rows_i + (size_t)posB * 64 + L.f0, rB0, rB1);
          const float dtB = warp_sum_f(fmaf(dx[2], eB.x, dx[3] * eB.y));
          if (L.lane == j + 1) dtau_l = dtB;
        }
      }
      if (L.lane < cnt) {
        ggamma = fmaf(dtau_l, me.pt, ggamma);
        const float dp = dtau_l * gamma * me.ht;   // d usert_emb[u, t]
        sq_acc = fmaf(dp, dp, sq_acc);
        ru[r0 + L.lane] = dp;
      }
    }
    for (int tt = ell + L.lane; tt < a.PU - 32; tt += 32) ru[tt] = 0.f;
  }

  __syncthreads();   // all warps are done with their staged rows: the area is reused for the reduction

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
      if (L.g == 0) { red[warp][k * 8 + 2 * L.t + jj] = r1; red[warp][72 + k * 8 + 2 * L.t + jj] = r2; }
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
    if (L.g == 0) { red[warp][64 + 2 * L.t + jj] = r1; red[warp][136 + 2 * L.t + jj] = r2; }
  }
  {
    const float r1 = warp_sum_f(ggamma), r2 = warp_sum_f(sq_acc);
    if (L.lane == 0) { red[warp][144] = r1; red[warp][145] = r2; }
  }
  __syncthreads();
  if (threadIdx.x < 146) {
    float r = 0.f;
#pragma unroll
    for (int wv = 0; wv < MMA_WARPS; ++wv) r += red[wv][threadIdx.x];
    const int dst = threadIdx.x < 144 ? TLSAN_OFF_W1L + threadIdx.x
                                      : (threadIdx.x == 144 ? TLSAN_OFF_GAMMA : TLSAN_PART_SUMSQ);
    a.part[(size_t)blockIdx.x * TLSAN_PART + dst] = r;
  }
}

// ------------------------------------------------------------------ dense 64x64 layer as batched GEMMs
// tf.layers.dense (model.py:347) over the whole batch on the tensor cores (3xTF32 mma.sync):
//   k_dense_fwd_mma : Z = O Wd + bd                       (scratch slot 1 -> slot 5)
//   k_dense_bwd_mma : dO = dZ Wd^T (slot 4 -> slot 0) ; dWd = O^T dZ, dbd = sum dZ -> part[cta]
// The B operand is pre-split into tf32 hi / lo images in shared memory, row stride 68 floats
// (fragment reads are bank-conflict free); A slot t <-> column 2t, slot t+4 <-> column 2t+1.
#define GEMM_LD 68
struct SmemGemm { float hi[64 * GEMM_LD]; float lo[64 * GEMM_LD]; };

__device__ __forceinline__ void split4(const float (&x)[4], uint32_t (&h)[4], uint32_t (&l)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = to_tf32(x[i]);
    l[i] = __float_as_uint(x[i] - __uint_as_float(h[i]));   // the tensor core drops the low 13 bits itself
  }
}
// acc += A(16x8) B(8x8), B fragment read from the hi/lo images at rows (k0+2t, k0+2t+1), column n0+g
// the big term (hi x hi) and the two small ones run on separate accumulators -- two independent HMMA chains per output
// tile instead of one three times as long; the caller adds `accs` to `acc` once, after the k loop
__device__ __forceinline__ void gemm_step(float (&acc)[4], float (&accs)[4], const uint32_t (&h)[4],
                                          const uint32_t (&l)[4], const SmemGemm& sb, int k0, int n0, int g, int t) {
  const int i0 = (k0 + 2 * t) * GEMM_LD + n0 + g, i1 = i0 + GEMM_LD;
  const uint32_t bh0 = __float_as_uint(sb.hi[i0]), bh1 = __float_as_uint(sb.hi[i1]);
  const uint32_t bl0 = __float_as_uint(sb.lo[i0]), bl1 = __float_as_uint(sb.lo[i1]);
  mma_tf32(accs, l[0], l[2], l[1], l[3], bh0, bh1);
  mma_tf32(acc, h[0], h[2], h[1], h[3], bh0, bh1);
  mma_tf32(accs, h[0], h[2], h[1], h[3], bl0, bl1);
}

// 64-row tiles of o_long travel global -> shared with cp.async, double-buffered: the next tile is in flight while this
// one is multiplied (the first version loaded each warp's A fragment into registers right before its 96 HMMAs:
// long_scoreboard 56 % of the stall samples, 10 % of the DRAM bandwidth).  Warp w: rows 16 (w / 2) .. +15 of the
// tile, output columns 32 (w % 2) .. +31.
#define TILE_LD 72
#define DF_ROWS 64
__global__ void __launch_bounds__(256, 2) k_dense_fwd_mma(const float* __restrict__ dense, float* __restrict__ scratch,
                                                       int B) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  SmemGemm& sb = *reinterpret_cast<SmemGemm*>(dyn_smem);
  float* sbd = reinterpret_cast<float*>(dyn_smem + sizeof(SmemGemm));
  float (*sa)[DF_ROWS * TILE_LD] = reinterpret_cast<float (*)[DF_ROWS * TILE_LD]>(dyn_smem + sizeof(SmemGemm) + 256);
  for (int e4 = threadIdx.x; e4 < 64 * 16; e4 += 256) {       // float4 loads: 4 independent per thread
    const float4 w4 = *reinterpret_cast<const float4*>(dense + TLSAN_OFF_WD + 4 * e4);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = 4 * e4 + q;
      const float h = __uint_as_float(to_tf32(wv[q]));
      sb.hi[(e >> 6) * GEMM_LD + (e & 63)] = h;
      sb.lo[(e >> 6) * GEMM_LD + (e & 63)] = __uint_as_float(to_tf32(wv[q] - h));
    }
  }
  if (threadIdx.x < 64) sbd[threadIdx.x] = dense[TLSAN_OFF_BD + threadIdx.x];
  pdl_wait();                                      // o_long of the long-term forward
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int half = warp & 1, r0 = (warp >> 1) * 16;
  const int ntiles = (B + DF_ROWS - 1) / DF_ROWS;
  // rows of a tile: 4 x 16-byte cp.async per thread, rows past B zero-filled
  auto issue = [&](int tile, int buf) {
    for (int e = threadIdx.x; e < DF_ROWS * 16; e += 256) {
      const int r = e >> 4, q = e & 15, row = tile * DF_ROWS + r;
      float* dst = sa[buf] + r * TILE_LD + 4 * q;
      if (row < B) cp16_async(dst, scratch + (size_t)row * (TLSAN_SCR * 64) + 64 + 4 * q);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int tile = blockIdx.x;
  if (tile < ntiles) issue(tile, 0);
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    cp_async_wait_all();
    __syncthreads();                                // this tile has landed (and sb is written); the previous one is consumed
    if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, buf ^ 1);
    const float* a = sa[buf] + (r0 + g) * TILE_LD + 2 * t;
    float acc[4][4], accs[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][2] = sbd[half * 32 + nt * 8 + 2 * t];
      acc[nt][1] = acc[nt][3] = sbd[half * 32 + nt * 8 + 2 * t + 1];
      accs[nt][0] = accs[nt][1] = accs[nt][2] = accs[nt][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const float2 xa = *reinterpret_cast<const float2*>(a + ks * 8);
      const float2 xb = *reinterpret_cast<const float2*>(a + 8 * TILE_LD + ks * 8);
      const float x[4] = {xa.x, xa.y, xb.x, xb.y};
      uint32_t h[4], l[4];
      split4(x, h, l);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) gemm_step(acc[nt], accs[nt], h, l, sb, ks * 8, half * 32 + nt * 8, g, t);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] += accs[nt][i];
    const int rA = tile * DF_ROWS + r0 + g, rB = rA + 8;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = 320 + half * 32 + nt * 8 + 2 * t;
      if (rA < B) st2(scratch + (size_t)rA * (TLSAN_SCR * 64) + col, acc[nt][0], acc[nt][1]);
      if (rB < B) st2(scratch + (size_t)rB * (TLSAN_SCR * 64) + col, acc[nt][2], acc[nt][3]);
    }
  }
}

__global__ void __launch_bounds__(128) k_dense_bwd_mma(const float* __restrict__ dense, float* __restrict__ scratch,
                                                       int B, float* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];   // 55 KB: over the static limit, opted in by the launcher
  SmemGemm& sb = *reinterpret_cast<SmemGemm*>(dyn_smem);      // Wd^T: image[j][f] = Wd[f][j]
  float (*so2)[16 * TILE_LD] = reinterpret_cast<float (*)[16 * TILE_LD]>(dyn_smem + sizeof(SmemGemm));   // double-buffered
  float (*sz2)[16 * TILE_LD] = so2 + 2;                       // tiles of o_long / dz (cp.async)
  for (int e4 = threadIdx.x; e4 < 64 * 16; e4 += 128) {       // float4 loads: 8 independent per thread
    const float4 w4 = *reinterpret_cast<const float4*>(dense + TLSAN_OFF_WD + 4 * e4);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = 4 * e4 + q;                      // Wd[f = e>>6][j = e&63]
      const float h = __uint_as_float(to_tf32(wv[q]));
      sb.hi[(e & 63) * GEMM_LD + (e >> 6)] = h;
      sb.lo[(e & 63) * GEMM_LD + (e >> 6)] = __uint_as_float(to_tf32(wv[q] - h));
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int ntiles = (B + 15) / 16;
  const int per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int t_lo = blockIdx.x * per, t_hi = min(ntiles, t_lo + per);
  float accw[8][4];                                 // dWd rows 16*warp .. +15, all 64 columns
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) accw[nt][0] = accw[nt][1] = accw[nt][2] = accw[nt][3] = 0.f;
  float bsum = 0.f;                                 // threads < 64: dbd[threadIdx.x]
  pdl_wait();                                       // dz of the short-term kernel
  pdl_trigger();
  // the 16 rows (o_long | dz) of a tile: 2 x 16-byte cp.async per thread and matrix, rows past B zero-filled
  auto issue = [&](int tile, int buf) {
    for (int e = threadIdx.x; e < 16 * 16; e += 128) {
      const int r = e >> 4, q = e & 15, row = tile * 16 + r;
      float* po = so2[buf] + r * TILE_LD + 4 * q;
      float* pz = sz2[buf] + r * TILE_LD + 4 * q;
      if (row < B) {
        const float* sc = scratch + (size_t)row * (TLSAN_SCR * 64);
        cp16_async(po, sc + 64 + 4 * q);
        cp16_async(pz, sc + 256 + 4 * q);
      } else {
        *reinterpret_cast<float4*>(po) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(pz) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (t_lo < t_hi) issue(t_lo, 0);
  for (int tile = t_lo; tile < t_hi; ++tile) {
    const int buf = (tile - t_lo) & 1;
    cp_async_wait_all();
    __syncthreads();                                // this tile has landed (and sb is written); the previous one is consumed
    if (tile + 1 < t_hi) issue(tile + 1, buf ^ 1);  // in flight while this tile is multiplied
    const float* so = so2[buf];
    const float* sz = sz2[buf];
    // (a) d o_long = dZ Wd^T : this warp computes output columns 16*warp .. +15 for the 16 rows
    {
      float acc[2][4], accs[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        accs[nt][0] = accs[nt][1] = accs[nt][2] = accs[nt][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const float2 a = *reinterpret_cast<const float2*>(sz + g * TILE_LD + ks * 8 + 2 * t);
        const float2 c = *reinterpret_cast<const float2*>(sz + (g + 8) * TILE_LD + ks * 8 + 2 * t);
        const float x[4] = {a.x, a.y, c.x, c.y};
        uint32_t h[4], l[4];
        split4(x, h, l);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) gemm_step(acc[nt], accs[nt], h, l, sb, ks * 8, warp * 16 + nt * 8, g, t);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] += accs[nt][i];
      const int rA = tile * 16 + g, rB = rA + 8;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = warp * 16 + nt * 8 + 2 * t;
        if (rA < B) st2(scratch + (size_t)rA * (TLSAN_SCR * 64) + col, acc[nt][0], acc[nt][1]);
        if (rB < B) st2(scratch + (size_t)rB * (TLSAN_SCR * 64) + col, acc[nt][2], acc[nt][3]);
      }
    }
    // (b) dWd[k][j] += sum_s O[s][k] dZ[s][j] : M = k (rows 16*warp..), N = j, K = the 16 samples
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int s0 = ks * 8, m0 = warp * 16;
      // A[m][slot] = O[s0 + slot][m0 + m] : a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4)
      const float xa[4] = {so[(s0 + t) * TILE_LD + m0 + g], so[(s0 + t + 4) * TILE_LD + m0 + g],
                           so[(s0 + t) * TILE_LD + m0 + g + 8], so[(s0 + t + 4) * TILE_LD + m0 + g + 8]};
      uint32_t h[4], l[4];
      split4(xa, h, l);   // mma3 order: a0 = x[0], a1 = x[2], a2 = x[1], a3 = x[3]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float b0 = sz[(s0 + t) * TILE_LD + nt * 8 + g], b1 = sz[(s0 + t + 4) * TILE_LD + nt * 8 + g];
        const uint32_t bh0 = to_tf32(b0), bh1 = to_tf32(b1);
        const uint32_t bl0 = __float_as_uint(b0 - __uint_as_float(bh0)), bl1 = __float_as_uint(b1 - __uint_as_float(bh1));
        mma_tf32(accw[nt], l[0], l[2], l[1], l[3], bh0, bh1);
        mma_tf32(accw[nt], h[0], h[2], h[1], h[3], bl0, bl1);
        mma_tf32(accw[nt], h[0], h[2], h[1], h[3], bh0, bh1);
      }
    }
    if (threadIdx.x < 64) {
#pragma unroll
      for (int r = 0; r < 16; ++r) bsum += sz[r * TILE_LD + threadIdx.x];
    }
  }
  float* p = part + (size_t)blockIdx.x * TLSAN_PART;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    st2(p + TLSAN_OFF_WD + (warp * 16 + g) * 64 + nt * 8 + 2 * t, accw[nt][0], accw[nt][1]);
    st2(p + TLSAN_OFF_WD + (warp * 16 + g + 8) * 64 + nt * 8 + 2 * t, accw[nt][2], accw[nt][3]);
  }
  if (threadIdx.x < 64) p[TLSAN_OFF_BD + threadIdx.x] = bsum;
}

// ------------------------------------------------------------------ launchers
FArgs tlsan_make_fargs(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b);


int tlsan_overlap_ctas() {
  static int v = 0;
  if (!v) { const char* e = getenv("TLSAN_OVERLAP_CTAS"); v = e ? atoi(e) : 2; if (v < 1 || v > 3) v = 2; }
  return v;
}
static int mma_grid(int B, int ctas_per_sm) {
  const int need = (B + MMA_WARPS - 1) / MMA_WARPS;
  const int cap = tlsan_num_sms() * ctas_per_sm;
  return need < cap ? need : cap;
}

int tlsan_launch_dense_fwd(const float* dense, float* scratch, int B, cudaStream_t st) {
  const int ntile = (B + DF_ROWS - 1) / DF_ROWS;
  const int gg = ntile < tlsan_num_sms() * 2 ? ntile : tlsan_num_sms() * 2;   // one resident wave: the B image is built once per CTA
  const size_t smem = sizeof(SmemGemm) + 256 + 2 * DF_ROWS * TILE_LD * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_dense_fwd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  tlsan_launch_k(k_dense_fwd_mma, dim3(gg), dim3(256), smem, st, dense, scratch, B);
  TLSAN_CHECK_LAUNCH("k_dense_fwd_mma");
  return TLSAN_OK;
}

int tlsan_launch_dense_bwd(const float* dense, float* scratch, int B, float* part, int* grid_c, cudaStream_t st) {
  const int ntile16 = (B + 15) / 16;
  const int gc = ntile16 < tlsan_num_sms() * 4 ? ntile16 : tlsan_num_sms() * 4;
  *grid_c = gc;
  const size_t smem = sizeof(SmemGemm) + 4 * 16 * TILE_LD * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_dense_bwd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  tlsan_launch_k(k_dense_bwd_mma, dim3(gc), dim3(128), smem, st, dense, scratch, B, part);
  TLSAN_CHECK_LAUNCH("k_dense_bwd_mma");
  return TLSAN_OK;
}

static const int kSmemLongMax = (int)(sizeof(float) * MMA_WARPS * 32 * 64);
// staging area of the long-term kernels: [warp][16 or 32 token rows][64]; never below the 8 x 160 floats the
// backward reuses for its end-of-kernel reduction
static int smem_long(int L) { return (int)(sizeof(float) * MMA_WARPS * (L <= 16 ? 16 : 32) * 64); }

static int set_long_attrs() {
  static bool done = false;
  if (!done) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_fwd_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLongMax));
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_long_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLongMax));
    done = true;
  }
  return TLSAN_OK;
}

// ctas_per_sm: 3 fills the register file; 2 leaves a third of it to the radix-sort kernels on the side stream
int tlsan_launch_long_fwd_mma(const FArgs& a, int ctas_per_sm, cudaStream_t st) {
  int rc = set_long_attrs();
  if (rc) return rc;
  k_fwd_mma<1><<<mma_grid(a.B, ctas_per_sm), MMA_THREADS, smem_long(a.L), st>>>(a, 1);
  TLSAN_CHECK_LAUNCH("k_fwd_mma<long>");
  return TLSAN_OK;
}

int tlsan_launch_bwd_long_mma(const FArgs& a, int* grid_b, cudaStream_t st) {
  int rc = set_long_attrs();
  if (rc) return rc;
  const int g = mma_grid(a.B, 2);
  *grid_b = g;
  k_bwd_long_mma<<<g, MMA_THREADS, smem_long(a.L), st>>>(a);
  TLSAN_CHECK_LAUNCH("k_bwd_long_mma");
  return TLSAN_OK;
}

int tlsan_launch_score_mma(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                           float* logits, float* ut, cudaStream_t st) {
  FArgs a = tlsan_make_fargs(d, p, b);
  a.logits = logits; a.ut = ut;
  static bool attr_set = false;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_fwd_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SmemMma)));
    attr_set = true;
  }
  k_fwd_mma<0><<<mma_grid(d.B, 2), MMA_THREADS, sizeof(SmemMma), st>>>(a, ncand);
  TLSAN_CHECK_LAUNCH("k_fwd_mma<score>");
  return TLSAN_OK;
}

int tlsan_launch_long_meta(const FArgs& a, void* meta, void* smeta, void* sscal, void* part, int fwd_ctas,
                           int score_ncand, cudaStream_t st);
size_t tlsan_score_meta_bytes(int B, int S);
int tlsan_launch_score_pf(const FArgs& a, const void* smeta, const void* sscal, const void* part, int ncand,
                          cudaStream_t st);                                                     // tlsan_fused_pf.cu
int tlsan_launch_long_fwd_pf(const FArgs& a, const void* meta, const void* part, int ctas_per_sm, cudaStream_t st);
int tlsan_launch_partition(const FArgs& a, int fwd_ctas, bool train, void* part, cudaStream_t st);
size_t tlsan_partition_bytes();

// scoring with a caller-provided scratch [B][TLSAN_SCR][64]: long FWA -> batched dense GEMM -> short FWA + logits
int tlsan_launch_score_ws(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b, int ncand,
                          float* logits, float* ut, float* scratch, cudaStream_t st) {
  FArgs a = tlsan_make_fargs(d, p, b);
  a.logits = logits; a.ut = ut; a.scratch = scratch;
  int rc;
  static int use_ws = -1;
  if (use_ws < 0) { const char* e = getenv("TLSAN_FUSED_IMPL"); use_ws = (e && *e && strcmp(e, "pf") != 0) ? 0 : 1; }
  void* meta = reinterpret_cast<char*>(scratch) + tlsan_align_up((size_t)d.B * TLSAN_SCR * 64 * sizeof(float), 256);
  void* part = reinterpret_cast<char*>(meta) + tlsan_align_up((size_t)d.B * d.L * 16, 256);
  // short-term metadata (scoring layout) behind the partition block: [B][S + 3] row pairs, [B] scalars
  void* smeta = reinterpret_cast<char*>(part) + tlsan_align_up(tlsan_partition_bytes(), 256);
  void* sscal = reinterpret_cast<char*>(smeta) + tlsan_align_up((size_t)d.B * (d.S + 3) * 8, 256);
  if (use_ws && (rc = tlsan_launch_long_meta(a, meta, smeta, sscal, part, 3, ncand, st))) return rc;
  if ((rc = use_ws ? tlsan_launch_long_fwd_pf(a, meta, part, 3, st) : tlsan_launch_long_fwd_mma(a, 3, st))) return rc;
  if ((rc = tlsan_launch_dense_fwd(p.dense, scratch, d.B, st))) return rc;
  if (use_ws) return tlsan_launch_score_pf(a, smeta, sscal, part, ncand, st);
  k_fwd_mma<3><<<mma_grid(d.B, 3), MMA_THREADS, 0, st>>>(a, ncand);
  TLSAN_CHECK_LAUNCH("k_fwd_mma<short score>");
  return TLSAN_OK;
}

int tlsan_launch_fwd_bwd_mma(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b,
                             const TlsanWs& w, char* ws, int* grid_a, int* grid_b, int* grid_c, cudaEvent_t sorted,
                             int long_ctas, cudaStream_t st) {
  FArgs a = tlsan_make_fargs(d, p, b);
  a.rows_i = reinterpret_cast<float*>(ws + w.rows_i);
  a.inv = reinterpret_cast<const int*>(ws + w.inv); a.spsh = w.SPSH;
  a.rows_u = reinterpret_cast<float*>(ws + w.rows_u);
  a.gscal = reinterpret_cast<float*>(ws + w.gscal);
  a.scratch = reinterpret_cast<float*>(ws + w.scratch);
  // forward: long FWA -> dense GEMM -> short FWA + loss + backward of logit / short FWA
  int rc;
  if ((rc = tlsan_launch_long_fwd_mma(a, long_ctas, st))) return rc;
  tlsan_profile_mark(TLSAN_PHASE_LONG_FWD, st);
  if ((rc = tlsan_launch_dense_fwd(p.dense, a.scratch, d.B, st))) return rc;
  const int g = mma_grid(d.B, 2);
  *grid_a = g; *grid_b = g;
  a.part = reinterpret_cast<float*>(ws + w.part_a);
  if (sorted) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, sorted, 0));   // gradient rows are written at sorted rank
  tlsan_profile_mark(TLSAN_PHASE_DENSE_FWD, st);                       // (phase includes the join with the sort stream)
  k_fwd_mma<2><<<g, MMA_THREADS, sizeof(float) * MMA_WARPS * 160, st>>>(a, 1);
  TLSAN_CHECK_LAUNCH("k_fwd_mma<short>");
  tlsan_profile_mark(TLSAN_PHASE_SHORT, st);
  // backward: dense GEMMs (d o_long, dWd, dbd) -> long FWA
  if ((rc = tlsan_launch_dense_bwd(p.dense, a.scratch, d.B, reinterpret_cast<float*>(ws + w.part_c), grid_c, st)))
    return rc;
  tlsan_profile_mark(TLSAN_PHASE_DENSE_BWD, st);
  a.part = reinterpret_cast<float*>(ws + w.part_b);
  if ((rc = tlsan_launch_bwd_long_mma(a, grid_b, st))) return rc;
  tlsan_profile_mark(TLSAN_PHASE_BWD_LONG, st);
  return TLSAN_OK;
}
