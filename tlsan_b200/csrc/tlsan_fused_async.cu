// Tensor-core TLSAN kernels with an asynchronous, register-free sample pipeline (default path).
//
// Same math and lane layout as tlsan_fused_mma.cu (one warp = one sample, 16-row mma tiles), but
// every global read of a sample is issued 1-3 samples AHEAD with cp.async (LDGSTS) into a private
// per-warp shared-memory ring, so the three dependent levels of the gather
//        ids / lengths  ->  icl[id], usert[u]  ->  embedding rows
// never stall the compute:
//   stage A (sample i+3): scalars, hist_i / hist_t / hist_i_new, the sample's scratch record
//   stage B (sample i+2): icl[id] of every token, usert[u][t]          (needs stage A values)
//   stage C (sample i+1): the 256-B token rows (item row | cate row)   (needs stage B values)
//   stage D (sample i)  : compute, reading everything from shared memory
// One cp.async.wait_all + __syncwarp per sample; a whole sample of compute separates the issue
// of a copy from its first use.  Tokens beyond the prefetch window (RL long / RS short rows,
// 32 ids) fall back to direct global loads, so any Ls <= TLSAN_MAX_L stays correct.
//
//   k_async<1>  long-term FWA forward -> o_long + softmax statistics (scratch)
//   k_async<2>  short FWA forward + logit + loss + backward of logit / short FWA
//   k_async<3>  backward of the long-term FWA and of the time-aware position term
#include "tlsan_mma_common.cuh"

#define RL 12   // long-term tokens whose rows are staged per sample
#define RS 6    // short-term items whose rows are staged per sample (+ candidate + user rows)

template <int KIND> struct Cfg;
template <> struct Cfg<1> { static constexpr int SCRF = 4, NROWS = RL, CTAS = 3, NINV = 4; };
template <> struct Cfg<2> { static constexpr int SCRF = 64, NROWS = RS + 2, CTAS = 2, NINV = 32; };
template <> struct Cfg<3> { static constexpr int SCRF = 256, NROWS = RL, CTAS = 2, NINV = 32; };

template <int KIND>
struct WarpBuf {
  struct A {
    int scal[8]; int hi[32]; float ht[32]; int hn[32]; float scr[Cfg<KIND>::SCRF];
    int inv[Cfg<KIND>::NINV];   // sorted rank of the sample's occurrence slots (KIND 2: slots L.., KIND 3: slots 0..)
  } a[4];
  struct Bs { int crl[32]; int crs[32]; int cc[4]; float ut[32]; } b[3];
  float rows[2][Cfg<KIND>::NROWS][64];
};

__device__ __forceinline__ void cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int KIND>
__device__ __forceinline__ void stage_a(const FArgs& a, int b, int lane, typename WarpBuf<KIND>::A& sa) {
  if (lane < 6) {
    const void* src = lane == 0 ? (const void*)(a.u + b) : lane == 1 ? (const void*)(a.sl + b)
                    : lane == 2 ? (const void*)(a.sl_new + b) : lane == 3 ? (const void*)(a.i + b)
                    : lane == 4 ? (const void*)(a.c + b) : (const void*)(a.y + b);
    cp4(&sa.scal[lane], src);
  }
  if (KIND != 2 && lane < a.L) {
    cp4(&sa.hi[lane], a.hist_i + (size_t)b * a.L + lane);
    cp4(&sa.ht[lane], a.hist_t + (size_t)b * a.L + lane);
  }
  if (KIND == 2) {
    if (lane < a.S) cp4(&sa.hn[lane], a.hist_i_new + (size_t)b * a.S + lane);
    if (lane < 16) cp16(&sa.scr[lane * 4], a.scratch + (size_t)b * (TLSAN_SCR * 64) + 320 + lane * 4);   // z
    if (lane < a.S + 2) cp4(&sa.inv[lane], a.inv + ((size_t)b << a.spsh) + a.L + lane);
  }
  if (KIND == 3) {   // do_long | o_long | max | 1/den : 256 contiguous floats
    const float* sc = a.scratch + (size_t)b * (TLSAN_SCR * 64);
    cp16(&sa.scr[lane * 4], sc + lane * 4);
    cp16(&sa.scr[128 + lane * 4], sc + 128 + lane * 4);
    if (lane < a.L) cp4(&sa.inv[lane], a.inv + ((size_t)b << a.spsh) + lane);
  }
}

template <int KIND>
__device__ __forceinline__ void stage_b(const FArgs& a, int lane, const typename WarpBuf<KIND>::A& sa,
                                        typename WarpBuf<KIND>::Bs& sb) {
  if (KIND != 2) {
    const int u = sa.scal[0], ell = sa.scal[1];
    if (lane < ell) {          // lane < 32 always: ids 32.. are fetched directly in the compute stage
      cp4(&sb.crl[lane], a.icl + sa.hi[lane]);
      cp4(&sb.ut[lane], a.usert + (size_t)u * a.L + lane);
    }
  } else {
    const int s = sa.scal[2];
    if (lane < s) cp4(&sb.crs[lane], a.icl + sa.hn[lane]);
    if (lane == 31) cp4(&sb.cc[0], a.icl + sa.scal[3]);
  }
}

template <int KIND>
__device__ __forceinline__ void stage_c(const FArgs& a, int lane, const typename WarpBuf<KIND>::A& sa,
                                        const typename WarpBuf<KIND>::Bs& sb, float (*rows)[64]) {
  const int hs = lane >> 4, c = lane & 15;   // 16 lanes x 16 B per token: chunks 0-7 item row, 8-15 cate row
  if (KIND != 2) {
    const int n = min(sa.scal[1], RL);
    for (int k = 0; k < n; k += 2) {
      const int t = k + hs;
      if (t < n) {
        const int row = c < 8 ? sa.hi[t] : a.NI + sb.crl[t];
        cp16(&rows[t][c * 4], a.emb + (size_t)row * 32 + (c & 7) * 4);
      }
    }
  } else {
    const int n = min(sa.scal[2], RS);
    for (int k = 0; k < n; k += 2) {
      const int t = k + hs;
      if (t < n) {
        const int row = c < 8 ? sa.hn[t] : a.NI + sb.crs[t];
        cp16(&rows[t][c * 4], a.emb + (size_t)row * 32 + (c & 7) * 4);
      }
    }
    // candidate e(i) -> rows[RS] ; user vector [user_emb[u] | cate_emb[u_cate]] -> rows[RS+1]
    const int row = hs == 0 ? (c < 8 ? sa.scal[3] : a.NI + sb.cc[0])
                            : (c < 8 ? a.NI + a.NC + sa.scal[0] : a.NI + sa.scal[4]);
    cp16(&rows[RS + hs][c * 4], a.emb + (size_t)row * 32 + (c & 7) * 4);
  }
}

// ---- accessors of the compute stage -------------------------------------------------------
// long-term token t (model.py:98-109): e = gathered slice (2 features of this lane), tau, P*hist_t, hist_t
template <int KIND>
struct LongTok {
  const FArgs& a; const LaneGeo& L; const typename WarpBuf<KIND>::A& sa; const typename WarpBuf<KIND>::Bs& sb;
  float (*rows)[64]; int b, u, ell; float gamma;
  __device__ __forceinline__ void meta(int t, int& id, int& crow, float& ht, float& pt) const {
    if (t < 32) {
      id = sa.hi[t]; crow = a.NI + sb.crl[t]; ht = sa.ht[t]; pt = sb.ut[t] * ht;
    } else {
      id = __ldg(a.hist_i + (size_t)b * a.L + t); crow = a.NI + __ldg(a.icl + id);
      ht = __ldg(a.hist_t + (size_t)b * a.L + t); pt = __ldg(a.usert + (size_t)u * a.L + t) * ht;
    }
  }
  __device__ __forceinline__ void get(int t, float2& e, float& tau, float& pt, float& ht) const {
    int id, crow;
    meta(t, id, crow, ht, pt);
    tau = gamma * pt;
    e = t < RL ? *reinterpret_cast<const float2*>(&rows[t][L.f0]) : ldg2(row_ptr(a, L, id, crow));
  }
};

template <int KIND>
__global__ void __launch_bounds__(MMA_THREADS, Cfg<KIND>::CTAS) k_async(const FArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using WB = WarpBuf<KIND>;
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  WB& wb = reinterpret_cast<WB*>(smem_raw)[warp];
  const float gamma = a.dense[TLSAN_OFF_GAMMA];
  const int nwarps = gridDim.x * MMA_WARPS;
  const int w0 = blockIdx.x * MMA_WARPS + warp;
  const int nmine = w0 < a.B ? (a.B - w0 + nwarps - 1) / nwarps : 0;

  FwaW w;
  FwaWT wt;
  FwaGrad G;
  float acc1 = 0.f, sq_acc = 0.f;   // KIND 2: loss sum ; KIND 3: d gamma
  w = load_fwa(a.dense, KIND == 2 ? TLSAN_OFF_W1S : TLSAN_OFF_W1L, L.g, L.t);
  if (KIND != 1) { wt = load_fwa_t(a.dense, KIND == 2 ? TLSAN_OFF_W1S : TLSAN_OFF_W1L, L.g, L.t); G.init(); }

  // ---- pipeline prologue
  for (int i = 0; i < 3; ++i)
    if (i < nmine) stage_a<KIND>(a, w0 + i * nwarps, L.lane, wb.a[i]);
  cp_wait_all(); __syncwarp();
  for (int i = 0; i < 2; ++i)
    if (i < nmine) stage_b<KIND>(a, L.lane, wb.a[i], wb.b[i]);
  cp_wait_all(); __syncwarp();
  if (0 < nmine) stage_c<KIND>(a, L.lane, wb.a[0], wb.b[0], wb.rows[0]);

  for (int i = 0; i < nmine; ++i) {
    cp_wait_all();
    __syncwarp();
    if (i + 3 < nmine) stage_a<KIND>(a, w0 + (i + 3) * nwarps, L.lane, wb.a[(i + 3) & 3]);
    if (i + 2 < nmine) stage_b<KIND>(a, L.lane, wb.a[(i + 2) & 3], wb.b[(i + 2) % 3]);
    if (i + 1 < nmine) stage_c<KIND>(a, L.lane, wb.a[(i + 1) & 3], wb.b[(i + 1) % 3], wb.rows[(i + 1) & 1]);

    const int b = w0 + i * nwarps;
    const typename WB::A& sa = wb.a[i & 3];
    const typename WB::Bs& sb = wb.b[i % 3];
    float (*rows)[64] = wb.rows[i & 1];
    const int u = sa.scal[0];

    if (KIND == 1) {
      // ================= long-term FWA forward (model.py:98-109, 334-345)
      const int ell = sa.scal[1];
      const LongTok<KIND> tok{a, L, sa, sb, rows, b, u, ell, gamma};
      Soft2 st; st.init();
      for (int j = 0; j < ell; j += 2) {
        const bool okB = j + 1 < ell;
        float2 eA, eB = make_float2(0.f, 0.f); float tA, tB = 0.f, pt, ht;
        tok.get(j, eA, tA, pt, ht);
        if (okB) tok.get(j + 1, eB, tB, pt, ht);
        const float x[4] = {eA.x * tA, eA.y * tA, eB.x * tB, eB.y * tB};
        float m1[4], m2[4];
        tile_maps(x, w, m1, m2);
        st.push(m2[0], m2[1], x[0], x[1]);
        if (okB) st.push(m2[2], m2[3], x[2], x[3]);
      }
      float* sc = a.scratch + (size_t)b * (TLSAN_SCR * 64) + L.f0;
      const float i0 = st.den[0] > 0.f ? 1.f / st.den[0] : 0.f, i1 = st.den[1] > 0.f ? 1.f / st.den[1] : 0.f;
      st2(sc + 64, st.acc[0] * i0, st.acc[1] * i1);
      st2(sc + 128, st.mx[0], st.mx[1]);
      st2(sc + 192, i0, i1);
    }

    if (KIND == 2) {
      // ================= short-term FWA forward over [z ; e(hist_i_new)] (model.py:350-364)
      const int s = sa.scal[2], cand = sa.scal[3];
      const int ntok = s + 1;
      const float2 zz = *reinterpret_cast<const float2*>(&sa.scr[L.f0]);
      const float z[2] = {zz.x, zz.y};
      // token n >= 1 is short item n-1: staged row (n-1 < RS), else direct gather
      auto item_x = [&](int it) -> float2 {
        if (it < RS) return *reinterpret_cast<const float2*>(&rows[it][L.f0]);
        int id, cr;
        if (it < 32) { id = sa.hn[it]; cr = a.NI + sb.crs[it]; }
        else { id = __ldg(a.hist_i_new + (size_t)b * a.S + it); cr = a.NI + __ldg(a.icl + id); }
        return ldg2(row_ptr(a, L, id, cr));
      };
      Soft2 ss; ss.init();
      for (int n = 0; n < ntok; n += 2) {
        const bool okB = n + 1 < ntok;
        float x[4];
        if (n == 0) { x[0] = z[0]; x[1] = z[1]; } else { const float2 e = item_x(n - 1); x[0] = e.x; x[1] = e.y; }
        if (okB) { const float2 e = item_x(n); x[2] = e.x; x[3] = e.y; } else { x[2] = 0.f; x[3] = 0.f; }
        float m1[4], m2[4];
        tile_maps(x, w, m1, m2);
        ss.push(m2[0], m2[1], x[0], x[1]);
        if (okB) ss.push(m2[2], m2[3], x[2], x[3]);
      }
      const float inv_s[2] = {1.f / ss.den[0], 1.f / ss.den[1]};
      const float v[2] = {ss.acc[0] * inv_s[0], ss.acc[1] * inv_s[1]};
      // ---- user vector, candidate, logit (model.py:84-95,135-137)
      const float2 q = *reinterpret_cast<const float2*>(&rows[RS][L.f0]);
      const float2 p = *reinterpret_cast<const float2*>(&rows[RS + 1][L.f0]);
      const float ut[2] = {v[0] + p.x, v[1] + p.y};
      const float logit = warp_sum_f(fmaf(ut[0], q.x, ut[1] * q.y)) + __ldg(a.item_b + cand);
      // ---- sigmoid cross entropy (model.py:171) and its gradient through reduce_mean
      const float yb = __int_as_float(sa.scal[5]);
      const float ex = expf(-fabsf(logit));
      const float bce = fmaxf(logit, 0.f) - logit * yb + log1pf(ex);
      const float sig = logit >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      const float gl = (sig - yb) * a.invB;
      if (L.lane == 0) { acc1 += bce; sq_acc = fmaf(gl, gl, sq_acc); a.gscal[b] = gl; }
      // gradient-row destinations: slot L + k has its sorted rank staged in sa.inv[k] (k < 32)
      auto slot_row = [&](int k) -> float* {
        const int pos = k < 32 ? sa.inv[k] : __ldg(a.inv + ((size_t)b << a.spsh) + a.L + k);
        return a.rows_i + (size_t)pos * 64 + L.f0;
      };
      float* rcand = slot_row(a.S);
      float* rvirt = slot_row(a.S + 1);
      const float dq[2] = {gl * ut[0], gl * ut[1]};
      const float du[2] = {gl * q.x, gl * q.y};
      sq_acc = fmaf(dq[0], dq[0], sq_acc); sq_acc = fmaf(dq[1], dq[1], sq_acc);
      sq_acc = fmaf(du[0], du[0], sq_acc); sq_acc = fmaf(du[1], du[1], sq_acc);
      st2(rcand, dq[0], dq[1]);                                           // -> item_emb[i] | cate_emb[icl[i]]
      if (L.half) st2(rvirt, du[0], du[1]);                               // -> cate_emb[u_cate]
      else { st2(rvirt, 0.f, 0.f); st2(a.rows_u + (size_t)b * a.PU + L.f0, du[0], du[1]); }  // -> user_emb[u]
      // ---- short-term FWA backward, d v = du
      float dz[2] = {0.f, 0.f};
      for (int n = 0; n < ntok; n += 2) {
        const bool okB = n + 1 < ntok;
        float x[4], dx[4];
        if (n == 0) { x[0] = z[0]; x[1] = z[1]; } else { const float2 e = item_x(n - 1); x[0] = e.x; x[1] = e.y; }
        if (okB) { const float2 e = item_x(n); x[2] = e.x; x[3] = e.y; } else { x[2] = 0.f; x[3] = 0.f; }
        tile_bwd(x, okB, v, du, ss.mx, inv_s, w, wt, L.lane, dx, G);
        if (n == 0) { dz[0] = dx[0]; dz[1] = dx[1]; }
        else {
          sq_acc = fmaf(dx[0], dx[0], sq_acc); sq_acc = fmaf(dx[1], dx[1], sq_acc);
          st2(slot_row(n - 1), dx[0], dx[1]);
        }
        if (okB) {
          sq_acc = fmaf(dx[2], dx[2], sq_acc); sq_acc = fmaf(dx[3], dx[3], sq_acc);
          st2(slot_row(n), dx[2], dx[3]);
        }
      }
      st2(a.scratch + (size_t)b * (TLSAN_SCR * 64) + 256 + L.f0, dz[0], dz[1]);   // -> k_dense_bwd_mma
    }

    if (KIND == 3) {
      // ================= long-term FWA backward + time-aware position term
      const int ell = sa.scal[1];
      const LongTok<KIND> tok{a, L, sa, sb, rows, b, u, ell, gamma};
      const float2 dol2 = *reinterpret_cast<const float2*>(&sa.scr[L.f0]);
      const float2 o2 = *reinterpret_cast<const float2*>(&sa.scr[64 + L.f0]);
      const float2 mx2 = *reinterpret_cast<const float2*>(&sa.scr[128 + L.f0]);
      const float2 inv2 = *reinterpret_cast<const float2*>(&sa.scr[192 + L.f0]);
      const float dol[2] = {dol2.x, dol2.y}, o[2] = {o2.x, o2.y}, mx[2] = {mx2.x, mx2.y}, inv[2] = {inv2.x, inv2.y};
      float* ru = a.rows_u + (size_t)b * a.PU + 32;
      float dtau_l = 0.f, pt_l = 0.f, ht_l = 0.f;     // lane (t & 31) owns token t of the current 32-block
      for (int j = 0; j < ell; j += 2) {
        const bool okB = j + 1 < ell;
        float2 eA, eB = make_float2(0.f, 0.f); float tA, tB = 0.f, ptA, htA, ptB = 0.f, htB = 0.f;
        tok.get(j, eA, tA, ptA, htA);
        if (okB) tok.get(j + 1, eB, tB, ptB, htB);
        const float x[4] = {eA.x * tA, eA.y * tA, eB.x * tB, eB.y * tB};
        float dx[4];
        tile_bwd(x, okB, o, dol, mx, inv, w, wt, L.lane, dx, G);
        // gradient of the gathered slices (tau * dX) and of tau (<dX, e>)
        const float rA0 = dx[0] * tA, rA1 = dx[1] * tA;
        sq_acc = fmaf(rA0, rA0, sq_acc); sq_acc = fmaf(rA1, rA1, sq_acc);
        st2(a.rows_i + (size_t)(j < 32 ? sa.inv[j] : __ldg(a.inv + ((size_t)b << a.spsh) + j)) * 64 + L.f0, rA0, rA1);
        const float dtA = warp_sum_f(fmaf(dx[0], eA.x, dx[1] * eA.y));
        if (L.lane == (j & 31)) { dtau_l = dtA; pt_l = ptA; ht_l = htA; }
        if (okB) {
          const float rB0 = dx[2] * tB, rB1 = dx[3] * tB;
          sq_acc = fmaf(rB0, rB0, sq_acc); sq_acc = fmaf(rB1, rB1, sq_acc);
          st2(a.rows_i + (size_t)(j + 1 < 32 ? sa.inv[j + 1] : __ldg(a.inv + ((size_t)b << a.spsh) + j + 1)) * 64 + L.f0, rB0, rB1);
          const float dtB = warp_sum_f(fmaf(dx[2], eB.x, dx[3] * eB.y));
          if (L.lane == ((j + 1) & 31)) { dtau_l = dtB; pt_l = ptB; ht_l = htB; }
        }
        if (((j + 2) & 31) == 0 || j + 2 >= ell) {   // a 32-token block (or the sequence) is complete
          const int t0 = j & ~31;
          if (t0 + L.lane < ell) {
            acc1 = fmaf(dtau_l, pt_l, acc1);                 // d gamma
            const float dp = dtau_l * gamma * ht_l;          // d usert_emb[u, t]
            sq_acc = fmaf(dp, dp, sq_acc);
            ru[t0 + L.lane] = dp;
          }
        }
      }
      for (int tt = ell + L.lane; tt < a.PU - 32; tt += 32) ru[tt] = 0.f;
    }
  }

  if (KIND != 1) {
    // ---- per-CTA partial sums, fixed order: butterfly over g -> warps 0..7 -> global
    cp_wait_all();
    __syncthreads();                                   // ring buffers are dead: reuse as staging
    float (*red)[160] = reinterpret_cast<float (*)[160]>(smem_raw);
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
      const float r1 = warp_sum_f(acc1), r2 = warp_sum_f(sq_acc);
      if (L.lane == 0) { red[warp][144] = r1; red[warp][145] = r2; }
    }
    __syncthreads();
    if (threadIdx.x < 146) {
      float r = 0.f;
#pragma unroll
      for (int wv = 0; wv < MMA_WARPS; ++wv) r += red[wv][threadIdx.x];
      const int base = KIND == 2 ? TLSAN_OFF_W1S : TLSAN_OFF_W1L;
      const int extra = KIND == 2 ? TLSAN_PART_LOSS : TLSAN_OFF_GAMMA;
      const int dst = threadIdx.x < 144 ? base + threadIdx.x : (threadIdx.x == 144 ? extra : TLSAN_PART_SUMSQ);
      a.part[(size_t)blockIdx.x * TLSAN_PART + dst] = r;
    }
  }
}

// ------------------------------------------------------------------ launcher
FArgs tlsan_make_fargs(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b);
int tlsan_launch_dense_fwd(const float* dense, float* scratch, int B, cudaStream_t st);
int tlsan_launch_dense_bwd(const float* dense, float* scratch, int B, float* part, int* grid_c, cudaStream_t st);

template <int KIND>
static int launch_async(const FArgs& a, int* grid_out, cudaStream_t st) {
  static bool attr_set = false;
  const int smem = (int)sizeof(WarpBuf<KIND>) * MMA_WARPS;
  if (!attr_set) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_async<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int need = (a.B + MMA_WARPS - 1) / MMA_WARPS;
  const int cap = tlsan_num_sms() * Cfg<KIND>::CTAS;
  const int g = need < cap ? need : cap;
  if (grid_out) *grid_out = g;
  k_async<KIND><<<g, MMA_THREADS, smem, st>>>(a);
  TLSAN_CHECK_LAUNCH("k_async");
  return TLSAN_OK;
}

int tlsan_launch_long_fwd_mma(const FArgs& a, int ctas_per_sm, cudaStream_t st);
int tlsan_overlap_ctas();
int tlsan_launch_bwd_long_mma(const FArgs& a, int* grid_b, cudaStream_t st);
int tlsan_launch_long_meta(const FArgs& a, void* meta, void* smeta, void* sscal, void* part, int fwd_ctas,
                           int score_ncand, cudaStream_t st);                                                     // tlsan_fused_pf.cu
int tlsan_launch_short_pf(const FArgs& a, const void* smeta, const void* sscal, const void* part, int* grid_a,
                          cudaStream_t st);
int tlsan_launch_partition(const FArgs& a, int fwd_ctas, bool train, void* part, cudaStream_t st);
int tlsan_launch_long_fwd_pf(const FArgs& a, const void* meta, const void* part, int ctas_per_sm, cudaStream_t st);
int tlsan_launch_bwd_long_pf(const FArgs& a, const void* meta, const void* part, int* grid_b, cudaStream_t st);

// `hybrid` (default): per kernel, whichever formulation measured faster on B200 (profiles/):
// synchronous gathers for the two long-term kernels (24 / 16 resident warps already hide the
// latency; the async pipeline only adds instructions there), cp.async pipeline for the short-term
// kernel (its per-sample dependent chain is otherwise exposed: 130 -> 113 us).
// `pf` (default since round 2): the long-term kernels are those of tlsan_fused_pf.cu (metadata pre-pass + in-warp
// prefetch pipeline); forward and backward must be the same variant (the saved softmax maximum is in the log2
// domain there).  variant: 0 = async everywhere, 1 = hybrid, 2 = pf.
int tlsan_launch_fwd_bwd_async(const tlsan_dims_t& d, const tlsan_params_t& p, const tlsan_batch_t& b,
                               const TlsanWs& w, char* ws, int* grid_a, int* grid_b, int* grid_c, int variant,
                               cudaEvent_t sorted, cudaEvent_t part_ready, bool part_early, int long_ctas,
                               cudaStream_t st) {
  const bool hybrid = variant == 1;
  FArgs a = tlsan_make_fargs(d, p, b);
  a.rows_i = reinterpret_cast<float*>(ws + w.rows_i);
  a.inv = reinterpret_cast<const int*>(ws + w.inv); a.spsh = w.SPSH;
  a.rows_u = reinterpret_cast<float*>(ws + w.rows_u);
  a.gscal = reinterpret_cast<float*>(ws + w.gscal);
  a.scratch = reinterpret_cast<float*>(ws + w.scratch);
  int rc;
  void* meta = ws + w.meta;
  void* smeta = ws + w.smeta;
  void* sscal = ws + w.sscal;
  void* part = ws + w.part;
  // the balanced partition (short-term and backward kernels; the long-term forward claims its samples dynamically)
  // depends on the batch only: the caller computed it behind the sort (part_ready), or asks for it here (no side stream)
  (void)part_early;
  if (variant == 2 && !part_ready && (rc = tlsan_launch_partition(a, long_ctas, true, part, st))) return rc;
  if (variant == 2 && (rc = tlsan_launch_long_meta(a, meta, smeta, sscal, part, long_ctas, 0, st))) return rc;
  if ((rc = variant == 2 ? tlsan_launch_long_fwd_pf(a, meta, part, long_ctas, st)
                         : hybrid ? tlsan_launch_long_fwd_mma(a, long_ctas, st) : launch_async<1>(a, nullptr, st)))
    return rc;
  tlsan_profile_mark(TLSAN_PHASE_LONG_FWD, st);
  if ((rc = tlsan_launch_dense_fwd(p.dense, a.scratch, d.B, st))) return rc;
  a.part = reinterpret_cast<float*>(ws + w.part_a);
  if (sorted) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, sorted, 0));   // gradient rows are written at sorted rank
  if (variant == 2 && part_ready) TLSAN_CHECK_CUDA(cudaStreamWaitEvent(st, part_ready, 0));
  tlsan_profile_mark(TLSAN_PHASE_DENSE_FWD, st);                       // (phase includes the join with the sort stream)
  if ((rc = variant == 2 ? tlsan_launch_short_pf(a, smeta, sscal, part, grid_a, st) : launch_async<2>(a, grid_a, st))) return rc;
  tlsan_profile_mark(TLSAN_PHASE_SHORT, st);
  if ((rc = tlsan_launch_dense_bwd(p.dense, a.scratch, d.B, reinterpret_cast<float*>(ws + w.part_c), grid_c, st)))
    return rc;
  tlsan_profile_mark(TLSAN_PHASE_DENSE_BWD, st);
  a.part = reinterpret_cast<float*>(ws + w.part_b);
  if ((rc = variant == 2 ? tlsan_launch_bwd_long_pf(a, meta, part, grid_b, st)
                         : hybrid ? tlsan_launch_bwd_long_mma(a, grid_b, st) : launch_async<3>(a, grid_b, st)))
    return rc;
  tlsan_profile_mark(TLSAN_PHASE_BWD_LONG, st);
  return TLSAN_OK;
}
