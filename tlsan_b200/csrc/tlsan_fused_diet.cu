// EXPERIMENTAL, DEFAULT OFF (TLSAN_BWD_LONG=diet): written at the end of round 1 after the GPU budget was spent --
// it compiles (80 registers, 4 B spill under __launch_bounds__(256, 3)) but has NEVER RUN.  Round 2: run the parity
// tests with TLSAN_BWD_LONG=diet (tests/test_gpu_experimental.py), then time it against k_bwd_long_mma.
//
// Register diet for the long-term FWA backward (the dominant kernel, latency-bound at 16 warps / SM):
//   * weight fragments (W1, W2, W2^T, W1^T as tf32 hi/lo B fragments: 16 words per lane position) live in shared
//     memory and are re-read per product instead of occupying 16+ registers per lane;
//   * the weight gradients dW1 = x^T dpre, dW2 = m1^T dm2 go through a per-warp shared-memory transpose and
//     mma.sync tiles (A = dpre^T / dm2^T, K = the 16 token-head rows of the tile): 8 accumulator registers
//     instead of the 32 FFMA2 accumulators + 32 quad shuffles of the shipped kernel.
// Question answered here: does the tile then fit 85 registers (three CTAs of 256 threads per SM)?
#include <stdlib.h>
#include <string.h>
#include "tlsan_mma_common.cuh"

struct WSm { uint32_t frag[4][32][4]; float b1[8], b2[8]; };   // [W1, W2, W2T, W1T][lane][h0 h1 l0 l1]

__device__ __forceinline__ BMat ldw(const WSm& w, int which, int lane) {
  const uint4 v = *reinterpret_cast<const uint4*>(w.frag[which][lane]);
  BMat m; m.h0 = v.x; m.h1 = v.y; m.l0 = v.z; m.l1 = v.w; return m;
}

// per-warp transpose area: 4 arrays x 16 rows x 8 (+1 pad)
struct TSm { float a[4][16][9]; };

__device__ __forceinline__ void acc_dw(float (&acc)[4], const float (*A)[9], const float (*Bm)[9], int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const float a0 = A[8 * ks + t][g], a2 = A[8 * ks + t + 4][g];
    const float b0 = Bm[8 * ks + t][g], b1 = Bm[8 * ks + t + 4][g];
    const uint32_t ah0 = to_tf32(a0), ah2 = to_tf32(a2), bh0 = to_tf32(b0), bh1 = to_tf32(b1);
    const uint32_t al0 = __float_as_uint(a0 - __uint_as_float(ah0)), al2 = __float_as_uint(a2 - __uint_as_float(ah2));
    const uint32_t bl0 = __float_as_uint(b0 - __uint_as_float(bh0)), bl1 = __float_as_uint(b1 - __uint_as_float(bh1));
    mma_tf32(acc, al0, 0u, al2, 0u, bh0, bh1);
    mma_tf32(acc, ah0, 0u, ah2, 0u, bl0, bl1);
    mma_tf32(acc, ah0, 0u, ah2, 0u, bh0, bh1);
  }
}

__device__ __forceinline__ void tile_bwd_diet(const float (&x)[4], bool okB, const float (&o)[2], const float (&dout)[2],
                                              const float (&mx)[2], const float (&inv)[2], const WSm& w, TSm& ts,
                                              int lane, int g, int t, float (&dx)[4], float (&aw1)[4], float (&aw2)[4],
                                              float (&gb)[4]) {
  float m1[4], m2[4];
  {
    const float b1a = w.b1[2 * t], b1b = w.b1[2 * t + 1];
    m1[0] = b1a; m1[1] = b1b; m1[2] = b1a; m1[3] = b1b;
    mma3(m1, x, ldw(w, 0, lane));
#pragma unroll
    for (int i = 0; i < 4; ++i) m1[i] = fmaxf(m1[i], 0.f);
    const float b2a = w.b2[2 * t], b2b = w.b2[2 * t + 1];
    m2[0] = b2a; m2[1] = b2b; m2[2] = b2a; m2[3] = b2b;
    mma3(m2, m1, ldw(w, 1, lane));
  }
  const float kf[2] = {inv[0] * dout[0], inv[1] * dout[1]};
  const float nmx[2] = {-mx[0] * 1.4426950408889634f, -mx[1] * 1.4426950408889634f};
  float ado[4], dm2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = i & 1;
    float ee;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ee) : "f"(fmaf(m2[i], 1.4426950408889634f, nmx[j])));
    const float a = ee * kf[j];
    ado[i] = (i < 2 || okB) ? a : 0.f;
    dm2[i] = ado[i] * (x[i] - o[j]);
  }
  gb[2] += dm2[0] + dm2[2]; gb[3] += dm2[1] + dm2[3];
  float dpre[4] = {0.f, 0.f, 0.f, 0.f};
  mma3(dpre, dm2, ldw(w, 2, lane));
#pragma unroll
  for (int i = 0; i < 4; ++i) dpre[i] = m1[i] > 0.f ? dpre[i] : 0.f;
  gb[0] += dpre[0] + dpre[2]; gb[1] += dpre[1] + dpre[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) dx[i] = ado[i];
  mma3(dx, dpre, ldw(w, 3, lane));
  // transposes through shared memory: rows g (token A) and g + 8 (token B), columns 2t, 2t+1
  __syncwarp();
  ts.a[0][g][2 * t] = x[0]; ts.a[0][g][2 * t + 1] = x[1]; ts.a[0][g + 8][2 * t] = x[2]; ts.a[0][g + 8][2 * t + 1] = x[3];
  ts.a[1][g][2 * t] = m1[0]; ts.a[1][g][2 * t + 1] = m1[1]; ts.a[1][g + 8][2 * t] = m1[2]; ts.a[1][g + 8][2 * t + 1] = m1[3];
  ts.a[2][g][2 * t] = dm2[0]; ts.a[2][g][2 * t + 1] = dm2[1]; ts.a[2][g + 8][2 * t] = dm2[2]; ts.a[2][g + 8][2 * t + 1] = dm2[3];
  ts.a[3][g][2 * t] = dpre[0]; ts.a[3][g][2 * t + 1] = dpre[1]; ts.a[3][g + 8][2 * t] = dpre[2]; ts.a[3][g + 8][2 * t + 1] = dpre[3];
  __syncwarp();
  acc_dw(aw1, ts.a[3], ts.a[0], g, t);   // dW1^T[j][k] += sum_rows dpre[row][j] x[row][k]
  acc_dw(aw2, ts.a[2], ts.a[1], g, t);   // dW2^T[j][k] += sum_rows dm2[row][j] m1[row][k]
}

__global__ void __launch_bounds__(MMA_THREADS, 3) k_bwd_long_diet(const FArgs a) {
  extern __shared__ __align__(16) unsigned char smem_b[];
  WSm& wsm = *reinterpret_cast<WSm*>(smem_b);
  TSm* tsm = reinterpret_cast<TSm*>(smem_b + sizeof(WSm));
  float (*rows)[64] = reinterpret_cast<float (*)[64]>(smem_b + sizeof(WSm) + MMA_WARPS * sizeof(TSm));
  LaneGeo L; L.init();
  const int warp = threadIdx.x >> 5;
  float (*rowsw)[64] = rows + (size_t)warp * 16;
  if (threadIdx.x < 32) {
    const int g = threadIdx.x >> 2, t = threadIdx.x & 3;
    const BMat w1 = load_b(a.dense + TLSAN_OFF_W1L, g, t), w2 = load_b(a.dense + TLSAN_OFF_W1L + 72, g, t);
    const BMat w2t = load_bt(a.dense + TLSAN_OFF_W1L + 72, g, t), w1t = load_bt(a.dense + TLSAN_OFF_W1L, g, t);
    const BMat all[4] = {w1, w2, w2t, w1t};
    for (int k = 0; k < 4; ++k) { wsm.frag[k][threadIdx.x][0] = all[k].h0; wsm.frag[k][threadIdx.x][1] = all[k].h1; wsm.frag[k][threadIdx.x][2] = all[k].l0; wsm.frag[k][threadIdx.x][3] = all[k].l1; }
    if (threadIdx.x < 8) { wsm.b1[threadIdx.x] = a.dense[TLSAN_OFF_W1L + 64 + threadIdx.x]; wsm.b2[threadIdx.x] = a.dense[TLSAN_OFF_W1L + 136 + threadIdx.x]; }
  }
  __syncthreads();
  const float gamma = a.dense[TLSAN_OFF_GAMMA];
  float aw1[4] = {0.f, 0.f, 0.f, 0.f}, aw2[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
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
    for (int r0 = 0; r0 < ell; r0 += 16) {
      const LongMeta me = load_long_meta(a, b, u, r0 + L.lane, min(ell, r0 + 16), gamma);
      const int cnt = min(16, ell - r0);
      float dtau_l = 0.f;
      const int inv_l = L.lane < cnt ? __ldg(a.inv + ((size_t)b << a.spsh) + r0 + L.lane) : 0;
      stage_round_rows(a, me, cnt, L.lane, rowsw);
      for (int j = 0; j < cnt; j += 2) {
        const bool okB = j + 1 < cnt;
        const float tA = __shfl_sync(0xffffffffu, me.tau, j), tB = __shfl_sync(0xffffffffu, me.tau, (j + 1) & 31);
        const float2 eA = *reinterpret_cast<const float2*>(&rowsw[j][L.f0]);
        const float2 eB = okB ? *reinterpret_cast<const float2*>(&rowsw[j + 1][L.f0]) : make_float2(0.f, 0.f);
        const int posA = __shfl_sync(0xffffffffu, inv_l, j), posB = __shfl_sync(0xffffffffu, inv_l, (j + 1) & 31);
        const float x[4] = {eA.x * tA, eA.y * tA, okB ? eB.x * tB : 0.f, okB ? eB.y * tB : 0.f};
        float dx[4];
        tile_bwd_diet(x, okB, o, dol, mx, inv, wsm, tsm[warp], L.lane, L.g, L.t, dx, aw1, aw2, gb);
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
        const float dp = dtau_l * gamma * me.ht;
        sq_acc = fmaf(dp, dp, sq_acc);
        ru[r0 + L.lane] = dp;
      }
    }
    for (int tt = ell + L.lane; tt < a.PU - 32; tt += 32) ru[tt] = 0.f;
  }
  __syncthreads();   // every warp is done with its staged rows: the area is reused for the reduction
  float (*red)[160] = reinterpret_cast<float (*)[160]>(rows);
  // aw?[0..1] = dW?^T[j = g][k = 2t, 2t+1]: the 32 lanes of a warp hold the whole 8x8 matrix, no butterfly needed
  red[warp][(2 * L.t) * 8 + L.g] = aw1[0]; red[warp][(2 * L.t + 1) * 8 + L.g] = aw1[1];
  red[warp][72 + (2 * L.t) * 8 + L.g] = aw2[0]; red[warp][72 + (2 * L.t + 1) * 8 + L.g] = aw2[1];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {   // biases: lanes with equal t hold the same columns
    float r1 = gb[jj], r2 = gb[2 + jj];
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

static const int kSmemDiet = (int)(sizeof(WSm) + MMA_WARPS * sizeof(TSm) + sizeof(float) * MMA_WARPS * 16 * 64);

bool tlsan_bwd_long_diet_selected() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TLSAN_BWD_LONG"); v = (e && strcmp(e, "diet") == 0) ? 1 : 0; }
  return v == 1;
}

int tlsan_launch_bwd_long_diet(const FArgs& a, int* grid_b, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    TLSAN_CHECK_CUDA(cudaFuncSetAttribute(k_bwd_long_diet, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDiet));
    attr = true;
  }
  const int need = (a.B + MMA_WARPS - 1) / MMA_WARPS, cap = tlsan_num_sms() * 3;
  const int g = need < cap ? need : cap;
  *grid_b = g;
  k_bwd_long_diet<<<g, MMA_THREADS, kSmemDiet, st>>>(a);
  TLSAN_CHECK_LAUNCH("k_bwd_long_diet");
  return TLSAN_OK;
}
