// Argument block shared by the fused forward / backward kernels (internal).
#pragma once
#include "tlsan_common.cuh"

struct FArgs {
  int B, L, S, NI, NC, NU, SI, PU;
  float invB;
  const float* emb;
  const float* usert;
  const float* item_b;
  const float* dense;
  const int* icl;
  const int *u, *i, *i2, *c, *sl, *sl_new, *hist_i, *hist_i_new;
  const float *y, *hist_t;
  const int* hist_d;   // raw day gaps (tlsan_batch_t.hist_d) or NULL
  // outputs
  float* logits;   // score: [B][ncand]
  float* ut;       // score: optional [B][64]
  float* rows_i;   // train: [nvalid][64] per-occurrence gradient rows (item half | cate half), SORTED order
  const int* inv;  // train: [B << spsh] occurrence id (b << spsh | slot) -> rank in the sorted order
  int spsh;
  float* rows_u;   // train: [B][PU]   user_emb grad (32) | usert_emb grad (L)
  float* gscal;    // train: [B] d loss / d logit  (item_b gradient per occurrence)
  float* scratch;  // train: [B][TLSAN_SCR][64]  do_long | o_long | max | 1/denominator | dz
  float* part;     // train: [grid][TLSAN_PART] per-CTA partial sums
};


// Destination of the gradient row of occurrence slot j of sample b.  Rows are written directly at
// their rank in the (stable) sorted order, so the segmented reduce streams contiguous memory.
__device__ __forceinline__ float* grad_row(const FArgs& a, int b, int j) {
  return a.rows_i + (size_t)__ldg(a.inv + ((size_t)b << a.spsh) + j) * 64;
}
