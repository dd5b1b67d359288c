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
  // outputs
  float* logits;   // score: [B][ncand]
  float* ut;       // score: optional [B][64]
  float* rows_i;   // train: [B*SI][64] per-occurrence gradient rows (item half | cate half)
  float* rows_u;   // train: [B][PU]   user_emb grad (32) | usert_emb grad (L)
  float* gscal;    // train: [B] d loss / d logit  (item_b gradient per occurrence)
  float* scratch;  // train: [B][TLSAN_SCR][64]  do_long | o_long | max | 1/denominator | dz
  float* part;     // train: [grid][TLSAN_PART] per-CTA partial sums
};

