// Persistent decode-step kernel (dstep.cu): argument block shared with engine.cu.
#pragma once
#include "decode.h"

namespace b2w {

struct DLayer {
  const __half *wqkv, *wo, *wq_x, *wo_x, *w1, *w2;
  const float *bqkv, *bo, *bq_x, *bo_x, *b1, *b2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
};

struct DStepArgs {
  const DLayer* layers;  // device array [L]
  int L;
  const __half* tok_emb;
  const __half* logit_w;
  const float* logit_b;
  const float* pos_emb;
  const float* lnf_g;
  const float* lnf_b;
  int R, d, H, n_ctx, slots, T, vpad, n_vocab, n_chunks, rows_per_chunk, xsplits;
  const RowInfo* rows;
  const int* tokens_in;
  float* x;        // [8][d] fp32 residual stream
  __half* q;       // [8][d]
  __half* ao;      // [8][d]
  __half* h;       // [8][4d]
  float* logits;   // [8][vpad]
  __half* kcache;
  __half* vcache;
  long long kv_layer_stride;
  const uint8_t* anc;
  long long anc_buf_stride;
  const DecBindings* bind;
  float* xpart;
  int* xcounters;
  unsigned* bar;
  int smem_xs_bytes;  // filled by dstep_launch
  unsigned long long* prof;  // optional: %globaltimer at every barrier exit (CTA 0), [1 + 8 L + 1]
};

size_t dstep_smem_bytes(const DStepArgs& a, size_t* xs_bytes);
void dstep_configure();
int dstep_max_grid(int num_sms, size_t smem);
void dstep_launch(DStepArgs a, int grid, cudaStream_t s);

}  // namespace b2w
