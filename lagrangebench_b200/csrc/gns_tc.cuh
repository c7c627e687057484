// Interface between gns.cu (forward orchestration) and gns_tc.cu (tcgen05 edge kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb {

struct EdgeTcArgs {
  int n;
  float inv_latent;  // 1 / (the model's true latent width <= 128): LayerNorm's divisor (columns beyond it are zero padding)
  const int32_t *rowptr, *snd, *rcv;
  const float* P;       // [n][256] per-node projections (sender half | receiver half + b1)
  const void* w_tc;     // W1e^T hi, lo, W2c^T hi, lo: four 128x128 fp16 operands in UMMA K-major layout
  const float* vec_tc;  // b2c[128] | ln_scale[128] | ln_offset[128]
  float *e, *agg, *carry_first, *carry_last;
  // encoder mode (gns.py:65-81, edge MLP): e = LN(relu(feat W0 + b0) W1c + b1c); no gather, no
  // residual, no aggregation.  w_tc then holds only W1c^T hi|lo, vec_tc = b1c | scale | offset,
  // enc_vec = W0[4][128] | b0[128], edge_feat in LIST order addressed through perm.
  int encoder;
  const float4* edge_feat;
  const int32_t* perm;
  const float* enc_vec;
  int pdl;      // launched with programmatic stream serialization: constants first, then griddepcontrol.wait;
                // otherwise the first tile's loads are issued before the weights are loaded
  int reverse;  // 1: tiles are visited from the end of the edge array (consecutive launches alternate, gns_tc2.cu)
  int l2_mode;  // 1: the residual read and the store of e carry an L2 evict-first hint (gns_tc2.cu)
};

int launch_edge_mp_tc(const EdgeTcArgs& a, int e_cap, cudaStream_t s);
// v2 (gns_tc2.cu): weights resident in TMEM, two-tile software pipeline per worker; processor steps only
int launch_edge_mp_tc2(const EdgeTcArgs& a, int e_cap, cudaStream_t s);

struct NodeTcArgs {
  int n;     // owned nodes
  int last;  // last message-passing step: decoder instead of the next step's projections
  int enc;   // node ENCODER (gns.py:65-81): h holds the zero-padded input features, no aggregate, no
             // residual; w_tc streams 4 operands W0pad^T, W1c^T, next W1s^T, next W1r^T (hi|lo each)
  int dim;
  const int32_t* rowptr;
  const float *agg, *carry_first, *carry_last;
  const void* w_tc;     // 5 (4 when last) streamed operands, each hi|lo = 64 KB, UMMA K-major layout
  const float* vec_tc;  // b1 | b2c | ln_scale | ln_offset | b_next | wd1[128][3] | bd1[4]
  float *h, *P, *out;
  // decomposed cloud: the sender half of the projections of boundary rows also goes into the neighbours'
  // arrays (peer-mapped): row v -> row dst_* + push_*[v] there.  NULL: no neighbour on that side.
  float *P_left, *P_right;
  const int32_t *push_left, *push_right;
  int dst_left, dst_right;
  int32_t* flag;  // OR-ed with 1 when a decoded output is NaN / Inf (or NULL)
  float inv_latent;  // as EdgeTcArgs
  int pdl;           // as EdgeTcArgs
  // encoder mode: input rows [n][enc_stride] (enc_stride <= 128, a multiple of 4; columns beyond it are zero)
  const float* enc_in;
  int enc_stride;
};

int launch_node_mp_tc(const NodeTcArgs& a, cudaStream_t s);
// v2 (node_tc2.cu): weights resident in TMEM, four workers x 32-node tiles, two passes (update, projections)
int launch_node_mp_tc2(const NodeTcArgs& a, cudaStream_t s);
// h[i] = [node_feat[i] | embedding[ptype[i]] | 0 ...] (128 wide): the encoder's input operand
int launch_node_embed(const float* node_feat, int node_in, int node_stride, const int32_t* ptype, const float* embedding,
                      int embed, int n_types, int n, float* h, cudaStream_t s);

}  // namespace lb
