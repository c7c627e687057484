// Interface between gns.cu (forward orchestration) and gns_tc.cu (tcgen05 edge kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb {

struct EdgeTcArgs {
  int n;
  const int32_t *rowptr, *snd, *rcv;
  const float* P;       // [n][256] per-node projections (sender half | receiver half + b1)
  const void* w_tc;     // W1e^T hi, lo, W2c^T hi, lo: four 128x128 fp16 operands in UMMA K-major layout
  const float* vec_tc;  // b2c[128] | ln_scale[128] | ln_offset[128]
  float *e, *agg, *carry_first, *carry_last;
};

int launch_edge_mp_tc(const EdgeTcArgs& a, int e_cap, cudaStream_t s);

}  // namespace lb
