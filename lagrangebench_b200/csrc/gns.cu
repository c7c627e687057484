// GNS forward: encoder, message passing (fused gather + edge MLP + LayerNorm + residual +
// deterministic segmented sum), node update (+ next-layer projection), decoder.
//
// Replaces lagrangebench/models/gns.py:65-171 (GNS._encoder/_processor/_decoder) with
// build_mlp of lagrangebench/models/utils.py:100-115 and the gather / segment_sum of
// jraph.GraphNetwork (third-party, call site gns.py:117-119).
//
// Data layout in HBM (all float32, row-major):
//   h    [N][128]      node latents
//   P    [N][256]      per-node projections for the NEXT edge update: cols 0..127 =
//                      h @ W1[0:128]  (sender part), cols 128..255 = h @ W1[128:256] + b1
//                      (receiver part).  The first edge-MLP layer is linear in its concat
//                      input [h_s, h_r, e] (gns.py:97-100), so its two node terms are hoisted
//                      out of the per-edge work (E ~ 6.6-12.8 N): 131 072 -> 65 536 FLOP/edge.
//   e    [E_cap][128]  edge latents in RECEIVER-MAJOR slot order (lb200_csr_build), so each
//                      receiver's incoming messages are contiguous and summed in ascending
//                      list order by exactly one tile -- no atomics, no zero-fill.
//   agg  [N][128]      aggregated messages of receivers whose bucket lies inside one tile;
//   carry_first/last [tiles][128]  partial sums of buckets that straddle a tile boundary,
//                      combined (in tile order) by the node kernel.
#include <stdlib.h>

#include "common.cuh"
#include "gns_tc.cuh"

namespace lb {

constexpr int kThreads = 256;
constexpr int kTM = 64;         // rows per tile (== kEdgeTile == kNodeTile)
constexpr int kWChunk = 32;     // weight rows per cp.async stage
constexpr int kLdA = kLatent + 4;
constexpr int kEncK = LB200_MAX_NODE_IN;  // node-encoder input width, zero padded
static_assert(kTM % kEdgeTile == 0 && kNodeTile == kTM, "tile constants");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// column owned by accumulator slot j of thread-column tx: two float4 groups 64 apart
__device__ __forceinline__ int col_of(int tx, int j) { return (j < 4 ? 0 : 64) + tx * 4 + (j & 3); }

__device__ __forceinline__ void load_w_chunk(float* ws, const float* __restrict__ wg) {
  // 32 x 128 floats = 1024 float4, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int f = (threadIdx.x + i * kThreads) * 4;
    cp_async16(ws + f, wg + f);
  }
}

// acc[4][8] += A[64][K] (smem, row stride lda) @ W[K][128] (global, streamed through ws[2][32*128])
template <int K>
__device__ __forceinline__ void gemm_tile(float (&acc)[4][8], const float* As, int lda,
                                          const float* __restrict__ wg, float* ws) {
  static_assert(K % kWChunk == 0, "K");
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  constexpr int NC = K / kWChunk;
  load_w_chunk(ws, wg);
  cp_async_commit();
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    if (c + 1 < NC) {
      load_w_chunk(ws + ((c + 1) & 1) * kWChunk * kLatent, wg + (int64_t)(c + 1) * kWChunk * kLatent);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* w = ws + (c & 1) * kWChunk * kLatent;
    const float* a_base = As + (ty * 4) * lda + c * kWChunk;
#pragma unroll
    for (int kk = 0; kk < kWChunk; kk += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a_base + i * lda + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 b0 = *reinterpret_cast<const float4*>(w + (kk + k4) * kLatent + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(w + (kk + k4) * kLatent + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = k4 == 0 ? a[i].x : (k4 == 1 ? a[i].y : (k4 == 2 ? a[i].z : a[i].w));
          acc[i][0] = fmaf(av, b0.x, acc[i][0]);
          acc[i][1] = fmaf(av, b0.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0.z, acc[i][2]);
          acc[i][3] = fmaf(av, b0.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ float half_warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

// hk.LayerNorm(axis=-1): (scale * rsqrt(var + 1e-5)) * (x - mean) + offset, biased variance
// `latent` <= 128 is the model's true width: columns beyond it are zero padding (zero weights, scale, offset),
// which adds nothing to the sum and (128 - latent) * mean^2 to the sum of squared deviations.
__device__ __forceinline__ void layer_norm_rows(float (&y)[4][8], const float* __restrict__ scale,
                                                const float* __restrict__ offset, int latent) {
  const int tx = threadIdx.x & 15;
  const float inv_n = 1.0f / (float)latent, n_pad = (float)(kLatent - latent);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += y[i][j];
    const float mean = half_warp_sum(s) * inv_n;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = y[i][j] - mean;
      v = fmaf(d, d, v);
    }
    const float var = fmaxf(half_warp_sum(v) - n_pad * mean * mean, 0.f) * inv_n;
    const float inv = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col_of(tx, j);
      y[i][j] = fmaf(scale[c] * inv, y[i][j] - mean, offset[c]);
    }
  }
}

__device__ __forceinline__ void add_bias(float (&y)[4][8], const float* __restrict__ b, bool relu) {
  const int tx = threadIdx.x & 15;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float bv = b[col_of(tx, j)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = y[i][j] + bv;
      y[i][j] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// registers -> smem tile (row stride lda), columns of this thread
__device__ __forceinline__ void store_tile_smem(const float (&y)[4][8], float* dst, int lda) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* r = dst + (ty * 4 + i) * lda;
    *reinterpret_cast<float4*>(r + tx * 4) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
    *reinterpret_cast<float4*>(r + 64 + tx * 4) = make_float4(y[i][4], y[i][5], y[i][6], y[i][7]);
  }
}

// registers -> global rows [row0, row0 + rows) of a [*][ld] matrix at column offset col0
__device__ __forceinline__ void store_tile_global(const float (&y)[4][8], float* __restrict__ dst, int64_t row0,
                                                  int rows, int ld, int col0) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r < rows) {
      float* p = dst + (row0 + r) * ld + col0;
      *reinterpret_cast<float4*>(p + tx * 4) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
      *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(y[i][4], y[i][5], y[i][6], y[i][7]);
    }
  }
}

struct MlpW {
  const float *w0, *b0, *w1, *b1, *lns, *lno;
};

// P = h @ [W1[0:128] | W1[128:256]] (+ b1 on the receiver half) for the next edge update.
__device__ __forceinline__ void project_next(const float* hs, int lda, const MlpW& nxt, float* ws,
                                             float* __restrict__ P, int64_t row0, int rows) {
  float acc[4][8];
  zero_acc(acc);
  gemm_tile<kLatent>(acc, hs, lda, nxt.w0, ws);
  store_tile_global(acc, P, row0, rows, 2 * kLatent, 0);
  zero_acc(acc);
  gemm_tile<kLatent>(acc, hs, lda, nxt.w0 + kLatent * kLatent, ws);
  add_bias(acc, nxt.b0, false);
  store_tile_global(acc, P, row0, rows, 2 * kLatent, kLatent);
}

// ------------------------------------------------------------------ node encoder
struct NodeEncArgs {
  int n, node_in, node_stride, embed, n_types, latent;
  const float* node_feat;
  const int32_t* ptype;
  const float* embedding;
  MlpW enc;   // w0 zero-padded to kEncK rows
  MlpW nxt;   // processor edge MLP of step 0 (for P)
  float *h, *P;
};

__global__ void __launch_bounds__(kThreads, 2) node_encoder_kernel(NodeEncArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [64][kEncK + 4]
  float* Hs = As + kTM * (kEncK + 4);     // [64][kLdA]
  float* ws = Hs + kTM * kLdA;            // [2][32][128]
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int rows = min(kTM, a.n - (int)row0);
  constexpr int lda = kEncK + 4;
  for (int f = threadIdx.x; f < kTM * kEncK; f += kThreads) {
    const int r = f / kEncK, c = f % kEncK;
    float v = 0.f;
    if (r < rows) {
      if (c < a.node_in) {
        v = a.node_feat[(row0 + r) * a.node_stride + c];
      } else if (c < a.node_in + a.embed) {
        int t = a.ptype[row0 + r];
        if (t < 0) t += a.n_types;  // hk.Embed indexes like NumPy: PAD_VALUE (-1) is the last row
        t = min(max(t, 0), a.n_types - 1);
        v = a.embedding[t * a.embed + (c - a.node_in)];
      }
    }
    As[r * lda + c] = v;
  }
  __syncthreads();
  float acc[4][8];
  zero_acc(acc);
  gemm_tile<kEncK>(acc, As, lda, a.enc.w0, ws);
  add_bias(acc, a.enc.b0, true);
  store_tile_smem(acc, Hs, kLdA);
  __syncthreads();
  zero_acc(acc);
  gemm_tile<kLatent>(acc, Hs, kLdA, a.enc.w1, ws);
  add_bias(acc, a.enc.b1, false);
  layer_norm_rows(acc, a.enc.lns, a.enc.lno, a.latent);
  store_tile_global(acc, a.h, row0, rows, kLatent, 0);
  store_tile_smem(acc, Hs, kLdA);  // gemm_tile ended with a barrier: Hs is free
  __syncthreads();
  project_next(Hs, kLdA, a.nxt, ws, a.P, row0, rows);
}

// ------------------------------------------------------------------ edge encoder
struct EdgeEncArgs {
  int n, latent;
  const int32_t* rowptr;  // rowptr[n] = number of real edges
  const int32_t* perm;
  const float4* edge_feat;  // LIST order
  MlpW enc;                 // w0 zero-padded to 4 rows
  float* e;
};

__global__ void __launch_bounds__(kThreads, 2) edge_encoder_kernel(EdgeEncArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Hs = smem;               // [64][kLdA]
  float* ws = Hs + kTM * kLdA;    // [2][32][128]
  float* w0s = ws + 2 * kWChunk * kLatent;  // [4][128]
  float4* efs = reinterpret_cast<float4*>(w0s + 4 * kLatent);  // [64]
  const int E = a.rowptr[a.n];
  const int64_t slot0 = (int64_t)blockIdx.x * kTM;
  if (slot0 >= E) return;
  const int rows = min(kTM, E - (int)slot0);
  for (int f = threadIdx.x; f < 4 * kLatent; f += kThreads) w0s[f] = a.enc.w0[f];
  if (threadIdx.x < kTM)
    efs[threadIdx.x] = threadIdx.x < rows ? a.edge_feat[a.perm ? a.perm[slot0 + threadIdx.x] : slot0 + threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = efs[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col_of(tx, j);
      float v = f.x * w0s[c];
      v = fmaf(f.y, w0s[kLatent + c], v);
      v = fmaf(f.z, w0s[2 * kLatent + c], v);
      v = fmaf(f.w, w0s[3 * kLatent + c], v);
      acc[i][j] = v;
    }
  }
  add_bias(acc, a.enc.b0, true);
  store_tile_smem(acc, Hs, kLdA);
  __syncthreads();
  zero_acc(acc);
  gemm_tile<kLatent>(acc, Hs, kLdA, a.enc.w1, ws);
  add_bias(acc, a.enc.b1, false);
  layer_norm_rows(acc, a.enc.lns, a.enc.lno, a.latent);
  store_tile_global(acc, a.e, slot0, rows, kLatent, 0);
}

// ------------------------------------------------------------------ message passing: edges
struct EdgeMpArgs {
  int n, latent;
  const int32_t *rowptr, *snd, *rcv;
  const float* P;  // [n][256]
  MlpW mlp;        // w0 = (384,128): rows 256..383 act on the edge latent; b0 is folded into P
  float *e, *agg, *carry_first, *carry_last;
};

__global__ void __launch_bounds__(kThreads, 2) edge_mp_kernel(EdgeMpArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* A1 = smem;                 // [64][kLdA]  e_old
  float* A2 = A1 + kTM * kLdA;      // [64][kLdA]  hidden, then e'
  float* ws = A2 + kTM * kLdA;      // [2][32][128]
  int* sidx = reinterpret_cast<int*>(ws + 2 * kWChunk * kLatent);  // [64]
  int* ridx = sidx + kTM;                                          // [64 + 2] (+ neighbours of the tile)
  const int E = a.rowptr[a.n];
  const int64_t slot0 = (int64_t)blockIdx.x * kTM;
  if (slot0 >= E) return;
  const int rows = min(kTM, E - (int)slot0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (threadIdx.x < kTM) {
    const bool ok = threadIdx.x < rows;
    sidx[threadIdx.x] = ok ? a.snd[slot0 + threadIdx.x] : 0;
    ridx[threadIdx.x] = ok ? a.rcv[slot0 + threadIdx.x] : -1;
  } else if (threadIdx.x == kTM) {
    ridx[kTM] = slot0 > 0 ? a.rcv[slot0 - 1] : -2;                 // receiver just before the tile
    ridx[kTM + 1] = slot0 + rows < E ? a.rcv[slot0 + rows] : -3;   // receiver just after the tile
  }
  for (int f = threadIdx.x; f < kTM * (kLatent / 4); f += kThreads) {
    const int r = f / (kLatent / 4), c4 = f % (kLatent / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) v = *reinterpret_cast<const float4*>(a.e + (slot0 + r) * kLatent + c4 * 4);
    *reinterpret_cast<float4*>(A1 + r * kLdA + c4 * 4) = v;
  }
  __syncthreads();
  float acc[4][8];
  zero_acc(acc);
  gemm_tile<kLatent>(acc, A1, kLdA, a.mlp.w0 + 2 * kLatent * kLatent, ws);
  // hidden = relu(e @ W1e + P_s[snd] + P_r[rcv])   (b1 is inside P_r)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const int s = sidx[r], rc = max(ridx[r], 0);
    const float* ps = a.P + (int64_t)s * (2 * kLatent);
    const float* pr = a.P + (int64_t)rc * (2 * kLatent) + kLatent;
    const float4 s0 = *reinterpret_cast<const float4*>(ps + tx * 4);
    const float4 s1 = *reinterpret_cast<const float4*>(ps + 64 + tx * 4);
    const float4 r0 = *reinterpret_cast<const float4*>(pr + tx * 4);
    const float4 r1 = *reinterpret_cast<const float4*>(pr + 64 + tx * 4);
    acc[i][0] = fmaxf(acc[i][0] + s0.x + r0.x, 0.f);
    acc[i][1] = fmaxf(acc[i][1] + s0.y + r0.y, 0.f);
    acc[i][2] = fmaxf(acc[i][2] + s0.z + r0.z, 0.f);
    acc[i][3] = fmaxf(acc[i][3] + s0.w + r0.w, 0.f);
    acc[i][4] = fmaxf(acc[i][4] + s1.x + r1.x, 0.f);
    acc[i][5] = fmaxf(acc[i][5] + s1.y + r1.y, 0.f);
    acc[i][6] = fmaxf(acc[i][6] + s1.z + r1.z, 0.f);
    acc[i][7] = fmaxf(acc[i][7] + s1.w + r1.w, 0.f);
  }
  store_tile_smem(acc, A2, kLdA);
  __syncthreads();
  zero_acc(acc);
  gemm_tile<kLatent>(acc, A2, kLdA, a.mlp.w1, ws);
  add_bias(acc, a.mlp.b1, false);
  layer_norm_rows(acc, a.mlp.lns, a.mlp.lno, a.latent);
  store_tile_smem(acc, A2, kLdA);  // e' (the message), for the column-wise segmented sum
  // residual: e <- e' + e (gns.py:120-122)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* o = A1 + (ty * 4 + i) * kLdA;
    const float4 o0 = *reinterpret_cast<const float4*>(o + tx * 4);
    const float4 o1 = *reinterpret_cast<const float4*>(o + 64 + tx * 4);
    acc[i][0] += o0.x; acc[i][1] += o0.y; acc[i][2] += o0.z; acc[i][3] += o0.w;
    acc[i][4] += o1.x; acc[i][5] += o1.y; acc[i][6] += o1.z; acc[i][7] += o1.w;
  }
  store_tile_global(acc, a.e, slot0, rows, kLatent, 0);
  __syncthreads();
  // deterministic segmented sum over receivers: one thread per column walks the tile's rows,
  // one carry sub-tile (kEdgeTile edges) at a time
  if (threadIdx.x < kLatent) {
    const int c = threadIdx.x;
    for (int r0 = 0; r0 < rows; r0 += kEdgeTile) {
      const int r1 = min(rows, r0 + kEdgeTile);
      const int before = r0 == 0 ? ridx[kTM] : ridx[r0 - 1];
      const int after = r1 == rows ? ridx[kTM + 1] : ridx[r1];
      const bool first_cont = before == ridx[r0];
      const bool last_cont = after == ridx[r1 - 1];
      const int64_t sub = (int64_t)blockIdx.x * (kTM / kEdgeTile) + r0 / kEdgeTile;
      float sum = 0.f;
      int seg_start = r0;
      for (int r = r0; r < r1; ++r) {
        sum += A2[r * kLdA + c];
        const bool end = (r == r1 - 1) || (ridx[r + 1] != ridx[r]);
        if (end) {
          if (seg_start == r0 && first_cont)
            a.carry_first[sub * kLatent + c] = sum;
          else if (r == r1 - 1 && last_cont)
            a.carry_last[sub * kLatent + c] = sum;
          else
            a.agg[(int64_t)ridx[r] * kLatent + c] = sum;
          sum = 0.f;
          seg_start = r + 1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ message passing: nodes
struct NodeMpArgs {
  int n, dim, last, latent;
  const int32_t* rowptr;
  const float *agg, *carry_first, *carry_last;
  MlpW mlp;  // w0 = (256,128): rows 0..127 act on h, 128..255 on the aggregate
  MlpW nxt;  // next step's edge MLP (projection) or the decoder when last
  float *h, *P, *out;
  int32_t* flag;  // OR-ed with 1 when a decoded output is NaN / Inf (or NULL)
};

__global__ void __launch_bounds__(kThreads, 2) node_mp_kernel(NodeMpArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int lda = 2 * kLatent + 4;
  float* As = smem;               // [64][260]: h | agg, later h_new | hidden
  float* ws = As + kTM * lda;     // [2][32][128]
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int rows = min(kTM, a.n - (int)row0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int f = threadIdx.x; f < kTM * (kLatent / 4); f += kThreads) {
    const int r = f / (kLatent / 4), c4 = f % (kLatent / 4);
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), av = hv;
    if (r < rows) {
      const int64_t v = row0 + r;
      hv = *reinterpret_cast<const float4*>(a.h + v * kLatent + c4 * 4);
      const int e0 = a.rowptr[v], e1 = a.rowptr[v + 1];
      if (e1 > e0) {
        const int ta = e0 / kEdgeTile, tb = (e1 - 1) / kEdgeTile;
        if (ta == tb) {
          av = *reinterpret_cast<const float4*>(a.agg + v * kLatent + c4 * 4);
        } else {  // bucket straddles tiles: partial sums in tile order
          av = *reinterpret_cast<const float4*>(a.carry_last + (int64_t)ta * kLatent + c4 * 4);
          for (int t = ta + 1; t <= tb; ++t) {
            const float4 p = *reinterpret_cast<const float4*>(a.carry_first + (int64_t)t * kLatent + c4 * 4);
            av.x += p.x; av.y += p.y; av.z += p.z; av.w += p.w;
          }
        }
      }
    }
    *reinterpret_cast<float4*>(As + r * lda + c4 * 4) = hv;
    *reinterpret_cast<float4*>(As + r * lda + kLatent + c4 * 4) = av;
  }
  __syncthreads();
  float acc[4][8];
  zero_acc(acc);
  gemm_tile<2 * kLatent>(acc, As, lda, a.mlp.w0, ws);
  add_bias(acc, a.mlp.b0, true);
  store_tile_smem(acc, As + kLatent, lda);  // hidden over the aggregate half
  __syncthreads();
  zero_acc(acc);
  gemm_tile<kLatent>(acc, As + kLatent, lda, a.mlp.w1, ws);
  add_bias(acc, a.mlp.b1, false);
  layer_norm_rows(acc, a.mlp.lns, a.mlp.lno, a.latent);
  // residual: h <- n' + h (gns.py:120-122); each thread owns its (row, col) elements of As
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* o = As + (ty * 4 + i) * lda;
    const float4 o0 = *reinterpret_cast<const float4*>(o + tx * 4);
    const float4 o1 = *reinterpret_cast<const float4*>(o + 64 + tx * 4);
    acc[i][0] += o0.x; acc[i][1] += o0.y; acc[i][2] += o0.z; acc[i][3] += o0.w;
    acc[i][4] += o1.x; acc[i][5] += o1.y; acc[i][6] += o1.z; acc[i][7] += o1.w;
  }
  store_tile_smem(acc, As, lda);
  if (!a.last) store_tile_global(acc, a.h, row0, rows, kLatent, 0);
  __syncthreads();
  if (!a.last) {
    project_next(As, lda, a.nxt, ws, a.P, row0, rows);
    return;
  }
  // decoder (gns.py:126-133): Linear(128,128) - ReLU - Linear(128,dim), no LayerNorm
  zero_acc(acc);
  gemm_tile<kLatent>(acc, As, lda, a.nxt.w0, ws);
  add_bias(acc, a.nxt.b0, true);
  for (int k = 0; k < a.dim; ++k) {
    float wk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wk[j] = a.nxt.w1[col_of(tx, j) * a.dim + k];
    const float bk = a.nxt.b1[k];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) p = fmaf(acc[i][j], wk[j], p);
      p = half_warp_sum(p);
      const int r = ty * 4 + i;
      if (tx == 0 && r < rows) {
        a.out[(row0 + r) * a.dim + k] = p + bk;
        if (a.flag != nullptr && !isfinite(p + bk)) atomicOr(a.flag, 1);
      }
    }
  }
}

// Which tensor-core message kernel runs: cfg.edge_impl 0 = default (v2, gns_tc2.cu), 2 = v1 (gns_tc.cu).
// LB200_EDGE_TC=1|2 overrides the default (A/B measurements).
static int edge_tc_version(int edge_impl) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("LB200_EDGE_TC");
    env = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  if (edge_impl == 2) return 1;
  return env ? env : 2;
}

// Which tensor-core node kernel runs: v2 (node_tc2.cu) unless LB200_NODE_TC=1 asks for v1 (A/B measurements).
static int node_tc_version() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("LB200_NODE_TC");
    env = (e && e[0] == '1') ? 1 : 2;
  }
  return env;
}
static bool launch_node_is_v2() { return node_tc_version() == 2; }
static int launch_node(const NodeTcArgs& a, cudaStream_t s) {
  return node_tc_version() == 1 ? launch_node_mp_tc(a, s) : launch_node_mp_tc2(a, s);
}

static MlpW mlp_ptrs(const float* w, const lb200_mlp_off& o) {
  MlpW m;
  m.w0 = w + o.w0;
  m.b0 = w + o.b0;
  m.w1 = w + o.w1;
  m.b1 = w + o.b1;
  m.lns = o.ln_scale >= 0 ? w + o.ln_scale : nullptr;
  m.lno = o.ln_offset >= 0 ? w + o.ln_offset : nullptr;
  return m;
}

constexpr int kSmemNodeEnc = (kTM * (kEncK + 4) + kTM * kLdA + 2 * kWChunk * kLatent) * 4;
constexpr int kSmemEdgeEnc = (kTM * kLdA + 2 * kWChunk * kLatent + 4 * kLatent) * 4 + kTM * 16;
constexpr int kSmemEdgeMp = (2 * kTM * kLdA + 2 * kWChunk * kLatent) * 4 + (2 * kTM + 2) * 4;
constexpr int kSmemNodeMp = (kTM * (2 * kLatent + 4) + 2 * kWChunk * kLatent) * 4;

static int set_smem_once() {
  static int state[kMaxDevices];  // 0: not yet, 1: done, < 0 or > 1: the error it ended with (offset by 2)
  int rc = 0;
  const int dev = device_slot(&rc);
  if (dev < 0) return rc;
  if (state[dev] == 1) return 0;
  cudaError_t e;
  e = cudaFuncSetAttribute(node_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNodeEnc);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(edge_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEdgeEnc);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(edge_mp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEdgeMp);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(node_mp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNodeMp);
  if (e == cudaSuccess) state[dev] = 1;
  return (int)e;
}

}  // namespace lb

using namespace lb;

extern "C" int64_t lb200_gns_scratch_bytes(int32_t n, int32_t e_cap) {
  int64_t nt = cdiv(e_cap, kEdgeTile) + 1;
  int64_t b = 0;
  b += align_up((int64_t)n * kLatent * 4, 256);        // h
  b += align_up((int64_t)n * 2 * kLatent * 4, 256);    // P
  b += align_up((int64_t)n * kLatent * 4, 256);        // agg
  b += align_up((int64_t)e_cap * kLatent * 4, 256);    // e
  b += 2 * align_up(nt * kLatent * 4, 256);            // carries
  return b + 4096;
}

extern "C" int lb200_gns_scratch_layout(int32_t n, int32_t e_cap, int64_t* off_h, int64_t* off_p, int64_t* off_agg,
                                        int64_t* off_e) {
  Arena ar(nullptr, 0);
  float* h = ar.take<float>((int64_t)n * kLatent);
  float* P = ar.take<float>((int64_t)n * 2 * kLatent);
  float* agg = ar.take<float>((int64_t)n * kLatent);
  float* e = ar.take<float>((int64_t)e_cap * kLatent);
  if (off_h) *off_h = (char*)h - (char*)nullptr;
  if (off_p) *off_p = (char*)P - (char*)nullptr;
  if (off_agg) *off_agg = (char*)agg - (char*)nullptr;
  if (off_e) *off_e = (char*)e - (char*)nullptr;
  return 0;
}

extern "C" int lb200_gns_forward(const lb200_gns_cfg* c, const float* weights_dev, const float* node_feat_dev,
                                 const float* edge_feat_dev, const int32_t* ptype_dev, const int32_t* rowptr_dev,
                                 const int32_t* perm_dev, const int32_t* snd_dev, const int32_t* rcv_dev,
                                 float* out_dev, void* scratch_dev, int64_t scratch_bytes, void* stream) {
  if (!c || !weights_dev || !node_feat_dev || !edge_feat_dev || !ptype_dev || !rowptr_dev ||
      !snd_dev || !rcv_dev || !out_dev || !scratch_dev)
    return LB200_EINVAL;
  if (c->num_mp_steps < 1 || c->node_in + c->embed_size > kEncK || (c->dim != 2 && c->dim != 3) || c->e_cap < 1 ||
      c->latent < 0 || c->latent > kLatent)
    return LB200_EUNSUPPORTED;
  int rc = set_smem_once();
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = c->n, e_cap = c->e_cap;
  const int n_own = c->n_owned > 0 ? c->n_owned : n;  // rows [n_own, n) are ghosts (halo exchange)
  const int nt = cdiv(e_cap, kTM);             // CUDA-core edge tiles
  const int n_sub = cdiv(e_cap, kEdgeTile);    // carry sub-tiles
  Arena ar(scratch_dev, scratch_bytes);
  float* h = ar.take<float>((int64_t)n * kLatent);
  float* P = ar.take<float>((int64_t)n * 2 * kLatent);
  float* agg = ar.take<float>((int64_t)n * kLatent);
  float* e = ar.take<float>((int64_t)e_cap * kLatent);
  float* cf = ar.take<float>((int64_t)(n_sub + 1) * kLatent);
  float* cl = ar.take<float>((int64_t)(n_sub + 1) * kLatent);
  if (!ar.ok()) return LB200_EINVAL;
  const float* w = weights_dev;
  const int latent = c->latent > 0 ? c->latent : kLatent;  // the model's true width; beyond it: zero padding
  const float inv_latent = 1.0f / (float)latent;
  // Decomposed cloud: the projections live in the peer heap, two arrays alternating by message-passing
  // step (a neighbour's store for step m + 1 can then never race this rank's message kernel of step m);
  // exchange 0 of a step is the ghost positions (rollout.cu), exchanges 1 .. num_mp_steps the projections.
  const lb200_shard* sh = c->shard;
  const int per_step = c->num_mp_steps + 1;
  if (sh != nullptr) {
    if (c->edge_impl == 1 || c->enc_node.tc_w < 0 || c->enc_node.tc_vec < 0) return LB200_EUNSUPPORTED;
    if (n > sh->n_cap || n_own != sh->n_owned) return LB200_EINVAL;
  }
  auto p_of = [&](int m) -> float* { return sh != nullptr ? shard_p_local(sh, m) : P; };
  auto set_push = [&](NodeTcArgs& na, int m_out) {  // the node kernel writing projection array m_out
    na.P_left = na.P_right = nullptr;
    na.push_left = na.push_right = nullptr;
    na.dst_left = na.dst_right = 0;
    na.flag = c->nonfinite_flag;
    na.inv_latent = inv_latent;
    if (sh == nullptr) return;
    if (sh->has_left && sh->n_send_left > 0) {
      na.P_left = shard_p_left(sh, m_out);
      na.push_left = sh->push_left;
      na.dst_left = sh->dst_row_left;
    }
    if (sh->has_right && sh->n_send_right > 0) {
      na.P_right = shard_p_right(sh, m_out);
      na.push_right = sh->push_right;
      na.dst_right = sh->dst_row_right;
    }
  };

  NodeEncArgs ne;
  ne.n = n_own;
  ne.latent = latent;
  ne.node_in = c->node_in;
  ne.node_stride = c->node_stride;
  ne.embed = c->embed_size;
  ne.n_types = c->num_particle_types;
  ne.node_feat = node_feat_dev;
  ne.ptype = ptype_dev;
  ne.embedding = w + c->embedding;
  ne.enc = mlp_ptrs(w, c->enc_node);
  ne.nxt = mlp_ptrs(w, c->proc_edge[0]);
  ne.h = h;
  ne.P = p_of(0);
  if (c->edge_impl != 1 && c->enc_node.tc_w >= 0 && c->enc_node.tc_vec >= 0) {
    // tensor-core encoder: the node-update kernel in encoder mode over the zero-padded input features
    const bool embedded = c->node_feat_embedded != 0 && launch_node_is_v2();
    if (c->node_feat_embedded && (!embedded || c->node_stride % 4 != 0 || c->node_stride > kLatent)) return LB200_EINVAL;
    if (!embedded) {
      rc = launch_node_embed(node_feat_dev, c->node_in, c->node_stride, ptype_dev, w + c->embedding, c->embed_size,
                             c->num_particle_types, n_own, h, s);
      if (rc) return rc;
    }
    NodeTcArgs na;
    na.n = n_own;
    na.last = 0;
    na.enc = 1;
    na.dim = c->dim;
    na.rowptr = rowptr_dev;
    na.agg = nullptr;
    na.carry_first = nullptr;
    na.carry_last = nullptr;
    na.w_tc = w + c->enc_node.tc_w;
    na.vec_tc = w + c->enc_node.tc_vec;
    na.h = h;
    na.P = p_of(0);
    na.out = out_dev;
    set_push(na, 0);
    na.enc_in = embedded ? node_feat_dev : h;
    na.enc_stride = embedded ? c->node_stride : kLatent;
    rc = launch_node(na, s);
    if (rc) return rc;
    if (sh != nullptr) shard_exchange(sh, 1, per_step, s);
  } else {
    if (c->node_feat_embedded) return LB200_EUNSUPPORTED;
    node_encoder_kernel<<<cdiv(n_own, kTM), kThreads, kSmemNodeEnc, s>>>(ne);
    LB_LAUNCHED(1);
  }

  if (c->edge_impl != 1 && c->enc_edge.tc_w >= 0 && c->enc_edge.tc_vec >= 0 &&
      c->enc_edge.b0 == c->enc_edge.w0 + 4 * kLatent) {
    EdgeTcArgs et;
    et.n = n_own;
    et.inv_latent = inv_latent;
    et.rowptr = rowptr_dev;
    et.snd = snd_dev;
    et.rcv = rcv_dev;
    et.P = nullptr;
    et.w_tc = w + c->enc_edge.tc_w;
    et.vec_tc = w + c->enc_edge.tc_vec;
    et.e = e;
    et.agg = nullptr;
    et.carry_first = nullptr;
    et.carry_last = nullptr;
    et.encoder = 1;
    et.reverse = 0;
    et.edge_feat = reinterpret_cast<const float4*>(edge_feat_dev);
    et.perm = perm_dev;
    et.enc_vec = w + c->enc_edge.w0;
    rc = edge_tc_version(c->edge_impl) == 2 ? launch_edge_mp_tc2(et, e_cap, s) : launch_edge_mp_tc(et, e_cap, s);
    if (rc) return rc;
  } else {
    EdgeEncArgs ee;
    ee.n = n_own;
    ee.latent = latent;
    ee.rowptr = rowptr_dev;
    ee.perm = perm_dev;
    ee.edge_feat = reinterpret_cast<const float4*>(edge_feat_dev);
    ee.enc = mlp_ptrs(w, c->enc_edge);
    ee.e = e;
    { edge_encoder_kernel<<<nt, kThreads, kSmemEdgeEnc, s>>>(ee); LB_LAUNCHED(1); }
  }

  for (int m = 0; m < c->num_mp_steps; ++m) {
    const lb200_mlp_off& eo = c->proc_edge[m];
    prof_begin(0, s);
    if (c->edge_impl != 1 && eo.tc_w >= 0 && eo.tc_vec >= 0) {
      EdgeTcArgs et;
      et.n = n_own;
      et.inv_latent = inv_latent;
      et.rowptr = rowptr_dev;
      et.snd = snd_dev;
      et.rcv = rcv_dev;
      et.P = p_of(m);
      et.w_tc = w + eo.tc_w;
      et.vec_tc = w + eo.tc_vec;
      et.e = e;
      et.agg = agg;
      et.carry_first = cf;
      et.carry_last = cl;
      et.encoder = 0;
      et.reverse = (m & 1) == 0;  // the encoder went forwards: each launch starts where the previous one ended (L2)
      et.edge_feat = nullptr;
      et.perm = nullptr;
      et.enc_vec = nullptr;
      rc = edge_tc_version(c->edge_impl) == 2 ? launch_edge_mp_tc2(et, e_cap, s) : launch_edge_mp_tc(et, e_cap, s);
      if (rc) return rc;
    } else {
      EdgeMpArgs em;
      em.n = n_own;
      em.latent = latent;
      em.rowptr = rowptr_dev;
      em.snd = snd_dev;
      em.rcv = rcv_dev;
      em.P = p_of(m);
      em.mlp = mlp_ptrs(w, eo);
      em.e = e;
      em.agg = agg;
      em.carry_first = cf;
      em.carry_last = cl;
      { edge_mp_kernel<<<nt, kThreads, kSmemEdgeMp, s>>>(em); LB_LAUNCHED(1); }
    }
    prof_end(0, s);

    const lb200_mlp_off& no = c->proc_node[m];
    const bool last = m == c->num_mp_steps - 1;
    prof_begin(1, s);
    if (c->edge_impl != 1 && no.tc_w >= 0 && no.tc_vec >= 0) {
      NodeTcArgs nt_args;
      nt_args.n = n_own;
      nt_args.last = last;
      nt_args.enc = 0;
      nt_args.dim = c->dim;
      nt_args.rowptr = rowptr_dev;
      nt_args.agg = agg;
      nt_args.carry_first = cf;
      nt_args.carry_last = cl;
      nt_args.w_tc = w + no.tc_w;
      nt_args.vec_tc = w + no.tc_vec;
      nt_args.h = h;
      nt_args.P = p_of(m + 1);
      nt_args.out = out_dev;
      nt_args.enc_in = nullptr;
      nt_args.enc_stride = 0;
      set_push(nt_args, m + 1);
      rc = launch_node(nt_args, s);
      if (rc) return rc;
      prof_end(1, s);  // the node kernel alone: the wait for the neighbours' stores is not its time
      if (sh != nullptr && !last) shard_exchange(sh, m + 2, per_step, s);
    } else {
      NodeMpArgs nm;
      nm.n = n_own;
      nm.latent = latent;
      nm.dim = c->dim;
      nm.last = last;
      nm.rowptr = rowptr_dev;
      nm.agg = agg;
      nm.carry_first = cf;
      nm.carry_last = cl;
      nm.mlp = mlp_ptrs(w, no);
      nm.nxt = last ? mlp_ptrs(w, c->dec) : mlp_ptrs(w, c->proc_edge[m + 1]);
      nm.h = h;
      nm.P = p_of(m + 1);
      nm.flag = c->nonfinite_flag;
      nm.out = out_dev;
      { node_mp_kernel<<<cdiv(n_own, kTM), kThreads, kSmemNodeMp, s>>>(nm); LB_LAUNCHED(1); }
      prof_end(1, s);
    }
  }
  LB_LAUNCH_CHECK();
  return 0;
}
