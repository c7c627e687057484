// Node kernel v2 (tcgen05 + TMEM), sm_100a: the processor's node update of
// lagrangebench/models/gns.py:103-122 -- n' = LN(W2 relu(W1 [h, agg] + b1) + b2), h <- n' + h -- plus the
// projections the next edge update gathers (P = h W1s | h W1r + b1, gns.py:97-100 restructured) or, after
// the last step, the decoder (gns.py:126-133); and the node ENCODER (gns.py:65-81) as a compile-time mode.
//
// Same skeleton as the message kernel (gns_tc2.cu): one persistent 512-thread CTA per SM = four
// independent 128-thread workers, 32-node tiles (MMA M = 128 features, N = 32 nodes, GEMMs transposed so
// that thread == feature == TMEM lane and every global access is a coalesced 128-byte segment), weights
// RESIDENT IN TENSOR MEMORY, one rescaled accumulator per GEMM (split-precision fp16, scale-input-d).
// The five 128x128 operands of a step (hi|lo each: 640 TMEM columns) do not fit 512 columns at once, so a
// CTA runs two passes over its tiles:
//   pass A  W1h | W1a | W2c resident (384 columns) + 4 x 32 accumulator columns:
//           [h rows, aggregate rows] -> GEMM 1 (K = 256) -> relu -> GEMM 2 -> LayerNorm + residual -> h
//   pass B  W1s | W1r of the NEXT edge MLP (or the decoder's first layer) resident + 4 x 64 accumulator columns:
//           h rows (just written, L2 / L1 hits) -> two GEMMs -> P (and the neighbours' ghost rows of P:
//           peer-mapped stores from this epilogue, see lb200_shard) or the decoded acceleration.
// v1 (gns_tc.cu) streamed the five operands through one shared-memory buffer for every 128-node tile:
// 320 KB of L2 traffic per tile and a serial chain of five dependent stages per CTA.
#include <stdlib.h>

#include "tc_common.cuh"

namespace lb {

constexpr int kN2Threads = 512;
constexpr int kN2Workers = 4;
constexpr int kN2WThreads = kN2Threads / kN2Workers;
constexpr uint32_t kN2ColAccA = 384;  // pass A: + worker * 32
constexpr uint32_t kN2ColAccB = 256;  // pass B: + worker * 64 (sender projection / decoder), + 32 (receiver projection)

// shared memory map (bytes)
constexpr uint32_t kN2OffX = 0;                               // X operand hi | lo: node rows (h), 32 rows per worker
constexpr uint32_t kN2OffY = kN2OffX + 2 * kBBytes;           // Y operand hi | lo: aggregate rows, then the hidden layer
constexpr uint32_t kN2VecFloats = 5 * 128 + 3 * 128 + 4;      // b1 | b2c | ln_scale | ln_offset | b_next | wd1[128][3] | bd1[4]
constexpr uint32_t kN2OffVec = kN2OffY + 2 * kBBytes;
constexpr uint32_t kN2OffRed = ((kN2OffVec + kN2VecFloats * 4 + 15) / 16) * 16;  // [worker][3][4 warps][32]
constexpr uint32_t kN2OffInv = kN2OffRed + kN2Workers * 3 * 128 * 4;             // [16 warps][32]
constexpr uint32_t kN2OffBar = kN2OffInv + 16 * 32 * 4;                          // mbarriers g1[4], g2[4]; tmem slot
constexpr uint32_t kN2Smem = kN2OffBar + 8 * 8 + 16;

#ifdef LB200_CROSSCHECK
// phase timeline of CTA 0 (cross-check builds only): [worker][event] = (id, SM clock); lb200_debug_node_trace reads it
__device__ long long g_node_trace[kN2Workers][64][2];
__device__ int g_node_trace_n[kN2Workers];
#define LB_TRACE(id)                                                                          \
  do {                                                                                        \
    if (!kEnc && !a.last && blockIdx.x == 0 && q == 0 && lane == 0 && g_node_trace_n[wk] < 64) { \
      const int _i = g_node_trace_n[wk]++;                                                    \
      g_node_trace[wk][_i][0] = (id);                                                         \
      g_node_trace[wk][_i][1] = clock64();                                                    \
    }                                                                                         \
  } while (0)
#else
#define LB_TRACE(id) do { } while (0)
#endif

template <bool kEnc>
__global__ void __launch_bounds__(kN2Threads, 1) node_mp_tc2_kernel(NodeTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int wk = warp >> 2;  // worker
  const int q = warp & 3;    // TMEM lane quarter == warp index inside the worker
  const int f = q * 32 + lane;
  float* vec = reinterpret_cast<float*>(smem + kN2OffVec);
  float* red = reinterpret_cast<float*>(smem + kN2OffRed) + wk * 3 * 128;
  float* invs = reinterpret_cast<float*>(smem + kN2OffInv) + warp * 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kN2OffBar + 64);
  const uint32_t bar_g1 = sbase + kN2OffBar + 8 * wk, bar_g2 = bar_g1 + 32;
  const uint32_t bar_worker = 1 + wk, bar_ln = 1 + kN2Workers + wk;

  const int n_tiles = (a.n + k2Tile - 1) / k2Tile;
  const int grid = (int)gridDim.x;
  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel's tail (it waits before reading)
#ifdef LB200_CROSSCHECK
  if (!kEnc && !a.last && blockIdx.x == 0 && q == 0 && lane == 0) g_node_trace_n[wk] = 0;
#endif
  LB_TRACE(0);
  if ((int)blockIdx.x >= n_tiles) {
    pdl_wait();  // a CTA never exits before the previous kernel is complete: the chain of waits stays transitive
    return;
  }

  if (!a.pdl && !kEnc) {
    // not a programmatic dependent launch: the previous kernel is complete, so this worker's first tile (h and
    // aggregate rows, bucket bounds) can travel towards L1 / L2 while the weights are loaded
    const int tile0 = (int)blockIdx.x + wk * grid;
    const int64_t row = (int64_t)tile0 * k2Tile + q * 8 + (lane >> 2);
    if (row < a.n) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a.h + row * kLatent + (lane & 3) * 32));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a.agg + row * kLatent + (lane & 3) * 32));
      if ((lane & 3) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.rowptr + row));
    }
  }
  if (tid == 0) {
    for (int w = 0; w < 2 * kN2Workers; ++w) mbar_init(sbase + kN2OffBar + 8 * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (int)kN2VecFloats; i += kN2Threads) vec[i] = a.vec_tc[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  LB_TRACE(1);
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const uint4* wsrc = reinterpret_cast<const uint4*>(a.w_tc);  // operand halves (hi, lo, hi, lo ...) of 2048 uint4
  // ---- pass A weights -> tensor memory: columns [0,128) W1h (encoder: W0 padded), [128,256) W1a, [256,384) W2c
  {
    constexpr int kHalves = kEnc ? 4 : 6;
    for (int hh = wk; hh < kHalves; hh += kN2Workers) {
      const uint32_t col = kEnc ? (hh < 2 ? (uint32_t)hh * 64 : 256u + (uint32_t)(hh - 2) * 64) : (uint32_t)hh * 64;
      weight_to_tmem(wsrc + (size_t)hh * 2048, f, tmem + lane_sel + col);
    }
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  LB_TRACE(2);
  pdl_wait();  // constants only so far (weights); h, the aggregates and the CSR are the previous kernels' results
  LB_TRACE(3);

  const float b1 = vec[f], b2c = vec[128 + f], ln_scale = vec[256 + f], ln_offset = vec[384 + f], b_next = vec[512 + f];
  const uint32_t w1h_hi = tmem, w1h_lo = tmem + 64, w1a_hi = tmem + 128, w1a_lo = tmem + 192, w2_hi = tmem + 256,
                 w2_lo = tmem + 320;
  // this worker's 32 operand rows inside every K slab
  const uint32_t w_off = (uint32_t)wk * (k2Tile * 16);
  unsigned char* const x_hi_p = smem + kN2OffX + w_off;
  unsigned char* const x_lo_p = x_hi_p + kBBytes;
  unsigned char* const y_hi_p = smem + kN2OffY + w_off;
  unsigned char* const y_lo_p = y_hi_p + kBBytes;
  const uint32_t x_hi = sbase + kN2OffX + w_off, x_lo = x_hi + kBBytes;
  const uint32_t y_hi = sbase + kN2OffY + w_off, y_lo = y_hi + kBBytes;
  const int r0 = q * 8;  // this warp's 8 rows when a tile's rows are loaded
  uint32_t ph1 = 0, ph2 = 0;

  // one fp32 row segment (4 consecutive features of row r) -> K-major hi/lo operand; activation low halves unscaled
  auto put_row4 = [&](unsigned char* hi_p, unsigned char* lo_p, int r, const float4& v) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
    const __half2 l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    const uint32_t off = (uint32_t)(lane >> 1) * kLboB + (uint32_t)r * 16 + (uint32_t)(lane & 1) * 8;
    *reinterpret_cast<uint2*>(hi_p + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(lo_p + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  };
  // every thread has written (and fenced) its part of the operands; the worker's first warp issues
  auto operand_ready = [&]() -> bool {
    fence_async_smem();
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(kN2WThreads) : "memory");
    return q == 0;
  };
  auto bdesc = [&](uint32_t base, int j) { return umma_desc(base + j * 2 * kLboB, kLboB); };

  // ================================================================ pass A
  for (int tile = (int)blockIdx.x + wk * grid; tile < n_tiles; tile += kN2Workers * grid) {
    const int64_t row0 = (int64_t)tile * k2Tile;
    const int rows = min(k2Tile, a.n - (int)row0);
    const uint32_t acc = tmem + kN2ColAccA + (uint32_t)wk * 32;
    // ---- A0: h rows and aggregate rows (bucket carries resolved) -> operands, 8 rows per warp; GEMM 1.
    //      Loads are issued in three independent batches (rows + bucket bounds, aggregates, then the rare
    //      partial sums of buckets that straddle carry sub-tiles), not row after row.
    {
      float4 hv[8], av[8];
      int e0[8], e1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + i;
        hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        e0[i] = e1[i] = 0;
        if (r < rows) {
          const int64_t v = row0 + r;
          if (kEnc) {
            if (lane * 4 < a.enc_stride) hv[i] = reinterpret_cast<const float4*>(a.enc_in + v * a.enc_stride)[lane];
          } else {
            hv[i] = reinterpret_cast<const float4*>(a.h + v * kLatent)[lane];
          }
          if (!kEnc) {
            e0[i] = __ldg(a.rowptr + v);
            e1[i] = __ldg(a.rowptr + v + 1);
          }
        }
      }
      LB_TRACE(5);
      if (!kEnc) {
        // the worker's NEXT tile: its h and aggregate rows -> L1 while this tile computes (lane = row / 128-byte line).
        // (Also prefetching the tile's bucket bounds and the carry rows of its straddling buckets, or prefetching
        // into L2 instead of L1, changed nothing: 117-123 us per launch at 125 k nodes in every variant.)
        const int64_t nrow = (int64_t)(tile + kN2Workers * grid) * k2Tile + r0 + (lane >> 2);
        if (nrow < a.n) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a.h + nrow * kLatent + (lane & 3) * 32));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a.agg + nrow * kLatent + (lane & 3) * 32));
        }
      }
      if (!kEnc) {
        // a bucket inside one carry sub-tile: its aggregate; a bucket that straddles sub-tiles (about every second
        // one at 14 in-edges per node): carry_last of its first sub-tile + carry_first of the following ones,
        // in slot order.  The first two terms of all 8 rows are requested together.
        float4 cv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          cv[i] = av[i];
          if (e1[i] > e0[i]) {
            const int ta = e0[i] / kEdgeTile, tb = (e1[i] - 1) / kEdgeTile;
            const float* src = ta == tb ? a.agg + (row0 + r0 + i) * kLatent : a.carry_last + (int64_t)ta * kLatent;
            av[i] = reinterpret_cast<const float4*>(src)[lane];
            if (tb > ta) cv[i] = reinterpret_cast<const float4*>(a.carry_first + (int64_t)(ta + 1) * kLatent)[lane];
          }
        }
        LB_TRACE(6);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          av[i].x += cv[i].x; av[i].y += cv[i].y; av[i].z += cv[i].z; av[i].w += cv[i].w;
          if (e1[i] > e0[i]) {
            const int ta = e0[i] / kEdgeTile, tb = (e1[i] - 1) / kEdgeTile;
            for (int t = ta + 2; t <= tb; ++t) {  // in-degree beyond 32: further sub-tiles
              const float4 p = reinterpret_cast<const float4*>(a.carry_first + (int64_t)t * kLatent)[lane];
              av[i].x += p.x; av[i].y += p.y; av[i].z += p.z; av[i].w += p.w;
            }
          }
        }
      }
      // range guard of the fp16 split (|x| <= 65504): the aggregate is a data-dependent sum, the encoder's input
      // is raw data; everything else is bounded by the weights (checked on the host when they are packed)
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 g = kEnc ? hv[i] : av[i];
        bad = bad || !(fabsf(g.x) <= 65504.f) || !(fabsf(g.y) <= 65504.f) || !(fabsf(g.z) <= 65504.f) ||
              !(fabsf(g.w) <= 65504.f);
      }
      if (bad && a.flag != nullptr) atomicOr(a.flag, 1);
      LB_TRACE(7);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        put_row4(x_hi_p, x_lo_p, r0 + i, hv[i]);
        if (!kEnc) put_row4(y_hi_p, y_lo_p, r0 + i, av[i]);
      }
    }
    LB_TRACE(10);
    if (operand_ready()) {
      tc_fence_after();
      // acc = (W1h_lo' X_hi + W1a_lo' Y_hi) 2^-11 + W1h_hi X_lo + W1a_hi Y_lo + W1h_hi X_hi + W1a_hi Y_hi
      // One dependent chain of 48 instructions: its latency is on the tile's critical path, so the issue is
      // unrolled and every B descriptor is the slab-0 descriptor plus one 32-bit add (tc_common.cuh).
      const uint64_t dxh = umma_desc(x_hi, kLboB), dxl = umma_desc(x_lo, kLboB), dyh = umma_desc(y_hi, kLboB),
                     dyl = umma_desc(y_lo, kLboB);
      const uint32_t dw = (uint32_t)(dxh >> 32), xh0 = (uint32_t)dxh, xl0 = (uint32_t)dxl, yh0 = (uint32_t)dyh,
                     yl0 = (uint32_t)dyl;
      constexpr uint32_t kStep = (2 * kLboB) >> 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_ts_d(acc, w1h_lo + j * 8, xh0 + j * kStep, dw, j > 0 ? 1u : 0u, k2Idesc);
      if (!kEnc) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_ts_d(acc, w1a_lo + j * 8, yh0 + j * kStep, dw, 1u, k2Idesc);
      }
      umma_ts_d_rescale11(acc, w1h_hi, xl0, dw, k2Idesc);
#pragma unroll
      for (int j = 1; j < 8; ++j) umma_ts_d(acc, w1h_hi + j * 8, xl0 + j * kStep, dw, 1u, k2Idesc);
      if (!kEnc) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_ts_d(acc, w1a_hi + j * 8, yl0 + j * kStep, dw, 1u, k2Idesc);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_ts_d(acc, w1h_hi + j * 8, xh0 + j * kStep, dw, 1u, k2Idesc);
      if (!kEnc) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_ts_d(acc, w1a_hi + j * 8, yh0 + j * kStep, dw, 1u, k2Idesc);
      }
      umma_commit(bar_g1);
    }
    LB_TRACE(11);
    // ---- E1: hidden = relu(acc + b1) -> Y, node-contiguous (MN-major core matrices, 16-byte stores); GEMM 2
    mbar_wait(bar_g1, ph1);
    ph1 ^= 1;
    tc_fence_after();
    LB_TRACE(12);
    {
      unsigned char* hi_p = y_hi_p + (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 16;
      unsigned char* lo_p = hi_p + kBBytes;
      bool hid_bad = false;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float av[16];
        tmem_ld16(acc + lane_sel + h * 16, av);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2) {
            const int j = g * 8 + p2 * 2;
            const float x0 = fmaxf(av[j] + b1, 0.f), x1 = fmaxf(av[j + 1] + b1, 0.f);
            hid_bad = hid_bad || !(x0 <= 65504.f) || !(x1 <= 65504.f);  // range guard of the split (data-dependent here)
            const __half2 hh = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            hw[p2] = *reinterpret_cast<const uint32_t*>(&hh);
            lw[p2] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t off = (uint32_t)(h * 2 + g) * 128;
          *reinterpret_cast<uint4*>(hi_p + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(lo_p + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      if (hid_bad && a.flag != nullptr) atomicOr(a.flag, 1);
    }
    LB_TRACE(13);
    if (operand_ready()) {
      tc_fence_after();
      issue_gemm_ts_d(w2_hi, w2_lo, y_hi, y_lo, acc, k2IdescBMn);
      umma_commit(bar_g2);
    }
    // ---- E2: LayerNorm (mean folded into the weights) + residual -> h
    float* const hrow = a.h + row0 * kLatent + f;
    float hold[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) hold[j] = (!kEnc && j < rows) ? hrow[(int64_t)j * kLatent] : 0.f;  // no residual in the encoder
    LB_TRACE(14);
    mbar_wait(bar_g2, ph2);
    ph2 ^= 1;
    tc_fence_after();
    LB_TRACE(15);
    {
      float yc[32];
      float part;
      {
        float sq[32];
        tmem_ld32(acc + lane_sel, yc);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          yc[j] += b2c;
          sq[j] = yc[j] * yc[j];
        }
        part = warp_transpose_reduce(sq);  // lane l: this warp's 32 features, node l
      }
      red[q * 32 + lane] = part;
      asm volatile("bar.sync %0, %1;" ::"r"(bar_ln), "n"(kN2WThreads) : "memory");
      {
        const float var = (red[lane] + red[32 + lane] + red[64 + lane] + red[96 + lane]) * a.inv_latent;
        invs[lane] = 1.0f / sqrtf(var + 1e-5f);
      }
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 inv4 = *reinterpret_cast<const float4*>(invs + j4);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = j4 + t;
          const float inv = t == 0 ? inv4.x : (t == 1 ? inv4.y : (t == 2 ? inv4.z : inv4.w));
          if (j < rows) hrow[(int64_t)j * kLatent] = fmaf(ln_scale * inv, yc[j], ln_offset) + hold[j];  // gns.py:120-122
        }
      }
    }
    tc_fence_before();  // the accumulator reads of this tile are ordered before the next GEMM 1 (worker barrier)
    LB_TRACE(16);
  }

  // ================================================================ switch the resident weights
  LB_TRACE(20);
  tc_fence_before();
  __syncthreads();  // every worker has waited for its last GEMM and written its rows of h
  tc_fence_after();
  LB_TRACE(21);
  {  // this worker's first pass-B tile: its h rows -> L1 under the weight reload
    const int64_t nrow = (int64_t)((int)blockIdx.x + wk * grid) * k2Tile + r0 + (lane >> 2);
    if (nrow < a.n) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.h + nrow * kLatent + (lane & 3) * 32));
  }
  {
    const int first = kEnc ? 4 : 6;             // operand halves of pass B in the blob
    const int halves = a.last ? 2 : 4;          // decoder: one operand
    if (wk < halves) weight_to_tmem(wsrc + (size_t)(first + wk) * 2048, f, tmem + lane_sel + (uint32_t)wk * 64);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  LB_TRACE(22);
  const uint32_t ws_hi = tmem, ws_lo = tmem + 64, wr_hi = tmem + 128, wr_lo = tmem + 192;
  const bool push = a.P_left != nullptr || a.P_right != nullptr;

  // ================================================================ pass B
  for (int tile = (int)blockIdx.x + wk * grid; tile < n_tiles; tile += kN2Workers * grid) {
    const int64_t row0 = (int64_t)tile * k2Tile;
    const int rows = min(k2Tile, a.n - (int)row0);
    const uint32_t acc_s = tmem + kN2ColAccB + (uint32_t)wk * 64, acc_r = acc_s + 32;
    // ---- B0: the new h rows -> X operand; both projections (or the decoder's first layer)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + i;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) hv = reinterpret_cast<const float4*>(a.h + (row0 + r) * kLatent)[lane];
      put_row4(x_hi_p, x_lo_p, r, hv);
    }
    LB_TRACE(30);
    {  // the worker's next pass-B tile: its (just written) h rows -> L1
      const int64_t nrow = (int64_t)(tile + kN2Workers * grid) * k2Tile + r0 + (lane >> 2);
      if (nrow < a.n) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.h + nrow * kLatent + (lane & 3) * 32));
    }
    if (operand_ready()) {
      tc_fence_after();
      if (a.last) {
        issue_gemm_ts_d(ws_hi, ws_lo, x_hi, x_lo, acc_s, k2Idesc);
      } else {
        // two independent accumulate chains (sender / receiver projection), issued interleaved so that the tensor
        // pipe overlaps them instead of running one dependent chain after the other
        const uint64_t dxh = umma_desc(x_hi, kLboB), dxl = umma_desc(x_lo, kLboB);
        const uint32_t dw = (uint32_t)(dxh >> 32), xh0 = (uint32_t)dxh, xl0 = (uint32_t)dxl;
        constexpr uint32_t kStep = (2 * kLboB) >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          umma_ts_d(acc_s, ws_lo + j * 8, xh0 + j * kStep, dw, j > 0 ? 1u : 0u, k2Idesc);
          umma_ts_d(acc_r, wr_lo + j * 8, xh0 + j * kStep, dw, j > 0 ? 1u : 0u, k2Idesc);
        }
        umma_ts_d_rescale11(acc_s, ws_hi, xl0, dw, k2Idesc);
        umma_ts_d_rescale11(acc_r, wr_hi, xl0, dw, k2Idesc);
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          umma_ts_d(acc_s, ws_hi + j * 8, xl0 + j * kStep, dw, 1u, k2Idesc);
          umma_ts_d(acc_r, wr_hi + j * 8, xl0 + j * kStep, dw, 1u, k2Idesc);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          umma_ts_d(acc_s, ws_hi + j * 8, xh0 + j * kStep, dw, 1u, k2Idesc);
          umma_ts_d(acc_r, wr_hi + j * 8, xh0 + j * kStep, dw, 1u, k2Idesc);
        }
      }
      umma_commit(bar_g1);
    }
    LB_TRACE(31);
    mbar_wait(bar_g1, ph1);
    ph1 ^= 1;
    tc_fence_after();
    LB_TRACE(32);
    if (!a.last) {
      // ---- next step's sender projection P[:, 0:128] (+ the neighbours' ghost rows), receiver projection P[:, 128:256]
      float* const prow = a.P + row0 * (2 * kLatent) + f;
      {
        float ps[32];
        tmem_ld32(acc_s + lane_sel, ps);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < rows) prow[(int64_t)j * (2 * kLatent)] = ps[j];
        if (push) {  // boundary rows: the same 128-byte segment goes into the neighbour's ghost row (NVLink store)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < rows) {
              const int64_t v = row0 + j;
              if (a.P_left != nullptr) {
                const int k = __ldg(a.push_left + v);
                if (k >= 0) a.P_left[(int64_t)(a.dst_left + k) * (2 * kLatent) + f] = ps[j];
              }
              if (a.P_right != nullptr) {
                const int k = __ldg(a.push_right + v);
                if (k >= 0) a.P_right[(int64_t)(a.dst_right + k) * (2 * kLatent) + f] = ps[j];
              }
            }
        }
      }
      {
        float pr[32];
        tmem_ld32(acc_r + lane_sel, pr);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < rows) prow[(int64_t)j * (2 * kLatent) + kLatent] = pr[j] + b_next;
      }
    } else {
      // ---- decoder (gns.py:126-133): out = relu(h Wd0 + bd0) Wd1 + bd1; the reduction over the 128 features
      //      runs across threads (31-shuffle transpose-reduce, then the 4 warps of the worker)
      float hid[32];
      tmem_ld32(acc_s + lane_sel, hid);
#pragma unroll
      for (int j = 0; j < 32; ++j) hid[j] = fmaxf(hid[j] + b_next, 0.f);
      for (int k = 0; k < a.dim; ++k) {
        const float wkk = vec[640 + f * 3 + k];
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = hid[j] * wkk;
        float* rb = red + (1 + (k & 1)) * 128;  // blocks 1 / 2 alternate over k
        rb[q * 32 + lane] = warp_transpose_reduce(t);
        asm volatile("bar.sync %0, %1;" ::"r"(bar_ln), "n"(kN2WThreads) : "memory");
        if (q == 0 && lane < rows) {
          const float o = rb[lane] + rb[32 + lane] + rb[64 + lane] + rb[96 + lane] + vec[640 + 384 + k];
          a.out[(row0 + lane) * a.dim + k] = o;
          if (a.flag != nullptr && !isfinite(o)) atomicOr(a.flag, 1);  // an activation left the fp16 split's range
        }
      }
    }
    tc_fence_before();
    LB_TRACE(33);
  }
  LB_TRACE(40);
  if (push) __threadfence_system();  // peer stores are performed before the kernel completes (exchange kernel follows)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int launch_node_mp_tc2(const NodeTcArgs& a_in, cudaStream_t s) {
  static int ready[kMaxDevices];
  NodeTcArgs a = a_in;
  int rc = 0;
  const int dev = device_slot(&rc);
  if (dev < 0) return rc;
  if (!ready[dev]) {
    rc = (int)cudaFuncSetAttribute(node_mp_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kN2Smem);
    if (rc == 0)
      rc = (int)cudaFuncSetAttribute(node_mp_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kN2Smem);
    if (rc) return rc;
    ready[dev] = 1;
  }
  const int sms = device_sm_count(&rc);
  if (rc) return rc;
  // tile t belongs to CTA t % grid, worker (t / grid) % 4: a small cloud spreads over as many SMs as it has tiles
  const int n_tiles = cdiv(a.n, k2Tile);
  const int grid = n_tiles < sms ? n_tiles : sms;
  const bool pdl = pdl_enabled((int64_t)a.n * 14);  // ~ edges of a 3-D cloud: the same size switch as the message kernel
  a.pdl = (pdl || !(early_issue_mask() & 2)) ? 1 : 0;
  rc = (int)(a.enc ? launch_maybe_pdl(node_mp_tc2_kernel<true>, grid, kN2Threads, kN2Smem, s, a, pdl)
                   : launch_maybe_pdl(node_mp_tc2_kernel<false>, grid, kN2Threads, kN2Smem, s, a, pdl));
  if (rc) return rc;
  LB_LAUNCHED(1);
  return 0;
}

}  // namespace lb

#ifdef LB200_CROSSCHECK
// debug (cross-check builds): copy the phase timeline of the last node-kernel launch's CTA 0 to the host
extern "C" int lb200_debug_node_trace(long long* out_4x64x2, int* n_out4) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out_4x64x2, lb::g_node_trace, sizeof(long long) * 4 * 64 * 2);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(n_out4, lb::g_node_trace_n, sizeof(int) * 4);
  return (int)e;
}
#endif
