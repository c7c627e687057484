// Message passing, edge side, on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces update_edge_fn + segment_sum of the jraph.GraphNetwork block
// (lagrangebench/models/gns.py:86-101,117-122) for tiles of 64 receiver-sorted edges (two
// independent 128-thread workers per persistent CTA, so one worker's epilogue overlaps the
// other's MMAs and memory latency):
//
//   hidden = relu(e @ W1e + P_s[snd] + P_r[rcv])        (P = per-node projections, gns.cu)
//   yc     = hidden @ W2c + b2c                          (W2c/b2c: LayerNorm mean folded in)
//   e'     = scale * yc * rsqrt(mean(yc^2) + 1e-5) + offset
//   e     <- e' + e ;  agg[rcv] = sum over the receiver's edges of e'  (ascending slot order)
//
// fp32-level accuracy on fp16 tensor cores: every operand x is split as
// x = hi + lo * 2^-11 with hi = fp16(x), lo = fp16((x - hi) * 2^11) (22+ significant bits);
// D = A_hi B_hi + 2^-11 (A_hi B_lo + A_lo B_hi) uses three kind::f16 MMA passes with fp32
// accumulation into two TMEM accumulators (the dropped lo*lo term is 2^-22 relative).
//
// Layout: the GEMMs are issued TRANSPOSED, D^T[feature][edge] = W^T[feature][k] * X^T[k][edge]:
// weights are the M-side operand (resident in shared memory for the whole persistent CTA),
// the edge tile is the N-side operand.  TMEM lane == output feature == epilogue thread, TMEM
// column == edge.  Consequences: every global access of the epilogue (gather of P rows by
// sender/receiver index, residual read, e store, aggregate store) is one coalesced 128-byte
// row segment per warp instruction; the per-receiver segmented sum runs in registers over the
// thread's columns; only LayerNorm's variance needs a cross-thread (butterfly) reduction.
//
// Operand tiles in shared memory use the no-swizzle K-major core-matrix layout: element
// (row, k) at (k / 8) * LBO + row * 16 + (k % 8) * 2 bytes, 8-row groups 128 B apart (SBO).
#include "tc_common.cuh"

namespace lb {

// The first-generation tensor-core kernels (message kernel with the weights in shared memory, node update with
// five streamed operands) are superseded by gns_tc2.cu / node_tc2.cu.  They are compiled only into a
// cross-check build (LB200_BUILD_CROSSCHECK=1 -> -DLB200_CROSSCHECK), where the tests compare v2 against them.
#ifdef LB200_CROSSCHECK
template <bool kEnc>
__global__ void __launch_bounds__(kTcThreads, 1) edge_mp_tc_kernel(EdgeTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int wk = warp / (4 * kChunks);   // worker: 4 * kChunks warps, one kTcTile-edge MMA tile at a time
  const int c = (warp >> 2) % kChunks;   // which 32-edge chunk (carry sub-tile) of the tile this warp finishes
  const int q = warp & 3;                // TMEM lane quarter of this warp
  const int wwarp = warp % (4 * kChunks);  // warp index inside the worker
  const int wtid = tid % kWThreads;      // thread index inside the worker
  const int f = q * 32 + lane;           // output feature == TMEM lane of this thread
  float* vec = reinterpret_cast<float*>(smem + kOffVec);
  int* sidx = reinterpret_cast<int*>(smem + kOffIdx) + wk * kIdxInts;
  int* rclamp = sidx + kTcTile;   // receivers clamped to a valid row (gather addresses)
  int* ridx = rclamp + kTcTile;   // ridx[0] = receiver before the tile, ridx[1 + i] = edge i, ridx[1 + rows] = after
  float* red = reinterpret_cast<float*>(smem + kOffRed) + (wk * kChunks + c) * 128;
  float* invs = reinterpret_cast<float*>(smem + kOffInv) + warp * 32;
  uint32_t* endm = reinterpret_cast<uint32_t*>(smem + kOffEnd) + wk * kChunks;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 48);
  const uint32_t bar_w = sbase + kOffBar, bar_mma = sbase + kOffBar + 8 + 8 * wk;
  const uint32_t bar_worker = 1 + wk;                       // named barrier: the worker's threads
  const uint32_t bar_chunk = 1 + kWorkers + wk * kChunks + c;  // named barrier: the 4 warps sharing a 32-edge chunk

  const int E = a.rowptr[a.n];
  const int n_tiles = (E + kTcTile - 1) / kTcTile;
  if ((int)blockIdx.x * kWorkers >= n_tiles) return;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int w = 0; w < kWorkers; ++w) mbar_init(sbase + kOffBar + 8 + 8 * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: [0,128) acc_hh (kTcTile per worker), [128,256) acc_x
  const uint32_t acc_hh = tmem + wk * kTcTile, acc_x = tmem + 128 + wk * kTcTile;

  // resident weights: 4 fp16 operands (128 KB) + centred bias, LayerNorm scale / offset
  if (tid == 0) {
    if (kEnc) {  // only the second-layer operand pair, into the W2 slots
      mbar_expect_tx(bar_w, 2 * kWBytes + 3 * 512);
      bulk_g2s(sbase + kOffW + 2 * kWBytes, a.w_tc, 2 * kWBytes, bar_w);
    } else {
      mbar_expect_tx(bar_w, 4 * kWBytes + 3 * 512);
      bulk_g2s(sbase + kOffW, a.w_tc, 4 * kWBytes, bar_w);
    }
    bulk_g2s(sbase + kOffVec, a.vec_tc, 3 * 512, bar_w);
  }
  mbar_wait(bar_w, 0);
  const float b2c = vec[f], ln_scale = vec[128 + f], ln_offset = vec[256 + f];
  float ew0 = 0.f, ew1 = 0.f, ew2 = 0.f, ew3 = 0.f, eb0 = 0.f;  // encoder first layer, this thread's feature
  if (kEnc) {
    ew0 = a.enc_vec[f];
    ew1 = a.enc_vec[128 + f];
    ew2 = a.enc_vec[256 + f];
    ew3 = a.enc_vec[384 + f];
    eb0 = a.enc_vec[512 + f];
  }

  const uint32_t w1_hi = sbase + kOffW, w1_lo = w1_hi + kWBytes, w2_hi = w1_lo + kWBytes, w2_lo = w2_hi + kWBytes;
  // this worker's kTcTile operand rows inside every K slab
  const uint32_t b_hi = sbase + kOffB + wk * (kTcTile * 16), b_lo = b_hi + kBBytes;
  unsigned char* b_hi_p = smem + kOffB + wk * (kTcTile * 16);
  unsigned char* b_lo_p = b_hi_p + kBBytes;
  const uint32_t t_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);  // lane quarter, chunk columns
  // this thread's element (k = f) of operand row `e` (edge) of its chunk:
  const uint32_t elem_off = (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 2 + (uint32_t)(c * 32) * 16;
  uint32_t phase = 0;

  // this warp's 8 edge rows of a tile (coalesced 512 B rows); requested one tile ahead
  const int r0 = wwarp * 8;
  float4 v[8];
  auto request_rows = [&](int t) {
    const int64_t s0 = (int64_t)t * kTcTile;
    const int nr = min(kTcTile, E - (int)s0);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = r0 + i < nr ? reinterpret_cast<const float4*>(a.e + (s0 + r0 + i) * kLatent)[lane]
                         : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  const int tile_stride = gridDim.x * kWorkers;

  for (int tile = blockIdx.x * kWorkers + wk; tile < n_tiles; tile += tile_stride) {
    const int64_t slot0 = (int64_t)tile * kTcTile;
    const int rows = min(kTcTile, E - (int)slot0);
    if (!kEnc) {
      request_rows(tile);
      // pull the next tile's rows into L2 now: its demand loads then skip the HBM round trip
      const int nt = tile + tile_stride;
      if (nt < n_tiles) {
        const float* nrow = a.e + ((int64_t)nt * kTcTile + r0) * kLatent + lane * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (int64_t)i * kLatent));
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(kWThreads) : "memory");  // previous tile is done with idx / operands
    if (wtid < kTcTile) {
      const bool ok = wtid < rows;
      const int r_here = ok ? a.rcv[slot0 + wtid] : -1;
      const int r_next = (slot0 + wtid + 1 < E) ? a.rcv[slot0 + wtid + 1] : -3;
      sidx[wtid] = ok ? (kEnc ? (a.perm ? a.perm[slot0 + wtid] : (int)(slot0 + wtid)) : a.snd[slot0 + wtid]) : 0;
      rclamp[wtid] = max(r_here, 0);
      ridx[1 + wtid] = ok ? r_here : (wtid == rows ? -3 : -1);  // -3: "no edge after the tile"
      // last edge of its receiver bucket inside its 32-edge carry sub-tile
      const bool end = ok && (r_next != r_here || lane == 31 || wtid == rows - 1);
      const uint32_t m = __ballot_sync(0xffffffffu, end);
      if (lane == 0) endm[wtid >> 5] = m;  // one mask per 32-edge chunk
      if (wtid == kTcTile - 1 && ok) ridx[1 + kTcTile] = r_next;  // receiver just after a full tile
    } else if (wtid == kTcTile) {
      ridx[0] = slot0 > 0 ? a.rcv[slot0 - 1] : -2;
    }
    if constexpr (kEnc) {
      // ---- encoder: first layer (K = dim + 1 <= 4) on CUDA cores straight into the layer-2 operand
      float4* feat_s = reinterpret_cast<float4*>(smem + kOffFeat) + wk * kTcTile;
      if (wtid < kTcTile)  // the tile's edge features, gathered once through perm (list order -> slot order)
        feat_s[wtid] = wtid < rows ? a.edge_feat[a.perm ? a.perm[slot0 + wtid] : slot0 + wtid] : make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(kWThreads) : "memory");
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float4 ft = feat_s[c * 32 + j];
        float v = ft.x * ew0;
        v = fmaf(ft.y, ew1, v);
        v = fmaf(ft.z, ew2, v);
        v = fmaf(ft.w, ew3, v);
        const float hval = fmaxf(v + eb0, 0.f);
        const __half hi = __float2half_rn(hval);
        const __half lo = __float2half_rn((hval - __half2float(hi)) * kLoScale);
        *reinterpret_cast<__half*>(b_hi_p + elem_off + (uint32_t)j * 16) = hi;
        *reinterpret_cast<__half*>(b_lo_p + elem_off + (uint32_t)j * 16) = lo;
      }
    } else {
    // ---- phase A: edge latents (already in registers) -> fp16 hi/lo N-side operand
    {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __half2 h01 = __floats2half2_rn(v[i].x, v[i].y), h23 = __floats2half2_rn(v[i].z, v[i].w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v[i].x - f01.x) * kLoScale, (v[i].y - f01.y) * kLoScale);
        const __half2 l23 = __floats2half2_rn((v[i].z - f23.x) * kLoScale, (v[i].w - f23.y) * kLoScale);
        const uint32_t off = (uint32_t)(lane >> 1) * kLboB + (uint32_t)(r0 + i) * 16 + (uint32_t)(lane & 1) * 8;
        *reinterpret_cast<uint2*>(b_hi_p + off) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(b_lo_p + off) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
    }
    fence_async_smem();
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(kWThreads) : "memory");
    if (wwarp == 0) {  // the worker's first warp issues (one elected lane per instruction)
      tc_fence_after();
      issue_gemm(w1_hi, w1_lo, b_hi, b_lo, acc_hh, acc_x);
      umma_commit(bar_mma);
    }
    // ---- epilogue 1: hidden = relu(acc + P_s[snd] + P_r[rcv]) -> fp16 hi/lo operand for layer 2.
    //      The gathers depend only on the indices: the first batch is in flight while the MMAs run,
    //      and batch b+1 is issued before batch b is consumed.
    {
      const int* sp = sidx + c * 32;
      const int* rp = rclamp + c * 32;
      float ps[32], pr[32];
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 4) {
        const int4 s4 = *reinterpret_cast<const int4*>(sp + j0);
        const int4 r4 = *reinterpret_cast<const int4*>(rp + j0);
        ps[j0 + 0] = __ldg(a.P + (int64_t)s4.x * (2 * kLatent) + f);
        ps[j0 + 1] = __ldg(a.P + (int64_t)s4.y * (2 * kLatent) + f);
        ps[j0 + 2] = __ldg(a.P + (int64_t)s4.z * (2 * kLatent) + f);
        ps[j0 + 3] = __ldg(a.P + (int64_t)s4.w * (2 * kLatent) + f);
        pr[j0 + 0] = __ldg(a.P + (int64_t)r4.x * (2 * kLatent) + kLatent + f);
        pr[j0 + 1] = __ldg(a.P + (int64_t)r4.y * (2 * kLatent) + kLatent + f);
        pr[j0 + 2] = __ldg(a.P + (int64_t)r4.z * (2 * kLatent) + kLatent + f);
        pr[j0 + 3] = __ldg(a.P + (int64_t)r4.w * (2 * kLatent) + kLatent + f);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float hh[16], xx[16];
        tmem_ld_pair16(acc_hh + t_addr + h * 16, acc_x + t_addr + h * 16, hh, xx);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int e = h * 16 + j;
          const float hval = fmaxf(fmaf(xx[j], kLoInv, hh[j]) + ps[e] + pr[e], 0.f);
          const __half hi = __float2half_rn(hval);
          const __half lo = __float2half_rn((hval - __half2float(hi)) * kLoScale);
          *reinterpret_cast<__half*>(b_hi_p + elem_off + (uint32_t)e * 16) = hi;
          *reinterpret_cast<__half*>(b_lo_p + elem_off + (uint32_t)e * 16) = lo;
        }
      }
    }
    }
    fence_async_smem();
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(kWThreads) : "memory");
    if (wwarp == 0) {
      tc_fence_after();
      issue_gemm(w2_hi, w2_lo, b_hi, b_lo, acc_hh, acc_x);
      umma_commit(bar_mma);
    }
    // ---- epilogue 2: LayerNorm (mean folded into the weights), residual, store, segmented sum.
    //      Each warp finishes its own 32-edge chunk == one carry sub-tile.  The residual rows are
    //      requested before waiting for the MMAs.
    {
      const int col0 = c * 32;
      const int valid = min(max(rows - col0, 0), 32);  // edges of this chunk that exist
      float* const erow = a.e + (slot0 + col0) * kLatent + f;
      float eold[32];
      if (kEnc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) eold[j] = 0.f;
      } else if (valid == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) eold[j] = erow[(int64_t)j * kLatent];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) eold[j] = j < valid ? erow[(int64_t)j * kLatent] : 0.f;
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      float yc[32];
      float part;
      {
        float sq[32];
        tmem_ld_pair(acc_hh + t_addr, acc_x + t_addr, yc, sq);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          yc[j] = fmaf(sq[j], kLoInv, yc[j]) + b2c;
          sq[j] = yc[j] * yc[j];
        }
        part = warp_transpose_reduce(sq);  // lane l: this warp's 32 features, edge col0 + l
      }
      red[q * 32 + lane] = part;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_chunk) : "memory");
      {
        const float var = (red[lane] + red[32 + lane] + red[64 + lane] + red[96 + lane]) * (1.0f / kLatent);
        invs[lane] = 1.0f / sqrtf(var + 1e-5f);  // once per edge per warp (same value in the 4 warps)
      }
      __syncwarp();
      if (valid > 0) {
        const uint32_t emask = endm[c];
        const bool first_cont = ridx[col0] == ridx[1 + col0];
        const bool last_cont = ridx[1 + col0 + valid] == ridx[col0 + valid];
        const int sub = tile * kChunks + c;  // carry sub-tile index (slot / kEdgeTile)
        float* const cfirst = a.carry_first + (int64_t)sub * kLatent + f;
        float* const clast = a.carry_last + (int64_t)sub * kLatent + f;
        float seg_sum = 0.f;
        bool seg_first = true;  // still inside the first bucket of the sub-tile
        auto finish = [&](auto full_tag) {
          constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 inv4 = *reinterpret_cast<const float4*>(invs + j4);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int j = j4 + t;
              const float inv = t == 0 ? inv4.x : (t == 1 ? inv4.y : (t == 2 ? inv4.z : inv4.w));
              const float msg = fmaf(ln_scale * inv, yc[j], ln_offset);  // e' : the message
              if (kFull || j < valid) {
                erow[(int64_t)j * kLatent] = msg + eold[j];  // residual (gns.py:120-122)
                seg_sum += msg;
                if (!kEnc && ((emask >> j) & 1u)) {  // bucket ends here (uniform across the chunk's 4 warps)
                  float* dst = a.agg + (int64_t)ridx[1 + col0 + j] * kLatent + f;
                  if (j == valid - 1 && last_cont) dst = clast;
                  if (seg_first && first_cont) dst = cfirst;
                  *dst = seg_sum;
                  seg_sum = 0.f;
                  seg_first = false;
                }
              }
            }
          }
        };
        if (valid == 32)
          finish(std::true_type{});
        else
          finish(std::false_type{});
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

// =====================================================================================
// Node update on the tensor cores (replaces update_node_fn + residual of gns.py:103-122 and
// the projection of the next step's first edge layer / the decoder of gns.py:126-133).
//
// One persistent CTA (512 threads) per 128-node tile; GEMMs transposed as in the edge kernel
// (TMEM lane = feature, column = node).  The node tile stays resident in shared memory as
// hi/lo fp16 operands while the five 128x128 weight operands of the step (64 KB each, hi|lo)
// are streamed through one shared-memory buffer by cp.async.bulk:
//   W1[0:128] (acts on h), W1[128:256] (acts on the aggregate), W2c (LayerNorm mean folded in),
//   then either the two halves of the next step's edge W1 (-> P) or the decoder's first layer.
constexpr int kNtThreads = 512;
constexpr int kNtTile = 128;
constexpr uint32_t kIdescN128 = (1u << 4) | ((uint32_t)(kNtTile >> 3) << 17) | (8u << 24);
constexpr uint32_t kNOffW = 0;                          // streamed weight operand: hi | lo
constexpr uint32_t kNOffBh = kNOffW + 2 * kWBytes;      // h (later h_new) operand: hi | lo
constexpr uint32_t kNOffBa = kNOffBh + 2 * kBBytes;     // aggregate (later hidden) operand: hi | lo
constexpr uint32_t kNOffVec = kNOffBa + 2 * kBBytes;    // b1 | b2c | ln_scale | ln_offset | b_next | wd1[128*3] | bd1[4]
constexpr uint32_t kNVecFloats = 5 * 128 + 3 * 128 + 4;
constexpr uint32_t kNOffRed = kNOffVec + ((kNVecFloats * 4 + 15) / 16) * 16;  // red[4 groups][4 warps][32] x 3
constexpr uint32_t kNOffInv = kNOffRed + 3 * 4 * 4 * 32 * 4;                   // inv[16 warps][32]
constexpr uint32_t kNOffBar = kNOffInv + 16 * 32 * 4;
constexpr uint32_t kSmemNodeTc = kNOffBar + 48;

// kEnc: node ENCODER mode (NodeTcArgs::enc) as a compile-time switch, so that the message-passing
// instantiation is exactly the kernel it was before the encoder existed
template <bool kEnc>
__global__ void __launch_bounds__(kNtThreads, 1) node_mp_tc_kernel(NodeTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int q = warp & 3;    // TMEM lane quarter
  const int g = warp >> 2;   // 32-node column group of the tile this warp finishes
  const int f = q * 32 + lane;
  float* vec = reinterpret_cast<float*>(smem + kNOffVec);
  float* red = reinterpret_cast<float*>(smem + kNOffRed) + g * 128;
  float* invs = reinterpret_cast<float*>(smem + kNOffInv) + warp * 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kNOffBar + 32);
  const uint32_t bar_w = sbase + kNOffBar, bar_mma = sbase + kNOffBar + 8, bar_mid = sbase + kNOffBar + 16;
  const uint32_t bar_group = 1 + g;
  const int n_tiles = (a.n + kNtTile - 1) / kNtTile;
  if ((int)blockIdx.x >= n_tiles) return;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_mid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < (int)kNVecFloats; i += kNtThreads) vec[i] = a.vec_tc[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc_hh = tmem, acc_x = tmem + 128;
  const float b1 = vec[f], b2c = vec[128 + f], ln_scale = vec[256 + f], ln_offset = vec[384 + f], b_next = vec[512 + f];

  const uint32_t w_hi = sbase + kNOffW, w_lo = w_hi + kWBytes;
  const uint32_t bh_hi = sbase + kNOffBh, bh_lo = bh_hi + kBBytes;
  const uint32_t ba_hi = sbase + kNOffBa, ba_lo = ba_hi + kBBytes;
  unsigned char* bh_hi_p = smem + kNOffBh;
  unsigned char* bh_lo_p = bh_hi_p + kBBytes;
  unsigned char* ba_hi_p = smem + kNOffBa;
  unsigned char* ba_lo_p = ba_hi_p + kBBytes;
  const uint32_t t_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32);
  const uint32_t elem_off = (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 2 + (uint32_t)(g * 32) * 16;
  const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.w_tc);
  uint32_t ph_w = 0, ph_mma = 0, ph_mid = 0;  // ph_w / ph_mid are only used by warp 0

  auto load_w = [&](int k) {  // warp 0 (one elected lane): stream weight operand k (hi|lo, 64 KB) into the buffer
    if (elect_one()) {
      mbar_expect_tx(bar_w, 2 * kWBytes);
      bulk_g2s(w_hi, wsrc + (size_t)k * 2 * kWBytes, 2 * kWBytes, bar_w);
    }
  };
  auto wait_w = [&]() {
    mbar_wait(bar_w, ph_w);
    ph_w ^= 1;
  };
  auto wait_all = [&]() {  // every thread: the MMAs committed to bar_mma are done
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
  };
  // split a float and store it as this thread's element (k = f) of operand row `col` (its column group)
  auto put_elem = [&](unsigned char* hi_p, unsigned char* lo_p, int col, float v) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
    *reinterpret_cast<__half*>(hi_p + elem_off + (uint32_t)col * 16) = hi;
    *reinterpret_cast<__half*>(lo_p + elem_off + (uint32_t)col * 16) = lo;
  };
  auto put_row4 = [&](unsigned char* hi_p, unsigned char* lo_p, int r, const float4& v) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((v.x - f01.x) * kLoScale, (v.y - f01.y) * kLoScale);
    const __half2 l23 = __floats2half2_rn((v.z - f23.x) * kLoScale, (v.w - f23.y) * kLoScale);
    const uint32_t off = (uint32_t)(lane >> 1) * kLboB + (uint32_t)r * 16 + (uint32_t)(lane & 1) * 8;
    *reinterpret_cast<uint2*>(hi_p + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(lo_p + off) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  };

  if (warp == 0) load_w(0);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * kNtTile;
    const int rows = min(kNtTile, a.n - (int)row0);
    const int valid = min(max(rows - g * 32, 0), 32);  // nodes of this warp's column group that exist
    // ---- phase 0: h and the aggregate (bucket carries resolved) -> hi/lo operands, 8 rows per warp
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = warp * 8 + i;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), av = hv;
      if (r < rows) {
        const int64_t v = row0 + r;
        hv = reinterpret_cast<const float4*>(a.h + v * kLatent)[lane];
        const int e0 = kEnc ? 0 : a.rowptr[v], e1 = kEnc ? 0 : a.rowptr[v + 1];
        if (e1 > e0) {
          const int ta = e0 / kEdgeTile, tb = (e1 - 1) / kEdgeTile;
          if (ta == tb) {
            av = reinterpret_cast<const float4*>(a.agg + v * kLatent)[lane];
          } else {  // bucket straddles carry sub-tiles: partial sums in slot order
            av = reinterpret_cast<const float4*>(a.carry_last + (int64_t)ta * kLatent)[lane];
            for (int t = ta + 1; t <= tb; ++t) {
              const float4 p = reinterpret_cast<const float4*>(a.carry_first + (int64_t)t * kLatent)[lane];
              av.x += p.x; av.y += p.y; av.z += p.z; av.w += p.w;
            }
          }
        }
      }
      put_row4(bh_hi_p, bh_lo_p, r, hv);
      if (!kEnc) put_row4(ba_hi_p, ba_lo_p, r, av);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 1: acc = W1h^T h + W1a^T agg   (two streamed operands)
    if (warp == 0) {
      tc_fence_after();
      wait_w();
      issue_gemm(w_hi, w_lo, bh_hi, bh_lo, acc_hh, acc_x, kIdescN128, false);
      if (kEnc) {  // encoder: a single input operand
        umma_commit(bar_mma);
      } else {
        umma_commit(bar_mid);
        mbar_wait(bar_mid, ph_mid);
        ph_mid ^= 1;
        load_w(1);
        wait_w();
        tc_fence_after();
        issue_gemm(w_hi, w_lo, ba_hi, ba_lo, acc_hh, acc_x, kIdescN128, true);
        umma_commit(bar_mma);
      }
    }
    wait_all();
    if (warp == 0) load_w(kEnc ? 1 : 2);  // W2c streams in while the hidden layer is written
    // hidden = relu(acc + b1) -> operand (over the aggregate, which is no longer needed)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float hh[16], xx[16];
      tmem_ld_pair16(acc_hh + t_addr + hf * 16, acc_x + t_addr + hf * 16, hh, xx);
#pragma unroll
      for (int j = 0; j < 16; ++j) put_elem(ba_hi_p, ba_lo_p, hf * 16 + j, fmaxf(fmaf(xx[j], kLoInv, hh[j]) + b1, 0.f));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2 (+ LayerNorm, residual)
    if (warp == 0) {
      tc_fence_after();
      wait_w();
      issue_gemm(w_hi, w_lo, ba_hi, ba_lo, acc_hh, acc_x, kIdescN128, false);
      umma_commit(bar_mma);
    }
    float* const hrow = a.h + (row0 + g * 32) * kLatent + f;
    float hold[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) hold[j] = (j < valid && !kEnc) ? hrow[(int64_t)j * kLatent] : 0.f;  // no residual in the encoder
    wait_all();
    if (warp == 0) load_w(kEnc ? 2 : 3);
    {
      float yc[32];
      float part;
      {
        float sq[32];
        tmem_ld_pair(acc_hh + t_addr, acc_x + t_addr, yc, sq);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          yc[j] = fmaf(sq[j], kLoInv, yc[j]) + b2c;
          sq[j] = yc[j] * yc[j];
        }
        part = warp_transpose_reduce(sq);
      }
      red[q * 32 + lane] = part;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_group) : "memory");
      {
        const float var = (red[lane] + red[32 + lane] + red[64 + lane] + red[96 + lane]) * (1.0f / kLatent);
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
          const float hnew = fmaf(ln_scale * inv, yc[j], ln_offset) + hold[j];  // residual (gns.py:120-122)
          if (!a.last && j < valid) hrow[(int64_t)j * kLatent] = hnew;
          put_elem(bh_hi_p, bh_lo_p, j, j < valid ? hnew : 0.f);
        }
      }
      __syncwarp();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- next step's sender projection P[:, 0:128], or the decoder's hidden layer
    if (warp == 0) {
      tc_fence_after();
      wait_w();
      issue_gemm(w_hi, w_lo, bh_hi, bh_lo, acc_hh, acc_x, kIdescN128, false);
      umma_commit(bar_mma);
    }
    wait_all();
    if (!a.last) {
      if (warp == 0) load_w(kEnc ? 3 : 4);
      float* const prow = a.P + (row0 + g * 32) * (2 * kLatent) + f;
      const bool push = a.P_left != nullptr || a.P_right != nullptr;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float hh[16], xx[16];
        tmem_ld_pair16(acc_hh + t_addr + hf * 16, acc_x + t_addr + hf * 16, hh, xx);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (hf * 16 + j < valid) {
            const float val = fmaf(xx[j], kLoInv, hh[j]);
            prow[(int64_t)(hf * 16 + j) * (2 * kLatent)] = val;
            if (push) {  // boundary rows: the same 128-byte segment goes into the neighbour's ghost row (NVLink store)
              const int64_t v = row0 + g * 32 + hf * 16 + j;
              if (a.P_left != nullptr) {
                const int k = __ldg(a.push_left + v);
                if (k >= 0) a.P_left[(int64_t)(a.dst_left + k) * (2 * kLatent) + f] = val;
              }
              if (a.P_right != nullptr) {
                const int k = __ldg(a.push_right + v);
                if (k >= 0) a.P_right[(int64_t)(a.dst_right + k) * (2 * kLatent) + f] = val;
              }
            }
          }
      }
      if (push) __threadfence_system();  // peer stores are performed before the kernel completes (exchange kernel follows)
      tc_fence_before();
      __syncthreads();  // every warp has drained the accumulators
      // ---- receiver projection P[:, 128:256] = h W1r + b1(next)
      if (warp == 0) {
        tc_fence_after();
        wait_w();
        issue_gemm(w_hi, w_lo, bh_hi, bh_lo, acc_hh, acc_x, kIdescN128, false);
        umma_commit(bar_mma);
      }
      wait_all();
      if (warp == 0 && tile + (int)gridDim.x < n_tiles) load_w(0);  // next tile's first operand
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float hh[16], xx[16];
        tmem_ld_pair16(acc_hh + t_addr + hf * 16, acc_x + t_addr + hf * 16, hh, xx);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (hf * 16 + j < valid)
            prow[(int64_t)(hf * 16 + j) * (2 * kLatent) + kLatent] = fmaf(xx[j], kLoInv, hh[j]) + b_next;
      }
    } else {
      if (warp == 0 && tile + (int)gridDim.x < n_tiles) load_w(0);
      // decoder (gns.py:126-133): out = relu(h Wd0 + bd0) Wd1 + bd1; the reduction over the 128
      // features runs across threads (31-shuffle transpose-reduce, then the 4 warps of the group)
      float hid[32];
      {
        float xx[32];
        tmem_ld_pair(acc_hh + t_addr, acc_x + t_addr, hid, xx);
#pragma unroll
        for (int j = 0; j < 32; ++j) hid[j] = fmaxf(fmaf(xx[j], kLoInv, hid[j]) + b_next, 0.f);
      }
      for (int k = 0; k < a.dim; ++k) {
        const float wk = vec[640 + f * 3 + k];
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = hid[j] * wk;
        red[(1 + (k & 1)) * 512 + q * 32 + lane] = warp_transpose_reduce(t);  // red block 1/2 alternate over k
        asm volatile("bar.sync %0, 128;" ::"r"(bar_group) : "memory");
        if (q == 0 && lane < valid) {
          const float* rb = red + (1 + (k & 1)) * 512;
          const float o = rb[lane] + rb[32 + lane] + rb[64 + lane] + rb[96 + lane] + vec[640 + 384 + k];
          a.out[(row0 + g * 32 + lane) * a.dim + k] = o;
          if (a.flag != nullptr && !isfinite(o)) atomicOr(a.flag, 1);
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // operands and accumulators are free for the next tile
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

#endif  // LB200_CROSSCHECK

__global__ void node_embed_kernel(const float* __restrict__ node_feat, int node_in, int node_stride,
                                  const int32_t* __restrict__ ptype, const float* __restrict__ embedding, int embed,
                                  int n_types, int n, float* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * kLatent) return;
  const int64_t r = i / kLatent;
  const int c = (int)(i % kLatent);
  float v = 0.f;
  if (c < node_in) {
    v = node_feat[r * node_stride + c];
  } else if (c < node_in + embed) {
    int t = ptype[r];
    if (t < 0) t += n_types;  // hk.Embed indexes like NumPy: PAD_VALUE (-1) is the last row
    t = min(max(t, 0), n_types - 1);
    v = embedding[t * embed + (c - node_in)];
  }
  h[i] = v;
}

int launch_node_embed(const float* node_feat, int node_in, int node_stride, const int32_t* ptype, const float* embedding,
                      int embed, int n_types, int n, float* h, cudaStream_t s) {
  const int64_t total = (int64_t)n * kLatent;
  node_embed_kernel<<<cdiv(total, 256), 256, 0, s>>>(node_feat, node_in, node_stride, ptype, embedding, embed, n_types, n, h);
  LB_LAUNCHED(1);
  return 0;
}

#ifdef LB200_CROSSCHECK
int launch_edge_mp_tc(const EdgeTcArgs& a, int e_cap, cudaStream_t s) {
  static int ready[kMaxDevices];
  int rc = 0;
  const int dev = device_slot(&rc);
  if (dev < 0) return rc;
  if (!ready[dev]) {
    rc = (int)cudaFuncSetAttribute(edge_mp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc);
    if (rc == 0)
      rc = (int)cudaFuncSetAttribute(edge_mp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc);
    if (rc) return rc;
    ready[dev] = 1;
  }
  const int sms = device_sm_count(&rc);
  if (rc) return rc;
  const int n_groups = cdiv(cdiv(e_cap, kTcTile), kWorkers);  // kWorkers tiles in flight per CTA
  const int grid = n_groups < sms ? n_groups : sms;
  if (a.encoder) {
    edge_mp_tc_kernel<true><<<grid, kTcThreads, kSmemTc, s>>>(a);
  } else {
    edge_mp_tc_kernel<false><<<grid, kTcThreads, kSmemTc, s>>>(a);
  }
  LB_LAUNCHED(1);
  return 0;
}

int launch_node_mp_tc(const NodeTcArgs& a, cudaStream_t s) {
  static int ready[kMaxDevices];
  int rc = 0;
  const int dev = device_slot(&rc);
  if (dev < 0) return rc;
  if (!ready[dev]) {
    rc = (int)cudaFuncSetAttribute(node_mp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNodeTc);
    if (rc == 0)
      rc = (int)cudaFuncSetAttribute(node_mp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNodeTc);
    if (rc) return rc;
    ready[dev] = 1;
  }
  const int sms = device_sm_count(&rc);
  if (rc) return rc;
  const int n_tiles = cdiv(a.n, kNtTile);
  const int grid = n_tiles < sms ? n_tiles : sms;
  if (a.enc) {
    node_mp_tc_kernel<true><<<grid, kNtThreads, kSmemNodeTc, s>>>(a);
  } else {
    node_mp_tc_kernel<false><<<grid, kNtThreads, kSmemNodeTc, s>>>(a);
  }
  LB_LAUNCHED(1);
  return 0;
}

#else
int launch_edge_mp_tc(const EdgeTcArgs&, int, cudaStream_t) { return LB200_EUNSUPPORTED; }
int launch_node_mp_tc(const NodeTcArgs&, cudaStream_t) { return LB200_EUNSUPPORTED; }
#endif  // LB200_CROSSCHECK

}  // namespace lb
