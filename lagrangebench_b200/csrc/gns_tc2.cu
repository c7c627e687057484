// Message kernel v2 (tcgen05 + TMEM), sm_100a: the processor's edge update + segmented sum of
// lagrangebench/models/gns.py:86-101,117-122 with
//   * the four 128x128 fp16 weight operands (W1e hi|lo, W2c hi|lo) resident in TENSOR MEMORY for
//     the whole persistent CTA (tcgen05.mma with the A operand in TMEM, 256 columns), which frees
//     128 KB of shared memory for
//   * double-buffered edge operands and accumulators, so that each 128-thread worker runs a
//     two-tile software pipeline:  E1(k) | A(k+1) | E2(k)  -- every GEMM overlaps the phase that
//     follows its issue instead of being waited for;
//   * ONE accumulator per GEMM: the cross terms A_hi B_lo' + A_lo' B_hi (lo' = lo * 2^11) are
//     accumulated first, and the first A_hi B_hi instruction rescales them with
//     scale-input-d = 11 (D = A B + D * 2^-11), so the epilogues read half the TMEM columns;
//   * the next tile's indices prefetched into registers one phase ahead, its fp32 rows brought in by
//     one 16 KB bulk async copy (cp.async.bulk + mbarrier) a pipeline round ahead, and the P rows it
//     will gather prefetched into L1.
// Math, operand layouts and the carry protocol are those of v1 (gns_tc.cu).
//
//   hidden = relu(e @ W1e + P_s[snd] + P_r[rcv])        (P = per-node projections)
//   yc     = hidden @ W2c + b2c                          (LayerNorm mean folded into W2c / b2c)
//   e'     = scale * yc * rsqrt(mean(yc^2) + 1e-5) + offset
//   e     <- e' + e ;  agg[rcv] = sum over the receiver's edges of e'  (ascending slot order)
#include <stdlib.h>

#include "tc_common.cuh"

namespace lb {

constexpr int k2Threads = 512;
constexpr int k2Workers = 4;
constexpr int kIssuer2 = 3;           // warp (of a worker) that issues GEMM 2
constexpr int k2DefaultL2Mode = 1;  // LB200_L2HINT=0 switches the eviction hints off (A/B)
constexpr int k2DefaultVariant = 15;  // see launch_edge_mp_tc2
constexpr int k2WThreads = k2Threads / k2Workers;

// TMEM columns: [0,256) weights (64 columns per 128x128 fp16 operand), [256,512) accumulators
constexpr uint32_t k2ColW1Hi = 0, k2ColW1Lo = 64, k2ColW2Hi = 128, k2ColW2Lo = 192, k2ColAcc = 256;

// shared memory map (bytes)
constexpr uint32_t k2OffB = 0;                                      // [buf 2][hi | lo], kBBytes each
constexpr uint32_t k2OffVec = k2OffB + 4 * kBBytes;                 // b2c[128], scale[128], offset[128]
constexpr uint32_t k2IdxInts = 32 + 32 + 36;                        // sidx[32], rclamp[32], ridx[34 (+2)]
constexpr uint32_t k2OffIdx = k2OffVec + 3 * 512;                   // [worker][3 bufs][k2IdxInts]
constexpr uint32_t k2OffRed = k2OffIdx + k2Workers * 3 * k2IdxInts * 4;  // [worker][2][4 warps][32]
constexpr uint32_t k2OffInv = k2OffRed + k2Workers * 2 * 128 * 4;   // [16 warps][32]
constexpr uint32_t k2OffEnd = k2OffInv + 16 * 32 * 4;               // [worker][4] end masks (3 used)
constexpr uint32_t k2OffBar = k2OffEnd + 64;                        // mbarriers: done_g1[4], done_g2[4], staged[4]
constexpr uint32_t k2OffState = k2OffBar + 16 * 8;                  // tmem base address
constexpr uint32_t k2OffStage = ((k2OffState + 48 + 127) / 128) * 128;  // [worker] fp32 edge-latent tile (TMA bulk copy)
constexpr uint32_t k2StageBytes = k2Tile * kLatent * 4;
constexpr uint32_t k2Smem = k2OffStage + k2Workers * k2StageBytes;

// kMn:      the hidden operand of GEMM 2 is written edge-contiguous (MN-major core matrices): a thread
//           packs 8 consecutive edges of its feature into ONE 16-byte store (fp16 pairs converted two
//           at a time) instead of 8 two-byte stores; same bytes, LBO / SBO as the K-major edge operand.
// kNoScale: the activations' low halves are stored unscaled (lo = fp16(x - hi), fp16 subnormals keep
//           them exact enough); only the weights' low halves carry the 2^11 factor, so the split costs
//           one multiply less per element.  GEMM order: A_lo' B_hi, rescale, A_hi B_lo, A_hi B_hi.
// kStage:   the fp32 edge-latent tile of phase A arrives by ONE bulk async copy (cp.async.bulk, 16 KB,
//           completion on an mbarrier) issued a whole pipeline round earlier -- right after the previous
//           tile's operand was built -- instead of 8 LDG.128 per thread followed by a wait.
// kPref:    prefetch.global.L1 of this warp's 128-byte segments of the tile's P rows, issued in phase A,
//           one phase before E1 gathers them.  Pays only together with kStage: the edge tiles then no
//           longer pass through L1 and the prefetched lines survive until they are used.
// Measured on LDC-3D 28k (us per launch): v1 176 -> TMEM weights + pipeline 149 -> kMn + kNoScale 141
// -> kStage 138 -> kPref 131 -> L2 evict-first hints on the streamed latents + alternating direction 128.6
// (53 % of the measured HBM roofline).  Tried and dropped (no gain):
// contiguous tile ranges per CTA; "last arriving warp issues the GEMM" instead of bar.sync (a dedicated
// 17th MMA warp is worse still: 5 warps on one SM sub-partition cap every thread at 96 registers);
// requesting the residual rows before phase A; ld.global.L1::no_allocate for the residual rows (slower).
// Round 2, measured and dropped as well (LDC-3D 28k, 132.1 us base): bar.arrive for the three non-issuing warps
// at "operand complete" with two alternating barrier ids (133.1 us); loading P_r[rcv] only at bucket starts
// through predicated loads (143.5 us: the select chain costs more issue slots than the L1 wavefronts it saves);
// deferring the issue of GEMM 1 into E2, behind the residual loads, where the issuing warp would otherwise wait for
// GEMM 2 (same-box A/B: 128.9 vs 128.8 us); requesting the next tile's indices after the proxy fence of phase A instead
// of before it (134.9 vs 129.9 us); GEMM 1 issued by warp 2 and the bulk copy started by warp 1, so that warp 0 keeps only
// the index work (130.7 vs 129.6 us).
#ifdef LB200_CROSSCHECK
// phase timeline of CTA 0, worker 0, its four warps (cross-check builds): (id, SM clock) pairs of
// the pipeline iterations 8..11; lb200_debug_edge_trace reads it
__device__ long long g_edge_trace[4][96][2];
__device__ int g_edge_trace_n[4];
#define ET(id)                                                                                          \
  do {                                                                                                  \
    if (!kEnc && trace_it && lane == 0 && g_edge_trace_n[q] < 96) {                                     \
      const int _i = g_edge_trace_n[q]++;                                                               \
      g_edge_trace[q][_i][0] = (id);                                                                    \
      g_edge_trace[q][_i][1] = clock64();                                                               \
    }                                                                                                   \
  } while (0)
#else
#define ET(id) do { } while (0)
#endif

template <bool kEnc, bool kMn, bool kNoScale, bool kStage, bool kPref>
__global__ void __launch_bounds__(k2Threads, 1) edge_mp_tc2_kernel(EdgeTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int wk = warp >> 2;              // worker: 4 warps, a two-tile software pipeline
  const int q = warp & 3;                // TMEM lane quarter of this warp == warp index inside the worker
  const int f = q * 32 + lane;           // output feature == TMEM lane of this thread
  float* vec = reinterpret_cast<float*>(smem + k2OffVec);
  int* idx_base = reinterpret_cast<int*>(smem + k2OffIdx) + wk * 3 * k2IdxInts;
  float* red_base = reinterpret_cast<float*>(smem + k2OffRed) + wk * 256;
  float* invs = reinterpret_cast<float*>(smem + k2OffInv) + warp * 32;
  uint32_t* endm = reinterpret_cast<uint32_t*>(smem + k2OffEnd) + wk * 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + k2OffState + 32);
  const uint32_t bar_g1 = sbase + k2OffBar + 8 * wk, bar_g2 = bar_g1 + 32, bar_st = bar_g1 + 64;
  const uint32_t bar_worker = 1 + wk;            // named barriers: the worker's 128 threads
  const uint32_t bar_ln = 1 + k2Workers + wk;

  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel's tail (it waits before reading)
#ifdef LB200_CROSSCHECK
  bool trace_it = false;
  if (!kEnc && blockIdx.x == 0 && wk == 0 && lane == 0) g_edge_trace_n[q] = 0;
#endif
  if (tid == 0) {
    for (int w = 0; w < 3 * k2Workers; ++w) mbar_init(sbase + k2OffBar + 8 * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 3 * 128; i += blockDim.x) vec[i] = a.vec_tc[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // resident weights -> TMEM: warp group g (4 warps, one per lane quarter) loads operand g.  CTA-wide (barrier).
  auto load_weights = [&]() {
    if (!kEnc || wk >= 2)  // encoder: w_tc holds only the second-layer pair, which goes to the W2 slots
      weight_to_tmem(reinterpret_cast<const uint4*>(a.w_tc) + (kEnc ? wk - 2 : wk) * 2048, f,
                     tmem + ((uint32_t)(q * 32) << 16) + wk * 64);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };
  // Programmatic dependent launch (small clouds): constants (the weights) first, then wait for the previous kernel.
  // Otherwise the first tile's rows and indices are requested first and travel while the weights are loaded.
  if (a.pdl) {
    load_weights();
    pdl_wait();
  }
  const int E = a.rowptr[a.n];
  const int n_tiles = (E + k2Tile - 1) / k2Tile;
  // this CTA's tiles: worker wk takes t_begin + wk + i * tile_stride
  const int t_begin = (int)blockIdx.x * k2Workers;

  const uint32_t w1_hi = tmem + k2ColW1Hi, w1_lo = tmem + k2ColW1Lo, w2_hi = tmem + k2ColW2Hi, w2_lo = tmem + k2ColW2Lo;
  const int tile_stride = (int)gridDim.x * k2Workers;
  // Consecutive launches walk the edge array in opposite directions: a launch starts with the tiles the previous
  // one wrote last, which are still in L2 (the array itself is larger than L2).  Everything below counts LOGICAL
  // tiles 0 .. n_tiles-1; phys() is the tile's place in the arrays.
  const int rev_base = a.reverse ? n_tiles - 1 : 0, rev_sign = a.reverse ? -1 : 1;
  auto phys = [&](int t) { return rev_base + rev_sign * t; };

  if (t_begin < n_tiles) {
  const float b2c = vec[f], ln_scale = vec[128 + f], ln_offset = vec[256 + f];
  const uint32_t acc0 = tmem + k2ColAcc + wk * 64;                       // + buf * 32
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  // this worker's 32 operand rows inside every K slab; buffer b at + b * 2 * kBBytes, lo at + kBBytes
  const uint32_t b_off = k2OffB + wk * (k2Tile * 16);
  // this thread's element (k = f) of operand row `e` (edge):
  const uint32_t elem_off = b_off + (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 2;
  const int r0 = q * 8;  // this warp's 8 edge rows in phase A
  uint32_t ph1 = 0, ph2 = 0, ph_st = 0;
  // The edge latents stream through L2 once per launch (207 MB at 28 k particles, L2 = 126 MB): their last read (the
  // residual) and the store of the new value are marked evict-first, so that the node arrays (h, aggregates, P)
  // survive in L2 until the node kernel reads them.  Measured: -1.7 % per step.  The bulk copy of a
  // tile must NOT carry the hint (its rows are re-read two tiles later: 153 vs 131 us); evict-last on the
  // aggregates / h / P gave nothing on top.
  const uint64_t pol_e = l2_policy(a.l2_mode & 1);
  // one elected lane of the calling warp: bulk-copy the tile's rows (contiguous in e) into the worker's stage
  auto stage_rows = [&](int tile) {
    if (tile >= n_tiles) return;
    if (elect_one()) {
      const int64_t s0 = (int64_t)phys(tile) * k2Tile;
      const uint32_t bytes = (uint32_t)min(k2Tile, E - (int)s0) * (kLatent * 4);
      mbar_expect_tx(bar_st, bytes);
      bulk_g2s(sbase + k2OffStage + wk * k2StageBytes, a.e + s0 * kLatent, bytes, bar_st);
    }
  };
  // next tile's indices, one per lane: every warp keeps snd / rcv of its lane's edge (gather
  // prefetch); warp 0 also rcv[i + 1] (bucket ends), warp 2 lane 0 the receiver before the tile
  int pre_s = 0, pre_r = -1, pre_b = 0;

  // Every thread has written its part of an operand (and fenced it for the async proxy): a worker-wide
  // bar.sync, after which the worker's first warp issues the GEMM.  Returns true in that warp.
  auto operand_ready = [&](int issuer = 0) -> bool {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_worker), "n"(k2WThreads) : "memory");
    return q == issuer;
  };

  auto prefetch_idx = [&](int tile) {
    if (tile >= n_tiles) return;
    const int64_t s = (int64_t)phys(tile) * k2Tile + lane;
    if (kPref || q == 0) pre_r = s < E ? __ldg(a.rcv + s) : -1;
    if (kPref || q == 1) pre_s = s < E ? __ldg(a.snd + s) : 0;
    if (q == 0) {
      pre_b = s + 1 < E ? __ldg(a.rcv + s + 1) : -3;
    } else if (q == 2) {
      pre_b = (lane == 0 && s > 0) ? __ldg(a.rcv + s - 1) : -2;
    }
  };

  // ---- phase A(tile -> operand buffer b, index buffer ib): indices -> smem, edge latents -> fp16
  //      hi/lo operand, GEMM 1
  auto phase_a = [&](int tile, int b, int ib) {
    const int64_t slot0 = (int64_t)phys(tile) * k2Tile;
    const int rows = min(k2Tile, E - (int)slot0);
    int* sidx = idx_base + ib * k2IdxInts;
    int* rclamp = sidx + 32;
    int* ridx = rclamp + 32;  // ridx[0] = receiver before the tile, ridx[1 + i] = edge i, ridx[1 + rows] = after
    float4 v[8];
    if (kStage) {
      ET(0);
      mbar_wait(bar_st, ph_st);  // the tile's rows were bulk-copied a pipeline round ago
      ET(1);
      ph_st ^= 1;
      const float4* srow = reinterpret_cast<const float4*>(smem + k2OffStage + wk * k2StageBytes) + r0 * 32 + lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = r0 + i < rows ? srow[i * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = r0 + i < rows ? reinterpret_cast<const float4*>(a.e + (slot0 + r0 + i) * kLatent)[lane]
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      // pull the rows of the tile after this one into L2
      const int nt = tile + tile_stride;
      if (nt < n_tiles) {
        const float* nrow = a.e + ((int64_t)phys(nt) * k2Tile + r0) * kLatent + lane * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + (int64_t)i * kLatent));
      }
    }
    if (kPref && lane < rows) {  // this warp's 128-byte segments of the tile's P rows -> L1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a.P + (int64_t)pre_s * (2 * kLatent) + q * 32));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a.P + (int64_t)pre_r * (2 * kLatent) + kLatent + q * 32));
    }
    if (q == 0) {
      const bool ok = lane < rows;
      const int r_here = pre_r, r_next = pre_b;
      rclamp[lane] = max(r_here, 0);
      ridx[1 + lane] = ok ? r_here : (lane == rows ? -3 : -1);  // -3: "no edge after the tile"
      // last edge of its receiver bucket inside the tile (== carry sub-tile)
      const bool end = ok && (r_next != r_here || lane == 31 || lane == rows - 1);
      const uint32_t m = __ballot_sync(0xffffffffu, end);
      if (lane == 0) endm[ib] = m;
      if (lane == 31 && ok) ridx[1 + 32] = r_next;  // receiver just after a full tile
    } else if (q == 1) {
      sidx[lane] = pre_s;
    } else if (q == 2) {
      if (lane == 0) ridx[0] = pre_b;
    }
    prefetch_idx(tile + tile_stride);
    unsigned char* hi_p = smem + k2OffB + b * 2 * kBBytes + wk * (k2Tile * 16);
    unsigned char* lo_p = hi_p + kBBytes;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 h01 = __floats2half2_rn(v[i].x, v[i].y), h23 = __floats2half2_rn(v[i].z, v[i].w);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      constexpr float kS = kNoScale ? 1.0f : kLoScale;
      const __half2 l01 = __floats2half2_rn((v[i].x - f01.x) * kS, (v[i].y - f01.y) * kS);
      const __half2 l23 = __floats2half2_rn((v[i].z - f23.x) * kS, (v[i].w - f23.y) * kS);
      const uint32_t off = (uint32_t)(lane >> 1) * kLboB + (uint32_t)(r0 + i) * 16 + (uint32_t)(lane & 1) * 8;
      *reinterpret_cast<uint2*>(hi_p + off) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint2*>(lo_p + off) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
    }
    fence_async_smem();
    tc_fence_before();
    ET(2);
    if (operand_ready()) {  // this warp issues (one elected lane per instruction)
      tc_fence_after();
      const uint32_t bh = sbase + b_off + b * 2 * kBBytes;
      if (kNoScale) issue_gemm_ts_d(w1_hi, w1_lo, bh, bh + kBBytes, acc0 + b * 32, k2Idesc);
      else issue_gemm_ts<true>(w1_hi, w1_lo, bh, bh + kBBytes, acc0 + b * 32, k2Idesc);
      umma_commit(bar_g1);
      if (kStage) stage_rows(tile + tile_stride);  // every warp is done reading the staged tile
    }
    ET(3);
  };

  // ---- E1(tile in buffers b / ib): hidden = relu(acc + P_s[snd] + P_r[rcv]) -> operand, GEMM 2
  auto phase_e1 = [&](int b, int ib) {
    const int* sp = idx_base + ib * k2IdxInts;
    const int* rp = sp + 32;
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
    ET(10);
    mbar_wait(bar_g1, ph1);
    ph1 ^= 1;
    tc_fence_after();
    ET(11);
    constexpr float kS = kNoScale ? 1.0f : kLoScale;
    if (kMn) {
      // edge-contiguous operand: 8 edges of feature k = f are one 16-byte chunk at
      // (k / 8) * LBO + (e / 8) * 128 + (k % 8) * 16 inside the worker's 512 bytes of every K slab
      unsigned char* hi_p = smem + b_off + b * 2 * kBBytes + (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 16;
      unsigned char* lo_p = hi_p + kBBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float acc[16];
        tmem_ld16(acc0 + b * 32 + lane_sel + h * 16, acc);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int p2 = 0; p2 < 4; ++p2) {
            const int j = g * 8 + p2 * 2, e = h * 16 + j;
            const float x0 = fmaxf(acc[j] + ps[e] + pr[e], 0.f);
            const float x1 = fmaxf(acc[j + 1] + ps[e + 1] + pr[e + 1], 0.f);
            const __half2 hh = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn((x0 - hf.x) * kS, (x1 - hf.y) * kS);
            hw[p2] = *reinterpret_cast<const uint32_t*>(&hh);
            lw[p2] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t off = (uint32_t)(h * 2 + g) * 128;
          *reinterpret_cast<uint4*>(hi_p + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(lo_p + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
    } else {
      unsigned char* hi_p = smem + elem_off + b * 2 * kBBytes;
      unsigned char* lo_p = hi_p + kBBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float acc[16];
        tmem_ld16(acc0 + b * 32 + lane_sel + h * 16, acc);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int e = h * 16 + j;
          const float hval = fmaxf(acc[j] + ps[e] + pr[e], 0.f);
          const __half hi = __float2half_rn(hval);
          const __half lo = __float2half_rn((hval - __half2float(hi)) * kS);
          *reinterpret_cast<__half*>(hi_p + (uint32_t)e * 16) = hi;
          *reinterpret_cast<__half*>(lo_p + (uint32_t)e * 16) = lo;
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    ET(12);
    if (operand_ready(kIssuer2)) {  // GEMM 2 is issued by another warp than GEMM 1: warp 0 already carries the index work
      tc_fence_after();
      const uint32_t bh = sbase + b_off + b * 2 * kBBytes;
      if (kNoScale) issue_gemm_ts_d(w2_hi, w2_lo, bh, bh + kBBytes, acc0 + b * 32, kMn ? k2IdescBMn : k2Idesc);
      else issue_gemm_ts<true>(w2_hi, w2_lo, bh, bh + kBBytes, acc0 + b * 32, kMn ? k2IdescBMn : k2Idesc);
      umma_commit(bar_g2);
    }
    ET(13);
  };

  float eold[32];  // residual rows of the tile E2 finishes
  auto load_eold = [&](int tile) {
    const int64_t slot0 = (int64_t)phys(tile) * k2Tile;
    const int valid = min(k2Tile, E - (int)slot0);
    const float* erow = a.e + slot0 * kLatent + f;
    if (valid == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) eold[j] = ld_hint(erow + (int64_t)j * kLatent, pol_e);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) eold[j] = j < valid ? ld_hint(erow + (int64_t)j * kLatent, pol_e) : 0.f;
    }
  };

  // ---- E2(tile in buffers b / ib): LayerNorm (mean folded into the weights), residual, store, segmented sum
  auto phase_e2 = [&](int tile, int b, int ib) {
    const int64_t slot0 = (int64_t)phys(tile) * k2Tile;
    const int valid = min(k2Tile, E - (int)slot0);  // edges of this tile that exist (>= 1)
    const int* ridx = idx_base + ib * k2IdxInts + 64;
    float* const erow = a.e + slot0 * kLatent + f;
    float* red = red_base + b * 128;
    if (kEnc) {
#pragma unroll
      for (int j = 0; j < 32; ++j) eold[j] = 0.f;
    } else {
      load_eold(tile);
    }
    if (kEnc && b) {  // encoder: two GEMMs of the same kind are in flight, one barrier per buffer
      mbar_wait(bar_g1, ph1);
      ph1 ^= 1;
    } else {
      ET(20);
      mbar_wait(bar_g2, ph2);
      ph2 ^= 1;
    }
    tc_fence_after();
    ET(21);
    float yc[32];
    float part;
    {
      float sq[32];
      tmem_ld32(acc0 + b * 32 + lane_sel, yc);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        yc[j] += b2c;
        sq[j] = yc[j] * yc[j];
      }
      part = warp_transpose_reduce(sq);  // lane l: this warp's 32 features, edge l
    }
    red[q * 32 + lane] = part;
    ET(22);
    asm volatile("bar.sync %0, %1;" ::"r"(bar_ln), "n"(k2WThreads) : "memory");
    ET(23);
    {
      const float var = (red[lane] + red[32 + lane] + red[64 + lane] + red[96 + lane]) * a.inv_latent;
      invs[lane] = 1.0f / sqrtf(var + 1e-5f);  // once per edge per warp (same value in the 4 warps)
    }
    __syncwarp();
    const uint32_t emask = kEnc ? 0u : endm[ib];
    const bool first_cont = !kEnc && ridx[0] == ridx[1];
    const bool last_cont = !kEnc && ridx[1 + valid] == ridx[valid];
    float* const cfirst = a.carry_first + (int64_t)phys(tile) * kLatent + f;
    float* const clast = a.carry_last + (int64_t)phys(tile) * kLatent + f;
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
            st_hint(erow + (int64_t)j * kLatent, msg + eold[j], pol_e);  // residual (gns.py:120-122)
            seg_sum += msg;
            if (!kEnc && ((emask >> j) & 1u)) {  // bucket ends here (uniform across the worker)
              float* dst = a.agg + (int64_t)ridx[1 + j] * kLatent + f;
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
    tc_fence_before();
  };

  // ---- encoder (gns.py:65-81, edge MLP): e = LN(relu(feat W0 + b0) W1c + b1c).  The first layer
  //      (K = dim + 1 <= 4) runs on CUDA cores straight into the operand of the one GEMM; no gather, no
  //      residual, no aggregation.  Pipeline per worker: A'(k0) ; { A'(k+1) ; E2(k) }.
  if constexpr (kEnc) {
    const float ew0 = a.enc_vec[f], ew1 = a.enc_vec[128 + f], ew2 = a.enc_vec[256 + f], ew3 = a.enc_vec[384 + f];
    const float eb0 = a.enc_vec[512 + f];
    float4* feat_w = reinterpret_cast<float4*>(smem + k2OffStage) + warp * 32;  // this warp's copy of the tile's features
    // edge features live in LIST order and are addressed through perm; the next tile's are fetched a phase ahead
    auto feat_of = [&](int tile) -> float4 {
      const int64_t s = (int64_t)phys(tile) * k2Tile + lane;
      return (tile < n_tiles && s < E) ? a.edge_feat[a.perm ? a.perm[s] : s] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 pre_f = make_float4(0.f, 0.f, 0.f, 0.f);
    auto phase_a_enc = [&](int tile, int b) {
      feat_w[lane] = pre_f;
      __syncwarp();
      pre_f = feat_of(tile + tile_stride);
      constexpr float kS = kNoScale ? 1.0f : kLoScale;
      unsigned char* hi_p = smem + b_off + b * 2 * kBBytes + (uint32_t)(f >> 3) * kLboB + (uint32_t)(f & 7) * 16;
      unsigned char* lo_p = hi_p + kBBytes;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) {
          float x[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float4 ft = feat_w[g * 8 + p2 * 2 + t];
            float v = ft.x * ew0;
            v = fmaf(ft.y, ew1, v);
            v = fmaf(ft.z, ew2, v);
            v = fmaf(ft.w, ew3, v);
            x[t] = fmaxf(v + eb0, 0.f);
          }
          const __half2 hh = __floats2half2_rn(x[0], x[1]);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn((x[0] - hf.x) * kS, (x[1] - hf.y) * kS);
          hw[p2] = *reinterpret_cast<const uint32_t*>(&hh);
          lw[p2] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(hi_p + g * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(lo_p + g * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      __syncwarp();  // feat_w is rewritten by the next call
      fence_async_smem();
      tc_fence_before();
      if (operand_ready()) {
        tc_fence_after();
        const uint32_t bh = sbase + b_off + b * 2 * kBBytes;
        if (kNoScale) issue_gemm_ts_d(w2_hi, w2_lo, bh, bh + kBBytes, acc0 + b * 32, k2IdescBMn);
        else issue_gemm_ts<true>(w2_hi, w2_lo, bh, bh + kBBytes, acc0 + b * 32, k2IdescBMn);
        // A'(k+1) commits before E2(k) waits: each buffer has its own barrier, so that no barrier
        // ever runs two phases ahead of a waiter (a parity wait cannot tell phase n from n + 2)
        umma_commit(b ? bar_g1 : bar_g2);
      }
    };
    const int k0 = t_begin + wk;
    if (k0 < n_tiles) pre_f = feat_of(k0);
    if (!a.pdl) load_weights();
    if (k0 < n_tiles) phase_a_enc(k0, 0);
    int buf = 0;
    for (int tile = k0; tile < n_tiles; tile += tile_stride) {
      if (tile + tile_stride < n_tiles) phase_a_enc(tile + tile_stride, buf ^ 1);
      phase_e2(tile, buf, 0);
      buf ^= 1;
    }
  } else {
  // two-tile software pipeline per worker:  A(k0) ; { E1(k) ; A(k+1) ; E2(k) }
  // operand / accumulator buffers alternate (i & 1); the index buffers rotate over three because a
  // sibling warp may still be finishing E2(k) when this warp writes the indices of tile k + 2... + 3
  const int k0 = t_begin + wk;
  if (k0 < n_tiles) {
    if (kStage && q == 0) stage_rows(k0);
    prefetch_idx(k0);
  }
  if (!a.pdl) load_weights();
  if (k0 < n_tiles) phase_a(k0, 0, 0);
  int buf = 0, ib = 0;
  for (int tile = k0; tile < n_tiles; tile += tile_stride) {
    const int ib_next = ib == 2 ? 0 : ib + 1;
#ifdef LB200_CROSSCHECK
    {
      const int it = (tile - k0) / tile_stride;
      trace_it = blockIdx.x == 0 && wk == 0 && it >= 8 && it < 12;
    }
#endif
    phase_e1(buf, ib);
    if (tile + tile_stride < n_tiles) phase_a(tile + tile_stride, buf ^ 1, ib_next);
    phase_e2(tile, buf, ib);
    ET(30);
    buf ^= 1;
    ib = ib_next;
  }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <bool kEnc, bool kMn, bool kNoScale, bool kStage, bool kPref>
static int launch_variant(const EdgeTcArgs& a_in, int grid, int e_cap, cudaStream_t s) {
  static int ready[kMaxDevices];
  int rc = 0;
  const int dev = device_slot(&rc);
  if (dev < 0) return rc;
  if (!ready[dev]) {
    rc = (int)cudaFuncSetAttribute(edge_mp_tc2_kernel<kEnc, kMn, kNoScale, kStage, kPref>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, k2Smem);
    if (rc) return rc;
    ready[dev] = 1;
  }
  EdgeTcArgs a = a_in;
  const bool pdl = pdl_enabled(e_cap);
  a.pdl = (pdl || !(early_issue_mask() & 1)) ? 1 : 0;
  rc = (int)launch_maybe_pdl(edge_mp_tc2_kernel<kEnc, kMn, kNoScale, kStage, kPref>, grid, k2Threads, k2Smem, s, a, pdl);
  if (rc) return rc;
  LB_LAUNCHED(1);
  return 0;
}

int launch_edge_mp_tc2(const EdgeTcArgs& a_in, int e_cap, cudaStream_t s) {
  static int variant = -1, l2_mode = -1, no_reverse = 0;
  if (variant < 0) {
    const char* e = getenv("LB200_TC2_VARIANT");  // bit 0: kMn, bit 1: kNoScale, bit 2: kStage, bit 3: kPref
    variant = e ? (atoi(e) & 15) : k2DefaultVariant;
    const char* h = getenv("LB200_L2HINT");  // A/B of the eviction hints (EdgeTcArgs::l2_mode)
    l2_mode = h ? (atoi(h) & 1) : k2DefaultL2Mode;
    const char* r = getenv("LB200_REVERSE");  // 0: every launch walks forwards (A/B)
    no_reverse = r && r[0] == '0';
  }
  EdgeTcArgs a = a_in;
  a.l2_mode = l2_mode;
  if (no_reverse) a.reverse = 0;
  int rc = 0;
  const int sms = device_sm_count(&rc);
  if (rc) return rc;
  const int n_groups = cdiv(cdiv(e_cap, k2Tile), k2Workers);
  const int grid = n_groups < sms ? n_groups : sms;
  if (a.encoder) return launch_variant<true, true, true, false, false>(a, grid, e_cap, s);
#ifdef LB200_CROSSCHECK
  switch (variant) {  // the measured ladder, cross-check builds only
    case 0: return launch_variant<false, false, false, false, false>(a, grid, e_cap, s);
    case 1: return launch_variant<false, true, false, false, false>(a, grid, e_cap, s);
    case 3: return launch_variant<false, true, true, false, false>(a, grid, e_cap, s);
    case 7: return launch_variant<false, true, true, true, false>(a, grid, e_cap, s);
    default: break;
  }
#endif
  return launch_variant<false, true, true, true, true>(a, grid, e_cap, s);
}

// =====================================================================================
// Hardware self-test of the two tcgen05 features v2 relies on (one CTA, 128 threads):
//   out[0] = max |D_ts - D_ss|   A operand read from TMEM (row m in lane m, column c = k 2c, 2c+1)
//                                 vs the same A read from shared memory
//   out[1] = max |D_scaled - (A B + (A B) 2^-11)|   scale-input-d = 11
//   out[2] = max |D_ss|          (sanity: non-zero)
//   out[3] = max |D_mn - D_ss|   B operand in the MN-major (edge-contiguous) core-matrix layout
//   out[4] = max |D_sub 2^18 - D_ss|   B scaled by 2^-18 into fp16 subnormals (must not be flushed)
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(float* out) {
  __shared__ __align__(128) unsigned char sm_a[128 * 32];   // A: 128 rows x K 16 fp16, K-major core matrices
  __shared__ __align__(128) unsigned char sm_b[32 * 32];    // B: 32 rows x K 16
  __shared__ __align__(128) unsigned char sm_b2[32 * 32];   // B, MN-major: (k/8)*512 + (n/8)*128 + (k%8)*16 + (n%8)*2
  __shared__ __align__(128) unsigned char sm_b3[32 * 32];   // B * 2^-18 (fp16 subnormals), K-major
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t slot;
  __shared__ float red[5][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mb = smem_u32(&bar);
  if (tid == 0) {
    mbar_init(mb, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A[m][k] = ((m * 7 + k * 3) % 17 - 8) / 8 ; B[n][k] = ((n * 5 + k) % 13 - 6) / 4   (exact in fp16)
  uint32_t arow[8];
  for (int k = 0; k < 16; ++k) {
    const float av = (float)((tid * 7 + k * 3) % 17 - 8) * 0.125f;
    const unsigned short ah = __half_as_ushort(__float2half_rn(av));
    *reinterpret_cast<unsigned short*>(sm_a + (k >> 3) * 2048 + tid * 16 + (k & 7) * 2) = ah;
    if (k & 1)
      arow[k >> 1] |= (uint32_t)ah << 16;
    else
      arow[k >> 1] = ah;
    if (tid < 32) {
      const float bv = (float)((tid * 5 + k) % 13 - 6) * 0.25f;
      *reinterpret_cast<__half*>(sm_b + (k >> 3) * 512 + tid * 16 + (k & 7) * 2) = __float2half_rn(bv);
      *reinterpret_cast<__half*>(sm_b2 + (k >> 3) * 512 + (tid >> 3) * 128 + (k & 7) * 16 + (tid & 7) * 2) = __float2half_rn(bv);
      *reinterpret_cast<__half*>(sm_b3 + (k >> 3) * 512 + tid * 16 + (k & 7) * 2) = __float2half_rn(bv * 3.814697265625e-06f);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  // A -> TMEM columns [96, 104)
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(tmem + lane_sel + 96),
               "r"(arow[0]), "r"(arow[1]), "r"(arow[2]), "r"(arow[3]), "r"(arow[4]), "r"(arow[5]), "r"(arow[6]),
               "r"(arow[7])
               : "memory");
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    const uint64_t ad = umma_desc(smem_u32(sm_a), 2048), bd = umma_desc(smem_u32(sm_b), 512);
    umma_f16(tmem + 0, ad, bd, 0u, k2Idesc);            // D_ss  (columns 0..31)
    umma_ts(tmem + 32, tmem + 96, bd, 0u, k2Idesc);     // D_ts  (columns 32..63)
    umma_f16(tmem + 64, ad, bd, 0u, k2Idesc);           // D_scaled = A B, then A B + D 2^-11
    umma_ss_rescale11(tmem + 64, ad, bd, k2Idesc);
    umma_f16(tmem + 128, ad, umma_desc(smem_u32(sm_b2), 512), 0u, k2IdescBMn);  // D_mn
    umma_f16(tmem + 160, ad, umma_desc(smem_u32(sm_b3), 512), 0u, k2Idesc);     // D_sub
    umma_commit(mb);
  }
  mbar_wait(mb, 0);
  tc_fence_after();
  float dss[32], dts[32];
  float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, e4 = 0.f;
  tmem_ld32(tmem + lane_sel + 0, dss);
  tmem_ld32(tmem + lane_sel + 32, dts);
  for (int j = 0; j < 32; ++j) {
    e0 = fmaxf(e0, fabsf(dts[j] - dss[j]));
    e2 = fmaxf(e2, fabsf(dss[j]));
  }
  tmem_ld32(tmem + lane_sel + 64, dts);
  for (int j = 0; j < 32; ++j) e1 = fmaxf(e1, fabsf(dts[j] - (dss[j] + dss[j] * (1.0f / 2048.0f))));
  tmem_ld32(tmem + lane_sel + 128, dts);
  for (int j = 0; j < 32; ++j) e3 = fmaxf(e3, fabsf(dts[j] - dss[j]));
  tmem_ld32(tmem + lane_sel + 160, dts);
  for (int j = 0; j < 32; ++j) e4 = fmaxf(e4, fabsf(dts[j] * 262144.0f - dss[j]));
  for (int off = 16; off >= 1; off >>= 1) {
    e0 = fmaxf(e0, __shfl_xor_sync(0xffffffffu, e0, off));
    e1 = fmaxf(e1, __shfl_xor_sync(0xffffffffu, e1, off));
    e2 = fmaxf(e2, __shfl_xor_sync(0xffffffffu, e2, off));
    e3 = fmaxf(e3, __shfl_xor_sync(0xffffffffu, e3, off));
    e4 = fmaxf(e4, __shfl_xor_sync(0xffffffffu, e4, off));
  }
  if (lane == 0) {
    red[0][warp] = e0;
    red[1][warp] = e1;
    red[2][warp] = e2;
    red[3][warp] = e3;
    red[4][warp] = e4;
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) out[i] = fmaxf(fmaxf(red[i][0], red[i][1]), fmaxf(red[i][2], red[i][3]));
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

}  // namespace lb

#ifdef LB200_CROSSCHECK
extern "C" int lb200_debug_edge_trace(long long* out_4x96x2, int* n_out4) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out_4x96x2, lb::g_edge_trace, sizeof(long long) * 4 * 96 * 2);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(n_out4, lb::g_edge_trace_n, sizeof(int) * 4);
  return (int)e;
}
#endif

extern "C" int lb200_tc_selftest(float* out5_dev, void* stream) {
  if (!out5_dev) return LB200_EINVAL;
  lb::tc_selftest_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(out5_dev);
  LB_LAUNCHED(1);
  LB_LAUNCH_CHECK();
  return 0;
}
