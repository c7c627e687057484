// tcgen05 / TMEM / mbarrier / bulk-copy helpers and the operand-layout constants shared by the
// tensor-core kernels (gns_tc.cu: message kernel v1, edge encoder, node update; gns_tc2.cu: the
// pipelined message kernel with TMEM-resident weights).  sm_100a only.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "gns_tc.cuh"

namespace lb {

constexpr int kTcThreads = 512;        // kWorkers independent workers per CTA; 16 warps hide the epilogue latency
constexpr int kWorkers = 4;            // 4 x 128 threads (32-edge tiles) or 2 x 256 threads (64-edge tiles)
constexpr int kTcTile = 128 / kWorkers;  // edges per worker tile (one MMA tile, N = kTcTile)
constexpr int kChunks = kTcTile / 32;    // 32-edge chunks (carry sub-tiles) per tile, one per group of 4 warps
constexpr int kWThreads = kTcThreads / kWorkers;
constexpr uint32_t kLboA = 2048;       // weights: 128 rows * 16 B per 8-wide K slab
constexpr uint32_t kLboB = 2064;       // edge operand: padded slab pitch -> conflict-free stores
constexpr uint32_t kSbo = 128;
constexpr uint32_t kWBytes = 16 * kLboA;  // one 128x128 fp16 weight operand (32 KB)
constexpr uint32_t kBBytes = 16 * kLboB;  // one 128-row fp16 edge operand (33 KB): rows 0..63 worker 0, 64..127 worker 1
// instruction descriptor: D=F32 (bit 4), A=B=F16 (0), both K-major, N=64 (>>3 at bit 17), M=128 (>>4 at bit 24)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTcTile >> 3) << 17) | (8u << 24);
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;
static_assert(kTcTile == kChunks * kEdgeTile && kEdgeTile == 32, "a warp's 32-edge chunk is one carry sub-tile");

// shared memory map (bytes)
constexpr uint32_t kOffW = 0;                         // W1e_hi, W1e_lo, W2c_hi, W2c_lo
constexpr uint32_t kOffB = kOffW + 4 * kWBytes;       // B_hi, B_lo
constexpr uint32_t kOffVec = kOffB + 2 * kBBytes;     // b2c[128], scale[128], offset[128]
constexpr uint32_t kOffIdx = kOffVec + 3 * 512;       // per worker: sidx[T], rclamp[T], ridx_ext[T + 4]
constexpr uint32_t kIdxInts = 3 * kTcTile + 4;
constexpr uint32_t kOffRed = kOffIdx + kWorkers * kIdxInts * 4;  // per 32-edge chunk: red[4 warps][32]
constexpr uint32_t kOffInv = kOffRed + 4 * 4 * 32 * 4;           // inv[16 warps][32]: 1/sqrt(var + eps) per edge
constexpr uint32_t kOffEnd = kOffInv + 16 * 32 * 4;              // per chunk: endmask
constexpr uint32_t kOffFeat = kOffEnd + 16;                      // per worker: feat[T] float4 (encoder inputs)
constexpr uint32_t kOffBar = kOffFeat + 128 * 16;                // mbarriers: weights, mma[kWorkers]; tmem base
constexpr uint32_t kSmemTc = kOffBar + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);           // start address  [0,14)
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;        // leading byte offset (K-adjacent core matrices)
  d |= (uint64_t)((kSbo >> 4) & 0x3FFFu) << 32;       // stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  return d;                                           // layout_type 0 = no swizzle
}

// One lane of a converged warp (elect.sync); operands computed by the whole warp stay warp-uniform,
// which lets the compiler feed tcgen05.mma from uniform registers without a per-instruction
// ELECT / R2UR "waterfall" loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Call from ALL lanes of one warp: a single elected lane issues the MMA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                         uint32_t idesc = kIdesc) {
  if (elect_one())
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t mbar) {  // call from ALL lanes of the issuing warp
  if (elect_one())
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

// L2 eviction policies (createpolicy): streamed data (the edge latents, larger than L2 at any interesting size) is
// marked evict-first so that it does not push the node arrays (h, aggregates, projections) out of L2.
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 normal, 1 evict-first, 2 evict-last
  uint64_t pol;
  if (kind == 1)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_hint(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float ld_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// two accumulators x 32 columns of this warp's 32 TMEM lanes -> registers (thread == lane)
__device__ __forceinline__ void tmem_ld_pair(uint32_t ta, uint32_t tb, float (&a)[32], float (&b)[32]) {
  uint32_t x[32], y[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
        "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]), "=r"(x[16]),
        "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]), "=r"(x[24]),
        "=r"(x[25]), "=r"(x[26]), "=r"(x[27]), "=r"(x[28]), "=r"(x[29]), "=r"(x[30]), "=r"(x[31]), "=r"(y[0]),
        "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]), "=r"(y[8]), "=r"(y[9]),
        "=r"(y[10]), "=r"(y[11]), "=r"(y[12]), "=r"(y[13]), "=r"(y[14]), "=r"(y[15]), "=r"(y[16]), "=r"(y[17]),
        "=r"(y[18]), "=r"(y[19]), "=r"(y[20]), "=r"(y[21]), "=r"(y[22]), "=r"(y[23]), "=r"(y[24]), "=r"(y[25]),
        "=r"(y[26]), "=r"(y[27]), "=r"(y[28]), "=r"(y[29]), "=r"(y[30]), "=r"(y[31])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    a[i] = __uint_as_float(x[i]);
    b[i] = __uint_as_float(y[i]);
  }
}

// two accumulators x 16 columns (half the registers of tmem_ld_pair)
__device__ __forceinline__ void tmem_ld_pair16(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
  uint32_t x[16], y[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
        "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]), "=r"(y[0]),
        "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]), "=r"(y[8]), "=r"(y[9]),
        "=r"(y[10]), "=r"(y[11]), "=r"(y[12]), "=r"(y[13]), "=r"(y[14]), "=r"(y[15])
      : "r"(ta), "r"(tb)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(x[i]);
    b[i] = __uint_as_float(y[i]);
  }
}

__device__ __forceinline__ void split_f16(float x, unsigned short& hi, unsigned short& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn((x - __half2float(h)) * kLoScale);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}

// 3-pass split-precision GEMM: acc_hh = A_hi B_hi ; acc_x = A_hi B_lo + A_lo B_hi   (K = 128)
__device__ __forceinline__ void issue_gemm(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           uint32_t acc_hh, uint32_t acc_x, uint32_t idesc = kIdesc,
                                           bool accumulate = false) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    umma_f16(acc_hh, umma_desc(a_hi + j * 2 * kLboA, kLboA), umma_desc(b_hi + j * 2 * kLboB, kLboB),
             (j > 0 || accumulate) ? 1u : 0u, idesc);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    umma_f16(acc_x, umma_desc(a_hi + j * 2 * kLboA, kLboA), umma_desc(b_lo + j * 2 * kLboB, kLboB),
             (j > 0 || accumulate) ? 1u : 0u, idesc);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    umma_f16(acc_x, umma_desc(a_lo + j * 2 * kLboA, kLboA), umma_desc(b_hi + j * 2 * kLboB, kLboB), 1u, idesc);
}

// ---- v2 kernels (gns_tc2.cu, node_tc2.cu): A operand in tensor memory, one rescaled accumulator, N = 32 tiles
constexpr int k2Tile = 32;  // edges (nodes) per worker tile == one carry sub-tile (kEdgeTile)
static_assert(k2Tile == kEdgeTile, "a tile is one carry sub-tile");
// instruction descriptor: D=F32, A=B=F16, K-major, N=32, M=128
constexpr uint32_t k2Idesc = (1u << 4) | ((uint32_t)(k2Tile >> 3) << 17) | (8u << 24);
constexpr uint32_t k2IdescBMn = k2Idesc | (1u << 16);  // B operand MN-major (edge-contiguous core matrices)

// D[tmem] (+)= A[tmem] * B[smem desc]; call from ALL lanes of one warp
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate,
                                        uint32_t idesc) {
  if (elect_one())
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D = A * B + D * 2^-11 (scale-input-d)
__device__ __forceinline__ void umma_ts_rescale11(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
  if (elect_one()) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p, 11;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(zero)
        : "memory");
  }
}

// SS twin of the rescaling instruction (self-test only)
__device__ __forceinline__ void umma_ss_rescale11(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (elect_one()) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %4, %4, %4}, p, 11;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(zero)
        : "memory");
  }
}

// split-precision GEMM (K = 128) into ONE accumulator.  Weights' low halves carry 2^11 (lo' = lo * 2^11).
//   b_scaled:  acc = (A_hi B_lo' + A_lo' B_hi) * 2^-11 + A_hi B_hi
//   !b_scaled: acc = (A_lo' B_hi) * 2^-11 + A_hi B_lo + A_hi B_hi        (activation lo unscaled)
template <bool kBScaled>
__device__ __forceinline__ void issue_gemm_ts(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t acc,
                                              uint32_t idesc) {
  if (kBScaled) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      umma_ts(acc, a_hi + j * 8, umma_desc(b_lo + j * 2 * kLboB, kLboB), j > 0 ? 1u : 0u, idesc);
#pragma unroll
    for (int j = 0; j < 8; ++j) umma_ts(acc, a_lo + j * 8, umma_desc(b_hi + j * 2 * kLboB, kLboB), 1u, idesc);
    umma_ts_rescale11(acc, a_hi, umma_desc(b_hi, kLboB), idesc);
#pragma unroll
    for (int j = 1; j < 8; ++j) umma_ts(acc, a_hi + j * 8, umma_desc(b_hi + j * 2 * kLboB, kLboB), 1u, idesc);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      umma_ts(acc, a_lo + j * 8, umma_desc(b_hi + j * 2 * kLboB, kLboB), j > 0 ? 1u : 0u, idesc);
    umma_ts_rescale11(acc, a_hi, umma_desc(b_lo, kLboB), idesc);
#pragma unroll
    for (int j = 1; j < 8; ++j) umma_ts(acc, a_hi + j * 8, umma_desc(b_lo + j * 2 * kLboB, kLboB), 1u, idesc);
#pragma unroll
    for (int j = 0; j < 8; ++j) umma_ts(acc, a_hi + j * 8, umma_desc(b_hi + j * 2 * kLboB, kLboB), 1u, idesc);
  }
}

// ---- the same GEMM with the B descriptor advanced by ONE 32-bit add per instruction: the descriptor of K slab j is
// the descriptor of slab 0 plus j * (2 LBO >> 4) in its low word (start-address field; no carry: the field has 14 bits
// and shared memory ends below 2^18).  Halves the uniform-datapath instructions the issuing warp spends per MMA.
__device__ __forceinline__ void umma_ts_d(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_hi,
                                          uint32_t accumulate, uint32_t idesc) {
  if (elect_one())
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 d;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_d_rescale11(uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_hi,
                                                    uint32_t idesc) {
  if (elect_one()) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 d;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, {%5, %5, %5, %5}, p, 11;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(zero)
        : "memory");
  }
}
// activation low halves unscaled (issue_gemm_ts<false>):  acc = (A_lo' B_hi) 2^-11 + A_hi B_lo + A_hi B_hi
__device__ __forceinline__ void issue_gemm_ts_d(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t acc,
                                                uint32_t idesc) {
  const uint64_t dh = umma_desc(b_hi, kLboB), dl = umma_desc(b_lo, kLboB);
  const uint32_t hi_word = (uint32_t)(dh >> 32), h0 = (uint32_t)dh, l0 = (uint32_t)dl;
  constexpr uint32_t kStep = (2 * kLboB) >> 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) umma_ts_d(acc, a_lo + j * 8, h0 + j * kStep, hi_word, j > 0 ? 1u : 0u, idesc);
  umma_ts_d_rescale11(acc, a_hi, l0, hi_word, idesc);
#pragma unroll
  for (int j = 1; j < 8; ++j) umma_ts_d(acc, a_hi + j * 8, l0 + j * kStep, hi_word, 1u, idesc);
#pragma unroll
  for (int j = 0; j < 8; ++j) umma_ts_d(acc, a_hi + j * 8, h0 + j * kStep, hi_word, 1u, idesc);
}

// 16 columns of this warp's 32 TMEM lanes -> registers (thread == lane)
__device__ __forceinline__ void tmem_ld16(uint32_t ta, float (&a)[16]) {
  uint32_t x[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
        "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
      : "r"(ta)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = __uint_as_float(x[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t ta, float (&a)[32]) {
  uint32_t x[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
        "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]), "=r"(x[16]),
        "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]), "=r"(x[24]),
        "=r"(x[25]), "=r"(x[26]), "=r"(x[27]), "=r"(x[28]), "=r"(x[29]), "=r"(x[30]), "=r"(x[31])
      : "r"(ta)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(x[i]);
}

// registers -> 16 columns of this warp's 32 TMEM lanes (thread == lane)
__device__ __forceinline__ void tmem_st16(uint32_t ta, const uint32_t (&x)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(ta),
      "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(x[8]), "r"(x[9]),
      "r"(x[10]), "r"(x[11]), "r"(x[12]), "r"(x[13]), "r"(x[14]), "r"(x[15])
      : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t mbar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One 128x128 fp16 operand, stored in global memory in the UMMA K-major layout [k/8][m][k%8]
// (models.py: umma_operand), -> 64 TMEM columns: lane m holds row m, column c holds k = 2c, 2c+1.
// Called by a warp for its own lane quarter; `lane_row` = the thread's row m.
__device__ __forceinline__ void weight_to_tmem(const uint4* op, int lane_row, uint32_t taddr) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t x[16];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const uint4 v = __ldg(op + (g * 4 + s) * 128 + lane_row);
      x[s * 4 + 0] = v.x;
      x[s * 4 + 1] = v.y;
      x[s * 4 + 2] = v.z;
      x[s * 4 + 3] = v.w;
    }
    tmem_st16(taddr + g * 16, x);
  }
}

// lane l ends with the sum over the warp's 32 lanes of v[l]   (31 shuffles, fixed order)
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

}  // namespace lb
