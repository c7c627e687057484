// Shared device/host helpers for the lb200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lb200.h"

#define LB_CHECK(expr)                            \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

#define LB_LAUNCH_CHECK()                         \
  do {                                            \
    cudaError_t _e = cudaGetLastError();          \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

namespace lb {

constexpr int kLatent = LB200_LATENT;
// Receiver-sorted edges per message tile / nodes per node tile.  The edge kernel and the
// node kernel agree on the carry protocol through these two constants only.
constexpr int kEdgeTile = 32;
constexpr int kNodeTile = 64;

// cumulative number of kernels launched by the library (lb200_launch_count)
extern std::atomic<int64_t> g_launches;
#define LB_LAUNCHED(k) (::lb::g_launches.fetch_add((k), std::memory_order_relaxed))

// One-time per-DEVICE setup (kernel attributes and the SM count belong to a device / context, not to
// the process): slot of the current device, or -1 with *rc set.
constexpr int kMaxDevices = 64;
static inline int device_slot(int* rc) {
  int dev = 0;
  const cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    *rc = e != cudaSuccess ? (int)e : LB200_EUNSUPPORTED;
    return -1;
  }
  return dev;
}
int device_sm_count(int* rc);  // cached per device (rollout.cu)

bool prof_enabled();

// integrate with the frame of `target` / `pred` selected by a device-side step counter
int integrate_indexed(const lb200_integrate_cfg* c, const float* out_dev, void* window_dev, const int32_t* ptype_dev,
                      const void* target_dev, void* pred_out_dev, const int32_t* skip_flag_dev,
                      const int32_t* step_counter_dev, cudaStream_t s);

// decomposed rollout (peer.cu): views into the peer heaps and the exchange / status kernels
float* shard_p_local(const lb200_shard* sh, int which);
float* shard_p_left(const lb200_shard* sh, int which);
float* shard_p_right(const lb200_shard* sh, int which);
void* shard_pos_local(const lb200_shard* sh);
const int32_t* shard_skip_flag(const lb200_shard* sh);
int shard_exchange(const lb200_shard* sh, int k, int per_step, cudaStream_t s);
int shard_push_positions(const lb200_shard* sh, const void* window, int tw, int dim, int pos_f64, cudaStream_t s);
int shard_flag_bcast(const lb200_shard* sh, const int32_t* nbr_stats, cudaStream_t s);
int shard_flag_wait(const lb200_shard* sh, cudaStream_t s);
int shard_step_done(const lb200_shard* sh, const int32_t* nbr_stats, int32_t* status, cudaStream_t s);
int shard_call_init(const lb200_shard* sh, cudaStream_t s);

// optional CUDA-event timing of a kernel class on its launch stream (lb200_profile)
void prof_begin(int cls, cudaStream_t s);
void prof_end(int cls, cudaStream_t s);

// ---- programmatic dependent launch: a kernel launched with the attribute may begin (its prologue: barrier
// init, TMEM allocation, weights -> tensor memory) while the previous kernel of the stream drains; it calls
// pdl_wait() before it touches anything that kernel produced.  Both instructions are no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Measured: the overlap pays for launch-bound clouds (TGV-2D: -7 % per step) and costs about 1 % once
// a kernel runs for tens of microseconds (the early prologue's weight traffic competes with the running kernel's
// tail), so it is applied to small problems only.  LB200_PDL=1 / 0 forces it on / off (A/B measurements).
bool pdl_enabled(int64_t work_items);
int early_issue_mask();  // LB200_EARLY: bit 0 message kernel, bit 1 node kernel issue their first loads before the weights (A/B)
constexpr int64_t kPdlMaxEdges = 100000;

template <typename Arg>
static inline cudaError_t launch_maybe_pdl(void (*kernel)(Arg), int grid, int block, size_t smem, cudaStream_t s,
                                           const Arg& arg, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Bump allocator over caller-provided scratch.
struct Arena {
  char* base;
  int64_t off, cap;
  Arena(void* p, int64_t bytes) : base((char*)p), off(0), cap(bytes) {}
  template <typename T>
  T* take(int64_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * (int64_t)sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---- rounding-exact arithmetic (no FMA contraction), matching NumPy / XLA-CPU elementwise ops
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float fmod_x(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double fmod_x(double a, double b) { return fmod(a, b); }

// jnp.mod for side > 0: C fmod, then + side where the remainder is negative.  fmod is exact, and
// for |x| < 2 side it is x, or x - side (exact by Sterbenz' lemma): those ranges -- every value a
// displacement or a one-step shift takes -- skip the (slow, iterative) library fmod bit for bit.
template <typename T>
__device__ __forceinline__ T floor_mod(T x, T side) {
  T r;
  if (x >= T(0)) {
    if (x < side) return x;
    r = x < add_rn(side, side) ? sub_rn(x, side) : fmod_x(x, side);
    return r;
  }
  r = x > -side ? x : fmod_x(x, side);
  return r < T(0) ? add_rn(r, side) : r;
}

// space.periodic / space.free displacement of one component.
template <typename T>
__device__ __forceinline__ T disp1(T a, T b, T side, T half, bool periodic) {
  T d = sub_rn(a, b);
  if (periodic) d = sub_rn(floor_mod(add_rn(d, half), side), half);
  return d;
}

template <typename T>
__device__ __forceinline__ T shift1(T r, T dr, T side, bool periodic) {
  T s = add_rn(r, dr);
  return periodic ? floor_mod(s, side) : s;
}

// Exclusive scan of int32 (n elements -> n+1 outputs, out[n] = total).
// One memset node + one kernel (decoupled look-back).  scratch: int32[scan_scratch_elems(n)], 8-byte aligned.
int exclusive_scan_i32(const int32_t* in, int32_t* out, int n, int32_t* scratch, cudaStream_t s);
static inline int64_t scan_scratch_elems(int64_t n) { return 2 * ((n + 2047) / 2048 + 2) + 2; }

}  // namespace lb
