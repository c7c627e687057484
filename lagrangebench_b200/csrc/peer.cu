// Peer-memory plumbing of the decomposed rollout (lb200_shard, include/lb200.h): the per-rank heap that
// neighbouring processes map through CUDA IPC, and the small kernels around the exchanges --
// ghost positions pushed into the neighbours' clouds, one signal/wait kernel per exchange, and the
// all-ranks OR of the status bits that decides whether a step's integrate takes effect.
//
// Everything here is stream-ordered device work: no host synchronisation, no collective library on
// the data path (the halo rows themselves are stored by the node kernel's epilogue, gns_tc.cu).
#include <string.h>

#include "common.cuh"

namespace lb {

// control words at the start of a heap (uint32): remote ranks write [0], [1] and [kCtrlFlags + r]
constexpr int kCtrlSigLeft = 0;    // exchanges completed by the LEFT neighbour (it bumps this)
constexpr int kCtrlSigRight = 1;   // ... by the RIGHT neighbour
constexpr int kCtrlEpoch = 2;      // rollout steps this rank has finished enqueuing-order-wise (local)
constexpr int kCtrlSticky = 3;     // status bits raised on this rank or seen from any rank (local, sticky per call)
constexpr int kCtrlGlobal = 4;     // OR over all ranks of this step's bits (the integrate kernel's skip flag)
// [kCtrlFlags + (epoch & 1) * LB200_MAX_RANKS + r]: (epoch + 1) << 8 | bits, written by rank r.  Two slots by step
// parity: only NEIGHBOURS are fenced against each other by the exchanges, a rank two slabs away can already
// publish step e + 1 while this rank has not read step e yet (it cannot reach step e + 2: that needs this
// rank's word of step e + 1).
constexpr int kCtrlFlags = 16;
constexpr int64_t kCtrlBytes = 4096;
constexpr long long kSpinTimeoutNs = 20ll * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Most recent positions of the owned rows -> this rank's cloud; boundary rows also into the
// neighbours' ghost blocks (coordinate shifted across the periodic wrap); drift check along the cut axis.
template <typename T, int DIM>
__global__ void shard_pos_kernel(const T* __restrict__ window, int tw, int n_owned, T* __restrict__ pos_local,
                                 T* pos_left, T* pos_right, const int32_t* __restrict__ push_left,
                                 const int32_t* __restrict__ push_right, int dst_left, int dst_right, int axis,
                                 T shift_left, T shift_right, const T* __restrict__ ref_coord, T drift_limit,
                                 T axis_length, uint32_t* ctrl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_owned) return;
  T p[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    p[k] = window[((int64_t)i * tw + (tw - 1)) * DIM + k];
    pos_local[(int64_t)i * DIM + k] = p[k];
  }
  if (ref_coord != nullptr) {
    T d = p[axis] - ref_coord[i];
    if (axis_length > T(0)) {  // periodic cut axis: shortest image
      if (d > axis_length * T(0.5)) d -= axis_length;
      if (d < -axis_length * T(0.5)) d += axis_length;
    }
    if (!(fabs((double)d) <= (double)drift_limit)) atomicOr(ctrl + kCtrlSticky, (uint32_t)LB200_OVF_DRIFT);
  }
  bool pushed = false;
  if (pos_left != nullptr) {
    const int k = push_left[i];
    if (k >= 0) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) pos_left[(int64_t)(dst_left + k) * DIM + c] = c == axis ? p[c] + shift_left : p[c];
      pushed = true;
    }
  }
  if (pos_right != nullptr) {
    const int k = push_right[i];
    if (k >= 0) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) pos_right[(int64_t)(dst_right + k) * DIM + c] = c == axis ? p[c] + shift_right : p[c];
      pushed = true;
    }
  }
  if (pushed) __threadfence_system();  // the stores are performed at the peer before this kernel is complete
}

// Exchange number k of the current step (per_step exchanges per step): tell both neighbours that
// everything this rank stored into their heaps so far is complete, then wait for the same from them.
// Counters only ever grow, so a replayed graph needs no reset: the expected value is computed from the
// step counter in this rank's control words.
__global__ void shard_exchange_kernel(uint32_t* ctrl, uint32_t* sig_at_left, uint32_t* sig_at_right, int k,
                                      int per_step) {
  if (threadIdx.x != 0) return;
  __threadfence_system();
  if (sig_at_left != nullptr) atomicAdd_system(sig_at_left, 1u);
  if (sig_at_right != nullptr) atomicAdd_system(sig_at_right, 1u);
  const uint32_t expect = ctrl[kCtrlEpoch] * (uint32_t)per_step + (uint32_t)k + 1u;
  const long long t0 = global_ns();
  bool ok_l = sig_at_left == nullptr, ok_r = sig_at_right == nullptr;
  while (!(ok_l && ok_r)) {
    if (!ok_l) ok_l = (int32_t)(ld_acquire_sys(ctrl + kCtrlSigLeft) - expect) >= 0;
    if (!ok_r) ok_r = (int32_t)(ld_acquire_sys(ctrl + kCtrlSigRight) - expect) >= 0;
    if (!(ok_l && ok_r) && global_ns() - t0 > kSpinTimeoutNs) {
      atomicOr(ctrl + kCtrlSticky, (uint32_t)LB200_OVF_PEER_TIMEOUT);
      break;
    }
  }
  __threadfence_system();
}

// This rank's status bits of the step (sticky bits | the neighbor search's overflow bits) -> every rank
struct CtrlAll {
  uint32_t* p[LB200_MAX_RANKS];
};

__global__ void shard_flag_bcast_kernel(uint32_t* ctrl, const int32_t* __restrict__ nbr_stats, int rank, int world,
                                        CtrlAll ctrl_all) {
  const int r = threadIdx.x;
  const uint32_t bits = (ctrl[kCtrlSticky] | (uint32_t)nbr_stats[2] | (nbr_stats[3] ? LB200_ERR_NONFINITE : 0u)) & 0xffu;
  const uint32_t epoch = ctrl[kCtrlEpoch];
  const uint32_t word = ((epoch + 1u) << 8) | bits;
  if (r < world) st_release_sys(ctrl_all.p[r] + kCtrlFlags + (epoch & 1u) * LB200_MAX_RANKS + rank, word);
}

// OR of all ranks' bits of this step -> ctrl[kCtrlGlobal] (integrate's skip flag), sticky for the later steps
__global__ void shard_flag_wait_kernel(uint32_t* ctrl, int world) {
  const int r = threadIdx.x;
  const uint32_t epoch = ctrl[kCtrlEpoch], tag = epoch + 1u;
  uint32_t bits = 0;
  if (r < world) {
    const long long t0 = global_ns();
    for (;;) {
      const uint32_t w = ld_acquire_sys(ctrl + kCtrlFlags + (epoch & 1u) * LB200_MAX_RANKS + r);
      if ((w >> 8) == tag) {
        bits = w & 0xffu;
        break;
      }
      if (global_ns() - t0 > kSpinTimeoutNs) {
        bits = LB200_OVF_PEER_TIMEOUT;
        break;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
  if (r == 0) {
    ctrl[kCtrlGlobal] = bits;
    ctrl[kCtrlSticky] |= bits;
  }
}

// status: [0] steps completed, [1] status bits, [2] last E of this rank
__global__ void shard_step_done_kernel(uint32_t* ctrl, const int32_t* __restrict__ nbr_stats, int32_t* status) {
  const uint32_t bits = ctrl[kCtrlGlobal];
  status[2] = nbr_stats[0];
  status[1] |= (int32_t)bits | (nbr_stats[3] ? LB200_ERR_NONFINITE : 0);
  if (bits == 0) status[0] += 1;
  ctrl[kCtrlEpoch] += 1;
}

__global__ void shard_call_init_kernel(uint32_t* ctrl) {
  ctrl[kCtrlSticky] = 0;
  ctrl[kCtrlGlobal] = 0;
}

static inline uint32_t* ctrl_of(void* heap) { return reinterpret_cast<uint32_t*>(heap); }

struct ShardPtrs {
  char *pos, *pos_left, *pos_right;
  float *p[2], *p_left[2], *p_right[2];
};

static ShardPtrs shard_ptrs(const lb200_shard* sh) {
  int64_t off_pos, off_p0, off_p1;
  lb200_peer_heap_layout(sh->n_cap, &off_pos, &off_p0, &off_p1);
  ShardPtrs r;
  auto at = [&](void* heap, int64_t off) { return heap ? (char*)heap + off : nullptr; };
  void* hl = sh->has_left ? sh->heap_left : nullptr;
  void* hr = sh->has_right ? sh->heap_right : nullptr;
  r.pos = at(sh->heap, off_pos);
  r.pos_left = at(hl, off_pos);
  r.pos_right = at(hr, off_pos);
  r.p[0] = (float*)at(sh->heap, off_p0);
  r.p[1] = (float*)at(sh->heap, off_p1);
  r.p_left[0] = (float*)at(hl, off_p0);
  r.p_left[1] = (float*)at(hl, off_p1);
  r.p_right[0] = (float*)at(hr, off_p0);
  r.p_right[1] = (float*)at(hr, off_p1);
  return r;
}

// ---- used by gns.cu / rollout.cu
float* shard_p_local(const lb200_shard* sh, int which) { return shard_ptrs(sh).p[which & 1]; }
float* shard_p_left(const lb200_shard* sh, int which) { return shard_ptrs(sh).p_left[which & 1]; }
float* shard_p_right(const lb200_shard* sh, int which) { return shard_ptrs(sh).p_right[which & 1]; }
void* shard_pos_local(const lb200_shard* sh) { return shard_ptrs(sh).pos; }
const int32_t* shard_skip_flag(const lb200_shard* sh) {
  return reinterpret_cast<const int32_t*>(ctrl_of(sh->heap) + kCtrlGlobal);
}

int shard_exchange(const lb200_shard* sh, int k, int per_step, cudaStream_t s) {
  if (!sh->has_left && !sh->has_right) return 0;
  uint32_t* at_left = sh->has_left ? ctrl_of(sh->heap_left) + kCtrlSigRight : nullptr;   // I am its right neighbour
  uint32_t* at_right = sh->has_right ? ctrl_of(sh->heap_right) + kCtrlSigLeft : nullptr;
  shard_exchange_kernel<<<1, 32, 0, s>>>(ctrl_of(sh->heap), at_left, at_right, k, per_step);
  LB_LAUNCHED(1);
  return 0;
}

int shard_push_positions(const lb200_shard* sh, const void* window, int tw, int dim, int pos_f64, cudaStream_t s) {
  const ShardPtrs p = shard_ptrs(sh);
  const int grid = cdiv(sh->n_owned, 128);
#define LB_POS(T, D)                                                                                               \
  shard_pos_kernel<T, D><<<grid, 128, 0, s>>>((const T*)window, tw, sh->n_owned, (T*)p.pos, (T*)p.pos_left,         \
                                              (T*)p.pos_right, sh->push_left, sh->push_right, sh->dst_row_left,    \
                                              sh->dst_row_right, sh->axis, (T)sh->shift_left, (T)sh->shift_right,  \
                                              (const T*)sh->ref_coord, (T)sh->drift_limit, (T)sh->axis_length,     \
                                              ctrl_of(sh->heap))
  if (pos_f64) {
    if (dim == 2) LB_POS(double, 2); else LB_POS(double, 3);
  } else {
    if (dim == 2) LB_POS(float, 2); else LB_POS(float, 3);
  }
#undef LB_POS
  LB_LAUNCHED(1);
  return 0;
}

int shard_flag_bcast(const lb200_shard* sh, const int32_t* nbr_stats, cudaStream_t s) {
  CtrlAll all;
  for (int r = 0; r < LB200_MAX_RANKS; ++r) all.p[r] = r < sh->world ? ctrl_of(r == sh->rank ? sh->heap : sh->heap_all[r]) : nullptr;
  shard_flag_bcast_kernel<<<1, 32, 0, s>>>(ctrl_of(sh->heap), nbr_stats, sh->rank, sh->world, all);
  LB_LAUNCHED(1);
  return 0;
}

int shard_flag_wait(const lb200_shard* sh, cudaStream_t s) {
  shard_flag_wait_kernel<<<1, 32, 0, s>>>(ctrl_of(sh->heap), sh->world);
  LB_LAUNCHED(1);
  return 0;
}

int shard_step_done(const lb200_shard* sh, const int32_t* nbr_stats, int32_t* status, cudaStream_t s) {
  shard_step_done_kernel<<<1, 1, 0, s>>>(ctrl_of(sh->heap), nbr_stats, status);
  LB_LAUNCHED(1);
  return 0;
}

int shard_call_init(const lb200_shard* sh, cudaStream_t s) {
  shard_call_init_kernel<<<1, 1, 0, s>>>(ctrl_of(sh->heap));
  LB_LAUNCHED(1);
  return 0;
}

}  // namespace lb

using namespace lb;

extern "C" int64_t lb200_peer_heap_layout(int32_t n_cap, int64_t* off_pos, int64_t* off_p0, int64_t* off_p1) {
  int64_t off = kCtrlBytes;
  if (off_pos) *off_pos = off;
  off += align_up((int64_t)n_cap * 3 * 8, 256);
  if (off_p0) *off_p0 = off;
  off += align_up((int64_t)n_cap * 2 * kLatent * 4, 256);
  if (off_p1) *off_p1 = off;
  off += align_up((int64_t)n_cap * 2 * kLatent * 4, 256);
  return off;
}

extern "C" int lb200_peer_heap_create(int64_t bytes, void** heap_out, void* handle64_out) {
  if (!heap_out || !handle64_out || bytes < kCtrlBytes) return LB200_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
  void* p = nullptr;
  LB_CHECK(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)kCtrlBytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle64_out, &h, 64);
  *heap_out = p;
  return 0;
}

extern "C" int lb200_peer_heap_open(const void* handle64, void** heap_out) {
  if (!handle64 || !heap_out) return LB200_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  LB_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *heap_out = p;
  return 0;
}

extern "C" int lb200_peer_heap_close(void* peer_heap) {
  if (!peer_heap) return 0;
  LB_CHECK(cudaIpcCloseMemHandle(peer_heap));
  return 0;
}

extern "C" int lb200_peer_heap_destroy(void* heap) {
  if (!heap) return 0;
  LB_CHECK(cudaFree(heap));
  return 0;
}
