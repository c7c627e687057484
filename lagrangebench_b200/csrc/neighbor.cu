// (i) Cell-list radius neighbor search + receiver-major CSR view.
//
// Produces the same (2, E_cap) int32 edge list as jax_md.partition.neighbor_list
// (Sparse, mask_self=False) -- third-party jax-sph 0.0.3, call sites
// lagrangebench/case_setup/case.py:120-130,184-190 -- in the same order (see
// oracle/partition.py for the restated rules), without materialising jax-md's
// (N, 3^d * capacity) candidate table: two sweeps (count, fill) around a prefix sum.
//
// HBM-bound integer/byte work: positions are re-laid out in cell order once so that the
// 3^d-cell sweeps read contiguous memory; the list itself is written as one contiguous run
// per particle.
#include <math.h>

#include "common.cuh"

namespace lb {

// ------------------------------------------------------------------ exclusive scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanBlock = kScanThreads * kScanItems;  // 1024

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// inclusive scan of one value per thread across a 256-thread block; returns inclusive value
// and the block total through *total.
__device__ __forceinline__ int block_incl_scan(int v, int* total) {
  __shared__ int wsum[kScanThreads / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < kScanThreads / 32 ? wsum[lane] : 0;
    s = warp_incl_scan(s);
    if (lane < kScanThreads / 32) wsum[lane] = s;
  }
  __syncthreads();
  int base = w > 0 ? wsum[w - 1] : 0;
  *total = wsum[kScanThreads / 32 - 1];
  __syncthreads();
  return inc + base;
}

__global__ void scan_block_sums(const int32_t* __restrict__ in, int n, int32_t* __restrict__ bsum) {
  int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_incl_scan(s, &total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void scan_sums(int32_t* bsum, int nb) {  // one block; exclusive in place, bsum[nb] = total
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += kScanThreads) {
    int i = base + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    int total;
    int inc = block_incl_scan(v, &total);
    int c = carry;
    if (i < nb) bsum[i] = c + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

__global__ void scan_final(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n,
                           const int32_t* __restrict__ bsum, int nb) {
  int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = base + i < n ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int inc = block_incl_scan(s, &total);
  int run = bsum[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int n, int32_t* scratch, cudaStream_t s) {
  int nb = cdiv(n, kScanBlock);
  if (nb == 0) nb = 1;
  { scan_block_sums<<<nb, kScanThreads, 0, s>>>(in, n, scratch); LB_LAUNCHED(1); }
  { scan_sums<<<1, kScanThreads, 0, s>>>(scratch, nb); LB_LAUNCHED(1); }
  { scan_final<<<nb, kScanThreads, 0, s>>>(in, out, n, scratch, nb); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ grid parameters on device
struct GridDev {
  int n, dim, periodic, use_cells, n_cells, cap;
  int nc[3];
  float cell_size[3];
  double box[3];
  double cutoff;
};

template <typename T>
struct Geo {  // per-thread typed copy of the geometry
  T side[3], half[3], cs[3], cutoff_sq;
  bool periodic[3];  // bit k of the mask: dimension k wraps
  __device__ Geo(const GridDev& g) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      periodic[k] = ((g.periodic >> k) & 1) != 0;
      side[k] = (T)g.box[k];
      half[k] = mul_rn(side[k], T(0.5));
      cs[k] = (T)g.cell_size[k];
    }
    T c = (T)g.cutoff;
    cutoff_sq = mul_rn(c, c);
  }
};

template <typename T, int DIM>
__device__ __forceinline__ int cell_hash(const T* p, const Geo<T>& geo, const GridDev& g) {
  int h = 0, mult = 1;
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    int c = (int)div_rn(p[k], geo.cs[k]);  // truncation toward zero, as jnp.array(..., dtype=i32)
    c = min(max(c, 0), g.nc[k] - 1);       // out-of-domain particles: clamp (documented deviation)
    h += c * mult;
    mult *= g.nc[k];
  }
  return h;
}

template <typename T, int DIM>
__global__ void hash_kernel(const T* __restrict__ pos, GridDev g, int32_t* __restrict__ hash,
                            int32_t* __restrict__ cell_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  Geo<T> geo(g);
  T p[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) p[k] = pos[(int64_t)i * DIM + k];
  int h = cell_hash<T, DIM>(p, geo, g);
  hash[i] = h;
  atomicAdd(&cell_count[h], 1);
}

__global__ void stats_init_kernel(int32_t* stats) {
  stats[0] = 0;
  stats[1] = 0;
}

__global__ void max_occ_kernel(const int32_t* __restrict__ cell_count, int n_cells, int32_t* stats) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int v = c < n_cells ? cell_count[c] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(&stats[1], v);
}

__global__ void cell_scatter_kernel(const int32_t* __restrict__ hash, int n,
                                    const int32_t* __restrict__ cell_start, int32_t* __restrict__ cursor,
                                    int32_t* __restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = hash[i];
  order[cell_start[h] + atomicAdd(&cursor[h], 1)] = i;
}

// stable argsort by hash == ascending particle id within each cell
__global__ void cell_sort_kernel(const int32_t* __restrict__ cell_start, int n_cells, int32_t* __restrict__ order) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  int a = cell_start[c], b = cell_start[c + 1];
  for (int i = a + 1; i < b; ++i) {
    int v = order[i], j = i - 1;
    while (j >= a && order[j] > v) {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = v;
  }
}

template <typename T, int DIM>
__global__ void gather_sorted_kernel(const T* __restrict__ pos, const int32_t* __restrict__ order, int n,
                                     T* __restrict__ spos) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int i = order[q];
#pragma unroll
  for (int k = 0; k < DIM; ++k) spos[(int64_t)q * DIM + k] = pos[(int64_t)i * DIM + k];
}

template <typename T, int DIM>
__device__ __forceinline__ bool within(const T* pi, const T* pj, const Geo<T>& geo) {
  // metric_sq(position[i], position[j]) = sum(disp(pos_i, pos_j)**2), left to right
  T d0 = disp1(pi[0], pj[0], geo.side[0], geo.half[0], geo.periodic[0]);
  T acc = mul_rn(d0, d0);
#pragma unroll
  for (int k = 1; k < DIM; ++k) {
    T d = disp1(pi[k], pj[k], geo.side[k], geo.half[k], geo.periodic[k]);
    acc = add_rn(acc, mul_rn(d, d));
  }
  return acc < geo.cutoff_sq;
}

// Visit the candidates of the particle of sorted rank q in jax-md's order and call
// f(particle id j) for every candidate within the cutoff.
template <typename T, int DIM, typename F>
__device__ __forceinline__ void for_each_neighbor(int q, const GridDev& g, const Geo<T>& geo,
                                                  const T* __restrict__ spos, const int32_t* __restrict__ order,
                                                  const int32_t* __restrict__ cell_start, F f) {
  T pi[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) pi[k] = spos[(int64_t)q * DIM + k];
  int h = cell_hash<T, DIM>(pi, geo, g);
  int cc[3] = {0, 0, 0};
  {
    int r = h;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      cc[k] = r % g.nc[k];
      r /= g.nc[k];
    }
  }
  constexpr int NOFF = DIM == 2 ? 9 : 27;
  // candidate cells: own cell, then ndindex(3,..,3)-1 (last component fastest) skipping 0;
  // component 0 acts on the slowest spatial axis; neighbour = coord - offset (wrap-around)
  for (int o = -1; o < NOFF; ++o) {
    int off[3] = {0, 0, 0};
    if (o >= 0) {
      if (DIM == 2) {
        off[1] = o / 3 - 1;  // slowest axis: y
        off[0] = o % 3 - 1;  // x
      } else {
        off[2] = o / 9 - 1;        // z
        off[1] = (o / 3) % 3 - 1;  // y
        off[0] = o % 3 - 1;        // x
      }
      if (off[0] == 0 && off[1] == 0 && off[2] == 0) continue;
    }
    int c = 0, mult = 1;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      int v = cc[k] - off[k];
      v = v < 0 ? v + g.nc[k] : (v >= g.nc[k] ? v - g.nc[k] : v);
      c += v * mult;
      mult *= g.nc[k];
    }
    int r0 = cell_start[c], m = cell_start[c + 1] - r0;
    if (m == 0) continue;
    // slot of sorted rank k inside its cell is k mod cap: ascending-slot order is a rotation
    int t0 = (g.cap - r0 % g.cap) % g.cap;
    if (t0 >= m) t0 = 0;
    for (int u = 0; u < m; ++u) {
      int t = t0 + u;
      if (t >= m) t -= m;
      int qq = r0 + t;
      T pj[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) pj[k] = spos[(int64_t)qq * DIM + k];
      if (within<T, DIM>(pi, pj, geo)) f(order[qq]);
    }
  }
}

template <typename T, int DIM>
__global__ void count_kernel(GridDev g, const T* __restrict__ spos, const int32_t* __restrict__ order,
                             const int32_t* __restrict__ cell_start, int32_t* __restrict__ cnt) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  Geo<T> geo(g);
  int c = 0;
  for_each_neighbor<T, DIM>(q, g, geo, spos, order, cell_start, [&](int) { ++c; });
  cnt[order[q]] = c;
}

template <typename T, int DIM>
__global__ void fill_kernel(GridDev g, const T* __restrict__ spos, const int32_t* __restrict__ order,
                            const int32_t* __restrict__ cell_start, const int32_t* __restrict__ off,
                            int32_t* __restrict__ idx, int e_cap) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.n) return;
  Geo<T> geo(g);
  int i = order[q];
  int w = off[i];
  for_each_neighbor<T, DIM>(q, g, geo, spos, order, cell_start, [&](int j) {
    if (w < e_cap) {
      idx[w] = j;          // row 0: receivers (candidate)
      idx[e_cap + w] = i;  // row 1: senders (center)
    }
    ++w;
  });
}

// all-pairs path (box smaller than 3 cutoffs): candidates 0..N-1 ascending
template <typename T, int DIM>
__global__ void count_allpairs_kernel(GridDev g, const T* __restrict__ pos, int32_t* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  Geo<T> geo(g);
  T pi[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) pi[k] = pos[(int64_t)i * DIM + k];
  int c = 0;
  for (int j = 0; j < g.n; ++j) {
    T pj[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) pj[k] = pos[(int64_t)j * DIM + k];
    c += within<T, DIM>(pi, pj, geo) ? 1 : 0;
  }
  cnt[i] = c;
}

template <typename T, int DIM>
__global__ void fill_allpairs_kernel(GridDev g, const T* __restrict__ pos, const int32_t* __restrict__ off,
                                     int32_t* __restrict__ idx, int e_cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  Geo<T> geo(g);
  T pi[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) pi[k] = pos[(int64_t)i * DIM + k];
  int w = off[i];
  for (int j = 0; j < g.n; ++j) {
    T pj[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) pj[k] = pos[(int64_t)j * DIM + k];
    if (within<T, DIM>(pi, pj, geo)) {
      if (w < e_cap) {
        idx[w] = j;
        idx[e_cap + w] = i;
      }
      ++w;
    }
  }
}

__global__ void nbr_finalize_kernel(const int32_t* __restrict__ off, int n, int cap, int e_cap, int have_list,
                                    int32_t* stats) {
  int e = off[n];
  stats[0] = e;
  int bits = 0;
  if (have_list && e > e_cap) bits |= LB200_OVF_NEIGHBOR_LIST;
  if (cap > 0 && stats[1] > cap) bits |= LB200_OVF_CELL_LIST;
  stats[2] |= bits;
}

__global__ void nbr_pad_kernel(int32_t* __restrict__ idx, int e_cap, int n, const int32_t* __restrict__ stats) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= e_cap) return;
  if (k >= stats[0]) {
    idx[k] = n;
    idx[e_cap + k] = n;
  }
}

template <typename T, int DIM>
static int nbr_build_t(const lb200_grid* gr, const T* pos, int cap, int32_t* idx, int e_cap, int32_t* stats,
                       void* scratch, int64_t scratch_bytes, cudaStream_t s) {
  const int n = gr->n;
  GridDev g;
  g.n = n;
  g.dim = gr->dim;
  g.periodic = gr->periodic;
  g.use_cells = gr->use_cells;
  g.n_cells = gr->n_cells;
  g.cap = cap > 0 ? cap : 1;
  for (int k = 0; k < 3; ++k) {
    g.nc[k] = gr->cells_per_side[k];
    g.cell_size[k] = gr->cell_size[k];
    g.box[k] = gr->box[k];
  }
  g.cutoff = gr->r_cutoff;

  Arena ar(scratch, scratch_bytes);
  int32_t* cnt = ar.take<int32_t>(n + 1);
  int32_t* off = ar.take<int32_t>(n + 1);
  int32_t* scan_tmp = ar.take<int32_t>(scan_scratch_elems(n > gr->n_cells ? n : gr->n_cells) + 1);
  const int tb = 128;
  { stats_init_kernel<<<1, 1, 0, s>>>(stats); LB_LAUNCHED(1); }
  if (gr->use_cells) {
    const int nc = gr->n_cells;
    int32_t* hash = ar.take<int32_t>(n);
    int32_t* order = ar.take<int32_t>(n);
    int32_t* cell_count = ar.take<int32_t>(2 * (nc + 1));
    int32_t* cursor = cell_count + (nc + 1);
    int32_t* cell_start = ar.take<int32_t>(nc + 1);
    T* spos = ar.take<T>((int64_t)n * DIM);
    if (!ar.ok()) return LB200_EINVAL;
    LB_CHECK(cudaMemsetAsync(cell_count, 0, sizeof(int32_t) * 2 * (nc + 1), s));
    { hash_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(pos, g, hash, cell_count); LB_LAUNCHED(1); }
    { max_occ_kernel<<<cdiv(nc, 256), 256, 0, s>>>(cell_count, nc, stats); LB_LAUNCHED(1); }
    int rc = exclusive_scan_i32(cell_count, cell_start, nc, scan_tmp, s);
    if (rc) return rc;
    { cell_scatter_kernel<<<cdiv(n, tb), tb, 0, s>>>(hash, n, cell_start, cursor, order); LB_LAUNCHED(1); }
    { cell_sort_kernel<<<cdiv(nc, tb), tb, 0, s>>>(cell_start, nc, order); LB_LAUNCHED(1); }
    { gather_sorted_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(pos, order, n, spos); LB_LAUNCHED(1); }
    { count_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(g, spos, order, cell_start, cnt); LB_LAUNCHED(1); }
    rc = exclusive_scan_i32(cnt, off, n, scan_tmp, s);
    if (rc) return rc;
    if (idx != nullptr && e_cap > 0)
      { fill_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(g, spos, order, cell_start, off, idx, e_cap); LB_LAUNCHED(1); }
  } else {
    if (!ar.ok()) return LB200_EINVAL;
    { count_allpairs_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(g, pos, cnt); LB_LAUNCHED(1); }
    int rc = exclusive_scan_i32(cnt, off, n, scan_tmp, s);
    if (rc) return rc;
    if (idx != nullptr && e_cap > 0)
      { fill_allpairs_kernel<T, DIM><<<cdiv(n, tb), tb, 0, s>>>(g, pos, off, idx, e_cap); LB_LAUNCHED(1); }
  }
  const int have_list = idx != nullptr && e_cap > 0;
  { nbr_finalize_kernel<<<1, 1, 0, s>>>(off, n, gr->use_cells ? cap : 0, e_cap, have_list, stats); LB_LAUNCHED(1); }
  if (have_list) { nbr_pad_kernel<<<cdiv(e_cap, 256), 256, 0, s>>>(idx, e_cap, n, stats); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ receiver-major CSR
__global__ void csr_degree_kernel(const int32_t* __restrict__ idx, int n, int e_cap, int32_t* __restrict__ deg) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= e_cap) return;
  int r = idx[k], s = idx[e_cap + k];
  if (r >= 0 && r < n && s >= 0 && s < n) atomicAdd(&deg[r], 1);
}

__global__ void csr_scatter_kernel(const int32_t* __restrict__ idx, int n, int e_cap,
                                   const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                   int32_t* __restrict__ perm) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= e_cap) return;
  int r = idx[k], s = idx[e_cap + k];
  if (r >= 0 && r < n && s >= 0 && s < n) perm[rowptr[r] + atomicAdd(&cursor[r], 1)] = k;
}

// ascending list position within every receiver bucket, then materialise snd / rcv
__global__ void csr_sort_rows_kernel(const int32_t* __restrict__ idx, int n, int e_cap,
                                     const int32_t* __restrict__ rowptr, int32_t* __restrict__ perm,
                                     int32_t* __restrict__ snd, int32_t* __restrict__ rcv) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  int a = rowptr[v], b = rowptr[v + 1];
  for (int i = a + 1; i < b; ++i) {
    int x = perm[i], j = i - 1;
    while (j >= a && perm[j] > x) {
      perm[j + 1] = perm[j];
      --j;
    }
    perm[j + 1] = x;
  }
  for (int i = a; i < b; ++i) {
    snd[i] = idx[e_cap + perm[i]];
    rcv[i] = v;
  }
}

__global__ void csr_pad_kernel(int n, int e_cap, const int32_t* __restrict__ rowptr, int32_t* __restrict__ perm,
                               int32_t* __restrict__ snd, int32_t* __restrict__ rcv) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= e_cap) return;
  if (k >= rowptr[n]) {
    perm[k] = 0;
    snd[k] = n;
    rcv[k] = n;
  }
}

}  // namespace lb

using namespace lb;

extern "C" int lb200_grid_init(lb200_grid* g, int32_t n, int32_t dim, int32_t pos_f64, int32_t periodic,
                               const double* box, double r_cutoff) {
  if (!g || !box || (dim != 2 && dim != 3) || n < 1 || !(r_cutoff > 0)) return LB200_EINVAL;
  g->n = n;
  g->dim = dim;
  g->pos_f64 = pos_f64 ? 1 : 0;
  g->periodic = periodic;  // bit k: dimension k is periodic (the reference uses all-or-none, case.py:104)
  g->r_cutoff = r_cutoff;
  g->n_cand_cells = dim == 2 ? 9 : 27;
  // jax-md: box = f32(box); use the cell list iff all(cutoff < box / 3)
  float cut = (float)r_cutoff;
  int use = 1;
  for (int k = 0; k < 3; ++k) {
    g->box[k] = k < dim ? box[k] : 1.0;
    g->cells_per_side[k] = 1;
    g->cell_size[k] = 1.0f;
  }
  for (int k = 0; k < dim; ++k) {
    float b = (float)box[k];
    if (!(cut < b / 3.0f)) use = 0;
  }
  g->use_cells = use;
  g->n_cells = 1;
  if (use) {
    int64_t total = 1;
    for (int k = 0; k < dim; ++k) {
      float b = (float)box[k];
      float cps = floorf(b / cut);  // _cell_dimensions: floor(box / minimum_cell_size) in f32
      g->cells_per_side[k] = (int32_t)cps;
      g->cell_size[k] = b / cps;
      if (g->cells_per_side[k] < 3) return LB200_EINVAL;
      total *= g->cells_per_side[k];
    }
    if (total > (int64_t)1 << 30) return LB200_EUNSUPPORTED;
    g->n_cells = (int32_t)total;
  }
  return 0;
}

extern "C" int64_t lb200_nbr_scratch_bytes(const lb200_grid* g) {
  int64_t n = g->n, nc = g->n_cells;
  int64_t b = 0;
  b += 2 * align_up((n + 1) * 4, 256);
  b += align_up((scan_scratch_elems(n > nc ? n : nc) + 1) * 4, 256);
  b += 2 * align_up(n * 4, 256);
  b += align_up(2 * (nc + 1) * 4, 256) + align_up((nc + 1) * 4, 256);
  b += align_up(n * 3 * 8, 256);
  return b + 4096;
}

extern "C" int lb200_nbr_build(const lb200_grid* g, const void* pos_dev, int32_t cell_capacity, int32_t* idx_dev,
                               int32_t e_cap, int32_t* stats_dev, void* scratch_dev, int64_t scratch_bytes,
                               void* stream) {
  if (!g || !pos_dev || !stats_dev || !scratch_dev || e_cap < 0 || cell_capacity < 0) return LB200_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (g->pos_f64) {
    if (g->dim == 2)
      return nbr_build_t<double, 2>(g, (const double*)pos_dev, cell_capacity, idx_dev, e_cap, stats_dev,
                                    scratch_dev, scratch_bytes, s);
    return nbr_build_t<double, 3>(g, (const double*)pos_dev, cell_capacity, idx_dev, e_cap, stats_dev, scratch_dev,
                                  scratch_bytes, s);
  }
  if (g->dim == 2)
    return nbr_build_t<float, 2>(g, (const float*)pos_dev, cell_capacity, idx_dev, e_cap, stats_dev, scratch_dev,
                                 scratch_bytes, s);
  return nbr_build_t<float, 3>(g, (const float*)pos_dev, cell_capacity, idx_dev, e_cap, stats_dev, scratch_dev,
                               scratch_bytes, s);
}

extern "C" int64_t lb200_csr_scratch_bytes(int32_t n, int32_t e_cap) {
  (void)e_cap;
  return 2 * align_up(((int64_t)n + 1) * 4, 256) + align_up((scan_scratch_elems(n) + 1) * 4, 256) + 4096;
}

extern "C" int lb200_csr_build(const int32_t* idx_dev, int32_t n, int32_t e_cap, int32_t* rowptr_dev,
                               int32_t* perm_dev, int32_t* snd_dev, int32_t* rcv_dev, void* scratch_dev,
                               int64_t scratch_bytes, void* stream) {
  if (!idx_dev || !rowptr_dev || !perm_dev || !snd_dev || !rcv_dev || !scratch_dev || n < 1 || e_cap < 1)
    return LB200_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(scratch_dev, scratch_bytes);
  int32_t* deg = ar.take<int32_t>(n + 1);
  int32_t* cursor = ar.take<int32_t>(n + 1);
  int32_t* scan_tmp = ar.take<int32_t>(scan_scratch_elems(n) + 1);
  if (!ar.ok()) return LB200_EINVAL;
  LB_CHECK(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (n + 1), s));
  LB_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (n + 1), s));
  { csr_degree_kernel<<<cdiv(e_cap, 256), 256, 0, s>>>(idx_dev, n, e_cap, deg); LB_LAUNCHED(1); }
  int rc = exclusive_scan_i32(deg, rowptr_dev, n, scan_tmp, s);
  if (rc) return rc;
  { csr_scatter_kernel<<<cdiv(e_cap, 256), 256, 0, s>>>(idx_dev, n, e_cap, rowptr_dev, cursor, perm_dev); LB_LAUNCHED(1); }
  { csr_sort_rows_kernel<<<cdiv(n, 128), 128, 0, s>>>(idx_dev, n, e_cap, rowptr_dev, perm_dev, snd_dev, rcv_dev); LB_LAUNCHED(1); }
  { csr_pad_kernel<<<cdiv(e_cap, 256), 256, 0, s>>>(n, e_cap, rowptr_dev, perm_dev, snd_dev, rcv_dev); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return 0;
}
