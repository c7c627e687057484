/*
 * lb200.h -- C ABI of the B200-native GNS rollout hot path (liblb200 / _lb200.so).
 *
 * The reference (tumaer/lagrangebench) has no FFI: its hot path is Python on JAX/XLA.
 * Each entry point below replaces one JAX-traced stage of that path; the citation names
 * the reference interface it stands in for (paths relative to the reference root).
 * A maintainer binds these with ctypes from the reference's Python.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer marked "dev" is DEVICE memory owned by
 *     the caller (the Python host uses torch tensors purely as allocations);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless stated;
 *   - return value: 0 on success, >0 a cudaError_t, <0 an LB200_E* argument error;
 *   - positions are `float` or `double` (`pos_f64` = 0/1): the reference preprocesses in
 *     float64 by default (lagrangebench/defaults.py:22) and runs the network in float32;
 *   - index arrays are int32; the pad value of the edge list is N (jax-md convention).
 */
#ifndef LB200_H
#define LB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB200_LATENT 128          /* latent_dim the kernels are specialised for (defaults.py:49) */
#define LB200_MAX_NODE_IN 64      /* max node-encoder input width (features + type embedding) */

#define LB200_EINVAL (-1)
#define LB200_EUNSUPPORTED (-2)

/* overflow flag bits, mirroring jax-md's PartitionErrorCode */
#define LB200_OVF_NEIGHBOR_LIST 1 /* E > E_cap: list truncated */
#define LB200_OVF_CELL_LIST 2     /* a cell holds more particles than cell_capacity */
/* status bits of a decomposed rollout (lb200_shard), OR-ed over all ranks before they take effect */
#define LB200_OVF_DRIFT 4         /* a particle moved further than the halo margin covers: choose new ghost sets */
#define LB200_OVF_PEER_TIMEOUT 8  /* a neighbour's signal did not arrive (a rank died): the rollout is void */
#define LB200_ERR_NONFINITE 16     /* the network produced NaN / Inf: an activation left the range of the fp16 split
                                      (|x| > 65504) -- rerun with edge_impl = 1 (float32 CUDA cores) */
#define LB200_MAX_RANKS 16

int lb200_version(void);
const char* lb200_error_string(int code);

/* ------------------------------------------------------------------------------------
 * Geometry of the search grid.  Replaces jax_md.partition._cell_dimensions /
 * neighbor_list's `use_cell_list` decision (third-party jax-sph 0.0.3; call site
 * lagrangebench/case_setup/case.py:120-130).  Host-only, no CUDA work.
 */
typedef struct {
  int32_t n;               /* particles (rows of the position array; the pad value of the list) */
  int32_t n_valid;         /* the first n_valid rows are real particles, the rest padding (particle type
                              PAD_VALUE, data.py:183-197) that must not enter the search: the reference's
                              `num_particles` argument (case.py:182-190).  lb200_grid_init sets it to n. */
  int32_t dim;             /* 2 or 3 */
  int32_t pos_f64;         /* 0: float positions, 1: double positions */
  int32_t periodic;        /* bit k set: dimension k is periodic.  The reference is all-or-none
                              (case.py:104-108); a slab of a decomposed domain is open along its cut axis */
  double box[3];
  double r_cutoff;
  int32_t use_cells;       /* out of lb200_grid_init: 1 cell list, 0 all-pairs */
  int32_t cells_per_side[3];
  float cell_size[3];
  int32_t n_cells;
  int32_t n_cand_cells;    /* 3^dim */
} lb200_grid;

int lb200_grid_init(lb200_grid* g, int32_t n, int32_t dim, int32_t pos_f64, int32_t periodic,
                    const double* box, double r_cutoff);

/* bytes of device scratch lb200_nbr_build / lb200_csr_build need for this grid / capacity */
int64_t lb200_nbr_scratch_bytes(const lb200_grid* g);
int64_t lb200_csr_scratch_bytes(int32_t n, int32_t e_cap);

/* ------------------------------------------------------------------------------------
 * (i) Radius neighbor search.  Replaces NeighborListFns.allocate / NeighborList.update of
 * jax_md.partition (Sparse, mask_self=False; call sites case.py:184-190) and yields the
 * same (2, E_cap) int32 array, row 0 = receivers, row 1 = senders
 * (lagrangebench/case_setup/features.py:110), in the same order, padded with N.
 *
 *   cell_capacity  jax-md's cell_list capacity (fixes the within-cell candidate rotation);
 *                  pass 0 to only measure: stats are written and no list is produced.
 *   idx            dev int32[2*e_cap] or NULL when e_cap == 0 (count only)
 *   stats          dev int32[4]: [0] true edge count E, [1] max cell occupancy,
 *                  [2] overflow bits (OR-ed into the previous value: sticky, like jax-md),
 *                  [3] reserved
 */
int lb200_nbr_build(const lb200_grid* g, const void* pos_dev, int32_t cell_capacity,
                    int32_t* idx_dev, int32_t e_cap, int32_t* stats_dev, void* scratch_dev,
                    int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Neighbor search straight into the receiver-major view (cell-list grids only): the same edges as
 * lb200_nbr_build followed by lb200_csr_build -- the in-edges of every receiver in ascending list
 * position -- without materialising the list; the step loop (lb200_rollout_steps) uses it.
 *
 *   pos_stride     elements between consecutive particles in pos_dev (dim for a compact array;
 *                  t_window * dim with pos_dev pointing at the most recent frame of a window)
 *   n_receivers    only receivers < n_receivers get a bucket (domain decomposition: ghosts send but
 *                  do not receive); 0 = all
 *   rowptr/snd/rcv as lb200_csr_build; rowptr is clamped to e_cap (edges past the capacity are dropped
 *                  and LB200_OVF_NEIGHBOR_LIST is raised, as for the truncated list)
 *   edge_feat      dev float[e_cap][4] or NULL: rel_disp | rel_dist | 0 per SLOT (features.py:115-124)
 *   tmp            dev int32[e_cap] scratch
 *   stats          as lb200_nbr_build
 */
int lb200_nbr_csr_build(const lb200_grid* g, const void* pos_dev, int64_t pos_stride, int32_t cell_capacity,
                        int32_t n_receivers, int32_t* rowptr_dev, int32_t* snd_dev, int32_t* rcv_dev,
                        float* edge_feat_dev, int32_t* tmp_dev, int32_t e_cap, int32_t* stats_dev,
                        void* scratch_dev, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Receiver-major view of an edge list for the deterministic segmented aggregation.
 * Replaces the scatter inside jraph.segment_sum (third-party jraph 0.0.6.dev0; call site
 * lagrangebench/models/gns.py:117-119).  Accepts ANY (2, e_cap) list (pad = index >= n):
 * real edges are bucketed by receiver, each bucket in ascending list position (the order a
 * sequential scatter-add would visit them).
 *
 *   rowptr   dev int32[n+1]   in-edges of receiver v occupy slots rowptr[v]..rowptr[v+1]
 *   perm     dev int32[e_cap] slot -> position in the input list
 *   snd, rcv dev int32[e_cap] sender / receiver per slot
 *   rowptr[n] is the number of real edges (device-side; kernels read it there).
 */
int lb200_csr_build(const int32_t* idx_dev, int32_t n, int32_t e_cap, int32_t* rowptr_dev,
                    int32_t* perm_dev, int32_t* snd_dev, int32_t* rcv_dev, void* scratch_dev,
                    int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Feature transform.  Replaces feature_transform (features.py:47-126) for the columns the
 * GNS consumes (models/gns.py:139-149), written as float32 (the jmp policy casts the
 * model inputs to float32, lagrangebench/runner.py:71-72).
 *
 *   window     dev T[n][t_window][dim]   position history, most recent last
 *   node_feat  dev float[n][node_stride] columns: vel_hist (K*dim) | vel_mag (K) |
 *                                        bound (2*dim) | force (dim), each optional
 *   edge_feat  dev float[e_cap][4]       rel_disp (dim) | rel_dist (1) | zero pad, in LIST order
 *   force: piecewise-constant along one axis (covers the reference datasets' force.py:
 *          RPF +-x by y, DAM gravity); force_mode 0 = none, 1 = piecewise,
 *          2 = caller-provided dev float[n][dim] in force_dev.
 */
typedef struct {
  int32_t n, dim, t_window, pos_f64, periodic;
  double box[3];
  double r_cutoff;
  double vel_mean[3], vel_std[3];
  int32_t magnitude_features;
  int32_t bound_features;    /* features.py:87: only when no dimension is periodic */
  double bounds_lo[3], bounds_hi[3];
  int32_t force_mode;
  int32_t force_axis;
  double force_threshold;    /* pos[axis] > threshold ? force_hi : force_lo */
  double force_lo[3], force_hi[3];
  int32_t node_stride;       /* floats per node_feat row (>= number of feature columns) */
  /* optional: append the particle-type embedding (gns.py:61-63,141-145) behind the feature columns, i.e. write
   * the node encoder's whole input row [features | Embed(ptype) | 0 ...] (embed_size = 0: features only) */
  int32_t embed_size, num_particle_types;
  const float* embedding_dev;  /* dev float[num_particle_types][embed_size] */
  const int32_t* ptype_dev;    /* dev int32[n] */
} lb200_feature_cfg;

int32_t lb200_node_feature_width(const lb200_feature_cfg* c);

int lb200_features(const lb200_feature_cfg* c, const void* window_dev, const float* force_dev,
                   const int32_t* idx_dev, int32_t e_cap, float* node_feat_dev,
                   float* edge_feat_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * GNS forward.  Replaces GNS.__call__ (models/gns.py:159-171): type embedding, encoder
 * (gns.py:65-81), num_mp_steps x jraph.GraphNetwork with residuals (gns.py:83-124),
 * decoder (gns.py:126-133); build_mlp = Linear-ReLU-Linear(+LayerNorm)
 * (models/utils.py:100-115).  All arithmetic float32.
 *
 * Weights: one device blob of floats; every matrix row-major (in, out) exactly as
 * hk.Linear stores `w`.  Offsets (in floats) into the blob:
 */
typedef struct {
  int64_t w0, b0, w1, b1, ln_scale, ln_offset; /* ln_* < 0: no LayerNorm */
  /* tensor-core operands of a processor edge MLP (< 0: absent): tc_w = four 128x128 fp16
   * matrices W1e^T hi|lo, W2c^T hi|lo in the UMMA no-swizzle K-major layout (element (m,k) at
   * half index (k/8)*1024 + m*8 + k%8), split as x = hi + lo * 2^-11; W2c = W2 with the
   * LayerNorm mean folded in (each row minus its mean).  tc_vec = b2c | ln_scale | ln_offset. */
  int64_t tc_w, tc_vec;
} lb200_mlp_off;

/* ------------------------------------------------------------------------------------
 * Domain decomposition of ONE cloud over the GPUs of a box (no reference counterpart: the reference
 * is single-device).  One process per GPU; the box is cut into slabs along `axis`; a
 * rank's local cloud is  [ owned | ghosts from the left neighbour | ghosts from the right neighbour ].
 * Ghost sets are chosen on the host (halo = cutoff + margin) and stay FIXED until a particle has
 * moved more than drift_limit along the cut axis (LB200_OVF_DRIFT), so that steps need no host
 * round trip: every exchange is a store into peer-mapped memory plus a device-side signal.
 *
 * Peer heap: one cudaMalloc block per rank, identical layout everywhere (lb200_peer_heap_layout),
 * shared with the other processes by CUDA IPC:  control words | positions of the local cloud
 * [n_cap][3] double | two projection arrays [n_cap][256] float (alternating by message-passing step).
 */
typedef struct lb200_shard_s {
  int32_t rank, world;
  int32_t has_left, has_right;         /* 0: open boundary on that side (non-periodic cut axis) */
  int32_t n_owned, n_ghost_left, n_ghost_right;
  int32_t n_send_left, n_send_right;
  const int32_t* push_left;            /* dev int32[n_owned]: k if row v is the k-th row the left neighbour holds */
  const int32_t* push_right;           /*                     as a ghost (ascending v), -1 otherwise */
  int32_t dst_row_left, dst_row_right; /* first row of my block inside the left / right neighbour's local cloud */
  int32_t axis;
  int32_t n_cap;                       /* rows per array of the heap layout (uniform across ranks) */
  double shift_left, shift_right;      /* added to the cut-axis coordinate of rows sent left / right (periodic wrap) */
  double axis_length;                  /* box length along the cut axis if it is periodic, else 0 */
  const void* ref_coord;               /* dev T[n_owned]: cut-axis coordinate when the ghost sets were chosen */
  double drift_limit;
  void* heap;                          /* this rank's heap */
  void* heap_left;                     /* the neighbours' heaps mapped into this process (or NULL) */
  void* heap_right;
  void* heap_all[LB200_MAX_RANKS];     /* every rank's heap ([rank] == heap): status bits go to all of them */
} lb200_shard;

/* bytes of a heap for n_cap rows, and the byte offsets of its parts */
int64_t lb200_peer_heap_layout(int32_t n_cap, int64_t* off_pos, int64_t* off_p0, int64_t* off_p1);
/* cudaMalloc + zeroed control words; handle64 receives the 64-byte CUDA IPC handle other processes open */
int lb200_peer_heap_create(int64_t bytes, void** heap_out, void* handle64_out);
int lb200_peer_heap_open(const void* handle64, void** heap_out);
int lb200_peer_heap_close(void* peer_heap);   /* a mapping obtained from lb200_peer_heap_open */
int lb200_peer_heap_destroy(void* heap);      /* a heap obtained from lb200_peer_heap_create */

typedef struct {
  int32_t n, dim, num_mp_steps;
  int32_t node_in;          /* node feature columns (without embedding) */
  int32_t node_stride;      /* floats per node_feat row */
  int32_t embed_size;       /* particle_type_embedding_size (16); 0 = no embedding */
  int32_t num_particle_types;
  int32_t e_cap;
  int64_t embedding;        /* offset of the (num_particle_types, embed_size) table */
  lb200_mlp_off enc_node, enc_edge, dec;
  const lb200_mlp_off* proc_edge; /* host array [num_mp_steps] */
  const lb200_mlp_off* proc_node; /* host array [num_mp_steps] */
  int32_t edge_impl;        /* 0: tcgen05 tensor-core message kernel (product path: weights in TMEM,
                                  pipelined tiles);
                               1: fp32 CUDA-core kernel (kept as the numerical cross-check);
                               2: first tensor-core kernel (weights in shared memory) */
  /* Domain decomposition (0 / NULL on a single GPU): rows [0, n_owned) of the node arrays are this
   * rank's particles, rows [n_owned, n) ghosts.  Node kernels run over owned rows, edge kernels over
   * the edges whose receiver is owned; the sender projections of the ghost rows arrive from the
   * neighbouring ranks through peer-mapped memory (lb200_shard below): the node kernel's epilogue stores
   * its boundary rows into the neighbours' arrays, one signal/wait kernel per message-passing step. */
  int32_t n_owned;
  const struct lb200_shard_s* shard;
  /* dev int32* or NULL: OR-ed with 1 when an output acceleration is NaN / Inf (see LB200_ERR_NONFINITE) */
  int32_t* nonfinite_flag;
  /* 1: the rows of node_feat already end with the particle-type embedding -- [features | Embed(ptype) | 0 ...],
   * lb200_features with embed_size > 0, node_stride a multiple of 4 -- and the node encoder reads them as they
   * are (no separate embedding pass).  Tensor-core path only. */
  int32_t node_feat_embedded;
  /* the model's latent width (gns.py:37 latent_size), 0 = 128.  Narrower models (the published GNS-5-64) run on the
   * 128-wide kernels: every latent dimension of the weights is zero-padded to 128 when they are packed, and
   * LayerNorm divides by this width (the padding columns are exactly zero throughout). */
  int32_t latent;
} lb200_gns_cfg;

/* device scratch the forward needs, in bytes (node latents, projections, edge latents ...) */
int64_t lb200_gns_scratch_bytes(int32_t n, int32_t e_cap);
/* byte offsets inside that scratch of h [n][128], P [n][256], agg [n][128], e [e_cap][128] */
int lb200_gns_scratch_layout(int32_t n, int32_t e_cap, int64_t* off_h, int64_t* off_p, int64_t* off_agg,
                             int64_t* off_e);

/*
 *   node_feat, edge_feat  as produced by lb200_features (edge_feat in LIST order)
 *   ptype                 dev int32[n]
 *   rowptr/perm/snd/rcv   as produced by lb200_csr_build; perm may be NULL when edge_feat is already in
 *                         SLOT order (lb200_nbr_csr_build)
 *   out                   dev float[n][dim]  normalised acceleration ({"acc": ...})
 */
int lb200_gns_forward(const lb200_gns_cfg* c, const float* weights_dev, const float* node_feat_dev,
                      const float* edge_feat_dev, const int32_t* ptype_dev,
                      const int32_t* rowptr_dev, const int32_t* perm_dev, const int32_t* snd_dev,
                      const int32_t* rcv_dev, float* out_dev, void* scratch_dev,
                      int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Integrate + kinematic override + window shift.  Replaces case.integrate
 * (case.py:230-259) followed by the tail of _forward_eval (evaluate/rollout.py:61-73):
 *   acc = mean + out*std; v = disp(p[-1], p[-2]) + acc; p' = shift(p[-1], v);
 *   p' = target where particle_type in {SOLID_WALL, MOVING_WALL, PAD};
 *   window <- concat(window[:, 1:], p').
 *   out_mode 0: "acc", 1: "vel", 2: "pos" (case.py:235-256)
 *   target    dev T[n][dim] or NULL (no override)
 *   pred_out  dev T[n][dim] or NULL: also receives p' (the rollout's prediction row)
 *   skip_flag dev int32* or NULL: if *skip_flag != 0 the kernel does nothing (the neighbor
 *             list overflowed; the host re-allocates and retries the same step,
 *             evaluate/rollout.py:135-151)
 */
typedef struct {
  int32_t n, dim, t_window, pos_f64, periodic, out_mode;
  double box[3];
  double mean[3], std[3]; /* acceleration (or velocity) normalisation */
} lb200_integrate_cfg;

int lb200_integrate(const lb200_integrate_cfg* c, const float* out_dev, void* window_dev,
                    const int32_t* ptype_dev, const void* target_dev, void* pred_out_dev,
                    const int32_t* skip_flag_dev, void* stream);

/* ------------------------------------------------------------------------------------
 * Device-resident rollout.  Replaces the body of the while-loop of _eval_batched_rollout
 * (evaluate/rollout.py:125-169) for `n_steps` consecutive steps without a host round trip:
 * neighbor update -> features -> forward -> integrate -> store prediction.  The blocking
 * overflow read of rollout.py:135 becomes a device flag: once a step overflows, that step
 * and all later ones are no-ops and status[0] holds the index of the first such step; the
 * host then re-allocates and calls again from that step (same retry contract).
 *
 *   targets   dev T[..][n][dim] or NULL        ground-truth positions for kinematic particles
 *   preds     dev T[..][n][dim]                predicted positions per step
 *   first_frame                                step k of this call reads targets[first_frame + k] and writes
 *                                              preds[first_frame + k]: a caller that advances a long rollout in
 *                                              chunks passes the same base pointers every time (one cached step graph)
 *   idx       dev int32[2][e_cap] or NULL     the jax-md ordered list of every step, when the caller wants it;
 *                                             NULL (cell-list grids): the step works on the receiver-major
 *                                             view alone (lb200_nbr_csr_build) -- same edges, same results,
 *                                             and the list of any state can be had from lb200_nbr_build
 *   status    dev int32[4]: [0] steps completed, [1] overflow bits, [2] last E, [3] first_frame
 */
typedef struct {
  lb200_grid grid;
  lb200_feature_cfg feat;
  lb200_gns_cfg gns;
  lb200_integrate_cfg integ;
  int32_t cell_capacity;
  int32_t e_cap;
  /* decomposed cloud (NULL: single GPU): grid.n = gns.n = local cloud (owned + ghosts), feat.n = integ.n =
   * gns.n_owned = owned rows; window / ptype / targets / preds hold the OWNED rows; gns.shard = this too */
  const lb200_shard* shard;
} lb200_rollout_cfg;

int64_t lb200_rollout_scratch_bytes(const lb200_rollout_cfg* c);

int lb200_rollout_steps(const lb200_rollout_cfg* c, int32_t n_steps, const float* weights_dev,
                        void* window_dev, const int32_t* ptype_dev, const float* force_dev,
                        const void* targets_dev, void* preds_dev, int32_t first_frame, int32_t* idx_dev,
                        int32_t* status_dev, void* scratch_dev, int64_t scratch_bytes,
                        void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): a cumulative count of kernels this library launched, and
 * optional CUDA-event timing of the two message-passing kernels on their launch stream.
 *   lb200_profile(1) enables + resets, lb200_profile(0) disables;
 *   lb200_profile_read synchronises the recorded events and returns the summed device
 *   time (ms) and launch count of [0] the edge (message+aggregate) kernel and [1] the
 *   node-update kernel since the last reset.
 */
/* Hardware self-test of the tcgen05 features the message kernel relies on: writes five floats to
 * out5_dev -- [0] max |D_ts - D_ss| (A operand read from tensor memory vs shared memory),
 * [1] max |D_scaled - expected| (scale-input-d accumulate), [2] max |D_ss| (must be non-zero),
 * [3] max |D_mn - D_ss| (B operand in the MN-major core-matrix layout), [4] max |D_sub 2^18 - D_ss|
 * (fp16 subnormal inputs are not flushed); [0], [1], [3], [4] must be exactly 0. */
int lb200_tc_selftest(float* out5_dev, void* stream);

int64_t lb200_launch_count(void);
int lb200_profile(int32_t enable);
int lb200_profile_read(double* ms_out2, int64_t* launches_out2);

#ifdef __cplusplus
}
#endif
#endif /* LB200_H */
