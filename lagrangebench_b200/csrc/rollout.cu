// Device-resident rollout: the body of _eval_batched_rollout's while-loop
// (lagrangebench/evaluate/rollout.py:125-169) for several consecutive steps, enqueued on one
// stream with no host round trip.  The reference's blocking overflow read (rollout.py:135)
// becomes a sticky device flag that turns the remaining integrate steps into no-ops.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace lb {

std::atomic<int64_t> g_launches{0};

int early_issue_mask() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LB200_EARLY");
    on = e ? (atoi(e) & 3) : 3;
  }
  return on;
}

bool pdl_enabled(int64_t work_items) {
  static int mode = -1;  // 0 off, 1 on, 2 by size
  if (mode < 0) {
    const char* e = getenv("LB200_PDL");
    mode = e ? (e[0] == '0' ? 0 : 1) : 2;
  }
  return mode == 1 || (mode == 2 && work_items <= kPdlMaxEdges);
}

int device_sm_count(int* rc) {
  static int sms[kMaxDevices];
  const int dev = device_slot(rc);
  if (dev < 0) return 0;
  if (sms[dev] == 0) {
    int v = 0;
    const cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
      *rc = (int)e;
      return 0;
    }
    sms[dev] = v;
  }
  return sms[dev];
}

// the graph cache and the profiling state are process-wide: one lock around every entry point that touches them
static std::mutex g_state_mutex;

// ---- optional per-kernel-class CUDA-event timing (bench.py's roofline leg)
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> pool;       // start/stop pairs, reused after each read
  std::vector<int> cls;                // class of pair i
  size_t used = 0;                     // pairs handed out since the last reset
  cudaEvent_t pending_start = nullptr;
};
static ProfState g_prof;

bool prof_enabled() { return g_prof.on; }

void prof_begin(int cls, cudaStream_t s) {
  if (!g_prof.on) return;
  if (g_prof.used * 2 + 2 > g_prof.pool.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof.pool.push_back(a);
    g_prof.pool.push_back(b);
    g_prof.cls.push_back(cls);
  }
  g_prof.cls[g_prof.used] = cls;
  cudaEventRecord(g_prof.pool[g_prof.used * 2], s);
}

void prof_end(int cls, cudaStream_t s) {
  (void)cls;
  if (!g_prof.on || g_prof.used * 2 + 2 > g_prof.pool.size()) return;
  cudaEventRecord(g_prof.pool[g_prof.used * 2 + 1], s);
  g_prof.used += 1;
}

template <typename T>
__global__ void extract_last_kernel(const T* __restrict__ window, int n, int tw, int dim, T* __restrict__ pos) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n * dim) return;
  int i = f / dim, k = f % dim;
  pos[f] = window[((int64_t)i * tw + (tw - 1)) * dim + k];
}

// status: [0] steps completed, [1] overflow bits, [2] last E
__global__ void rollout_step_done_kernel(const int32_t* __restrict__ nbr_stats, int32_t* status) {
  status[2] = nbr_stats[0];
  status[1] |= nbr_stats[2] | (nbr_stats[3] ? LB200_ERR_NONFINITE : 0);
  if (nbr_stats[2] == 0) status[0] += 1;
}

__global__ void rollout_init_kernel(int32_t* nbr_stats, int32_t* status, int first_frame) {
  nbr_stats[0] = nbr_stats[1] = nbr_stats[2] = nbr_stats[3] = 0;
  status[0] = status[1] = status[2] = 0;
  status[3] = first_frame;
}


// ---- CUDA graphs of one rollout step, kept across lb200_rollout_steps calls
struct GraphEntry {
  std::string key;
  cudaGraphExec_t exec = nullptr;
  int64_t launches_per_step = 0;
};
static std::vector<GraphEntry> g_graphs;       // small LRU (front = oldest)
static std::vector<std::string> g_seen_once;   // keys of single-step calls seen once, not yet captured
constexpr size_t kMaxGraphs = 64, kMaxSeen = 128;  // one graph per engine of a trajectory batch and per chunk shape

template <typename T>
static void key_put(std::string& k, const T& v) {
  k.append(reinterpret_cast<const char*>(&v), sizeof(T));
}

// Everything the launches of a step depend on: the configuration (including the per-step weight
// offset tables it points to), every device pointer, the stream.
static std::string graph_key(const lb200_rollout_cfg* c, const void* w, const void* window, const void* ptype,
                             const void* force, const void* targets, const void* preds, const void* idx,
                             const void* status, const void* scratch, int64_t scratch_bytes, const void* stream) {
  std::string k;
  k.reserve(sizeof(*c) + 1024);
  // raw bytes of the caller's struct (a struct assignment need not copy padding), with the two host
  // pointers blanked: the tables they point to are appended instead
  k.append(reinterpret_cast<const char*>(c), sizeof(*c));
  const size_t off_gns = offsetof(lb200_rollout_cfg, gns);
  memset(&k[off_gns + offsetof(lb200_gns_cfg, proc_edge)], 0, sizeof(void*));
  memset(&k[off_gns + offsetof(lb200_gns_cfg, proc_node)], 0, sizeof(void*));
  memset(&k[off_gns + offsetof(lb200_gns_cfg, shard)], 0, sizeof(void*));
  memset(&k[offsetof(lb200_rollout_cfg, shard)], 0, sizeof(void*));
  for (int m = 0; m < c->gns.num_mp_steps; ++m) {
    key_put(k, c->gns.proc_edge[m]);
    key_put(k, c->gns.proc_node[m]);
  }
  if (c->shard != nullptr) key_put(k, *c->shard);  // counts, tables and heap addresses of the decomposition
  const void* ptrs[] = {w, window, ptype, force, targets, preds, idx, status, scratch, stream};
  key_put(k, ptrs);
  key_put(k, scratch_bytes);
  int dev = 0;
  cudaGetDevice(&dev);
  key_put(k, dev);
  return k;
}

static GraphEntry* graph_find(const std::string& key) {
  for (size_t i = 0; i < g_graphs.size(); ++i)
    if (g_graphs[i].key == key) {
      if (i + 1 != g_graphs.size()) {  // move to the back (most recently used)
        GraphEntry e = std::move(g_graphs[i]);
        g_graphs.erase(g_graphs.begin() + i);
        g_graphs.push_back(std::move(e));
      }
      return &g_graphs.back();
    }
  return nullptr;
}

static bool graph_seen_before(const std::string& key) {
  for (size_t i = 0; i < g_seen_once.size(); ++i)
    if (g_seen_once[i] == key) {
      g_seen_once.erase(g_seen_once.begin() + i);
      return true;
    }
  if (g_seen_once.size() >= kMaxSeen) g_seen_once.erase(g_seen_once.begin());
  g_seen_once.push_back(key);
  return false;
}

// Capture one step (nothing executes) and keep the instantiated graph.  nullptr if capture fails.
template <typename Step>
static GraphEntry* graph_capture(const std::string& key, cudaStream_t s, Step& one_step) {
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  const int64_t launches0 = g_launches.load();
  const int rc = one_step();
  const int64_t per_step = g_launches.load() - launches0;
  g_launches.fetch_sub(per_step);  // the captured launches did not execute
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaStreamEndCapture(s, &graph);
  const bool ok = rc == 0 && e == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();  // capture failed: clear the error, the caller finishes eagerly
    return nullptr;
  }
  if (g_graphs.size() >= kMaxGraphs) {
    cudaGraphExecDestroy(g_graphs.front().exec);
    g_graphs.erase(g_graphs.begin());
  }
  GraphEntry ent;
  ent.key = key;
  ent.exec = exec;
  ent.launches_per_step = per_step;
  g_graphs.push_back(std::move(ent));
  return &g_graphs.back();
}

constexpr int kNodeFeatStride = LB200_MAX_NODE_IN;  // room for [features | embedding] rows (the encoder's input width limit)

struct RolloutBufs {
  void* pos;
  int32_t *stats, *rowptr, *perm, *snd, *rcv, *idx;
  float *node_feat, *edge_feat, *out;
  void *nbr_scratch, *csr_scratch, *gns_scratch;
  int64_t nbr_bytes, csr_bytes, gns_bytes;
};

static bool carve(const lb200_rollout_cfg* c, void* scratch, int64_t bytes, RolloutBufs* b, int64_t* need) {
  const int64_t n = c->grid.n, e_cap = c->e_cap;
  Arena ar(scratch, bytes);
  b->pos = ar.take<double>(n * 3);
  b->stats = ar.take<int32_t>(8);
  b->rowptr = ar.take<int32_t>(n + 1);
  b->perm = ar.take<int32_t>(e_cap);  // cell-list grids: scratch of the direct CSR build
  b->snd = ar.take<int32_t>(e_cap);
  b->rcv = ar.take<int32_t>(e_cap);
  b->idx = c->grid.use_cells ? nullptr : ar.take<int32_t>(2 * e_cap);  // all-pairs grids go through the list
  b->node_feat = ar.take<float>(n * (c->feat.node_stride > kNodeFeatStride ? c->feat.node_stride : kNodeFeatStride));
  b->edge_feat = ar.take<float>(e_cap * 4);
  b->out = ar.take<float>(n * 3);
  b->nbr_bytes = lb200_nbr_scratch_bytes(&c->grid);
  b->csr_bytes = lb200_csr_scratch_bytes((int32_t)n, (int32_t)e_cap);
  b->gns_bytes = lb200_gns_scratch_bytes((int32_t)n, (int32_t)e_cap);
  b->nbr_scratch = ar.take<char>(b->nbr_bytes);
  b->csr_scratch = ar.take<char>(b->csr_bytes);
  b->gns_scratch = ar.take<char>(b->gns_bytes);
  if (need) *need = ar.off + 4096;
  return scratch == nullptr ? true : ar.ok();
}

}  // namespace lb

using namespace lb;

// 200: round-2 ABI; odd: a cross-check build that also carries the first-generation tensor-core kernels
extern "C" int lb200_version(void) {
#ifdef LB200_CROSSCHECK
  return 201;
#else
  return 200;
#endif
}

extern "C" int64_t lb200_launch_count(void) { return g_launches.load(); }

extern "C" int lb200_profile(int32_t enable) {
  std::lock_guard<std::mutex> lock(g_state_mutex);
  g_prof.on = enable != 0;
  g_prof.used = 0;
  return 0;
}

extern "C" int lb200_profile_read(double* ms_out2, int64_t* launches_out2) {
  if (!ms_out2 || !launches_out2) return LB200_EINVAL;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  ms_out2[0] = ms_out2[1] = 0.0;
  launches_out2[0] = launches_out2[1] = 0;
  for (size_t i = 0; i < g_prof.used; ++i) {
    LB_CHECK(cudaEventSynchronize(g_prof.pool[2 * i + 1]));
    float ms = 0.f;
    LB_CHECK(cudaEventElapsedTime(&ms, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]));
    int c = g_prof.cls[i] == 0 ? 0 : 1;
    ms_out2[c] += ms;
    launches_out2[c] += 1;
  }
  g_prof.used = 0;
  return 0;
}

extern "C" const char* lb200_error_string(int code) {
  if (code == 0) return "success";
  if (code == LB200_EINVAL) return "lb200: invalid argument";
  if (code == LB200_EUNSUPPORTED) return "lb200: unsupported configuration";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "lb200: unknown error";
}

extern "C" int64_t lb200_rollout_scratch_bytes(const lb200_rollout_cfg* c) {
  RolloutBufs b;
  int64_t need = 0;
  carve(c, nullptr, 0, &b, &need);
  return need;
}

extern "C" int lb200_rollout_steps(const lb200_rollout_cfg* c, int32_t n_steps, const float* weights_dev,
                                   void* window_dev, const int32_t* ptype_dev, const float* force_dev,
                                   const void* targets_dev, void* preds_dev, int32_t first_frame, int32_t* idx_dev,
                                   int32_t* status_dev, void* scratch_dev, int64_t scratch_bytes, void* stream) {
  if (!c || !weights_dev || !window_dev || !ptype_dev || !status_dev || !scratch_dev || n_steps < 0 || first_frame < 0)
    return LB200_EINVAL;
  RolloutBufs b;
  if (!carve(c, scratch_dev, scratch_bytes, &b, nullptr)) return LB200_EINVAL;
  std::lock_guard<std::mutex> lock(g_state_mutex);  // graph cache, profiling events
  cudaStream_t s = (cudaStream_t)stream;
  const int n = c->grid.n, dim = c->grid.dim, tw = c->feat.t_window;
  const int64_t esz = c->grid.pos_f64 ? 8 : 4;
  { rollout_init_kernel<<<1, 1, 0, s>>>(b.stats, status_dev, first_frame); LB_LAUNCHED(1); }
  // One step; targets / predictions are indexed on the device by status[0] (steps completed in this
  // call), so the very same launches serve every step -- and can be replayed from a CUDA graph.
  const bool direct = c->grid.use_cells != 0;  // cell-list grids: receiver-major view straight from the cells
  const lb200_shard* sh = c->shard;
  lb200_gns_cfg gns = c->gns;
  gns.nonfinite_flag = b.stats + 3;  // NaN / Inf accelerations (fp16 split out of range) -> status bit
  // tensor-core node encoder: the feature kernel appends the particle-type embedding to its rows and the encoder
  // reads them as they are (one kernel and one round trip through the latent array less)
  lb200_feature_cfg feat = c->feat;
  {
    const int wide = (lb200_node_feature_width(&feat) + gns.embed_size + 3) / 4 * 4;
    if (direct && gns.edge_impl != 1 && gns.enc_node.tc_w >= 0 && gns.enc_node.tc_vec >= 0 && feat.force_mode != 2 &&
        getenv("LB200_NODE_TC") == nullptr && wide <= kNodeFeatStride) {
      feat.node_stride = wide;
      feat.embed_size = gns.embed_size;
      feat.num_particle_types = gns.num_particle_types;
      feat.embedding_dev = weights_dev + gns.embedding;
      feat.ptype_dev = ptype_dev;
      gns.node_stride = wide;
      gns.node_feat_embedded = 1;
    }
  }
  if (sh != nullptr) {
    if (!direct || c->gns.shard != sh || c->gns.n_owned != sh->n_owned || c->integ.n != sh->n_owned ||
        c->feat.n != sh->n_owned || n != sh->n_owned + sh->n_ghost_left + sh->n_ghost_right || n > sh->n_cap ||
        sh->world > LB200_MAX_RANKS)
      return LB200_EINVAL;
    int rc = shard_call_init(sh, s);
    if (rc) return rc;
  }
  int32_t* const list = idx_dev != nullptr ? idx_dev : b.idx;
  auto one_step = [&]() -> int {
    int rc;
    if (sh != nullptr) {
      // decomposed cloud: [owned | ghosts]; every exchange is a store into the neighbours' heaps + one
      // signal/wait kernel; the status bits are OR-ed over all ranks before integrate takes effect
      rc = shard_push_positions(sh, window_dev, tw, dim, c->grid.pos_f64, s);
      if (rc) return rc;
      rc = shard_exchange(sh, 0, c->gns.num_mp_steps + 1, s);
      if (rc) return rc;
      rc = lb200_nbr_csr_build(&c->grid, shard_pos_local(sh), dim, c->cell_capacity, sh->n_owned, b.rowptr, b.snd,
                               b.rcv, b.edge_feat, b.perm, c->e_cap, b.stats, b.nbr_scratch, b.nbr_bytes, stream);
      if (rc) return rc;
      rc = shard_flag_bcast(sh, b.stats, s);
      if (rc) return rc;
      rc = lb200_features(&feat, window_dev, force_dev, nullptr, 0, b.node_feat, nullptr, stream);
      if (rc) return rc;
      rc = lb200_gns_forward(&gns, weights_dev, b.node_feat, b.edge_feat, ptype_dev, b.rowptr, nullptr, b.snd,
                             b.rcv, b.out, b.gns_scratch, b.gns_bytes, stream);
      if (rc) return rc;
      rc = shard_flag_wait(sh, s);
      if (rc) return rc;
      rc = integrate_indexed(&c->integ, b.out, window_dev, ptype_dev, targets_dev, preds_dev, shard_skip_flag(sh),
                             status_dev, s);
      if (rc) return rc;
      return shard_step_done(sh, b.stats, status_dev, s);
    }
    if (!direct || idx_dev != nullptr) {  // the jax-md ordered list: all-pairs grids, or a caller that asked for it
      if (c->grid.pos_f64)
        { extract_last_kernel<double><<<cdiv(n * dim, 256), 256, 0, s>>>((const double*)window_dev, n, tw, dim, (double*)b.pos); LB_LAUNCHED(1); }
      else
        { extract_last_kernel<float><<<cdiv(n * dim, 256), 256, 0, s>>>((const float*)window_dev, n, tw, dim, (float*)b.pos); LB_LAUNCHED(1); }
      rc = lb200_nbr_build(&c->grid, b.pos, c->cell_capacity, list, c->e_cap, b.stats, b.nbr_scratch, b.nbr_bytes, stream);
      if (rc) return rc;
    }
    const int32_t* perm = nullptr;
    if (direct) {
      const char* last = (const char*)window_dev + (int64_t)(tw - 1) * dim * esz;
      rc = lb200_nbr_csr_build(&c->grid, last, (int64_t)tw * dim, c->cell_capacity, 0, b.rowptr, b.snd, b.rcv,
                               b.edge_feat, b.perm, c->e_cap, b.stats, b.nbr_scratch, b.nbr_bytes, stream);
      if (rc) return rc;
      rc = lb200_features(&feat, window_dev, force_dev, nullptr, 0, b.node_feat, nullptr, stream);
      if (rc) return rc;
    } else {
      rc = lb200_csr_build(list, n, c->e_cap, b.rowptr, b.perm, b.snd, b.rcv, b.csr_scratch, b.csr_bytes, stream);
      if (rc) return rc;
      rc = lb200_features(&c->feat, window_dev, force_dev, list, c->e_cap, b.node_feat, b.edge_feat, stream);
      if (rc) return rc;
      perm = b.perm;
    }
    rc = lb200_gns_forward(&gns, weights_dev, b.node_feat, b.edge_feat, ptype_dev, b.rowptr, perm, b.snd, b.rcv,
                           b.out, b.gns_scratch, b.gns_bytes, stream);
    if (rc) return rc;
    rc = integrate_indexed(&c->integ, b.out, window_dev, ptype_dev, targets_dev, preds_dev, b.stats + 2, status_dev, s);
    if (rc) return rc;
    { rollout_step_done_kernel<<<1, 1, 0, s>>>(b.stats, status_dev); LB_LAUNCHED(1); }
    return 0;
  };
  // Launch-bound inner loop (about 100 launches per step): steps are replayed from a CUDA graph, which is
  // kept across calls -- every argument a step depends on is part of the key, so a per-step caller
  // (steps_per_sync = 1) replays too.  Not on the legacy default stream (capture is unsupported there),
  // not while per-kernel events are on.
  const bool graph_ok = s != nullptr && s != cudaStreamLegacy && !prof_enabled() &&
                        getenv("LB200_NO_GRAPH") == nullptr;
  int t = 0;
  if (graph_ok && n_steps > 0) {
    std::string key = graph_key(c, weights_dev, window_dev, ptype_dev, force_dev, targets_dev, preds_dev, idx_dev,
                                status_dev, scratch_dev, scratch_bytes, stream);
    GraphEntry* hit = graph_find(key);
    if (!hit) {
      // first sight of this configuration: one eager step (it also performs every one-time kernel attribute
      // setup), then capture -- at once when the call has steps left to replay, else when the key comes back
      int rc = one_step();
      if (rc) return rc;
      t = 1;
      if (n_steps - t >= 2 || graph_seen_before(key)) hit = graph_capture(key, s, one_step);
    }
    if (hit) {
      for (; t < n_steps; ++t) {
        if (cudaGraphLaunch(hit->exec, s) != cudaSuccess) {
          cudaGetLastError();
          break;
        }
        g_launches.fetch_add(hit->launches_per_step);
      }
    }
  }
  for (; t < n_steps; ++t) {
    int rc = one_step();
    if (rc) return rc;
  }
  LB_LAUNCH_CHECK();
  return 0;
}
