// Feature transform and integrate/window-shift kernels (elementwise, HBM-bound).
//
// Replaces lagrangebench/case_setup/features.py:47-126 (feature_transform) and
// lagrangebench/case_setup/case.py:230-259 (integrate) + evaluate/rollout.py:61-73
// (kinematic override, window shift).  Arithmetic is done in the position dtype with
// round-to-nearest ops in the reference's order (no FMA contraction), results stored as
// float32 -- the values the float32 network consumes.
#include "common.cuh"

namespace lb {

struct FeatDev {
  int n, dim, tw, periodic, mag, bound, force_mode, force_axis, stride, embed, n_types;
  const float* embedding;
  const int32_t* ptype;
  double box[3], r, vmean[3], vstd[3], lo[3], hi[3], fthr, flo[3], fhi[3];
};

template <typename T, int DIM>
__global__ void node_feature_kernel(FeatDev c, const T* __restrict__ window, const float* __restrict__ force_in,
                                    float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  T side[DIM], half[DIM], vm[DIM], vs[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    side[k] = (T)c.box[k];
    half[k] = mul_rn(side[k], T(0.5));
    vm[k] = (T)c.vmean[k];
    vs[k] = (T)c.vstd[k];
  }
  const T* w = window + (int64_t)i * c.tw * DIM;
  float* o = out + (int64_t)i * c.stride;
  const int K = c.tw - 1;
  int col = 0;
  T prev[DIM], cur[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) prev[k] = w[k];
  for (int t = 0; t < K; ++t) {
    T ss = T(0);
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      cur[k] = w[(t + 1) * DIM + k];
      T v = disp1(cur[k], prev[k], side[k], half[k], ((c.periodic >> k) & 1) != 0);
      T nv = div_rn(sub_rn(v, vm[k]), vs[k]);
      o[t * DIM + k] = (float)nv;
      ss = k == 0 ? mul_rn(nv, nv) : add_rn(ss, mul_rn(nv, nv));
      prev[k] = cur[k];
    }
    if (c.mag) o[K * DIM + t] = (float)sqrt_rn(ss);
  }
  col = K * DIM + (c.mag ? K : 0);
  // prev now holds the most recent position
  if (c.bound) {
    T r = (T)c.r;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      T dl = div_rn(sub_rn(prev[k], (T)c.lo[k]), r);
      T du = div_rn(sub_rn((T)c.hi[k], prev[k]), r);
      dl = dl < T(-1) ? T(-1) : (dl > T(1) ? T(1) : dl);
      du = du < T(-1) ? T(-1) : (du > T(1) ? T(1) : du);
      o[col + k] = (float)dl;
      o[col + DIM + k] = (float)du;
    }
    col += 2 * DIM;
  }
  if (c.force_mode == 1) {
    bool hi = prev[c.force_axis] > (T)c.fthr;
#pragma unroll
    for (int k = 0; k < DIM; ++k) o[col + k] = (float)(T)(hi ? c.fhi[k] : c.flo[k]);
    col += DIM;
  } else if (c.force_mode == 2) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) o[col + k] = force_in[(int64_t)i * DIM + k];
    col += DIM;
  }
  if (c.embed > 0) {  // hk.Embed (gns.py:61-63): row lookup, NumPy indexing (-1 = last row)
    int t = c.ptype[i];
    if (t < 0) t += c.n_types;
    t = min(max(t, 0), c.n_types - 1);
    for (int k = 0; k < c.embed; ++k) o[col + k] = c.embedding[t * c.embed + k];
    col += c.embed;
  }
  for (; col < c.stride; ++col) o[col] = 0.f;
}

template <typename T, int DIM>
__global__ void edge_feature_kernel(FeatDev c, const T* __restrict__ window, const int32_t* __restrict__ idx,
                                    int e_cap, float4* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_cap) return;
  int r = min(idx[e], c.n - 1), s = min(idx[e_cap + e], c.n - 1);  // JAX gathers clamp the pad index
  r = max(r, 0);
  s = max(s, 0);
  const T* pr = window + ((int64_t)r * c.tw + (c.tw - 1)) * DIM;
  const T* ps = window + ((int64_t)s * c.tw + (c.tw - 1)) * DIM;
  T radius = (T)c.r;
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  T ss = T(0);
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    T side = (T)c.box[k];
    T d = disp1(pr[k], ps[k], side, mul_rn(side, T(0.5)), ((c.periodic >> k) & 1) != 0);
    T nd = div_rn(d, radius);
    f[k] = (float)nd;
    ss = k == 0 ? mul_rn(nd, nd) : add_rn(ss, mul_rn(nd, nd));
  }
  f[DIM] = ss > T(0) ? (float)sqrt_rn(ss) : 0.f;
  out[e] = make_float4(f[0], f[1], f[2], f[3]);
}

struct IntegDev {
  int n, dim, tw, periodic, mode;
  double box[3], mean[3], std[3];
};

template <typename T, int DIM>
__global__ void integrate_kernel(IntegDev c, const float* __restrict__ net_out, T* __restrict__ window,
                                 const int32_t* __restrict__ ptype, const T* __restrict__ target,
                                 T* __restrict__ pred, const int32_t* __restrict__ skip,
                                 const int32_t* __restrict__ step_counter) {
  if (skip != nullptr && *skip != 0) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.n) return;
  if (step_counter != nullptr) {  // device-resident loop: frame index = steps completed so far
    const int64_t frame = (int64_t)(step_counter[0] + step_counter[3]) * c.n * DIM;  // [3]: first frame of this call
    if (target != nullptr) target += frame;
    if (pred != nullptr) pred += frame;
  }
  T* w = window + (int64_t)i * c.tw * DIM;
  T np[DIM];
  int pt = ptype[i];
  bool kin = (pt == 1) || (pt == 2) || (pt == -1);  // SOLID_WALL, MOVING_WALL, PAD (utils.py:28-35)
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    T side = (T)c.box[k];
    T half = mul_rn(side, T(0.5));
    const bool per = ((c.periodic >> k) & 1) != 0;
    T last = w[(c.tw - 1) * DIM + k];
    T x = (T)net_out[(int64_t)i * DIM + k];
    T res;
    if (c.mode == 2) {
      res = x;  // "pos": zeroth Euler step
    } else {
      T nv = add_rn((T)c.mean[k], mul_rn(x, (T)c.std[k]));
      if (c.mode == 0) {  // "acc": second Euler step
        T v = disp1(last, w[(c.tw - 2) * DIM + k], side, half, per);
        nv = add_rn(v, nv);
      }
      res = shift1(last, nv, side, per);
    }
    if (kin && target != nullptr) res = target[(int64_t)i * DIM + k];
    np[k] = res;
  }
  for (int t = 0; t + 1 < c.tw; ++t)
#pragma unroll
    for (int k = 0; k < DIM; ++k) w[t * DIM + k] = w[(t + 1) * DIM + k];
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    w[(c.tw - 1) * DIM + k] = np[k];
    if (pred != nullptr) pred[(int64_t)i * DIM + k] = np[k];
  }
}

template <typename T, int DIM>
static int features_t(const FeatDev& c, const void* window, const float* force, const int32_t* idx, int e_cap,
                      float* node_feat, float* edge_feat, cudaStream_t s) {
  if (node_feat)
    { node_feature_kernel<T, DIM><<<cdiv(c.n, 128), 128, 0, s>>>(c, (const T*)window, force, node_feat); LB_LAUNCHED(1); }
  if (edge_feat && idx && e_cap > 0)
    { edge_feature_kernel<T, DIM><<<cdiv(e_cap, 256), 256, 0, s>>>(c, (const T*)window, idx, e_cap, (float4*)edge_feat); LB_LAUNCHED(1); }
  LB_LAUNCH_CHECK();
  return 0;
}

}  // namespace lb

using namespace lb;

extern "C" int32_t lb200_node_feature_width(const lb200_feature_cfg* c) {
  int K = c->t_window - 1;
  return K * c->dim + (c->magnitude_features ? K : 0) + (c->bound_features ? 2 * c->dim : 0) +
         (c->force_mode ? c->dim : 0);
}

extern "C" int lb200_features(const lb200_feature_cfg* c, const void* window_dev, const float* force_dev,
                              const int32_t* idx_dev, int32_t e_cap, float* node_feat_dev, float* edge_feat_dev,
                              void* stream) {
  if (!c || !window_dev || (c->dim != 2 && c->dim != 3) || c->t_window < 1) return LB200_EINVAL;
  if (node_feat_dev && (c->t_window < 2 || c->node_stride < lb200_node_feature_width(c) + (c->embed_size > 0 ? c->embed_size : 0)))
    return LB200_EINVAL;
  if (c->embed_size > 0 && (!c->embedding_dev || !c->ptype_dev || c->num_particle_types < 1)) return LB200_EINVAL;
  if (c->force_mode == 2 && !force_dev) return LB200_EINVAL;
  FeatDev d;
  d.n = c->n;
  d.dim = c->dim;
  d.tw = c->t_window;
  d.periodic = c->periodic;
  d.mag = c->magnitude_features;
  d.bound = c->bound_features;
  d.force_mode = c->force_mode;
  d.force_axis = c->force_axis;
  d.stride = c->node_stride;
  d.embed = c->embed_size > 0 ? c->embed_size : 0;
  d.n_types = c->num_particle_types;
  d.embedding = c->embedding_dev;
  d.ptype = c->ptype_dev;
  d.r = c->r_cutoff;
  d.fthr = c->force_threshold;
  for (int k = 0; k < 3; ++k) {
    d.box[k] = c->box[k];
    d.vmean[k] = c->vel_mean[k];
    d.vstd[k] = c->vel_std[k];
    d.lo[k] = c->bounds_lo[k];
    d.hi[k] = c->bounds_hi[k];
    d.flo[k] = c->force_lo[k];
    d.fhi[k] = c->force_hi[k];
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (c->pos_f64)
    return c->dim == 2 ? features_t<double, 2>(d, window_dev, force_dev, idx_dev, e_cap, node_feat_dev, edge_feat_dev, s)
                       : features_t<double, 3>(d, window_dev, force_dev, idx_dev, e_cap, node_feat_dev, edge_feat_dev, s);
  return c->dim == 2 ? features_t<float, 2>(d, window_dev, force_dev, idx_dev, e_cap, node_feat_dev, edge_feat_dev, s)
                     : features_t<float, 3>(d, window_dev, force_dev, idx_dev, e_cap, node_feat_dev, edge_feat_dev, s);
}

namespace lb {
int integrate_indexed(const lb200_integrate_cfg* c, const float* out_dev, void* window_dev, const int32_t* ptype_dev,
                      const void* target_dev, void* pred_out_dev, const int32_t* skip_flag_dev,
                      const int32_t* step_counter_dev, cudaStream_t s) {
  if (!c || !out_dev || !window_dev || !ptype_dev || (c->dim != 2 && c->dim != 3) || c->t_window < 2)
    return LB200_EINVAL;
  IntegDev d;
  d.n = c->n;
  d.dim = c->dim;
  d.tw = c->t_window;
  d.periodic = c->periodic;
  d.mode = c->out_mode;
  for (int k = 0; k < 3; ++k) {
    d.box[k] = c->box[k];
    d.mean[k] = c->mean[k];
    d.std[k] = c->std[k];
  }
  int grid = cdiv(c->n, 128);
#define LB_INTEG(T, D)                                                                                       \
  do {                                                                                                       \
    integrate_kernel<T, D><<<grid, 128, 0, s>>>(d, out_dev, (T*)window_dev, ptype_dev, (const T*)target_dev, \
                                                (T*)pred_out_dev, skip_flag_dev, step_counter_dev);          \
    LB_LAUNCHED(1);                                                                                          \
  } while (0)
  if (c->pos_f64) {
    if (c->dim == 2) LB_INTEG(double, 2); else LB_INTEG(double, 3);
  } else {
    if (c->dim == 2) LB_INTEG(float, 2); else LB_INTEG(float, 3);
  }
#undef LB_INTEG
  LB_LAUNCH_CHECK();
  return 0;
}
}  // namespace lb

extern "C" int lb200_integrate(const lb200_integrate_cfg* c, const float* out_dev, void* window_dev,
                               const int32_t* ptype_dev, const void* target_dev, void* pred_out_dev,
                               const int32_t* skip_flag_dev, void* stream) {
  return lb::integrate_indexed(c, out_dev, window_dev, ptype_dev, target_dev, pred_out_dev, skip_flag_dev, nullptr,
                               (cudaStream_t)stream);
}
