// features.cu -- per-edge and per-node feature extraction.
//   edge features: GeometricGraph.extract_node_pair_features (reference
//                  graph_constructor/graph.py:139-223) with get_En_equivariant_point_pair_metrics
//                  (graph_constructor/features.py:6-122), all arithmetic in fp64;
//   degree:        Graph.get_degree (graph.py:93-96), the degree of the undirected graph
//                  networkx builds from the adjacency matrix;
//   node features: GeometricGraph.extract_single_node_features (graph.py:225-275).
#include <math.h>

#include "features.cuh"

namespace rgnn {
namespace {

template <typename T>
__device__ __forceinline__ void load_vec(const T* __restrict__ base, int64_t row, int dims, double* out) {
  if (dims == 2) {
    if constexpr (sizeof(T) == 4) {
      const float2 t = *reinterpret_cast<const float2*>(base + row * 2);
      out[0] = t.x; out[1] = t.y;
    } else {
      const double2 t = *reinterpret_cast<const double2*>(base + row * 2);
      out[0] = t.x; out[1] = t.y;
    }
    out[2] = out[3] = 0.0;
  } else {
    for (int d = 0; d < 4; ++d) out[d] = d < dims ? static_cast<double>(base[row * dims + d]) : 0.0;
  }
}

__device__ __forceinline__ double norm_of(const double* a, int dims) {
  double acc = 0.0;
  for (int d = 0; d < dims; ++d) acc = __dadd_rn(acc, __dmul_rn(a[d], a[d]));
  return sqrt(acc);
}

__device__ __forceinline__ double dot_of(const double* a, const double* b, int dims) {
  double acc = 0.0;
  for (int d = 0; d < dims; ++d) acc = __dadd_rn(acc, __dmul_rn(a[d], b[d]));
  return acc;
}

// features.py:24-40: a velocity whose components are all exactly zero stays zero
__device__ __forceinline__ void unit_or_zero(const double* a, int dims, double* out) {
  bool zero = true;
  for (int d = 0; d < dims; ++d) zero = zero && (a[d] == 0.0);
  const double n = zero ? 1.0 : norm_of(a, dims);
  for (int d = 0; d < 4; ++d) out[d] = (zero || d >= dims) ? 0.0 : a[d] / n;
}

// features.py:62-65: the connection vector is zeroed when its norm is 0
__device__ __forceinline__ void unit_or_zero_by_norm(const double* a, int dims, double* out) {
  const double n = norm_of(a, dims);
  const bool zero = n == 0.0;
  for (int d = 0; d < 4; ++d) out[d] = (zero || d >= dims) ? 0.0 : a[d] / n;
}

// features.py:46-56: |dot| in (1, 1 + 1e-3) snaps to +-1, anything further is an error
__device__ __forceinline__ double clamp_dot(double dot, int32_t* error_flag) {
  if (fabs(dot) > 1.0) {
    if (fabs(dot) - 1.0 < 1e-3) return dot > 0.0 ? 1.0 : -1.0;
    atomicExch(error_flag, RGNN_ERR_DOT_PRODUCT);
  }
  return dot;
}

__device__ __forceinline__ double to_degrees(double cosine) { return acos(cosine) * 180.0 / 3.141592653589793; }

__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }  // Python min(a, b)
__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }

__device__ void point_pair_features(const double* p1, const double* p2, const double* v1, const double* v2,
                                    int pos_dims, int vel_dims, int mode, int32_t* error_flag, double* out4) {
  double v1n[4], v2n[4], diff[4], dn[4];
  unit_or_zero(v1, vel_dims, v1n);
  unit_or_zero(v2, vel_dims, v2n);
  for (int d = 0; d < 4; ++d) diff[d] = d < pos_dims ? p1[d] - p2[d] : 0.0;
  out4[0] = norm_of(diff, pos_dims);
  out4[1] = to_degrees(clamp_dot(dot_of(v1n, v2n, vel_dims), error_flag));
  const int dd = pos_dims < vel_dims ? pos_dims : vel_dims;
  if (mode == RGNN_DIRECTED) {
    for (int d = 0; d < 4; ++d) diff[d] = d < pos_dims ? p2[d] - p1[d] : 0.0;
    unit_or_zero_by_norm(diff, pos_dims, dn);
    out4[2] = to_degrees(clamp_dot(dot_of(v1n, dn, dd), error_flag));
    out4[3] = to_degrees(clamp_dot(dot_of(v2n, dn, dd), error_flag));
  } else {
    double d1[4], d2[4];
    unit_or_zero_by_norm(diff, pos_dims, d1);  // p1 - p2
    for (int d = 0; d < 4; ++d) diff[d] = d < pos_dims ? p2[d] - p1[d] : 0.0;
    unit_or_zero_by_norm(diff, pos_dims, d2);
    // features.py:105-120: no clamp in undirected mode (NaN propagates like numpy's arccos)
    const double t_d1_v1 = to_degrees(dot_of(v1n, d1, dd)), t_d1_v2 = to_degrees(dot_of(v2n, d1, dd));
    const double t_d2_v1 = to_degrees(dot_of(v1n, d2, dd)), t_d2_v2 = to_degrees(dot_of(v2n, d2, dd));
    const double t1 = py_min(t_d1_v1, t_d2_v1), t2 = py_min(t_d1_v2, t_d2_v2);
    out4[2] = py_min(t1, t2);
    out4[3] = py_max(t1, t2);
  }
}

template <typename T, typename OT>
__global__ void __launch_bounds__(256)
edge_features_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int pos_dims, int vel_dims,
                     const int64_t* __restrict__ edge_index, int64_t n_edges, EdgeFeatureSpec spec,
                     OT* __restrict__ out, int32_t* __restrict__ error_flag) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t i = edge_index[e], j = edge_index[n_edges + e];
  double xi[4], xj[4], vi[4], vj[4];
  load_vec(pos, i, pos_dims, xi);
  load_vec(pos, j, pos_dims, xj);
  load_vec(vel, i, vel_dims, vi);
  load_vec(vel, j, vel_dims, vj);
  OT* row = out + e * spec.width;
  int col = 0;
  for (int f = 0; f < spec.n; ++f) {
    switch (spec.feature[f]) {
      case RGNN_EF_POINT_PAIR_FEATURES: {
        double ppf[4];
        point_pair_features(xi, xj, vi, vj, pos_dims, vel_dims, spec.edge_mode, error_flag, ppf);
        for (int c = 0; c < 4; ++c) row[col + c] = static_cast<OT>(ppf[c]);
        col += 4;
        break;
      }
      case RGNN_EF_SPATIAL_EUCLIDEAN_DISTANCE: {
        double d[4];
        for (int c = 0; c < 4; ++c) d[c] = c < pos_dims ? xi[c] - xj[c] : 0.0;
        row[col++] = static_cast<OT>(norm_of(d, pos_dims));
        break;
      }
      case RGNN_EF_VELOCITY_EUCLIDEAN_DISTANCE: {
        double d[4];
        for (int c = 0; c < 4; ++c) d[c] = c < vel_dims ? vi[c] - vj[c] : 0.0;
        row[col++] = static_cast<OT>(norm_of(d, vel_dims));
        break;
      }
      case RGNN_EF_RELATIVE_POSITION: {  // graph.py:198-207: components 0 and 1 only
        double dx = xi[0] - xj[0], dy = xi[1] - xj[1];
        if (spec.edge_mode == RGNN_UNDIRECTED) { dx = fabs(dx); dy = fabs(dy); }
        row[col] = static_cast<OT>(dx); row[col + 1] = static_cast<OT>(dy);
        col += 2;
        break;
      }
      case RGNN_EF_RELATIVE_VELOCITY: {
        double du = vi[0] - vj[0], dv = vi[1] - vj[1];
        if (spec.edge_mode == RGNN_UNDIRECTED) { du = fabs(du); dv = fabs(dv); }
        row[col] = static_cast<OT>(du); row[col + 1] = static_cast<OT>(dv);
        col += 2;
        break;
      }
      default: break;
    }
  }
}

// first edge row whose source is >= v (edge_index[0] ascending)
__device__ __forceinline__ int64_t row_lower_bound(const int64_t* __restrict__ src, int64_t n_edges, int64_t v) {
  int64_t lo = 0, hi = n_edges;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (src[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// degree of the undirected graph: every unordered pair {i, j} with i->j or j->i counts once at
// both ends.  Edge (i -> j) adds 1 at i, and 1 at j unless the reverse edge exists (in which case
// j's own edge adds it).
__global__ void __launch_bounds__(256)
undirected_degree_kernel(const int64_t* __restrict__ edge_index, int64_t n_edges, int32_t* __restrict__ degree) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + n_edges;
  const int64_t i = src[e], j = dst[e];
  if (e > 0 && src[e - 1] == i) {
    // parallel edges (not produced by the builders) must not be counted twice
    for (int64_t p = e - 1; p >= 0 && src[p] == i; --p)
      if (dst[p] == j) return;
  }
  if (i == j) { atomicAdd(&degree[i], 2); return; }  // networkx counts a self loop twice
  atomicAdd(&degree[i], 1);
  bool reverse = false;
  for (int64_t p = row_lower_bound(src, n_edges, j); p < n_edges && src[p] == j; ++p)
    if (dst[p] == i) { reverse = true; break; }
  if (!reverse) atomicAdd(&degree[j], 1);
}

template <typename OT>
__global__ void __launch_bounds__(256)
node_features_kernel(const double* __restrict__ rcs, const double* __restrict__ time_index,
                     const int32_t* __restrict__ degree, const double* __restrict__ pos,
                     const double* __restrict__ vel, int64_t n, NodeFeatureSpec spec, OT* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  OT* row = out + i * spec.width;
  int col = 0;
  for (int f = 0; f < spec.n; ++f) {
    switch (spec.feature[f]) {
      case RGNN_NF_RCS: row[col++] = static_cast<OT>(rcs[i]); break;
      case RGNN_NF_TIME_INDEX: row[col++] = static_cast<OT>(time_index[i]); break;
      case RGNN_NF_DEGREE: row[col++] = static_cast<OT>(degree[i]); break;
      case RGNN_NF_VELOCITY_VECTOR_LENGTH: {
        const double v[2] = {vel[i * 2], vel[i * 2 + 1]};
        row[col++] = static_cast<OT>(norm_of(v, 2));
        break;
      }
      case RGNN_NF_VELOCITY_VECTOR:
        row[col] = static_cast<OT>(vel[i * 2]); row[col + 1] = static_cast<OT>(vel[i * 2 + 1]);
        col += 2;
        break;
      case RGNN_NF_SPATIAL_COORDINATES:
        row[col] = static_cast<OT>(pos[i * 2]); row[col + 1] = static_cast<OT>(pos[i * 2 + 1]);
        col += 2;
        break;
      default: break;
    }
  }
}

}  // namespace

int make_edge_feature_spec(const int32_t* features_host, int32_t n_features, int32_t edge_mode, EdgeFeatureSpec* spec) {
  if (n_features < 0 || n_features > RGNN_MAX_EDGE_FEATURES || (n_features > 0 && features_host == nullptr))
    return RGNN_ERR_INVALID_ARGUMENT;
  if (edge_mode != RGNN_DIRECTED && edge_mode != RGNN_UNDIRECTED) return RGNN_ERR_INVALID_ARGUMENT;
  spec->n = n_features;
  spec->edge_mode = edge_mode;
  spec->width = 0;
  for (int f = 0; f < n_features; ++f) {
    spec->feature[f] = features_host[f];
    switch (features_host[f]) {
      case RGNN_EF_POINT_PAIR_FEATURES: spec->width += 4; break;
      case RGNN_EF_SPATIAL_EUCLIDEAN_DISTANCE:
      case RGNN_EF_VELOCITY_EUCLIDEAN_DISTANCE: spec->width += 1; break;
      case RGNN_EF_RELATIVE_POSITION:
      case RGNN_EF_RELATIVE_VELOCITY: spec->width += 2; break;
      default: return RGNN_ERR_INVALID_FEATURE;
    }
  }
  return RGNN_OK;
}

int launch_edge_features(const void* pos, const void* vel, int32_t in_dtype, int32_t pos_dims, int32_t vel_dims,
                         const int64_t* edge_index, int64_t n_edges, const EdgeFeatureSpec& spec,
                         void* edge_attr, int32_t out_dtype, int32_t* error_flag, cudaStream_t stream) {
  RGNN_CUDA_CHECK(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), stream));
  if (n_edges == 0 || spec.width == 0) return RGNN_OK;
  RGNN_PROFILE("edge_features", stream);
  const unsigned blocks = div_up(n_edges, 256);
#define RGNN_EF_LAUNCH(T, OT)                                                                          \
  edge_features_kernel<T, OT><<<blocks, 256, 0, stream>>>(static_cast<const T*>(pos),                  \
      static_cast<const T*>(vel), pos_dims, vel_dims, edge_index, n_edges, spec,                       \
      static_cast<OT*>(edge_attr), error_flag)
  if (in_dtype == RGNN_F32 && out_dtype == RGNN_F32) RGNN_EF_LAUNCH(float, float);
  else if (in_dtype == RGNN_F32) RGNN_EF_LAUNCH(float, double);
  else if (out_dtype == RGNN_F32) RGNN_EF_LAUNCH(double, float);
  else RGNN_EF_LAUNCH(double, double);
#undef RGNN_EF_LAUNCH
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // namespace rgnn

using namespace rgnn;

extern "C" {

int32_t rgnn_edge_feature_width(const int32_t* features_host, int32_t n_features) {
  EdgeFeatureSpec spec;
  if (make_edge_feature_spec(features_host, n_features, RGNN_DIRECTED, &spec) != RGNN_OK) return -1;
  return spec.width;
}

int rgnn_edge_features(const void* pos, const void* vel, int32_t in_dtype, int32_t pos_dims, int32_t vel_dims,
                       int64_t n_points, const int64_t* edge_index, int64_t n_edges,
                       const int32_t* features_host, int32_t n_features, int32_t edge_mode, void* edge_attr,
                       int32_t out_dtype, int32_t* error_flag, rgnn_stream_t stream) {
  EdgeFeatureSpec spec;
  RGNN_RETURN_IF_ERROR(make_edge_feature_spec(features_host, n_features, edge_mode, &spec));
  if (pos_dims < 2 || pos_dims > 4 || vel_dims < 2 || vel_dims > 4) return RGNN_ERR_INVALID_ARGUMENT;
  if ((in_dtype != RGNN_F32 && in_dtype != RGNN_F64) || (out_dtype != RGNN_F32 && out_dtype != RGNN_F64))
    return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges < 0 || n_points < 0 || error_flag == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && (pos == nullptr || vel == nullptr || edge_index == nullptr || edge_attr == nullptr))
    return RGNN_ERR_INVALID_ARGUMENT;
  return launch_edge_features(pos, vel, in_dtype, pos_dims, vel_dims, edge_index, n_edges, spec, edge_attr,
                              out_dtype, error_flag, static_cast<cudaStream_t>(stream));
}

int rgnn_undirected_degree(const int64_t* edge_index, int64_t n_edges, int64_t n_points, int32_t* degree,
                           rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_points < 0 || n_edges < 0 || (n_points > 0 && degree == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_points == 0) return RGNN_OK;
  RGNN_CUDA_CHECK(cudaMemsetAsync(degree, 0, sizeof(int32_t) * n_points, stream));
  if (n_edges == 0) return RGNN_OK;
  if (edge_index == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  undirected_degree_kernel<<<div_up(n_edges, 256), 256, 0, stream>>>(edge_index, n_edges, degree);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int32_t rgnn_node_feature_width(const int32_t* features_host, int32_t n_features) {
  if (n_features < 0 || n_features > RGNN_MAX_NODE_FEATURES || (n_features > 0 && features_host == nullptr)) return -1;
  int32_t w = 0;
  for (int f = 0; f < n_features; ++f) {
    switch (features_host[f]) {
      case RGNN_NF_RCS: case RGNN_NF_TIME_INDEX: case RGNN_NF_DEGREE: case RGNN_NF_VELOCITY_VECTOR_LENGTH: w += 1; break;
      case RGNN_NF_VELOCITY_VECTOR: case RGNN_NF_SPATIAL_COORDINATES: w += 2; break;
      default: return -1;
    }
  }
  return w;
}

int rgnn_node_features(const double* rcs, const double* time_index, const int32_t* degree, const double* pos,
                       const double* vel, int64_t n_points, const int32_t* features_host, int32_t n_features,
                       void* out, int32_t out_dtype, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NodeFeatureSpec spec;
  spec.width = rgnn_node_feature_width(features_host, n_features);
  if (spec.width < 0) return RGNN_ERR_INVALID_FEATURE;
  spec.n = n_features;
  for (int f = 0; f < n_features; ++f) {
    spec.feature[f] = features_host[f];
    const int32_t ft = features_host[f];
    if ((ft == RGNN_NF_RCS && rcs == nullptr) || (ft == RGNN_NF_TIME_INDEX && time_index == nullptr) ||
        (ft == RGNN_NF_DEGREE && degree == nullptr) ||
        ((ft == RGNN_NF_VELOCITY_VECTOR || ft == RGNN_NF_VELOCITY_VECTOR_LENGTH) && vel == nullptr) ||
        (ft == RGNN_NF_SPATIAL_COORDINATES && pos == nullptr))
      return n_points > 0 ? RGNN_ERR_INVALID_ARGUMENT : RGNN_OK;
  }
  if (out_dtype != RGNN_F32 && out_dtype != RGNN_F64) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_points == 0 || spec.width == 0) return RGNN_OK;
  if (out == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (out_dtype == RGNN_F32)
    node_features_kernel<float><<<div_up(n_points, 256), 256, 0, stream>>>(rcs, time_index, degree, pos, vel,
                                                                          n_points, spec, static_cast<float*>(out));
  else
    node_features_kernel<double><<<div_up(n_points, 256), 256, 0, stream>>>(rcs, time_index, degree, pos, vel,
                                                                           n_points, spec, static_cast<double*>(out));
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // extern "C"
