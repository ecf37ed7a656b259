// features.cu -- per-edge and per-node feature extraction.
//   edge features: GeometricGraph.extract_node_pair_features (reference
//                  graph_constructor/graph.py:139-223) with get_En_equivariant_point_pair_metrics
//                  (graph_constructor/features.py:6-122), all arithmetic in fp64;
//   degree:        Graph.get_degree (graph.py:93-96), the degree of the undirected graph
//                  networkx builds from the adjacency matrix;
//   node features: GeometricGraph.extract_single_node_features (graph.py:225-275).
#include <math.h>

#include "edge_feature_math.cuh"
#include "features.cuh"

namespace rgnn {
namespace {

template <typename T, typename OT>
__global__ void __launch_bounds__(256)
edge_features_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int pos_dims, int vel_dims,
                     const int64_t* __restrict__ edge_index, int64_t n_edges, int64_t n_points, EdgeFeatureSpec spec,
                     OT* __restrict__ out, int32_t* __restrict__ error_flag) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t i = edge_index[e], j = edge_index[n_edges + e];
  if (static_cast<uint64_t>(i) >= static_cast<uint64_t>(n_points) || static_cast<uint64_t>(j) >= static_cast<uint64_t>(n_points)) {
    atomicExch(error_flag, RGNN_ERR_INDEX_OUT_OF_RANGE);   // never used as an address
    return;
  }
  double xi[4], xj[4], vi[4], vj[4];
  efm::load_vec(pos, i, pos_dims, xi);
  efm::load_vec(pos, j, pos_dims, xj);
  efm::load_vec(vel, i, vel_dims, vi);
  efm::load_vec(vel, j, vel_dims, vj);
  efm::write_edge_row<OT>(xi, xj, vi, vj, pos_dims, vel_dims, spec, error_flag, out + e * spec.width, nullptr);
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
        row[col++] = static_cast<OT>(efm::norm_of(v, 2));
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
                         const int64_t* edge_index, int64_t n_edges, int64_t n_points, const EdgeFeatureSpec& spec,
                         void* edge_attr, int32_t out_dtype, int32_t* error_flag, cudaStream_t stream) {
  if (n_edges == 0 || spec.width == 0) return RGNN_OK;
  RGNN_PROFILE("edge_features", stream);
  const unsigned blocks = div_up(n_edges, 256);
#define RGNN_EF_LAUNCH(T, OT)                                                                          \
  edge_features_kernel<T, OT><<<blocks, 256, 0, stream>>>(static_cast<const T*>(pos),                  \
      static_cast<const T*>(vel), pos_dims, vel_dims, edge_index, n_edges, n_points, spec,             \
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
  RGNN_CUDA_CHECK(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream)));
  return launch_edge_features(pos, vel, in_dtype, pos_dims, vel_dims, edge_index, n_edges, n_points, spec, edge_attr,
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
