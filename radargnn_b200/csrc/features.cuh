// features.cuh -- internal interface of the feature kernels (features.cu), shared with pipeline.cu.
#pragma once

#include "common.cuh"

namespace rgnn {

struct EdgeFeatureSpec {
  int32_t n;
  int32_t feature[RGNN_MAX_EDGE_FEATURES];
  int32_t edge_mode;
  int32_t width;  // columns of edge_attr
};

struct NodeFeatureSpec {
  int32_t n;
  int32_t feature[RGNN_MAX_NODE_FEATURES];
  int32_t width;
};

int make_edge_feature_spec(const int32_t* features_host, int32_t n_features, int32_t edge_mode, EdgeFeatureSpec* spec);

int launch_edge_features(const void* pos, const void* vel, int32_t in_dtype, int32_t pos_dims, int32_t vel_dims,
                         const int64_t* edge_index, int64_t n_edges, int64_t n_points, const EdgeFeatureSpec& spec,
                         void* edge_attr, int32_t out_dtype, int32_t* error_flag, cudaStream_t stream);

}  // namespace rgnn
