// edge_feature_math.cuh -- device-side arithmetic of the per-edge features (reference
// graph_constructor/graph.py:172-223, features.py:24-122), shared by the standalone feature kernel
// (features.cu) and the fused CSC-fill kernel (csc.cu).  All arithmetic in fp64.
#pragma once

#include <math.h>

#include "features.cuh"

namespace rgnn {
namespace efm {

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

__device__ __forceinline__ void point_pair_features(const double* p1, const double* p2, const double* v1, const double* v2,
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


// One edge (i -> j): all requested feature columns, written to `row` (and, when non-null, to `row2`).
template <typename OT>
__device__ __forceinline__ void write_edge_row(const double* xi, const double* xj, const double* vi, const double* vj,
                                               int pos_dims, int vel_dims, const EdgeFeatureSpec& spec,
                                               int32_t* error_flag, OT* __restrict__ row, OT* __restrict__ row2) {
  int col = 0;
  auto put = [&](double v) {
    row[col] = static_cast<OT>(v);
    if (row2 != nullptr) row2[col] = static_cast<OT>(v);
    ++col;
  };
  for (int f = 0; f < spec.n; ++f) {
    switch (spec.feature[f]) {
      case RGNN_EF_POINT_PAIR_FEATURES: {
        double ppf[4];
        point_pair_features(xi, xj, vi, vj, pos_dims, vel_dims, spec.edge_mode, error_flag, ppf);
        for (int c = 0; c < 4; ++c) put(ppf[c]);
        break;
      }
      case RGNN_EF_SPATIAL_EUCLIDEAN_DISTANCE: {
        double d[4];
        for (int c = 0; c < 4; ++c) d[c] = c < pos_dims ? xi[c] - xj[c] : 0.0;
        put(norm_of(d, pos_dims));
        break;
      }
      case RGNN_EF_VELOCITY_EUCLIDEAN_DISTANCE: {
        double d[4];
        for (int c = 0; c < 4; ++c) d[c] = c < vel_dims ? vi[c] - vj[c] : 0.0;
        put(norm_of(d, vel_dims));
        break;
      }
      case RGNN_EF_RELATIVE_POSITION: {  // graph.py:198-207: components 0 and 1 only
        double dx = xi[0] - xj[0], dy = xi[1] - xj[1];
        if (spec.edge_mode == RGNN_UNDIRECTED) { dx = fabs(dx); dy = fabs(dy); }
        put(dx); put(dy);
        break;
      }
      case RGNN_EF_RELATIVE_VELOCITY: {
        double du = vi[0] - vj[0], dv = vi[1] - vj[1];
        if (spec.edge_mode == RGNN_UNDIRECTED) { du = fabs(du); dv = fabs(dv); }
        put(du); put(dv);
        break;
      }
      default: break;
    }
  }
}

}  // namespace efm
}  // namespace rgnn
