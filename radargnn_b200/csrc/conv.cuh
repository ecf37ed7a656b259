// conv.cuh -- internal interface of the graph-convolution layer (conv.cu), shared with pipeline.cu.
#pragma once

#include "common.cuh"

namespace rgnn {

// How a layer reads its input rows: raw x, or the previous layer's output with that layer's
// training-mode BatchNorm + ReLU applied on load (gnn_models.py:124-128 fused into the consumer).
struct ConvInput {
  const float* x = nullptr;  // [N, C]
  int64_t ldx = 0;
  const float* mean = nullptr;   // [C] or null (no normalisation)
  const float* scale = nullptr;
  const float* beta = nullptr;
  int32_t relu = 0;
};

struct ConvShape {
  int32_t c, c_out, de, de_eff, p, pp;  // pp = p rounded up to a multiple of 4 (16-byte rows)
  bool general;                          // pre_layers > 1: per-edge MLP cannot be factored
};

struct ConvWorkspace {
  float* a;        // [N, pp]  x W_t^T  (MPNN only)
  float* b;        // [N, pp]  x W_s^T
  float* m;        // [N, pp]  aggregated messages
  float* t1;       // [N, c_out] post_mlp ping-pong (post_layers > 1)
  float* t2;
  float* w_eff;    // [p, de] folded edge-encoder weight
  float* b_eff;    // [p]
  float* ea_csc;   // [E, de] edge attributes in CSC slot order (when csc_eid is given)
  float* u1;       // [E, pp] per-edge activations (general path)
  float* u2;
};

int conv_shape(const rgnn_conv_desc& d, ConvShape* s);

template <typename ArenaT>
inline ConvWorkspace carve_conv_workspace(ArenaT& a, const rgnn_conv_desc& d, const ConvShape& s,
                                          int64_t n_nodes, int64_t n_edges, bool need_ea_gather) {
  ConvWorkspace w{};
  const size_t npp = static_cast<size_t>(n_nodes) * s.pp;
  w.a = d.conv_type == RGNN_CONV_MPNN ? a.template take<float>(npp) : nullptr;
  w.b = a.template take<float>(npp);
  w.m = a.template take<float>(npp);
  if (d.post_layers > 1) {
    w.t1 = a.template take<float>(static_cast<size_t>(n_nodes) * s.c_out);
    w.t2 = a.template take<float>(static_cast<size_t>(n_nodes) * s.c_out);
  }
  if (d.use_edge_encoder) {
    w.w_eff = a.template take<float>(static_cast<size_t>(s.p) * s.de);
    w.b_eff = a.template take<float>(s.p);
  }
  if (need_ea_gather) w.ea_csc = a.template take<float>(static_cast<size_t>(n_edges) * s.de);
  if (s.general) {
    w.u1 = a.template take<float>(static_cast<size_t>(n_edges) * s.pp);
    w.u2 = a.template take<float>(static_cast<size_t>(n_edges) * s.pp);
  }
  return w;
}

// edge_attr rows are read through csc_eid when it is non-null (original edge order), else
// edge_attr must already be in CSC slot order.
int conv_forward(const rgnn_conv_desc& d, const ConvShape& s, const ConvInput& in, int64_t n_nodes,
                 const int32_t* csc_ptr, const int32_t* csc_src, const int32_t* csc_eid,
                 const float* edge_attr, int64_t n_edges, float* out, const ConvWorkspace& w,
                 cudaStream_t stream);

int gather_edge_rows(const float* edge_attr, const int32_t* csc_eid, int64_t n_edges, int32_t de,
                     float* ea_csc, cudaStream_t stream);

}  // namespace rgnn
