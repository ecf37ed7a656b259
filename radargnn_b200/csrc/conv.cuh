// conv.cuh -- internal interface of the graph-convolution layer (conv.cu), shared with pipeline.cu.
#pragma once

#include "common.cuh"
#include "node_gemm.cuh"

namespace rgnn {

// How a layer reads its input rows: raw x, or the previous layer's output with that layer's
// training-mode BatchNorm + ReLU applied on load (gnn_models.py:124-128 fused into the consumer).
struct ConvInput {
  const float* x = nullptr;  // [N, C]
  int64_t ldx = 0;
  const int32_t* rows = nullptr;  // optional gather: node r of the layer reads x[rows[r]]
  const float* mean = nullptr;   // [C] or null (no normalisation)
  const float* scale = nullptr;
  const float* beta = nullptr;
  int32_t relu = 0;
};

struct ConvShape {
  int32_t c, c_out, de, de_eff, p, pp;  // pp = p rounded up to a multiple of 4 (16-byte rows)
  bool general;                          // pre_layers > 1: per-edge MLP cannot be factored
  // Split message layout (tensor-core path only): the message rows B / M' are kept as a main array
  // [N, pm] whose rows are whole 128-byte lines (pm = 128: one float4 per lane of a warp) plus a narrow
  // tail array [N, pt4] for the remaining p - pm channels, so that the per-edge gathers are aligned
  // 512-byte row reads.  split == false: one [N, pp] array.
  bool split;
  int32_t pm, pt4;
};

struct ConvWorkspace {
  float* a;        // [N, pp]  x W_t^T  (MPNN only)
  float* b;        // [N, pp]  x W_s^T            (split layout: [N, pm])
  float* m;        // [N, pp]  aggregated messages (split layout: pm / 32 swizzled panels per 128-row tile)
  float* bt;       // [N, pt4] tail channels of B  (split layout only)
  float* mt;       // [N, pt4] tail channels of M
  float* t1;       // [N, c_out] post_mlp ping-pong (post_layers > 1)
  float* t2;
  float* w_eff;    // [p, de] folded edge-encoder weight
  float* b_eff;    // [p]
  float* ea_csc;   // [E, de] edge attributes in CSC slot order (when csc_eid is given)
  float* u1;       // [E, pp] per-edge activations (general path)
  float* u2;
  // tensor-core path (node_gemm.cu)
  bool tc_pre, tc_post;  // which contractions run on tcgen05
  float* wpack_pre;      // packed W_s image
  float* wpack_post;     // packed [W_x | W_m | W_fold] image
  float* w_fold;         // [c_out, c]
  double* bn_partial;    // [2][c_out][tc_tiles(N)] column sums / squares of the layer output
  int32_t* tc_status;    // device flag
};

// what a node without incoming edge contributes on the folded tensor-core path (null w_t: nothing)
struct IsolatedNodeTerm {
  const float* w_t = nullptr;  // W_pre[:, 0:C], row stride ldw
  int64_t ldw = 0;
  const float* x = nullptr;    // layer input rows
  int64_t ldx = 0;
  const int32_t* rows = nullptr;
  int32_t c = 0;
  const float* mean = nullptr; const float* scale = nullptr; const float* beta = nullptr;
  int32_t relu = 0;
};

// Fused edge aggregate + node update (fused_layer.cu): the segmented reduce over the CSC view writes the
// aggregated messages M' of a 128-row tile straight into shared memory, from where they feed the tcgen05
// update contraction -- M' never exists in HBM.  Split layout only (C = 64 MPNNConv, max / min / mean).
struct FusedLayerArgs {
  // aggregate side (what launch_edge_aggregate_split takes)
  int32_t aggr = 0;
  const float* bm = nullptr; const float* bt = nullptr;   // B main [N, 128], B tail [N, 4]
  float* mt = nullptr;                                    // scratch [N, 4]: tail channels of M' (reduced before the fused kernel)
  int32_t p = 0, de = 0;
  const float* bias_msg = nullptr; const float* w_e = nullptr; int64_t ldwe = 0;
  const float* ea = nullptr; const int32_t* csc_ptr = nullptr; const int32_t* csc_src = nullptr;
  IsolatedNodeTerm iso;
  // update side
  const float* x = nullptr; int64_t ldx = 0; const int32_t* x_rows = nullptr;   // layer input [N, 64] (optional gather map)
  const float* x_mean = nullptr; const float* x_scale = nullptr; const float* x_beta = nullptr; int32_t relu_x = 0;
  const float* wpack = nullptr;          // tc_pack_weights image of the update weights: K blocks [x | M' main | tail]
  const float* w_tail = nullptr; int64_t ld_wtail = 0;   // update weights of the tail channels: post_weight[:, C + 128 ...]
  const float* bias_post = nullptr;
  int32_t c_out = 0;
  float* y = nullptr; int64_t ldy = 0;
  double* bn_partial = nullptr;          // [2][c_out][partials] column sums / squares, one partial per CTA, or null
  int32_t* status = nullptr;
  int64_t n_nodes = 0;
};
bool fused_layer_supported(const rgnn_conv_desc& d, const ConvShape& s);
int64_t fused_layer_partials(int64_t n_nodes);   // partial sums per channel the kernel writes (= CTAs launched, <= SM count)
int launch_fused_layer(const FusedLayerArgs& a, cudaStream_t stream);

// shapes of the two node contractions of a layer when they run on the tensor cores
inline TcGemmShape conv_pre_shape(const ConvShape& s) { TcGemmShape t; t.k1 = s.c; t.n = s.p; return t; }
inline TcGemmShape conv_post_shape(const rgnn_conv_desc& d, const ConvShape& s) {
  TcGemmShape t; t.k1 = s.c; t.k2 = s.split ? s.pm : s.pp; t.kt = s.split ? s.pt4 : 0;
  t.k3 = (d.conv_type == RGNN_CONV_MPNN && d.aggr == RGNN_AGGR_ADD) ? s.c : 0;  // deg * x only for add
  t.n = s.c_out; return t;
}

int conv_shape(const rgnn_conv_desc& d, ConvShape* s);

template <typename ArenaT>
inline ConvWorkspace carve_conv_workspace(ArenaT& a, const rgnn_conv_desc& d, const ConvShape& s,
                                          int64_t n_nodes, int64_t n_edges, bool need_ea_gather) {
  ConvWorkspace w{};
  const size_t npp = static_cast<size_t>(n_nodes) * (s.split ? s.pm : s.pp);
  w.a = (d.conv_type == RGNN_CONV_MPNN && !s.split) ? a.template take<float>(npp) : nullptr;  // unused on the tensor-core path
  w.b = a.template take<float>(npp);
  // split layout: M' is panel-major (whole 128-row tiles), see node_gemm.cuh
  w.m = a.template take<float>(s.split ? tc_panel_major_floats(n_nodes, s.pm) : npp);
  if (s.split) {
    w.bt = a.template take<float>(static_cast<size_t>(n_nodes) * s.pt4);
    w.mt = a.template take<float>(static_cast<size_t>(n_nodes) * s.pt4);
  }
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
  // the factored path puts both node contractions on the tensor cores when their operands fit
  w.tc_pre = !s.general && tc_gemm_supported(conv_pre_shape(s));
  w.tc_post = w.tc_pre && tc_gemm_supported(conv_post_shape(d, s));
  w.tc_pre = w.tc_post;
  if (w.tc_post) {
    w.wpack_pre = a.template take<float>(tc_pack_floats(conv_pre_shape(s)));
    w.wpack_post = a.template take<float>(tc_pack_floats(conv_post_shape(d, s)));
    w.w_fold = a.template take<float>(static_cast<size_t>(s.c_out) * s.c);
    // one partial per 128-row tile (node_gemm.cu) or per CTA (fused_layer.cu, at most one per SM)
    w.bn_partial = a.template take<double>(static_cast<size_t>(tc_tiles(n_nodes) + sm_count()) * 2 * s.c_out);
    w.tc_status = a.template take<int32_t>(64);
  }
  return w;
}

// edge_attr rows are read through csc_eid when it is non-null (original edge order), else
// edge_attr must already be in CSC slot order.
int conv_forward(const rgnn_conv_desc& d, const ConvShape& s, const ConvInput& in, int64_t n_nodes,
                 const int32_t* csc_ptr, const int32_t* csc_src, const int32_t* csc_eid,
                 const float* edge_attr, int64_t n_edges, float* out, const ConvWorkspace& w,
                 cudaStream_t stream, int64_t* bn_partials = nullptr);

int gather_edge_rows(const float* edge_attr, const int32_t* csc_eid, int64_t n_edges, int32_t de,
                     float* ea_csc, cudaStream_t stream);

}  // namespace rgnn
