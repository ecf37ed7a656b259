// csc.cuh -- internal interface of the CSC build (csc.cu), shared with pipeline.cu.
#pragma once

#include "common.cuh"
#include "features.cuh"

namespace rgnn {

struct CscWorkspace {
  int32_t* count;         // [N + 1] in-degree (may be pre-filled by the k-NN kernel)
  int32_t* cursor;        // [N + 1]
  int32_t* scan_scratch;  // scan_scratch_ints(N)
};

template <typename ArenaT>
inline CscWorkspace carve_csc_workspace(ArenaT& a, int64_t n_nodes) {
  CscWorkspace w{};
  w.count = a.template take<int32_t>(n_nodes + 1);
  w.cursor = a.template take<int32_t>(n_nodes + 1);
  w.scan_scratch = a.template take<int32_t>(scan_scratch_ints(n_nodes));
  return w;
}

// counts_ready: w.count already holds the in-degree histogram; ordered: sort every segment by
// edge id (needed for a deterministic sum / mean; max / min do not care).
// node_map (optional): targets and sources are renumbered through node_map[] (the conv stack runs in
// cell-sorted node order); csc_eid always refers to the caller's edge order.
int csc_build(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, bool counts_ready, bool ordered,
              const CscWorkspace& w, int32_t* csc_ptr, int32_t* csc_src, int32_t* csc_eid, cudaStream_t stream,
              const int32_t* node_map = nullptr);

// Fused variant for the pipeline (unordered segments only): the slot-fill pass also computes the edge
// attributes (fp64 arithmetic, f32 output) and writes them twice -- edge_attr [E, De] in the caller's
// edge order and ea_csc [E, De] in slot order -- saving the separate feature and gather passes.
struct FusedEdgeAttr {
  const float* pos = nullptr;   // [N, 2] f32
  const float* vel = nullptr;   // [N, 2] f32
  EdgeFeatureSpec spec{};
  float* edge_attr = nullptr;
  float* ea_csc = nullptr;
  int32_t* error_flag = nullptr;
};
int csc_build_fused(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, bool counts_ready,
                    const CscWorkspace& w, int32_t* csc_ptr, int32_t* csc_src, int32_t* csc_eid, cudaStream_t stream,
                    const int32_t* node_map, const FusedEdgeAttr& fea);

}  // namespace rgnn
