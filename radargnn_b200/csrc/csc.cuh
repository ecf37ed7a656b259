// csc.cuh -- internal interface of the CSC build (csc.cu), shared with pipeline.cu.
#pragma once

#include "common.cuh"
#include "features.cuh"
#include "graph_build.cuh"

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
              const int32_t* node_map = nullptr, int32_t* error_flag = nullptr);

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

// Same result for the k-NN graph the pipeline has just built (edge rows i*k .. i*k+k-1 per query i), but the
// edges are visited per CELL-SORTED query: the targets of neighbouring threads are neighbours in the sorted
// node order, so the row-pointer / cursor / slot / position accesses are local instead of one lone 32-byte
// sector each; the source of every edge is the query's own sorted position (no lookup).  Requires
// counts_ready (the k-NN kernel's in-degree histogram) and f32 [N, 2] sorted positions (distance_dims == 2)
// or [N, 4] = [pos | vel] (distance_dims == 4) in graph.sorted_pts.
int csc_build_fused_knn(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int32_t k, int32_t dims,
                        const GraphWorkspace& graph, const CscWorkspace& w, int32_t* csc_ptr, int32_t* csc_src,
                        int32_t* csc_eid, cudaStream_t stream, const FusedEdgeAttr& fea);

}  // namespace rgnn
