// graph_build.cuh -- internal interface of the neighbour-search stage (cell list + k-NN /
// radius queries).  Shared between graph_build.cu (C ABI of the stage) and pipeline.cu.
#pragma once

#include "common.cuh"

namespace rgnn {

// Uniform cell list of ONE frame.  Cells are `h` wide, gx * gy of them, stored row-major
// (cell = cell_off + cy * gx + cx) in one global cell array shared by all frames.
struct FrameGrid {
  double x0, y0;     // lower-left corner (min of the frame's first two coordinates)
  double h, inv_h;   // cell edge, 1 / h
  int32_t gx, gy;
  int32_t cell_off;  // first global cell id of this frame
  int32_t active;    // 1 when the frame takes part in the search (n_f > 1)
  int64_t pt_begin, pt_end;  // global point range of the frame
  int64_t edge_off;          // k-NN: first edge row of the frame's first point
};

constexpr int kPointsPerCell = 4;  // target occupancy of a cell

// Device buffers of the stage, carved out of the caller's workspace.
struct GraphWorkspace {
  int64_t* frame_ptr;      // [F + 1] device copy
  int64_t* frame_edge_off; // [F + 1] k-NN edge offsets per frame
  int32_t* frame_cell_off; // [F + 1]
  long long* bbox;         // [F][4] ordered-int64 min x, min y, max x, max y
  FrameGrid* grids;        // [F]
  int32_t* point_cell;     // [N] cell of every point (original order)
  int32_t* cell_count;     // [cells + 1]
  int32_t* cell_start;     // [cells + 1] exclusive scan of cell_count
  int32_t* cell_cursor;    // [cells]
  int32_t* sorted_idx;     // [N] original point id at each sorted position
  int32_t* sorted_cell;    // [N]
  int32_t* sorted_frame;   // [N]
  int32_t* rank;           // [N] sorted position of every original point (inverse of sorted_idx)
  void* sorted_pts;        // [N, dims] of the basis dtype, cell-sorted
  int32_t* row_count;      // [N + 1] radius: neighbours per point (original order)
  int64_t* row_ptr;        // [N + 1] radius: exclusive scan
  int32_t* scan_scratch;   // scan_scratch_ints(max(N, cells)) * 2 (int64 variant)
  int32_t* status;         // [64] device flag of the stage: RGNN_ERR_NON_FINITE_INPUT when a coordinate is NaN / inf
  int32_t total_cells;
};

int64_t graph_total_cells_bound(int64_t n_points, int32_t n_frames);

template <typename ArenaT>
inline GraphWorkspace carve_graph_workspace(ArenaT& a, int64_t n, int32_t f) {
  GraphWorkspace w{};
  const int64_t cells = graph_total_cells_bound(n, f);
  w.total_cells = static_cast<int32_t>(cells);
  w.frame_ptr = a.template take<int64_t>(f + 1);
  w.frame_edge_off = a.template take<int64_t>(f + 1);
  w.frame_cell_off = a.template take<int32_t>(f + 1);
  w.bbox = a.template take<long long>(static_cast<size_t>(f) * 4);
  w.grids = a.template take<FrameGrid>(f);
  w.point_cell = a.template take<int32_t>(n);
  w.cell_count = a.template take<int32_t>(cells + 1);
  w.cell_cursor = a.template take<int32_t>(cells + 1);   // directly after cell_count: both are zeroed by one memset
  w.cell_start = a.template take<int32_t>(cells + 1);
  w.sorted_idx = a.template take<int32_t>(n);
  w.sorted_cell = a.template take<int32_t>(n);
  w.sorted_frame = a.template take<int32_t>(n);
  w.rank = a.template take<int32_t>(n);
  w.sorted_pts = a.template take<double>(static_cast<size_t>(n) * 4);
  w.row_count = a.template take<int32_t>(n + 1);
  w.row_ptr = a.template take<int64_t>(n + 1);
  const int64_t m = n > cells ? n : cells;
  w.scan_scratch = a.template take<int32_t>(scan_scratch_ints(m) * 2);
  w.status = a.template take<int32_t>(64);
  return w;
}

// Bins the points of all frames into their cell lists (bbox -> grid -> count -> scan -> scatter).
// `k` > 0 additionally fills frame_edge_off for the k-NN row layout.
// status (device int32, optional): set to RGNN_ERR_NON_FINITE_INPUT when some coordinate is NaN or infinite
// (sklearn raises "Input contains NaN"); zeroed first when zero_status is set.
int build_cell_lists(const void* basis, int32_t basis_dtype, int32_t dims,
                     const int64_t* frame_ptr_host, int32_t n_frames, int32_t k,
                     const GraphWorkspace& w, cudaStream_t stream, int32_t* status = nullptr,
                     bool zero_status = false);

// k-NN query over the cell lists.  Optional fused outputs (may be null): in-degree
// histogram of the targets (edge_index[1]) for the CSC build.
// degree_map (optional): the histogram is indexed by degree_map[target] (the conv stack runs in
// cell-sorted node order, degree_map = rank).
int knn_query(int32_t basis_dtype, int32_t dims, int64_t n_points, int32_t k,
              int64_t* edge_index, int64_t n_edges, int32_t* in_degree, const int32_t* degree_map,
              const GraphWorkspace& w, cudaStream_t stream);

}  // namespace rgnn
