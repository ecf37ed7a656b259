// graph_build.cu -- neighbour search on the GPU: uniform cell lists + exact k-NN / radius
// queries.  Replaces Graph.build (reference graph_constructor/graph.py:32-82), i.e.
// sklearn's KD-tree behind kneighbors_graph / radius_neighbors_graph.
//
// Exactness contract (hard parts 1-2): distances are the fp64 reduced
// distances sklearn computes -- sum over the dimensions, left to right, of (a - b) * (a - b)
// with no fused multiply-add -- so membership (radius: d2 <= r*r, inclusive) and order
// (k-NN: ascending (d2, index)) are bit-identical to the CPU oracle.  The cell list only
// prunes: a ring search stops when the k-th best distance is strictly inside the distance to
// the unsearched region (with a safety slack against cell-assignment rounding).
#include <math.h>

#include <vector>

#include <stdlib.h>

#include "graph_build.cuh"

namespace rgnn {

int64_t graph_total_cells_bound(int64_t n_points, int32_t n_frames) {
  return 2 * (n_points / kPointsPerCell) + 6 * static_cast<int64_t>(n_frames) + 8;
}

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int find_frame(const int64_t* __restrict__ frame_ptr, int n_frames, int64_t i) {
  int lo = 0, hi = n_frames;  // frame f holds [ptr[f], ptr[f+1])
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (frame_ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
frame_bbox_kernel(const T* __restrict__ basis, int dims, int64_t n, const int64_t* __restrict__ frame_ptr,
                  int n_frames, long long* __restrict__ bbox, int32_t* __restrict__ status) {
  // warp -> block reduction: a warp whose lanes lie in one frame reduces with shuffles, and when all warps
  // of the block lie in the same frame the block issues ONE set of atomics (a single large frame would
  // otherwise serialise thousands of warps on four addresses)
  __shared__ double red[kThreads / 32][4];
  __shared__ int red_frame[kThreads / 32];   // frame of a frame-uniform warp, -1 otherwise
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  int f = -1;
  double x = 0.0, y = 0.0;
  bool finite = true;
  if (valid) {
    f = n_frames == 1 ? 0 : find_frame(frame_ptr, n_frames, i);
    x = static_cast<double>(basis[i * dims]);
    y = static_cast<double>(basis[i * dims + 1]);
    finite = isfinite(x) && isfinite(y);
    for (int d = 2; d < dims; ++d) finite = finite && isfinite(static_cast<double>(basis[i * dims + d]));
    // sklearn's check_array rejects such input ("Input contains NaN" / "infinity"): flag it and keep the
    // point out of the bounding box so that the grid stays sane
    if (!finite && status != nullptr) atomicExch(status, RGNN_ERR_NON_FINITE_INPUT);
  }
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f0 = __shfl_sync(full, f, 0);
  // lanes beyond n (last warp only) take part as neutral elements of frame f0
  const bool uniform = f0 >= 0 && __all_sync(full, f == f0 || !valid);
  double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  if (uniform) {
    if (valid && finite) { mnx = x; mny = y; mxx = x; mxy = y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin(mnx, __shfl_xor_sync(full, mnx, o));
      mny = fmin(mny, __shfl_xor_sync(full, mny, o));
      mxx = fmax(mxx, __shfl_xor_sync(full, mxx, o));
      mxy = fmax(mxy, __shfl_xor_sync(full, mxy, o));
    }
  } else if (valid && finite) {
    atomicMin(&bbox[f * 4 + 0], double_to_ordered(x));
    atomicMin(&bbox[f * 4 + 1], double_to_ordered(y));
    atomicMax(&bbox[f * 4 + 2], double_to_ordered(x));
    atomicMax(&bbox[f * 4 + 3], double_to_ordered(y));
  }
  if (lane == 0) {
    red[wid][0] = mnx; red[wid][1] = mny; red[wid][2] = mxx; red[wid][3] = mxy;
    red_frame[wid] = uniform ? f0 : -1;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  constexpr int kWarps = kThreads / 32;
  int w = 0;
  while (w < kWarps) {   // runs of warps of the same frame -> one set of atomics per run
    const int fr = red_frame[w];
    if (fr < 0) { ++w; continue; }
    double a0 = red[w][0], a1 = red[w][1], a2 = red[w][2], a3 = red[w][3];
    int e = w + 1;
    for (; e < kWarps && red_frame[e] == fr; ++e) {
      a0 = fmin(a0, red[e][0]); a1 = fmin(a1, red[e][1]); a2 = fmax(a2, red[e][2]); a3 = fmax(a3, red[e][3]);
    }
    if (a0 <= a2) {   // at least one finite point
      atomicMin(&bbox[fr * 4 + 0], double_to_ordered(a0));
      atomicMin(&bbox[fr * 4 + 1], double_to_ordered(a1));
      atomicMax(&bbox[fr * 4 + 2], double_to_ordered(a2));
      atomicMax(&bbox[fr * 4 + 3], double_to_ordered(a3));
    }
    w = e;
  }
}

// Host -> device upload of the short per-frame tables through kernel parameters (unlike a pageable
// cudaMemcpyAsync this never synchronises the stream on the host side and can be captured into a CUDA
// graph), fused with the initialisation of the bounding boxes and of the stage's status flag.
constexpr int kTableChunk = 160;   // entries of each of the three tables per launch (3 * 160 * 8 = 3840 bytes of parameters)
struct FrameTableChunk { int64_t frame_ptr[kTableChunk], edge_off[kTableChunk], cell_off[kTableChunk]; };
__global__ void __launch_bounds__(256)
init_tables_kernel(const __grid_constant__ FrameTableChunk c, int first, int count, int64_t* __restrict__ frame_ptr,
                   int64_t* __restrict__ edge_off, int32_t* __restrict__ cell_off, long long* __restrict__ bbox,
                   int n_frames, int32_t* __restrict__ status) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    frame_ptr[first + i] = c.frame_ptr[i];
    edge_off[first + i] = c.edge_off[i];
    cell_off[first + i] = static_cast<int32_t>(c.cell_off[i]);
  }
  // bounding boxes of the frames this chunk covers (entry j of frame_ptr = start of frame j)
  for (int i = threadIdx.x; i < count * 4; i += blockDim.x) {
    const int fr = first + (i >> 2);
    if (fr < n_frames) bbox[fr * 4 + (i & 3)] = (i & 3) < 2 ? 0x7fffffffffffffffLL : (long long)0x8000000000000000ULL;
  }
  if (status != nullptr && first == 0 && threadIdx.x == 0) *status = 0;
}

// One thread per frame: choose the cell size so that a cell holds ~kPointsPerCell points.
__global__ void frame_grid_kernel(const long long* __restrict__ bbox, const int64_t* __restrict__ frame_ptr,
                                  const int64_t* __restrict__ frame_edge_off,
                                  const int32_t* __restrict__ frame_cell_off, int n_frames,
                                  FrameGrid* __restrict__ grids, double points_per_cell) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  FrameGrid g;
  g.pt_begin = frame_ptr[f];
  g.pt_end = frame_ptr[f + 1];
  g.edge_off = frame_edge_off[f];
  g.cell_off = frame_cell_off[f];
  const int64_t nf = g.pt_end - g.pt_begin;
  g.active = nf > 1 ? 1 : 0;
  const int cap = frame_cell_off[f + 1] - frame_cell_off[f];
  if (nf <= 0) {
    g.x0 = g.y0 = 0.0; g.h = 1.0; g.inv_h = 1.0; g.gx = g.gy = 1;
    grids[f] = g;
    return;
  }
  const double mnx = ordered_to_double(bbox[f * 4 + 0]), mny = ordered_to_double(bbox[f * 4 + 1]);
  const double mxx = ordered_to_double(bbox[f * 4 + 2]), mxy = ordered_to_double(bbox[f * 4 + 3]);
  const double w = mxx - mnx, hgt = mxy - mny;
  double target = static_cast<double>(nf) / points_per_cell;
  if (target < 1.0) target = 1.0;
  double h;
  if (!(w > 0.0) && !(hgt > 0.0)) h = 1.0;
  else if (!(hgt > 0.0)) h = w / target;
  else if (!(w > 0.0)) h = hgt / target;
  else h = sqrt(w * hgt / target);
  if (!(h > 0.0) || !isfinite(h)) h = 1.0;
  int gx, gy;
  for (int it = 0; it < 200; ++it) {
    const double fx = floor(w / h) + 1.0, fy = floor(hgt / h) + 1.0;
    if (fx * fy <= static_cast<double>(cap)) { gx = static_cast<int>(fx); gy = static_cast<int>(fy); break; }
    h *= 1.25;
    gx = gy = 1;
  }
  if (static_cast<int64_t>(gx) * gy > cap) { gx = gy = 1; h = fmax(w, hgt) * 2.0 + 1.0; }
  g.x0 = mnx; g.y0 = mny; g.h = h; g.inv_h = 1.0 / h; g.gx = gx; g.gy = gy;
  grids[f] = g;
}

__device__ __forceinline__ int cell_coord(double v, double origin, double inv_h, int n) {
  int c = static_cast<int>(floor((v - origin) * inv_h));
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
bin_count_kernel(const T* __restrict__ basis, int dims, int64_t n, const int64_t* __restrict__ frame_ptr,
                 int n_frames, const FrameGrid* __restrict__ grids, int32_t* __restrict__ point_cell,
                 int32_t* __restrict__ cell_count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = n_frames == 1 ? 0 : find_frame(frame_ptr, n_frames, i);
  const FrameGrid g = grids[f];
  const double x = static_cast<double>(basis[i * dims]), y = static_cast<double>(basis[i * dims + 1]);
  const int cx = cell_coord(x, g.x0, g.inv_h, g.gx), cy = cell_coord(y, g.y0, g.inv_h, g.gy);
  const int cell = g.cell_off + cy * g.gx + cx;
  point_cell[i] = cell;
  atomicAdd(&cell_count[cell], 1);
}

// counting-sort scatter: only the ids; the order inside a cell is made deterministic afterwards
__global__ void __launch_bounds__(kThreads)
bin_scatter_kernel(int64_t n, const int32_t* __restrict__ point_cell, const int32_t* __restrict__ cell_start,
                   int32_t* __restrict__ cell_cursor, int32_t* __restrict__ sorted_idx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = point_cell[i];
  sorted_idx[cell_start[cell] + atomicAdd(&cell_cursor[cell], 1)] = static_cast<int32_t>(i);
}

// one thread per cell: ascending point id inside the cell, so that the cell-sorted order (and with it
// every reduction order downstream) does not depend on the atomics' arrival order
__global__ void __launch_bounds__(128)
cell_sort_kernel(const int32_t* __restrict__ cell_start, int total_cells, int32_t* __restrict__ sorted_idx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= total_cells) return;
  int32_t* a = sorted_idx + cell_start[c];
  const int n = cell_start[c + 1] - cell_start[c];
  if (n < 2) return;
  if (n <= 8) {
    // the common case (~4 points per cell): registers + a fixed odd-even transposition network instead of a
    // chain of dependent global loads and stores
    int32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < n ? a[j] : 0x7fffffff;
#pragma unroll
    for (int round = 0; round < 8; ++round) {
#pragma unroll
      for (int j = round & 1; j + 1 < 8; j += 2) {
        const int32_t lo = min(v[j], v[j + 1]), hi = max(v[j], v[j + 1]);
        v[j] = lo; v[j + 1] = hi;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < n) a[j] = v[j];
    return;
  }
  if (n <= 32) {
    for (int j = 1; j < n; ++j) {
      const int32_t v = a[j];
      int m = j - 1;
      while (m >= 0 && a[m] > v) { a[m + 1] = a[m]; --m; }
      a[m + 1] = v;
    }
    return;
  }
  auto sift = [&](int start, int end) {
    int root = start;
    while (2 * root + 1 <= end) {
      int child = 2 * root + 1;
      if (child + 1 <= end && a[child] < a[child + 1]) ++child;
      if (a[root] < a[child]) { const int32_t t = a[root]; a[root] = a[child]; a[child] = t; root = child; }
      else return;
    }
  };
  for (int s = (n - 2) / 2; s >= 0; --s) sift(s, n - 1);
  for (int end = n - 1; end > 0; --end) {
    const int32_t t = a[0]; a[0] = a[end]; a[end] = t;
    sift(0, end - 1);
  }
}

template <typename T, int DIMS>
__global__ void __launch_bounds__(kThreads)
gather_sorted_kernel(const T* __restrict__ basis, int64_t n, const int64_t* __restrict__ frame_ptr, int n_frames,
                     const int32_t* __restrict__ point_cell, const int32_t* __restrict__ sorted_idx,
                     int32_t* __restrict__ sorted_cell, int32_t* __restrict__ sorted_frame,
                     int32_t* __restrict__ rank, T* __restrict__ sorted_pts) {
  const int64_t pos = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int64_t i = sorted_idx[pos];
  rank[i] = static_cast<int32_t>(pos);
  sorted_cell[pos] = point_cell[i];
  sorted_frame[pos] = n_frames == 1 ? 0 : find_frame(frame_ptr, n_frames, i);
#pragma unroll
  for (int d = 0; d < DIMS; ++d) sorted_pts[pos * DIMS + d] = basis[i * DIMS + d];
}

// ---- exact reduced distance ---------------------------------------------------------
template <typename T, int DIMS>
struct Point {
  double v[DIMS];
  __device__ __forceinline__ void load(const T* __restrict__ p) {
    if constexpr (sizeof(T) == 4 && DIMS == 2) {
      const float2 t = *reinterpret_cast<const float2*>(p);
      v[0] = t.x; v[1] = t.y;
    } else if constexpr (sizeof(T) == 4 && DIMS == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (sizeof(T) == 8 && DIMS == 2) {
      const double2 t = *reinterpret_cast<const double2*>(p);
      v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
      for (int d = 0; d < DIMS; ++d) v[d] = static_cast<double>(p[d]);
    }
  }
};

template <int DIMS>
__device__ __forceinline__ double rdist(const double* a, const double* b) {
  // sklearn euclidean_rdist: d = 0; for j: tmp = a[j] - b[j]; d += tmp * tmp   (no FMA)
  double acc = 0.0;
#pragma unroll
  for (int d = 0; d < DIMS; ++d) {
    const double t = __dsub_rn(a[d], b[d]);
    acc = __dadd_rn(acc, __dmul_rn(t, t));
  }
  return acc;
}

__device__ __forceinline__ bool cand_less(double d, int id, double wd, int wid) {
  return d < wd || (d == wd && id < wid);
}

// ---- k-NN query: one thread per point, in cell-sorted order ------------------------------
// OCC: resident 128-thread CTAs per SM the kernel is compiled for (register cap); the search is bound by
// instruction issue at few warps per scheduler, so for k <= 16 the cap is lowered to 80 registers.
template <typename T, int DIMS, int KMAX, int OCC>
__global__ void __launch_bounds__(128, OCC)
knn_query_kernel(const T* __restrict__ sorted_pts, const int32_t* __restrict__ sorted_idx,
                 const int32_t* __restrict__ sorted_cell, const int32_t* __restrict__ sorted_frame,
                 const int32_t* __restrict__ cell_start, const FrameGrid* __restrict__ grids,
                 int64_t n_points, int k, int64_t* __restrict__ edge_index, int64_t n_edges,
                 int32_t* __restrict__ in_degree, const int32_t* __restrict__ degree_map) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_points) return;
  const FrameGrid g = grids[sorted_frame[q]];
  if (!g.active) return;
  Point<T, DIMS> me;
  me.load(sorted_pts + q * DIMS);
  const int c = sorted_cell[q] - g.cell_off;
  const int cy = c / g.gx, cx = c - cy * g.gx;

  // best-k list kept in DESCENDING order: slot 0 is always the current k-th (worst) entry, so
  // the pruning bound needs no dynamic register index.  Slots >= k hold a -inf sentinel that
  // no candidate can pass.
  double bd[KMAX];
  int bi[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    bd[j] = j < k ? INFINITY : -INFINITY;
    bi[j] = j < k ? 0x7fffffff : -1;
  }

  const double slack = 1e-7 * g.h + 1e-14 * (fabs(me.v[0]) + fabs(me.v[1]) + fabs(g.x0) + fabs(g.y0));
  const int max_ring = g.gx > g.gy ? g.gx : g.gy;
  for (int r = 0; r <= max_ring; ++r) {
    const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r;
    // ring r = top row, bottom row, then the left / right cell of every row in between:
    // one loop over the 4r spans so that the scan body exists once (keeps bd / bi in registers)
    const int n_spans = r == 0 ? 1 : 4 * r;
    for (int s = 0; s < n_spans; ++s) {
      int row, xa, xb;
      if (s < 2) {
        row = s == 0 ? y0 : y1;
        xa = x0 < 0 ? 0 : x0;
        xb = x1 >= g.gx ? g.gx - 1 : x1;
        if (r == 0) { xa = cx; xb = cx; }
      } else {
        row = y0 + 1 + ((s - 2) >> 1);
        xa = xb = ((s - 2) & 1) ? x1 : x0;
        if (xa < 0 || xa >= g.gx) continue;
      }
      if (row < 0 || row >= g.gy) continue;
      const int base = g.cell_off + row * g.gx;
      const int p0 = cell_start[base + xa], p1 = cell_start[base + xb + 1];
      for (int p = p0; p < p1; ++p) {
        if (p == q) continue;  // self is excluded by index, never by distance
        Point<T, DIMS> o;
        o.load(sorted_pts + static_cast<int64_t>(p) * DIMS);
        const double d = rdist<DIMS>(me.v, o.v);
        if (d <= bd[0]) {
          const int id = sorted_idx[p];
          if (cand_less(d, id, bd[0], bi[0])) {
            bool prev = true;  // candidate beats slot j
#pragma unroll
            for (int j = 0; j < KMAX - 1; ++j) {
              const bool sh = cand_less(d, id, bd[j + 1], bi[j + 1]);  // slot j+1 moves down to j
              bd[j] = sh ? bd[j + 1] : (prev ? d : bd[j]);
              bi[j] = sh ? bi[j + 1] : (prev ? id : bi[j]);
              prev = sh;
            }
            if (prev) { bd[KMAX - 1] = d; bi[KMAX - 1] = id; }
          }
        }
      }
    }
    // distance from the query to the nearest unsearched region (nothing lies beyond the grid)
    double gap = INFINITY;
    if (x0 > 0) gap = fmin(gap, me.v[0] - (g.x0 + x0 * g.h));
    if (x1 < g.gx - 1) gap = fmin(gap, (g.x0 + (x1 + 1) * g.h) - me.v[0]);
    if (y0 > 0) gap = fmin(gap, me.v[1] - (g.y0 + y0 * g.h));
    if (y1 < g.gy - 1) gap = fmin(gap, (g.y0 + (y1 + 1) * g.h) - me.v[1]);
    if (gap == INFINITY) break;  // whole frame searched
    gap -= slack;
    if (gap > 0.0 && bd[0] < gap * gap) break;
  }

  const int i = sorted_idx[q];
  const int64_t e0 = g.edge_off + (static_cast<int64_t>(i) - g.pt_begin) * k;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < k) {  // slot j is the (k-1-j)-th nearest
      edge_index[e0 + (k - 1 - j)] = i;
      edge_index[n_edges + e0 + (k - 1 - j)] = bi[j];
      // a slot that never received a candidate (non-finite coordinates: the stage's status flag is set)
      // still holds its sentinel id and must not be used as an index
      if (in_degree != nullptr && static_cast<unsigned>(bi[j]) < static_cast<unsigned>(n_points))
        atomicAdd(&in_degree[degree_map != nullptr ? degree_map[bi[j]] : bi[j]], 1);
    }
  }
}

// ---- k-NN query with fp32 list keys -------------------------------------------------------------------
// Same search, same results (ascending (fp64 squared distance, index), bit-exact against sklearn), but the
// best-k list holds (float key, sorted position) instead of (double distance, point id): 2 registers per slot
// instead of 3, and the 16-step insertion cascade -- what the kernel spends its time in, it runs for the whole
// warp whenever ANY lane accepts a candidate -- compares with one FSETP per slot instead of two DSETP + ISETP +
// predicate logic.  Rounding to fp32 is monotone, so key_a < key_b implies d_a < d_b; only candidates whose
// keys are EQUAL need the exact comparison, and for those the fp64 distance of the list entry is recomputed
// from its coordinates (rare: a handful per million insertions on continuous data, every time on duplicates).
template <typename T, int DIMS, int KMAX, int OCC>
__global__ void __launch_bounds__(128, OCC)
knn_query_f32key_kernel(const T* __restrict__ sorted_pts, const int32_t* __restrict__ sorted_idx,
                        const int32_t* __restrict__ sorted_cell, const int32_t* __restrict__ sorted_frame,
                        const int32_t* __restrict__ cell_start, const FrameGrid* __restrict__ grids,
                        int64_t n_points, int k, int64_t* __restrict__ edge_index, int64_t n_edges,
                        int32_t* __restrict__ in_degree, const int32_t* __restrict__ degree_map) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_points) return;
  const FrameGrid g = grids[sorted_frame[q]];
  if (!g.active) return;
  Point<T, DIMS> me;
  me.load(sorted_pts + q * DIMS);
  const int c = sorted_cell[q] - g.cell_off;
  const int cy = c / g.gx, cx = c - cy * g.gx;

  // DESCENDING list: slot 0 = current k-th (worst).  Empty slots: key +inf, position INT_MAX; slots >= k:
  // key -inf (no candidate passes).
  constexpr int kEmpty = 0x7fffffff;
  float bf[KMAX];
  int bp[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    bf[j] = j < k ? INFINITY : -INFINITY;
    bp[j] = j < k ? kEmpty : -1;
  }
  // exact order of candidate (d, id) against the list entry at sorted position pj (equal fp32 keys only)
  auto exact_beats = [&](double d, int id, int pj) -> bool {
    if (pj == kEmpty) return true;
    if (pj < 0) return false;
    Point<T, DIMS> o;
    o.load(sorted_pts + static_cast<int64_t>(pj) * DIMS);
    return cand_less(d, id, rdist<DIMS>(me.v, o.v), sorted_idx[pj]);
  };

  const double slack = 1e-7 * g.h + 1e-14 * (fabs(me.v[0]) + fabs(me.v[1]) + fabs(g.x0) + fabs(g.y0));
  const int max_ring = g.gx > g.gy ? g.gx : g.gy;
  for (int r = 0; r <= max_ring; ++r) {
    const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r;
    const int n_spans = r == 0 ? 1 : 4 * r;
    for (int s = 0; s < n_spans; ++s) {
      int row, xa, xb;
      if (s < 2) {
        row = s == 0 ? y0 : y1;
        xa = x0 < 0 ? 0 : x0;
        xb = x1 >= g.gx ? g.gx - 1 : x1;
        if (r == 0) { xa = cx; xb = cx; }
      } else {
        row = y0 + 1 + ((s - 2) >> 1);
        xa = xb = ((s - 2) & 1) ? x1 : x0;
        if (xa < 0 || xa >= g.gx) continue;
      }
      if (row < 0 || row >= g.gy) continue;
      const int base = g.cell_off + row * g.gx;
      const int p0 = cell_start[base + xa], p1 = cell_start[base + xb + 1];
      for (int p = p0; p < p1; ++p) {
        if (p == q) continue;  // self is excluded by index, never by distance
        Point<T, DIMS> o;
        o.load(sorted_pts + static_cast<int64_t>(p) * DIMS);
        const double d = rdist<DIMS>(me.v, o.v);
        const float df = __double2float_rn(d);
        if (df <= bf[0]) {
          bool take = df < bf[0];
          if (!take) take = exact_beats(d, sorted_idx[p], bp[0]);   // equal keys with the k-th: exact order
          if (take) {
            bool prev = true, eq = false;   // prev: the candidate beats slot j; eq: some list key equals the candidate's
#pragma unroll
            for (int j = 0; j < KMAX - 1; ++j) {
              const bool sh = df < bf[j + 1];  // slot j+1 moves down to j
              eq = eq || (df == bf[j + 1]);
              bf[j] = sh ? bf[j + 1] : (prev ? df : bf[j]);
              bp[j] = sh ? bp[j + 1] : (prev ? p : bp[j]);
              prev = sh;
            }
            if (prev) { bf[KMAX - 1] = df; bp[KMAX - 1] = p; }
            if (eq) {
              // the candidate sits below every entry with the same key: bubble it up past those it beats
              const int id = sorted_idx[p];
#pragma unroll
              for (int j = 0; j < KMAX - 1; ++j) {
                if (bp[j] == p && bf[j + 1] == df && exact_beats(d, id, bp[j + 1])) {
                  bp[j] = bp[j + 1]; bp[j + 1] = p;   // keys are equal: only the positions swap
                }
              }
            }
          }
        }
      }
    }
    double gap = INFINITY;
    if (x0 > 0) gap = fmin(gap, me.v[0] - (g.x0 + x0 * g.h));
    if (x1 < g.gx - 1) gap = fmin(gap, (g.x0 + (x1 + 1) * g.h) - me.v[0]);
    if (y0 > 0) gap = fmin(gap, me.v[1] - (g.y0 + y0 * g.h));
    if (y1 < g.gy - 1) gap = fmin(gap, (g.y0 + (y1 + 1) * g.h) - me.v[1]);
    if (gap == INFINITY) break;  // whole frame searched
    gap -= slack;
    // upper bound of the exact k-th distance: one fp32 ulp above its key
    const double kth_ub = bf[0] < INFINITY ? static_cast<double>(__uint_as_float(__float_as_uint(bf[0]) + 1u)) : INFINITY;
    if (gap > 0.0 && kth_ub < gap * gap) break;
  }

  const int i = sorted_idx[q];
  const int64_t e0 = g.edge_off + (static_cast<int64_t>(i) - g.pt_begin) * k;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < k) {  // slot j is the (k-1-j)-th nearest
      const bool filled = static_cast<unsigned>(bp[j]) < static_cast<unsigned>(n_points);
      const int nb = filled ? sorted_idx[bp[j]] : 0x7fffffff;   // a slot that never received a candidate keeps the sentinel
      edge_index[e0 + (k - 1 - j)] = i;
      edge_index[n_edges + e0 + (k - 1 - j)] = nb;
      if (in_degree != nullptr && filled) atomicAdd(&in_degree[degree_map != nullptr ? degree_map[nb] : nb], 1);
    }
  }
}

// ---- radius query: count, then fill (rows ascending in j after sort_rows_kernel) -------------
template <typename T, int DIMS, bool FILL>
__global__ void __launch_bounds__(128)
radius_query_kernel(const T* __restrict__ sorted_pts, const int32_t* __restrict__ sorted_idx,
                    const int32_t* __restrict__ sorted_cell, const int32_t* __restrict__ sorted_frame,
                    const int32_t* __restrict__ cell_start, const FrameGrid* __restrict__ grids,
                    int64_t n_points, double r, double r2, int32_t* __restrict__ row_count,
                    const int64_t* __restrict__ row_ptr, int64_t* __restrict__ edge_index, int64_t n_edges) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_points) return;
  const int i = sorted_idx[q];
  const FrameGrid g = grids[sorted_frame[q]];
  if (!g.active) { if (!FILL) row_count[i] = 0; return; }
  Point<T, DIMS> me;
  me.load(sorted_pts + q * DIMS);
  const double slack = 1e-7 * g.h + 1e-14 * (fabs(me.v[0]) + fabs(me.v[1]) + fabs(g.x0) + fabs(g.y0));
  const double reach = r + slack;
  const int xa = cell_coord(me.v[0] - reach, g.x0, g.inv_h, g.gx), xb = cell_coord(me.v[0] + reach, g.x0, g.inv_h, g.gx);
  const int ya = cell_coord(me.v[1] - reach, g.y0, g.inv_h, g.gy), yb = cell_coord(me.v[1] + reach, g.y0, g.inv_h, g.gy);
  int count = 0;
  int64_t out = FILL ? row_ptr[i] : 0;
  for (int yy = ya; yy <= yb; ++yy) {
    const int base = g.cell_off + yy * g.gx;
    const int p0 = cell_start[base + xa], p1 = cell_start[base + xb + 1];
    for (int p = p0; p < p1; ++p) {
      if (p == q) continue;
      Point<T, DIMS> o;
      o.load(sorted_pts + static_cast<int64_t>(p) * DIMS);
      if (rdist<DIMS>(me.v, o.v) <= r2) {  // inclusive, like sklearn
        if (FILL) {
          edge_index[out] = i;
          edge_index[n_edges + out] = sorted_idx[p];
          ++out;
        }
        ++count;
      }
    }
  }
  if (!FILL) row_count[i] = count;
}

// one thread per row: in-place heapsort of the neighbour ids (rows are short)
__global__ void __launch_bounds__(128)
sort_rows_kernel(const int64_t* __restrict__ row_ptr, int64_t n_points, int64_t* __restrict__ cols) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  int64_t* a = cols + row_ptr[i];
  const int n = static_cast<int>(row_ptr[i + 1] - row_ptr[i]);
  if (n < 2) return;
  if (n <= 24) {  // insertion sort
    for (int j = 1; j < n; ++j) {
      const int64_t v = a[j];
      int m = j - 1;
      while (m >= 0 && a[m] > v) { a[m + 1] = a[m]; --m; }
      a[m + 1] = v;
    }
    return;
  }
  auto sift = [&](int start, int end) {
    int root = start;
    while (2 * root + 1 <= end) {
      int child = 2 * root + 1;
      if (child + 1 <= end && a[child] < a[child + 1]) ++child;
      if (a[root] < a[child]) { const int64_t t = a[root]; a[root] = a[child]; a[child] = t; root = child; }
      else return;
    }
  };
  for (int s = (n - 2) / 2; s >= 0; --s) sift(s, n - 1);
  for (int end = n - 1; end > 0; --end) {
    const int64_t t = a[0]; a[0] = a[end]; a[end] = t;
    sift(0, end - 1);
  }
}

int host_frame_tables(const int64_t* frame_ptr_host, int32_t n_frames, int32_t k,
                      int64_t* edge_off, int32_t* cell_off) {
  int64_t e = 0;
  int64_t c = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int64_t nf = frame_ptr_host[f + 1] - frame_ptr_host[f];
    if (nf < 0) return RGNN_ERR_INVALID_ARGUMENT;
    edge_off[f] = e;
    cell_off[f] = static_cast<int32_t>(c);
    if (nf > 1 && k > 0) e += nf * k;
    int64_t target = nf / kPointsPerCell;
    if (target < 1) target = 1;
    c += 2 * target + 4;
  }
  edge_off[n_frames] = e;
  cell_off[n_frames] = static_cast<int32_t>(c);
  return RGNN_OK;
}

template <typename T>
int build_cell_lists_t(const T* basis, int32_t dims, const int64_t* frame_ptr_host, int32_t n_frames, int32_t k,
                       const GraphWorkspace& w, cudaStream_t stream, int32_t* status, bool zero_status) {
  const int64_t n = frame_ptr_host[n_frames] - frame_ptr_host[0];
  RGNN_PROFILE("cell_lists", stream);
  // small per-frame tables: computed on the host, uploaded without a host-side stream sync
  {
    std::vector<int64_t> edge_off(n_frames + 1);
    std::vector<int32_t> cell_off(n_frames + 1);
    RGNN_RETURN_IF_ERROR(host_frame_tables(frame_ptr_host, n_frames, k, edge_off.data(), cell_off.data()));
    if (cell_off[n_frames] > w.total_cells) return RGNN_ERR_WORKSPACE_TOO_SMALL;
    for (int first = 0; first <= n_frames; first += kTableChunk) {
      FrameTableChunk c;
      const int m = n_frames + 1 - first < kTableChunk ? n_frames + 1 - first : kTableChunk;
      for (int i = 0; i < m; ++i) {
        c.frame_ptr[i] = frame_ptr_host[first + i];
        c.edge_off[i] = edge_off[first + i];
        c.cell_off[i] = cell_off[first + i];
      }
      init_tables_kernel<<<1, 256, 0, stream>>>(c, first, m, w.frame_ptr, w.frame_edge_off, w.frame_cell_off, w.bbox,
                                                n_frames, zero_status ? status : nullptr);
      RGNN_LAUNCH_CHECK();
    }
  }
  if (n == 0) return RGNN_OK;
  // cell_count and cell_cursor are adjacent in the workspace: one memset
  RGNN_CUDA_CHECK(cudaMemsetAsync(w.cell_count, 0, reinterpret_cast<char*>(w.cell_cursor + w.total_cells + 1) -
                                                       reinterpret_cast<char*>(w.cell_count), stream));
  const unsigned blocks = div_up(n, kThreads);
  frame_bbox_kernel<T><<<blocks, kThreads, 0, stream>>>(basis, dims, n, w.frame_ptr, n_frames, w.bbox, status);
  RGNN_LAUNCH_CHECK();
  static double points_per_cell = 0.0;   // RGNN_CELL_POINTS: experiments on the grid's target occupancy (>= 2: the cell array holds 2 N / 4 cells)
  if (points_per_cell == 0.0) {
    const char* e = getenv("RGNN_CELL_POINTS");
    points_per_cell = e != nullptr ? atof(e) : static_cast<double>(kPointsPerCell);
    if (!(points_per_cell >= 2.0)) points_per_cell = kPointsPerCell;
  }
  frame_grid_kernel<<<div_up(n_frames, 64), 64, 0, stream>>>(w.bbox, w.frame_ptr, w.frame_edge_off,
                                                             w.frame_cell_off, n_frames, w.grids, points_per_cell);
  RGNN_LAUNCH_CHECK();
  bin_count_kernel<T><<<blocks, kThreads, 0, stream>>>(basis, dims, n, w.frame_ptr, n_frames, w.grids,
                                                       w.point_cell, w.cell_count);
  RGNN_LAUNCH_CHECK();
  RGNN_RETURN_IF_ERROR(exclusive_scan_i32(w.cell_count, w.cell_start, w.total_cells, w.scan_scratch, stream));
  bin_scatter_kernel<<<blocks, kThreads, 0, stream>>>(n, w.point_cell, w.cell_start, w.cell_cursor, w.sorted_idx);
  RGNN_LAUNCH_CHECK();
  cell_sort_kernel<<<div_up(w.total_cells, 128), 128, 0, stream>>>(w.cell_start, w.total_cells, w.sorted_idx);
  RGNN_LAUNCH_CHECK();
  if (dims == 2) {
    gather_sorted_kernel<T, 2><<<blocks, kThreads, 0, stream>>>(basis, n, w.frame_ptr, n_frames, w.point_cell,
        w.sorted_idx, w.sorted_cell, w.sorted_frame, w.rank, static_cast<T*>(w.sorted_pts));
  } else {
    gather_sorted_kernel<T, 4><<<blocks, kThreads, 0, stream>>>(basis, n, w.frame_ptr, n_frames, w.point_cell,
        w.sorted_idx, w.sorted_cell, w.sorted_frame, w.rank, static_cast<T*>(w.sorted_pts));
  }
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

template <typename T, int DIMS>
int knn_query_t(int64_t n, int32_t k, int64_t* edge_index, int64_t n_edges, int32_t* in_degree,
                const int32_t* degree_map, const GraphWorkspace& w, cudaStream_t stream) {
  const unsigned blocks = div_up(n, 128);
  const T* pts = static_cast<const T*>(w.sorted_pts);
  RGNN_PROFILE("knn_query", stream);
#define RGNN_KNN_LAUNCH(KMAX, OCC)                                                                    \
  knn_query_kernel<T, DIMS, KMAX, OCC><<<blocks, 128, 0, stream>>>(pts, w.sorted_idx, w.sorted_cell,  \
      w.sorted_frame, w.cell_start, w.grids, n, k, edge_index, n_edges, in_degree, degree_map)
  static int occ16 = 0;   // RGNN_KNN_OCC = 4 | 6 | 8: experiments on the k <= 16 variant
  if (occ16 == 0) { const char* e = getenv("RGNN_KNN_OCC"); occ16 = e != nullptr ? atoi(e) : -1; }
  static int keys = 0;    // RGNN_KNN_KEYS = 64: the fp64-key kernel (experiments / cross-check); default fp32 keys
  if (keys == 0) { const char* e = getenv("RGNN_KNN_KEYS"); keys = (e != nullptr && atoi(e) == 64) ? 64 : 32; }
  if (keys == 32) {
#define RGNN_KNN32_LAUNCH(KMAX, OCC)                                                                         \
  knn_query_f32key_kernel<T, DIMS, KMAX, OCC><<<blocks, 128, 0, stream>>>(pts, w.sorted_idx, w.sorted_cell,  \
      w.sorted_frame, w.cell_start, w.grids, n, k, edge_index, n_edges, in_degree, degree_map)
    if (k <= 4) RGNN_KNN32_LAUNCH(4, 8);
    else if (k <= 8) RGNN_KNN32_LAUNCH(8, 8);
    else if (k <= 16) { if (occ16 == 4) RGNN_KNN32_LAUNCH(16, 4); else if (occ16 == 8) RGNN_KNN32_LAUNCH(16, 8); else RGNN_KNN32_LAUNCH(16, 6); }   // 6: 78 registers, no spills (8: 64 + spills, measured equal)
    else if (k <= 24) RGNN_KNN32_LAUNCH(24, 5);
    else if (k <= 32) RGNN_KNN32_LAUNCH(32, 4);
    else RGNN_KNN32_LAUNCH(64, 2);
#undef RGNN_KNN32_LAUNCH
    RGNN_LAUNCH_CHECK();
    return RGNN_OK;
  }
  if (k <= 4) RGNN_KNN_LAUNCH(4, 8);
  else if (k <= 8) RGNN_KNN_LAUNCH(8, 8);
  else if (k <= 16) { if (occ16 == 8) RGNN_KNN_LAUNCH(16, 8); else if (occ16 == 4) RGNN_KNN_LAUNCH(16, 4); else RGNN_KNN_LAUNCH(16, 6); }
  else if (k <= 24) RGNN_KNN_LAUNCH(24, 4);
  else if (k <= 32) RGNN_KNN_LAUNCH(32, 3);
  else RGNN_KNN_LAUNCH(64, 2);
#undef RGNN_KNN_LAUNCH
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

template <typename T, int DIMS>
int radius_query_t(bool fill, int64_t n, double r, int64_t* edge_index, int64_t n_edges,
                   const GraphWorkspace& w, cudaStream_t stream) {
  const unsigned blocks = div_up(n, 128);
  const T* pts = static_cast<const T*>(w.sorted_pts);
  const double r2 = r * r;  // sklearn: reduced radius = r * r in fp64
  RGNN_PROFILE("radius_query", stream);
  if (!fill) {
    radius_query_kernel<T, DIMS, false><<<blocks, 128, 0, stream>>>(pts, w.sorted_idx, w.sorted_cell,
        w.sorted_frame, w.cell_start, w.grids, n, r, r2, w.row_count, w.row_ptr, edge_index, n_edges);
  } else {
    radius_query_kernel<T, DIMS, true><<<blocks, 128, 0, stream>>>(pts, w.sorted_idx, w.sorted_cell,
        w.sorted_frame, w.cell_start, w.grids, n, r, r2, w.row_count, w.row_ptr, edge_index, n_edges);
  }
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int check_graph_args(const void* basis, int32_t basis_dtype, int32_t dims, const int64_t* frame_ptr_host,
                     int32_t n_frames, void* workspace, size_t workspace_bytes, int64_t* n_out) {
  if (frame_ptr_host == nullptr || n_frames < 1) return RGNN_ERR_INVALID_ARGUMENT;
  if (dims != 2 && dims != 4) return RGNN_ERR_INVALID_ARGUMENT;
  if (basis_dtype != RGNN_F32 && basis_dtype != RGNN_F64) return RGNN_ERR_INVALID_ARGUMENT;
  if (frame_ptr_host[0] != 0) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t n = frame_ptr_host[n_frames];
  if (n < 0 || n > 0x7ffffff0LL) return RGNN_ERR_INVALID_ARGUMENT;
  if (n > 0 && basis == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_graph_workspace_bytes(n, n_frames)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  *n_out = n;
  return RGNN_OK;
}

}  // namespace

int build_cell_lists(const void* basis, int32_t basis_dtype, int32_t dims, const int64_t* frame_ptr_host,
                     int32_t n_frames, int32_t k, const GraphWorkspace& w, cudaStream_t stream, int32_t* status,
                     bool zero_status) {
  if (basis_dtype == RGNN_F32)
    return build_cell_lists_t<float>(static_cast<const float*>(basis), dims, frame_ptr_host, n_frames, k, w, stream,
                                     status, zero_status);
  return build_cell_lists_t<double>(static_cast<const double*>(basis), dims, frame_ptr_host, n_frames, k, w, stream,
                                    status, zero_status);
}

int knn_query(int32_t basis_dtype, int32_t dims, int64_t n_points, int32_t k, int64_t* edge_index,
              int64_t n_edges, int32_t* in_degree, const int32_t* degree_map, const GraphWorkspace& w,
              cudaStream_t stream) {
  if (n_points == 0 || n_edges == 0) return RGNN_OK;
  if (basis_dtype == RGNN_F32) {
    if (dims == 2) return knn_query_t<float, 2>(n_points, k, edge_index, n_edges, in_degree, degree_map, w, stream);
    return knn_query_t<float, 4>(n_points, k, edge_index, n_edges, in_degree, degree_map, w, stream);
  }
  if (dims == 2) return knn_query_t<double, 2>(n_points, k, edge_index, n_edges, in_degree, degree_map, w, stream);
  return knn_query_t<double, 4>(n_points, k, edge_index, n_edges, in_degree, degree_map, w, stream);
}

static int radius_query(bool fill, int32_t basis_dtype, int32_t dims, int64_t n, double r, int64_t* edge_index,
                        int64_t n_edges, const GraphWorkspace& w, cudaStream_t stream) {
  if (basis_dtype == RGNN_F32) {
    if (dims == 2) return radius_query_t<float, 2>(fill, n, r, edge_index, n_edges, w, stream);
    return radius_query_t<float, 4>(fill, n, r, edge_index, n_edges, w, stream);
  }
  if (dims == 2) return radius_query_t<double, 2>(fill, n, r, edge_index, n_edges, w, stream);
  return radius_query_t<double, 4>(fill, n, r, edge_index, n_edges, w, stream);
}

}  // namespace rgnn

using namespace rgnn;

extern "C" {

size_t rgnn_graph_workspace_bytes(int64_t n_points, int32_t n_frames) {
  if (n_points < 0 || n_frames < 0) return 0;
  SizeArena a;
  carve_graph_workspace(a, n_points, n_frames);
  return a.used;
}

int64_t rgnn_knn_edge_count(const int64_t* frame_ptr_host, int32_t n_frames, int32_t k, int* status) {
  int st = RGNN_OK;
  int64_t e = 0;
  if (frame_ptr_host == nullptr || n_frames < 0 || k < 1 || k > RGNN_MAX_K) {
    st = RGNN_ERR_INVALID_ARGUMENT;
  } else {
    for (int f = 0; f < n_frames; ++f) {
      const int64_t nf = frame_ptr_host[f + 1] - frame_ptr_host[f];
      if (nf < 0) { st = RGNN_ERR_INVALID_ARGUMENT; break; }
      if (nf > 1) {
        if (k >= nf) { st = RGNN_ERR_K_NOT_SMALLER_THAN_N; break; }
        e += nf * k;
      }
    }
  }
  if (status) *status = st;
  return st == RGNN_OK ? e : -1;
}

int rgnn_graph_build_knn(const void* basis, int32_t basis_dtype, int32_t dims, const int64_t* frame_ptr_host,
                         int32_t n_frames, int32_t k, int64_t* edge_index, int64_t n_edges, int32_t* error_flag,
                         void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t n = 0;
  RGNN_RETURN_IF_ERROR(check_graph_args(basis, basis_dtype, dims, frame_ptr_host, n_frames, workspace, workspace_bytes, &n));
  int st = RGNN_OK;
  const int64_t expect = rgnn_knn_edge_count(frame_ptr_host, n_frames, k, &st);
  if (st != RGNN_OK) return st;
  if (expect != n_edges || (n_edges > 0 && edge_index == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  Arena arena(workspace, workspace_bytes);
  GraphWorkspace w = carve_graph_workspace(arena, n, n_frames);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_RETURN_IF_ERROR(build_cell_lists(basis, basis_dtype, dims, frame_ptr_host, n_frames, k, w, stream,
                                        error_flag != nullptr ? error_flag : w.status, true));
  return knn_query(basis_dtype, dims, n, k, edge_index, n_edges, nullptr, nullptr, w, stream);
}

int rgnn_graph_build_radius_count(const void* basis, int32_t basis_dtype, int32_t dims,
                                  const int64_t* frame_ptr_host, int32_t n_frames, double r,
                                  int64_t* n_edges_host, void* workspace, size_t workspace_bytes,
                                  rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t n = 0;
  RGNN_RETURN_IF_ERROR(check_graph_args(basis, basis_dtype, dims, frame_ptr_host, n_frames, workspace, workspace_bytes, &n));
  if (n_edges_host == nullptr || !(r >= 0.0)) return RGNN_ERR_INVALID_ARGUMENT;
  Arena arena(workspace, workspace_bytes);
  GraphWorkspace w = carve_graph_workspace(arena, n, n_frames);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  *n_edges_host = 0;
  RGNN_RETURN_IF_ERROR(build_cell_lists(basis, basis_dtype, dims, frame_ptr_host, n_frames, 0, w, stream, w.status, true));
  if (n == 0) return RGNN_OK;
  RGNN_RETURN_IF_ERROR(radius_query(false, basis_dtype, dims, n, r, nullptr, 0, w, stream));
  RGNN_RETURN_IF_ERROR(exclusive_scan_i32_to_i64(w.row_count, w.row_ptr, n,
                                                 reinterpret_cast<int64_t*>(w.scan_scratch), stream));
  int32_t status_host = 0;
  RGNN_CUDA_CHECK(cudaMemcpyAsync(n_edges_host, w.row_ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  RGNN_CUDA_CHECK(cudaMemcpyAsync(&status_host, w.status, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  RGNN_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_host != 0 ? status_host : RGNN_OK;   // RGNN_ERR_NON_FINITE_INPUT
}

int rgnn_graph_build_radius_fill(const void* basis, int32_t basis_dtype, int32_t dims,
                                 const int64_t* frame_ptr_host, int32_t n_frames, double r,
                                 int64_t* edge_index, int64_t n_edges, void* workspace,
                                 size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int64_t n = 0;
  RGNN_RETURN_IF_ERROR(check_graph_args(basis, basis_dtype, dims, frame_ptr_host, n_frames, workspace, workspace_bytes, &n));
  if (n_edges < 0 || (n_edges > 0 && edge_index == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  if (n == 0 || n_edges == 0) return RGNN_OK;
  // the cell lists and row_ptr of the preceding count call are still in the workspace
  Arena arena(workspace, workspace_bytes);
  GraphWorkspace w = carve_graph_workspace(arena, n, n_frames);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_RETURN_IF_ERROR(radius_query(true, basis_dtype, dims, n, r, edge_index, n_edges, w, stream));
  sort_rows_kernel<<<div_up(n, 128), 128, 0, stream>>>(w.row_ptr, n, edge_index + n_edges);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // extern "C"
