// detection.cu -- the callers either side of the hot path:
//   * training / validation loss of the detection heads: class-weighted cross entropy + Huber box loss over the
//     foreground nodes (reference gnn/trainer.py:184-206, 276-298: torch CrossEntropyLoss(weight) and a per-node
//     Python loop over torch HuberLoss -- the reference's worst GPU stall), one deterministic reduction;
//   * non-maximum suppression of predicted boxes, axis-aligned (torchvision.ops.nms semantics, fp32) and rotated
//     (detectron2 nms_rotated semantics, fp64) -- postprocessor/postprocessing.py:336-435;
//   * nearest neighbour of every point (the "en" box representation's reference direction:
//     preprocessor/radarscenes/dataset_creation.py:314-318, nuscenes/conversion.py:133-137,
//     postprocessing.py:233-237, 468-472) on top of the k-NN search with k = 1;
//   * time_index node feature (dense rank of the point's timestamp inside its frame,
//     dataset_creation.py:214-223) and the node-id offsets of PyG's disjoint-union collate
//     (utils/data_handling.py:30) on the device.
#include <math.h>

#include "common.cuh"
#include "graph_build.cuh"

namespace rgnn {
namespace {

// ---- loss -----------------------------------------------------------------------------------------------
constexpr int kLossBlocks = 296;   // 2 x 148 SMs
constexpr int kLossThreads = 256;

// partial[b][0..4] = sum w_y * nll, sum w_y, sum huber (foreground), foreground count, invalid labels
__global__ void __launch_bounds__(kLossThreads)
loss_partial_kernel(const float* __restrict__ cls, int k, const float* __restrict__ bb, int nb, const float* __restrict__ y,
                    int64_t ldy, int64_t n, const float* __restrict__ class_weight, int bg_index, double delta,
                    double* __restrict__ partial) {
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int label = static_cast<int>(static_cast<long long>(y[i * ldy]));   // trainer.py:184: graph_batch.y[:, 0].long()
    if (label == -100) continue;                                               // CrossEntropyLoss ignore_index
    if (label < 0 || label >= k) { acc[4] += 1.0; continue; }
    // -log softmax(cls_i)[label], max-shifted, in fp64
    const float* c = cls + i * k;
    double mx = -INFINITY;
    for (int j = 0; j < k; ++j) mx = fmax(mx, static_cast<double>(c[j]));
    double se = 0.0;
    for (int j = 0; j < k; ++j) se += exp(static_cast<double>(c[j]) - mx);
    const double nll = (mx + log(se)) - static_cast<double>(c[label]);
    const double w = class_weight != nullptr ? static_cast<double>(class_weight[label]) : 1.0;
    acc[0] += w * nll;
    acc[1] += w;
    if (label != bg_index && nb > 0) {
      // torch HuberLoss(delta, reduction = mean) over the node's box parameters (trainer.py:196)
      double h = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double d = fabs(static_cast<double>(y[i * ldy + 1 + j]) - static_cast<double>(bb[i * nb + j]));
        h += d < delta ? 0.5 * d * d : delta * (d - 0.5 * delta);   // NaN falls through to the linear branch and stays NaN
      }
      acc[2] += h / static_cast<double>(nb);
      acc[3] += 1.0;
    }
  }
  __shared__ double red[kLossThreads / 32][5];
#pragma unroll
  for (int q = 0; q < 5; ++q) acc[q] = warp_sum(acc[q]);
  if ((threadIdx.x & 31) == 0)
    for (int q = 0; q < 5; ++q) red[threadIdx.x >> 5][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) t += red[w][threadIdx.x];
    partial[blockIdx.x * 5 + threadIdx.x] = t;
  }
}

// out[0..4] = loss, loss_cls, loss_bb, num_bb, invalid labels
__global__ void loss_final_kernel(const double* __restrict__ partial, int blocks, double alpha, double beta, int nan_to_zero,
                                  double* __restrict__ out) {
  const int q = threadIdx.x;
  __shared__ double tot[5];
  if (q < 5) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += partial[b * 5 + q];   // fixed order: deterministic
    tot[q] = t;
  }
  __syncthreads();
  if (q != 0) return;
  const double loss_cls = tot[0] / tot[1];               // weighted mean (0 / 0 = NaN, as torch)
  double loss_bb = tot[3] > 0.0 ? tot[2] / tot[3] : 0.0;  // trainer.py:201-204
  if (nan_to_zero && isnan(loss_bb)) loss_bb = 0.0;       // trainer.py:206-216 (training only)
  out[0] = alpha * loss_cls + beta * loss_bb;
  out[1] = loss_cls;
  out[2] = loss_bb;
  out[3] = tot[3];
  out[4] = tot[4];
}

// ---- non-maximum suppression --------------------------------------------------------------------------------
// shift[0] = |min| + 100 if any coordinate is negative, else 0 (postprocessing.py:361-365, 400-404: the
// reference moves all boxes into the positive quadrant first; in fp32 the shift takes part in the rounding
// of every IoU, so it is reproduced)
__global__ void nms_shift_kernel(const void* __restrict__ boxes, int is_f64, int64_t count, int64_t row, int cols_used,
                                 double* __restrict__ shift) {
  __shared__ double red[256];
  double mn = INFINITY;
  for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
    if (static_cast<int>(i % row) >= cols_used) continue;
    const double v = is_f64 ? static_cast<const double*>(boxes)[i] : static_cast<double>(static_cast<const float*>(boxes)[i]);
    mn = fmin(mn, v);
  }
  red[threadIdx.x] = mn;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) shift[0] = red[0] < 0.0 ? fabs(red[0]) + 100.0 : 0.0;
}

// rank of box i in (frame ascending, score descending, index ascending) order; order[rank] = i
template <typename T>
__global__ void __launch_bounds__(256)
nms_rank_kernel(const T* __restrict__ scores, const int32_t* __restrict__ frame, int n, int32_t* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T si = scores[i];
  const int fi = frame != nullptr ? frame[i] : 0;
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const int fj = frame != nullptr ? frame[j] : 0;
    const T sj = scores[j];
    const bool before = fj < fi || (fj == fi && (sj > si || (sj == si && j < i)));
    rank += before ? 1 : 0;
  }
  order[rank] = i;
}

__device__ __forceinline__ float iou_aligned(const float* a, const float* b, float shift) {
  // torchvision nms kernel (fp32): boxes (x1, y1, x2, y2)
  const float ax1 = a[0] + shift, ay1 = a[1] + shift, ax2 = a[2] + shift, ay2 = a[3] + shift;
  const float bx1 = b[0] + shift, by1 = b[1] + shift, bx2 = b[2] + shift, by2 = b[3] + shift;
  const float left = fmaxf(ax1, bx1), right = fminf(ax2, bx2);
  const float top = fmaxf(ay1, by1), bottom = fminf(ay2, by2);
  const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  const float inter = w * h;
  const float area_a = (ax2 - ax1) * (ay2 - ay1), area_b = (bx2 - bx1) * (by2 - by1);
  return inter / (area_a + area_b - inter);
}

struct P2 { double x, y; };
__device__ __forceinline__ double cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }

// corners of a rotated box (cx, cy, w, h, angle in degrees, counter-clockwise), detectron2 get_rotated_vertices
__device__ __forceinline__ void rotated_vertices(const double* b, double sx, double sy, P2* pts) {
  const double theta = b[4] * 0.01745329251994329577;
  const double c2 = cos(theta) * 0.5, s2 = sin(theta) * 0.5;
  const double cx = b[0] - sx, cy = b[1] - sy;
  pts[0].x = cx + s2 * b[3] + c2 * b[2];
  pts[0].y = cy + c2 * b[3] - s2 * b[2];
  pts[1].x = cx - s2 * b[3] + c2 * b[2];
  pts[1].y = cy - c2 * b[3] - s2 * b[2];
  pts[2].x = 2 * cx - pts[0].x;
  pts[2].y = 2 * cy - pts[0].y;
  pts[3].x = 2 * cx - pts[1].x;
  pts[3].y = 2 * cy - pts[1].y;
}

// IoU of two rotated boxes: the intersection polygon by Sutherland-Hodgman clipping of one rectangle against
// the four half-planes of the other (both convex), area by the shoelace formula.  fp64, box centres shifted to
// their midpoint first (as detectron2 does for precision).
__device__ double iou_rotated(const double* a, const double* b) {
  const double area_a = a[2] * a[3], area_b = b[2] * b[3];
  if (area_a < 1e-14 || area_b < 1e-14) return 0.0;
  const double sx = (a[0] + b[0]) * 0.5, sy = (a[1] + b[1]) * 0.5;
  P2 pa[4], pb[4];
  rotated_vertices(a, sx, sy, pa);
  rotated_vertices(b, sx, sy, pb);
  // orientation of the clip rectangle (the vertex order above is clockwise or counter-clockwise depending on the axes)
  const double orient = cross2(P2{pb[1].x - pb[0].x, pb[1].y - pb[0].y}, P2{pb[2].x - pb[1].x, pb[2].y - pb[1].y}) >= 0.0 ? 1.0 : -1.0;
  P2 poly[16], tmp[16];
  int np = 4;
  for (int i = 0; i < 4; ++i) poly[i] = pa[i];
  for (int e = 0; e < 4 && np > 0; ++e) {
    const P2 c0 = pb[e], c1 = pb[(e + 1) & 3];
    const P2 edge{c1.x - c0.x, c1.y - c0.y};
    int nt = 0;
    for (int i = 0; i < np; ++i) {
      const P2 p = poly[i], q = poly[(i + 1) % np];
      const double dp = orient * cross2(edge, P2{p.x - c0.x, p.y - c0.y});
      const double dq = orient * cross2(edge, P2{q.x - c0.x, q.y - c0.y});
      if (dp >= 0.0) tmp[nt++] = p;
      if ((dp >= 0.0) != (dq >= 0.0)) {
        const double t = dp / (dp - dq);
        tmp[nt++] = P2{p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)};
      }
    }
    np = nt;
    for (int i = 0; i < np; ++i) poly[i] = tmp[i];
  }
  if (np < 3) return 0.0;
  double area2 = 0.0;
  for (int i = 0; i < np; ++i) area2 += cross2(poly[i], poly[(i + 1) % np]);
  const double inter = fabs(area2) * 0.5;
  return inter / (area_a + area_b - inter);
}

// mask[a][w] bit b set: the box at sorted position 64 w + b (> a, same frame) overlaps the one at position a
// by more than the threshold
template <bool ROTATED>
__global__ void __launch_bounds__(64)
nms_mask_kernel(const void* __restrict__ boxes, const int32_t* __restrict__ frame, const int32_t* __restrict__ order, int n,
                double thr, const double* __restrict__ shift, unsigned long long* __restrict__ mask) {
  const int a = blockIdx.y, w = blockIdx.x, words = (n + 63) >> 6;
  if (64 * w + 63 <= a) { if (threadIdx.x == 0) mask[static_cast<int64_t>(a) * words + w] = 0ull; return; }
  const int bpos = 64 * w + threadIdx.x;
  bool hit = false;
  if (bpos < n && bpos > a) {
    const int ia = order[a], ib = order[bpos];
    if (frame == nullptr || frame[ia] == frame[ib]) {
      if (ROTATED) {
        double ba[5], bbx[5];
        const double* src = static_cast<const double*>(boxes);
        for (int j = 0; j < 5; ++j) { ba[j] = src[ia * 5 + j]; bbx[j] = src[ib * 5 + j]; }
        ba[0] += shift[0]; ba[1] += shift[0]; bbx[0] += shift[0]; bbx[1] += shift[0];
        hit = iou_rotated(ba, bbx) > thr;
      } else {
        const float* src = static_cast<const float*>(boxes);
        hit = iou_aligned(src + ia * 4, src + ib * 4, static_cast<float>(shift[0])) > static_cast<float>(thr);
      }
    }
  }
  const unsigned lo = __ballot_sync(0xffffffffu, hit);
  __shared__ unsigned halves[2];
  if ((threadIdx.x & 31) == 0) halves[threadIdx.x >> 5] = lo;
  __syncthreads();
  if (threadIdx.x == 0) mask[static_cast<int64_t>(a) * words + w] = static_cast<unsigned long long>(halves[0]) | (static_cast<unsigned long long>(halves[1]) << 32);
}

// greedy pass in sorted order: one block, removed[] in shared memory (words <= 1024: n <= 65536)
__global__ void __launch_bounds__(256)
nms_reduce_kernel(const unsigned long long* __restrict__ mask, const int32_t* __restrict__ order, int n,
                  int64_t* __restrict__ keep, int32_t* __restrict__ keep_count, uint8_t* __restrict__ keep_flag) {
  extern __shared__ unsigned long long removed[];
  const int words = (n + 63) >> 6;
  for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] = 0ull;
  for (int i = threadIdx.x; i < n; i += blockDim.x) keep_flag[i] = 0;
  __shared__ int count;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  for (int a = 0; a < n; ++a) {
    const bool alive = !((removed[a >> 6] >> (a & 63)) & 1ull);   // uniform: every thread reads the same word
    __syncthreads();
    if (alive) {
      for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] |= mask[static_cast<int64_t>(a) * words + w];
      if (threadIdx.x == 0) { keep[count] = order[a]; keep_flag[order[a]] = 1; ++count; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *keep_count = count;
}

// ---- nearest neighbour ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_nn_kernel(const T* __restrict__ basis, int dims, const int64_t* __restrict__ edge_index, int64_t n,
                 int64_t* __restrict__ nn_index, T* __restrict__ nn_points) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int64_t i = edge_index[e], j = edge_index[n + e];   // k = 1: one edge per point
  if (nn_index != nullptr) nn_index[i] = j;
  if (nn_points != nullptr)
    for (int d = 0; d < dims; ++d) nn_points[i * dims + d] = basis[j * dims + d];
}

// ---- time index ---------------------------------------------------------------------------------------------
// dense rank of the point's timestamp among the distinct timestamps of its frame (np.unique order): one block per
// frame, bitonic sort of the frame's timestamps in shared memory, a timestamp's rank = number of distinct smaller ones
constexpr int kTimeMax = 8192;
__global__ void __launch_bounds__(512)
time_index_kernel(const double* __restrict__ ts, const int64_t* __restrict__ frame_ptr, double* __restrict__ out) {
  extern __shared__ double sm[];   // [m2] sorted values, then [m2] int ranks
  const int64_t beg = frame_ptr[blockIdx.x], end = frame_ptr[blockIdx.x + 1];
  const int m = static_cast<int>(end - beg);
  if (m <= 0) return;
  int m2 = 1;
  while (m2 < m) m2 <<= 1;
  int* rank = reinterpret_cast<int*>(sm + m2);
  for (int i = threadIdx.x; i < m2; i += blockDim.x) sm[i] = i < m ? ts[beg + i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= m2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const double a = sm[i], b = sm[l];
          if ((a > b) == up) { sm[i] = b; sm[l] = a; }
        }
      }
      __syncthreads();
    }
  // rank[i] = number of distinct values before sorted position i (serial prefix over "is a new value"; m is small)
  for (int i = threadIdx.x; i < m2; i += blockDim.x) rank[i] = (i > 0 && i < m && sm[i] != sm[i - 1]) ? 1 : 0;
  __syncthreads();
  for (int o = 1; o < m2; o <<= 1) {   // Hillis-Steele inclusive scan
    int v[kTimeMax / 512];
    int c = 0;
    for (int i = threadIdx.x; i < m2; i += blockDim.x) v[c++] = i >= o ? rank[i - o] : 0;
    __syncthreads();
    c = 0;
    for (int i = threadIdx.x; i < m2; i += blockDim.x) rank[i] += v[c++];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double t = ts[beg + i];
    int lo = 0, hi = m - 1;   // first sorted position holding t
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm[mid] < t) lo = mid + 1; else hi = mid; }
    out[beg + i] = static_cast<double>(rank[lo]);
  }
}

// Frames above kTimeMax points: the same three steps in global memory.  The sort is the all-ascending bitonic
// network (per merge of size k a mirror stage l = i ^ (k - 1), then l = i ^ j for j = k/4 .. 1): every comparator
// orders its pair ascending, so the virtual +inf padding of a frame that is no power of two never moves and
// comparators that reach past the frame's end are skipped; one launch per stage over all points of the batch.
__device__ __forceinline__ int frame_of(const int64_t* __restrict__ frame_ptr, int n_frames, int64_t i) {
  int lo = 0, hi = n_frames - 1;   // last f with frame_ptr[f] <= i
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (frame_ptr[mid] <= i) lo = mid; else hi = mid - 1; }
  return lo;
}

__global__ void __launch_bounds__(256)
time_sort_stage_kernel(double* __restrict__ v, const int64_t* __restrict__ frame_ptr, int n_frames, int64_t n, int64_t k, int64_t j) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = frame_of(frame_ptr, n_frames, i);
  const int64_t beg = frame_ptr[f], m = frame_ptr[f + 1] - beg, li = i - beg;
  const int64_t l = j == 0 ? (li ^ (k - 1)) : (li ^ j);   // j == 0: the mirror stage of the merge of size k
  if (l > li && l < m) {
    const double a = v[beg + li], b = v[beg + l];
    if (a > b) { v[beg + li] = b; v[beg + l] = a; }
  }
}

__global__ void __launch_bounds__(256)
time_new_value_kernel(const double* __restrict__ v, const int64_t* __restrict__ frame_ptr, int n_frames, int64_t n, int32_t* __restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t beg = frame_ptr[frame_of(frame_ptr, n_frames, i)];
  flag[i] = (i > beg && v[i] != v[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
time_rank_kernel(const double* __restrict__ ts, const double* __restrict__ v, const int32_t* __restrict__ prefix,
                 const int64_t* __restrict__ frame_ptr, int n_frames, int64_t n, double* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = frame_of(frame_ptr, n_frames, i);
  const int64_t beg = frame_ptr[f], end = frame_ptr[f + 1];
  const double t = ts[i];
  int64_t lo = beg, hi = end - 1;   // first sorted position holding t
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (v[mid] < t) lo = mid + 1; else hi = mid; }
  out[i] = static_cast<double>(prefix[lo + 1] - prefix[beg]);   // distinct values before it (flag[beg] = 0)
}

struct TimeIndexWorkspace {
  double* sorted;      // [n]
  int32_t* flag;       // [n]
  int32_t* prefix;     // [n + 1]
  int32_t* scan;       // scan_scratch_ints(n)
};
template <typename ArenaT>
TimeIndexWorkspace carve_time_index(ArenaT& a, int64_t n) {
  TimeIndexWorkspace w;
  w.sorted = a.template take<double>(n);
  w.flag = a.template take<int32_t>(n);
  w.prefix = a.template take<int32_t>(n + 1);
  w.scan = a.template take<int32_t>(scan_scratch_ints(n));
  return w;
}

// ---- disjoint-union collate -----------------------------------------------------------------------------------
// edge_index [2, E] holds frame-local node ids, the edges of frame f are columns edge_ptr[f] .. edge_ptr[f+1]:
// add the frame's node offset to both rows (what PyG's Batch.from_data_list does to edge_index)
__global__ void __launch_bounds__(256)
collate_offsets_kernel(int64_t* __restrict__ edge_index, int64_t n_edges, const int64_t* __restrict__ edge_ptr,
                       const int64_t* __restrict__ node_ptr, int n_frames, int64_t* __restrict__ batch_of_node) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_edges) {
    int lo = 0, hi = n_frames;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (edge_ptr[mid] <= e) lo = mid; else hi = mid; }
    const int64_t off = node_ptr[lo];
    edge_index[e] += off;
    edge_index[n_edges + e] += off;
  }
  const int64_t n_nodes = node_ptr[n_frames];
  if (batch_of_node != nullptr && e < n_nodes) {
    int lo = 0, hi = n_frames;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (node_ptr[mid] <= e) lo = mid; else hi = mid; }
    batch_of_node[e] = lo;
  }
}

}  // namespace
}  // namespace rgnn

using namespace rgnn;

extern "C" {

size_t rgnn_detection_loss_workspace_bytes(void) { return sizeof(double) * kLossBlocks * 5 + kAlign; }

int rgnn_detection_loss(const float* cls, int32_t n_classes, const float* bb, int32_t n_box, const float* y, int64_t ldy,
                        int64_t n_nodes, const float* class_weight, int32_t bg_index, float cls_loss_weight,
                        float bb_loss_weight, float huber_delta, int32_t nan_to_zero, double* out, void* workspace,
                        size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_nodes < 0 || n_classes < 1 || n_box < 0 || out == nullptr || ldy < 1 + n_box || !(huber_delta > 0.f)) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_nodes > 0 && (cls == nullptr || y == nullptr || (n_box > 0 && bb == nullptr))) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_detection_loss_workspace_bytes()) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  Arena arena(workspace, workspace_bytes);
  double* partial = arena.take<double>(kLossBlocks * 5);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_PROFILE("detection_loss", stream);
  loss_partial_kernel<<<kLossBlocks, kLossThreads, 0, stream>>>(cls, n_classes, bb, n_box, y, ldy, n_nodes, class_weight, bg_index,
                                                               static_cast<double>(huber_delta), partial);
  RGNN_LAUNCH_CHECK();
  loss_final_kernel<<<1, 32, 0, stream>>>(partial, kLossBlocks, static_cast<double>(cls_loss_weight),
                                          static_cast<double>(bb_loss_weight), nan_to_zero, out);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

size_t rgnn_nms_workspace_bytes(int64_t n_boxes) {
  if (n_boxes < 0 || n_boxes > 65536) return 0;
  SizeArena a;
  a.take<int32_t>(static_cast<size_t>(n_boxes));
  a.take<unsigned long long>(static_cast<size_t>(n_boxes) * ((n_boxes + 63) / 64));
  a.take<double>(4);
  return a.used;
}

int rgnn_nms(const void* boxes, int32_t rotated, const void* scores, const int32_t* box_frame, int64_t n_boxes,
             double iou_threshold, int32_t shift_negative, int64_t* keep, int32_t* keep_count, uint8_t* keep_flag,
             void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_boxes < 0 || n_boxes > 65536 || keep_count == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_boxes > 0 && (boxes == nullptr || scores == nullptr || keep == nullptr || keep_flag == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_nms_workspace_bytes(n_boxes)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  if (n_boxes == 0) { RGNN_CUDA_CHECK(cudaMemsetAsync(keep_count, 0, sizeof(int32_t), stream)); return RGNN_OK; }
  const int n = static_cast<int>(n_boxes), words = (n + 63) / 64;
  Arena arena(workspace, workspace_bytes);
  int32_t* order = arena.take<int32_t>(n);
  unsigned long long* mask = arena.take<unsigned long long>(static_cast<size_t>(n) * words);
  double* shift = arena.take<double>(4);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_PROFILE("nms", stream);
  if (shift_negative) {
    // rotated: only the centres are shifted (postprocessing.py:362-365), aligned: all four coordinates (:401-404)
    nms_shift_kernel<<<1, 256, 0, stream>>>(boxes, rotated ? 1 : 0, n_boxes * (rotated ? 5 : 4), rotated ? 5 : 4, rotated ? 2 : 4, shift);
    RGNN_LAUNCH_CHECK();
  } else {
    RGNN_CUDA_CHECK(cudaMemsetAsync(shift, 0, sizeof(double), stream));
  }
  if (rotated) nms_rank_kernel<double><<<div_up(n, 256), 256, 0, stream>>>(static_cast<const double*>(scores), box_frame, n, order);
  else nms_rank_kernel<float><<<div_up(n, 256), 256, 0, stream>>>(static_cast<const float*>(scores), box_frame, n, order);
  RGNN_LAUNCH_CHECK();
  const dim3 grid(words, n);
  if (rotated) nms_mask_kernel<true><<<grid, 64, 0, stream>>>(boxes, box_frame, order, n, iou_threshold, shift, mask);
  else nms_mask_kernel<false><<<grid, 64, 0, stream>>>(boxes, box_frame, order, n, iou_threshold, shift, mask);
  RGNN_LAUNCH_CHECK();
  nms_reduce_kernel<<<1, 256, sizeof(unsigned long long) * words, stream>>>(mask, order, n, keep, keep_count, keep_flag);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

size_t rgnn_nearest_neighbor_workspace_bytes(int64_t n_points, int32_t n_frames) {
  if (n_points < 0 || n_frames < 1) return 0;
  return rgnn_graph_workspace_bytes(n_points, n_frames) + align_up(sizeof(int64_t) * 2 * static_cast<size_t>(n_points)) + kAlign;
}

int rgnn_nearest_neighbor(const void* basis, int32_t basis_dtype, int32_t dims, const int64_t* frame_ptr_host, int32_t n_frames,
                          int64_t* nn_index, void* nn_points, void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (frame_ptr_host == nullptr || n_frames < 1 || (nn_index == nullptr && nn_points == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t n = frame_ptr_host[n_frames];
  if (n < 0) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_nearest_neighbor_workspace_bytes(n, n_frames)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  if (n == 0) return RGNN_OK;
  // every frame needs a second point (sklearn: "Expected n_neighbors < n_samples_fit"): checked by the k-NN build
  const size_t graph_bytes = rgnn_graph_workspace_bytes(n, n_frames);
  char* base = static_cast<char*>(workspace);
  int64_t* edges = reinterpret_cast<int64_t*>(base + align_up(graph_bytes));
  int st = RGNN_OK;
  const int64_t e = rgnn_knn_edge_count(frame_ptr_host, n_frames, 1, &st);   // points of one-point frames have no neighbour
  if (st != RGNN_OK) return st;
  if (nn_index != nullptr) RGNN_CUDA_CHECK(cudaMemsetAsync(nn_index, 0xff, sizeof(int64_t) * n, stream));   // -1: none
  if (e == 0) return RGNN_OK;
  RGNN_RETURN_IF_ERROR(rgnn_graph_build_knn(basis, basis_dtype, dims, frame_ptr_host, n_frames, 1, edges, e, nullptr, base, graph_bytes, stream_));
  RGNN_PROFILE("nearest_neighbor", stream);
  if (basis_dtype == RGNN_F64)
    gather_nn_kernel<double><<<div_up(e, 256), 256, 0, stream>>>(static_cast<const double*>(basis), dims, edges, e, nn_index, static_cast<double*>(nn_points));
  else
    gather_nn_kernel<float><<<div_up(e, 256), 256, 0, stream>>>(static_cast<const float*>(basis), dims, edges, e, nn_index, static_cast<float*>(nn_points));
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

size_t rgnn_time_index_workspace_bytes(int64_t n_points) {
  if (n_points < 0 || n_points > 0x7ffffff0LL) return 0;
  SizeArena a;
  carve_time_index(a, n_points);
  return a.used;
}

int rgnn_time_index(const double* timestamp, const int64_t* frame_ptr, const int64_t* frame_ptr_host, int32_t n_frames,
                    double* time_index, void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (frame_ptr == nullptr || frame_ptr_host == nullptr || n_frames < 1) return RGNN_ERR_INVALID_ARGUMENT;
  int64_t longest = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int64_t m = frame_ptr_host[f + 1] - frame_ptr_host[f];
    if (m < 0) return RGNN_ERR_INVALID_ARGUMENT;
    longest = m > longest ? m : longest;
  }
  if (longest == 0) return RGNN_OK;
  if (timestamp == nullptr || time_index == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (longest > kTimeMax) {
    // a frame too long for the shared-memory sort: the global-memory path over the whole batch
    const int64_t n = frame_ptr_host[n_frames];
    if (n > 0x7ffffff0LL) return RGNN_ERR_INVALID_ARGUMENT;
    if (workspace == nullptr || workspace_bytes < rgnn_time_index_workspace_bytes(n)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
    Arena arena(workspace, workspace_bytes);
    TimeIndexWorkspace w = carve_time_index(arena, n);
    if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
    RGNN_PROFILE("time_index", stream);
    RGNN_CUDA_CHECK(cudaMemcpyAsync(w.sorted, timestamp, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    const unsigned blocks = div_up(n, 256);
    for (int64_t k = 2; (k >> 1) < longest; k <<= 1) {
      time_sort_stage_kernel<<<blocks, 256, 0, stream>>>(w.sorted, frame_ptr, n_frames, n, k, 0);
      for (int64_t j = k >> 2; j > 0; j >>= 1) time_sort_stage_kernel<<<blocks, 256, 0, stream>>>(w.sorted, frame_ptr, n_frames, n, k, j);
    }
    time_new_value_kernel<<<blocks, 256, 0, stream>>>(w.sorted, frame_ptr, n_frames, n, w.flag);
    RGNN_LAUNCH_CHECK();
    RGNN_RETURN_IF_ERROR(exclusive_scan_i32(w.flag, w.prefix, n, w.scan, stream));
    time_rank_kernel<<<blocks, 256, 0, stream>>>(timestamp, w.sorted, w.prefix, frame_ptr, n_frames, n, time_index);
    RGNN_LAUNCH_CHECK();
    return RGNN_OK;
  }
  int m2 = 1;
  while (m2 < longest) m2 <<= 1;
  const size_t smem = static_cast<size_t>(m2) * (sizeof(double) + sizeof(int));
  static bool configured[kMaxDevices] = {};
  if (smem > 48 * 1024) RGNN_CUDA_CHECK(opt_in_dynamic_smem(time_index_kernel, configured, kTimeMax * 12));
  RGNN_PROFILE("time_index", stream);
  time_index_kernel<<<n_frames, 512, smem, stream>>>(timestamp, frame_ptr, time_index);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int rgnn_collate_offsets(int64_t* edge_index, int64_t n_edges, const int64_t* edge_ptr, const int64_t* node_ptr, int32_t n_frames,
                         int64_t n_nodes, int64_t* batch_of_node, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_edges < 0 || n_nodes < 0 || n_frames < 1 || edge_ptr == nullptr || node_ptr == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && edge_index == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t work = n_edges > n_nodes ? n_edges : n_nodes;
  if (work == 0) return RGNN_OK;
  RGNN_PROFILE("collate", stream);
  collate_offsets_kernel<<<div_up(work, 256), 256, 0, stream>>>(edge_index, n_edges, edge_ptr, node_ptr, n_frames, batch_of_node);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // extern "C"
