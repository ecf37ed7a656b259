// pipeline.cu -- the north-star hot path in one call: neighbour search -> edge_index +
// edge_attr -> CSC view -> L x (graph convolution, training-mode BatchNorm, ReLU).
// Fuses what the reference runs as two offline/online stages joined by .pt files
// (preprocessor/radarscenes/dataset_creation.py:187-229 -> gnn/gnn_models.py:124-128).
#include <stdlib.h>

#include <mutex>

#include "conv.cuh"
#include "csc.cuh"
#include "features.cuh"
#include "graph_build.cuh"

namespace rgnn {
namespace {

struct PipelineWorkspace {
  GraphWorkspace graph;
  CscWorkspace csc;
  float* basis4;       // [N, 4] = [pos | vel] when distance_dims == 4
  int32_t* csc_ptr;    // [N + 1]
  int32_t* csc_src;    // [E]
  int32_t* csc_eid;    // [E]
  float* ea_csc;       // [E, De]
  float* h[2];         // [N, c_max] ping-pong layer outputs
  float* stats;        // [L][3 * c_max]: mean | scale | beta per layer
  double* bn_scratch;
  size_t conv_mark;    // arena offset where the per-layer conv workspace starts
  size_t conv_bytes;   // its size (max over layers)
};

__global__ void __launch_bounds__(256)
concat_basis_kernel(const float* __restrict__ pos, const float* __restrict__ vel, int64_t n, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = *reinterpret_cast<const float2*>(pos + i * 2);
  const float2 v = *reinterpret_cast<const float2*>(vel + i * 2);
  *reinterpret_cast<float4*>(out + i * 4) = make_float4(p.x, p.y, v.x, v.y);
}

int validate(const rgnn_pipeline_desc* d, int32_t* de_out, int32_t* c_max_out) {
  if (d == nullptr || d->layers == nullptr || d->n_layers < 1 || d->n_layers > 64) return RGNN_ERR_INVALID_ARGUMENT;
  if (d->search != 0 && d->search != 1) return RGNN_ERR_INVALID_ARGUMENT;
  if (d->distance_dims != 2 && d->distance_dims != 4) return RGNN_ERR_INVALID_ARGUMENT;
  EdgeFeatureSpec spec;
  RGNN_RETURN_IF_ERROR(make_edge_feature_spec(d->edge_features, d->n_edge_features, d->edge_mode, &spec));
  int32_t c_max = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    const rgnn_conv_desc& c = d->layers[l];
    if (c.edge_dim != spec.width) return RGNN_ERR_INVALID_ARGUMENT;
    if (l > 0 && c.in_channels != d->layers[l - 1].out_channels) return RGNN_ERR_INVALID_ARGUMENT;
    if (c.out_channels > c_max) c_max = c.out_channels;
  }
  *de_out = spec.width;
  *c_max_out = c_max;
  return RGNN_OK;
}

template <typename ArenaT>
int carve(ArenaT& a, const rgnn_pipeline_desc* d, int64_t n, int32_t n_frames, int64_t e, int32_t de,
          int32_t c_max, bool with_weights, PipelineWorkspace* w) {
  w->graph = carve_graph_workspace(a, n, n_frames);
  w->csc = carve_csc_workspace(a, n);
  w->basis4 = d->distance_dims == 4 ? a.template take<float>(static_cast<size_t>(n) * 4) : nullptr;
  w->csc_ptr = a.template take<int32_t>(n + 1);
  w->csc_src = a.template take<int32_t>(e);
  w->csc_eid = a.template take<int32_t>(e);
  w->ea_csc = a.template take<float>(static_cast<size_t>(e) * de);
  w->h[0] = a.template take<float>(static_cast<size_t>(n) * c_max);
  w->h[1] = a.template take<float>(static_cast<size_t>(n) * c_max);
  w->stats = a.template take<float>(static_cast<size_t>(d->n_layers) * 3 * c_max);
  w->bn_scratch = a.template take<double>(bn_scratch_doubles(n, c_max));
  w->conv_mark = a.used;
  size_t worst = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    rgnn_conv_desc c = d->layers[l];
    if (!with_weights) {
      static const float dummy = 0.f;
      for (int i = 0; i < RGNN_MAX_MLP_LAYERS; ++i) c.pre_weight[i] = c.pre_bias[i] = c.post_weight[i] = c.post_bias[i] = &dummy;
      c.edge_encoder_weight = c.edge_encoder_bias = &dummy;
    }
    ConvShape s;
    RGNN_RETURN_IF_ERROR(conv_shape(c, &s));
    SizeArena sa;
    sa.used = 0;
    carve_conv_workspace(sa, c, s, n, e, false);
    if (sa.used > worst) worst = sa.used;
  }
  w->conv_bytes = worst + kAlign;
  a.template take<char>(w->conv_bytes);
  return RGNN_OK;
}

// Streams, events and the replay cache of the host-buffer entry point, one set per device.  The entry point
// runs on its own stream (ordered after the caller's, synchronised before returning), which makes the call
// capturable whatever stream the caller passes (the legacy default stream cannot be captured).
struct HostPathStreams {
  cudaStream_t main, copy;
  cudaEvent_t start, x0_ready, graph_done, copies_done;
  int32_t* flag_pinned;        // error flag read back by the (possibly replayed) copy node
  uint64_t graph_key;          // arguments the cached graph was captured for (0 = none)
  cudaGraphExec_t graph_exec;
  bool ok;
};
HostPathStreams* host_path_streams() {
  static HostPathStreams per_device[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostPathStreams& h = per_device[dev];
  if (!h.ok) {
    if (cudaStreamCreateWithFlags(&h.main, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&h.start, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.x0_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.graph_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.copies_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaMallocHost(reinterpret_cast<void**>(&h.flag_pinned), 64) != cudaSuccess) return nullptr;
    h.graph_key = 0;
    h.graph_exec = nullptr;
    h.ok = true;
  }
  return &h;
}

inline void hash_bytes(uint64_t& h, const void* p, size_t n) {   // FNV-1a
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
}

}  // namespace
}  // namespace rgnn

using namespace rgnn;

extern "C" {

size_t rgnn_pipeline_workspace_bytes(const rgnn_pipeline_desc* desc, int64_t n_points, int32_t n_frames,
                                     int64_t n_edges) {
  int32_t de = 0, c_max = 0;
  if (n_points < 0 || n_frames < 1 || n_edges < 0 || validate(desc, &de, &c_max) != RGNN_OK) return 0;
  SizeArena a;
  PipelineWorkspace w;
  if (carve(a, desc, n_points, n_frames, n_edges, de, c_max, false, &w) != RGNN_OK) return 0;
  return a.used;
}

}  // extern "C"

// x0_ready (optional): event the stream waits for before the first layer reads x0; graph_done (optional):
// recorded once edge_index / edge_attr are final -- the host entry point overlaps its copies with them
static int pipeline_forward_impl(const rgnn_pipeline_desc* desc, const float* pos, const float* vel, const float* x0,
                                 const int64_t* frame_ptr_host, int32_t n_frames, int64_t* edge_index, int64_t n_edges,
                                 float* edge_attr, float* h, int32_t* error_flag, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream, cudaEvent_t x0_ready, cudaEvent_t graph_done) {
  rgnn_stream_t stream_ = static_cast<rgnn_stream_t>(stream);
  int32_t de = 0, c_max = 0;
  RGNN_RETURN_IF_ERROR(validate(desc, &de, &c_max));
  if (frame_ptr_host == nullptr || n_frames < 1 || frame_ptr_host[0] != 0 || n_edges < 0) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t n = frame_ptr_host[n_frames];
  if (n < 0 || n > 0x7ffffff0LL || n_edges > 0x7ffffff0LL) return RGNN_ERR_INVALID_ARGUMENT;
  if (error_flag == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (n > 0 && (pos == nullptr || vel == nullptr || x0 == nullptr || h == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && (edge_index == nullptr || (de > 0 && edge_attr == nullptr))) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_pipeline_workspace_bytes(desc, n, n_frames, n_edges))
    return RGNN_ERR_WORKSPACE_TOO_SMALL;
  Arena arena(workspace, workspace_bytes);
  PipelineWorkspace w;
  RGNN_RETURN_IF_ERROR(carve(arena, desc, n, n_frames, n_edges, de, c_max, true, &w));
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_CUDA_CHECK(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), stream));
  if (n == 0) return RGNN_OK;

  // ---- 1. neighbour search -> edge_index ---------------------------------------------
  const float* basis = pos;
  if (desc->distance_dims == 4) {
    concat_basis_kernel<<<div_up(n, 256), 256, 0, stream>>>(pos, vel, n, w.basis4);
    RGNN_LAUNCH_CHECK();
    basis = w.basis4;
  }
  bool counts_ready = false;
  if (desc->search == 0) {
    int st = RGNN_OK;
    const int64_t expect = rgnn_knn_edge_count(frame_ptr_host, n_frames, desc->k, &st);
    if (st != RGNN_OK) return st;
    if (expect != n_edges) return RGNN_ERR_INVALID_ARGUMENT;
    RGNN_RETURN_IF_ERROR(build_cell_lists(basis, RGNN_F32, desc->distance_dims, frame_ptr_host, n_frames, desc->k, w.graph, stream,
                                          error_flag, false));   // non-finite coordinates -> RGNN_ERR_NON_FINITE_INPUT
    RGNN_CUDA_CHECK(cudaMemsetAsync(w.csc.count, 0, sizeof(int32_t) * (n + 1), stream));
    RGNN_RETURN_IF_ERROR(knn_query(RGNN_F32, desc->distance_dims, n, desc->k, edge_index, n_edges, w.csc.count,
                                   w.graph.rank, w.graph, stream));
    counts_ready = true;
  } else {
    // radius: the edge count is data dependent; the caller learned it from
    // rgnn_graph_build_radius_count, which is repeated here on this workspace (one host sync)
    int64_t counted = 0;
    RGNN_RETURN_IF_ERROR(rgnn_graph_build_radius_count(basis, RGNN_F32, desc->distance_dims, frame_ptr_host, n_frames,
                                                       desc->r, &counted, workspace, workspace_bytes, stream_));
    if (counted != n_edges) return RGNN_ERR_INVALID_ARGUMENT;
    RGNN_RETURN_IF_ERROR(rgnn_graph_build_radius_fill(basis, RGNN_F32, desc->distance_dims, frame_ptr_host, n_frames,
                                                      desc->r, edge_index, n_edges, workspace, workspace_bytes, stream_));
  }

  // ---- 2 + 3. edge attributes, CSC view, edge attributes in slot order -------------------------
  EdgeFeatureSpec spec;
  RGNN_RETURN_IF_ERROR(make_edge_feature_spec(desc->edge_features, desc->n_edge_features, desc->edge_mode, &spec));
  bool ordered = false;
  for (int l = 0; l < desc->n_layers; ++l)
    if (desc->layers[l].aggr == RGNN_AGGR_ADD || desc->layers[l].aggr == RGNN_AGGR_MEAN) ordered = true;
  if (!ordered && spec.width > 0) {
    // max / min only: slot order inside a segment is free, so one pass fills the slots and writes the
    // attributes in both orders (conv stack in CELL-SORTED node order, see below)
    FusedEdgeAttr fea;
    fea.pos = pos; fea.vel = vel; fea.spec = spec; fea.edge_attr = edge_attr; fea.ea_csc = w.ea_csc; fea.error_flag = error_flag;
    if (desc->search == 0 && counts_ready)
      RGNN_RETURN_IF_ERROR(csc_build_fused_knn(edge_index, n_edges, n, desc->k, desc->distance_dims, w.graph, w.csc,
                                               w.csc_ptr, w.csc_src, w.csc_eid, stream, fea));
    else
      RGNN_RETURN_IF_ERROR(csc_build_fused(edge_index, n_edges, n, counts_ready, w.csc, w.csc_ptr, w.csc_src, w.csc_eid,
                                           stream, w.graph.rank, fea));
  } else {
  RGNN_RETURN_IF_ERROR(launch_edge_features(pos, vel, RGNN_F32, 2, 2, edge_index, n_edges, n, spec, edge_attr, RGNN_F32,
                                            error_flag, stream));
  // The conv stack runs in CELL-SORTED node order (node r = original point sorted_idx[r]): spatial
  // neighbours are then neighbours in memory, so the per-edge gathers of B[source] hit L1 / L2.
  RGNN_RETURN_IF_ERROR(csc_build(edge_index, n_edges, n, counts_ready, ordered, w.csc, w.csc_ptr, w.csc_src, w.csc_eid,
                                 stream, w.graph.rank));
  RGNN_RETURN_IF_ERROR(gather_edge_rows(edge_attr, w.csc_eid, n_edges, de, w.ea_csc, stream));
  }

  if (graph_done != nullptr) RGNN_CUDA_CHECK(cudaEventRecord(graph_done, stream));
  if (x0_ready != nullptr) RGNN_CUDA_CHECK(cudaStreamWaitEvent(stream, x0_ready, 0));

  // ---- 4. conv -> BatchNorm(train) -> ReLU, L times ------------------------------------------
  ConvInput in;
  in.x = x0; in.ldx = desc->layers[0].in_channels;
  in.rows = w.graph.sorted_idx;  // layer 0 gathers its input rows into sorted order
  for (int l = 0; l < desc->n_layers; ++l) {
    const rgnn_conv_desc& c = desc->layers[l];
    ConvShape s;
    RGNN_RETURN_IF_ERROR(conv_shape(c, &s));
    Arena sub(static_cast<char*>(workspace) + w.conv_mark, w.conv_bytes);
    ConvWorkspace cw = carve_conv_workspace(sub, c, s, n, n_edges, false);
    if (sub.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
    float* out = w.h[l & 1];
    int64_t fused_partials = 0;
    RGNN_RETURN_IF_ERROR(conv_forward(c, s, in, n, w.csc_ptr, w.csc_src, nullptr, w.ea_csc, n_edges, out, cw, stream,
                                      &fused_partials));
    float* st = w.stats + static_cast<size_t>(l) * 3 * c_max;
    const float* bw = desc->bn_weight != nullptr ? desc->bn_weight[l] : nullptr;
    const float* bb = desc->bn_bias != nullptr ? desc->bn_bias[l] : nullptr;
    float* rmean = desc->bn_running_mean != nullptr ? desc->bn_running_mean[l] : nullptr;
    float* rvar = desc->bn_running_var != nullptr ? desc->bn_running_var[l] : nullptr;
    if (rmean == nullptr || rvar == nullptr) rmean = rvar = nullptr;
    if (fused_partials > 0) {
      // the node-update contraction already produced the per-tile column sums
      RGNN_RETURN_IF_ERROR(bn_finalize_partials(cw.bn_partial, fused_partials, n, s.c_out, bw, bb, desc->bn_eps,
                                                desc->bn_momentum, rmean, rvar, st, st + c_max, st + 2 * c_max, stream));
    } else {
      RGNN_RETURN_IF_ERROR(bn_statistics(out, s.c_out, n, s.c_out, bw, bb, desc->bn_eps, desc->bn_momentum, rmean, rvar,
                                         st, st + c_max, st + 2 * c_max, w.bn_scratch, stream));
    }
    in.x = out; in.ldx = s.c_out; in.rows = nullptr;
    in.mean = st; in.scale = st + c_max; in.beta = st + 2 * c_max; in.relu = 1;
  }
  const int32_t c_last = desc->layers[desc->n_layers - 1].out_channels;
  // last BatchNorm + ReLU, scattered back to the caller's node order
  return bn_apply(in.x, in.ldx, n, c_last, in.mean, in.scale, in.beta, 1, h, c_last, stream, w.graph.sorted_idx);
}

extern "C" {

int rgnn_pipeline_forward(const rgnn_pipeline_desc* desc, const float* pos, const float* vel, const float* x0,
                          const int64_t* frame_ptr_host, int32_t n_frames, int64_t* edge_index, int64_t n_edges,
                          float* edge_attr, float* h, int32_t* error_flag, void* workspace,
                          size_t workspace_bytes, rgnn_stream_t stream_) {
  return pipeline_forward_impl(desc, pos, vel, x0, frame_ptr_host, n_frames, edge_index, n_edges, edge_attr, h, error_flag,
                               workspace, workspace_bytes, static_cast<cudaStream_t>(stream_), nullptr, nullptr);
}

size_t rgnn_pipeline_host_workspace_bytes(const rgnn_pipeline_desc* desc, int64_t n_points, int32_t n_frames,
                                          int64_t n_edges, int32_t c0) {
  int32_t de = 0, c_max = 0;
  if (n_points < 0 || n_frames < 1 || n_edges < 0 || c0 < 1 || validate(desc, &de, &c_max) != RGNN_OK) return 0;
  const size_t inner = rgnn_pipeline_workspace_bytes(desc, n_points, n_frames, n_edges);
  if (inner == 0) return 0;
  SizeArena a;
  a.take<float>(static_cast<size_t>(n_points) * 2);
  a.take<float>(static_cast<size_t>(n_points) * 2);
  a.take<float>(static_cast<size_t>(n_points) * c0);
  a.take<int64_t>(static_cast<size_t>(n_edges) * 2);
  a.take<float>(static_cast<size_t>(n_edges) * de);
  a.take<float>(static_cast<size_t>(n_points) * desc->layers[desc->n_layers - 1].out_channels);
  a.take<int32_t>(64);
  a.take<char>(inner);
  return a.used;
}

int rgnn_pipeline_forward_host(const rgnn_pipeline_desc* desc, const float* pos_host, const float* vel_host,
                               const float* x0_host, int32_t c0, const int64_t* frame_ptr_host, int32_t n_frames,
                               int64_t* edge_index_host, int64_t n_edges, float* edge_attr_host, float* h_host,
                               void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int32_t de = 0, c_max = 0;
  RGNN_RETURN_IF_ERROR(validate(desc, &de, &c_max));
  if (frame_ptr_host == nullptr || n_frames < 1 || n_edges < 0 || c0 != desc->layers[0].in_channels) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t n = frame_ptr_host[n_frames];
  if (n < 0) return RGNN_ERR_INVALID_ARGUMENT;
  if (n > 0 && (pos_host == nullptr || vel_host == nullptr || x0_host == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  const size_t need = rgnn_pipeline_host_workspace_bytes(desc, n, n_frames, n_edges, c0);
  if (workspace == nullptr || need == 0 || workspace_bytes < need) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  const int32_t c_last = desc->layers[desc->n_layers - 1].out_channels;
  Arena a(workspace, workspace_bytes);
  float* pos = a.take<float>(static_cast<size_t>(n) * 2);
  float* vel = a.take<float>(static_cast<size_t>(n) * 2);
  float* x0 = a.take<float>(static_cast<size_t>(n) * c0);
  int64_t* edge_index = a.take<int64_t>(static_cast<size_t>(n_edges) * 2);
  float* edge_attr = a.take<float>(static_cast<size_t>(n_edges) * de);
  float* h = a.take<float>(static_cast<size_t>(n) * c_last);
  int32_t* flag = a.take<int32_t>(64);
  const size_t inner = rgnn_pipeline_workspace_bytes(desc, n, n_frames, n_edges);
  char* inner_ws = a.take<char>(inner);
  if (a.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  // Copies overlap the compute: x0 (the bulk of the input) travels on a side stream while the main stream
  // builds the graph, and edge_index / edge_attr travel back while the layers run.  The whole sequence
  // (copies, ~40 kernels, cross-stream events) is captured into a CUDA graph the first time it is seen and
  // replayed while the arguments stay the same: the eager launches cost ~0.1 ms of host time per call.
  // one call at a time per process: the per-device streams, events and the cached graph are shared state
  static std::mutex host_path_mutex;
  std::lock_guard<std::mutex> host_path_lock(host_path_mutex);
  HostPathStreams* hs = host_path_streams();
  if (hs == nullptr) return RGNN_ERR_CUDA;
  cudaStream_t ms = hs->main;
  RGNN_CUDA_CHECK(cudaEventRecord(hs->start, stream));   // order our stream after the caller's
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ms, hs->start, 0));

  auto enqueue = [&]() -> int {
    if (n > 0) {
      RGNN_CUDA_CHECK(cudaMemcpyAsync(pos, pos_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ms));
      RGNN_CUDA_CHECK(cudaMemcpyAsync(vel, vel_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ms));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->x0_ready, ms));                 // fork the side stream
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(hs->copy, hs->x0_ready, 0));
      RGNN_CUDA_CHECK(cudaMemcpyAsync(x0, x0_host, sizeof(float) * n * c0, cudaMemcpyHostToDevice, hs->copy));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->x0_ready, hs->copy));
    }
    RGNN_RETURN_IF_ERROR(pipeline_forward_impl(desc, pos, vel, x0, frame_ptr_host, n_frames, edge_index, n_edges,
                                               edge_attr, h, flag, inner_ws, inner, ms, n > 0 ? hs->x0_ready : nullptr,
                                               n > 0 ? hs->graph_done : nullptr));
    if (n > 0) {
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(hs->copy, hs->graph_done, 0));
      if (edge_index_host != nullptr && n_edges > 0)
        RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_index_host, edge_index, sizeof(int64_t) * n_edges * 2, cudaMemcpyDeviceToHost, hs->copy));
      if (edge_attr_host != nullptr && n_edges > 0 && de > 0)
        RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_attr_host, edge_attr, sizeof(float) * n_edges * de, cudaMemcpyDeviceToHost, hs->copy));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->copies_done, hs->copy));
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(ms, hs->copies_done, 0));        // join
    }
    if (h_host != nullptr && n > 0)
      RGNN_CUDA_CHECK(cudaMemcpyAsync(h_host, h, sizeof(float) * n * c_last, cudaMemcpyDeviceToHost, ms));
    RGNN_CUDA_CHECK(cudaMemcpyAsync(hs->flag_pinned, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, ms));
    return RGNN_OK;
  };

  // replay key: every argument the enqueued work depends on (pointers are compared, not their targets --
  // buffers are read when the graph runs)
  static int graph_enabled = -1;
  if (graph_enabled < 0) { const char* e = getenv("RGNN_HOST_GRAPH"); graph_enabled = (e != nullptr && e[0] == '0') ? 0 : 1; }
  uint64_t key = 1469598103934665603ull;
  hash_bytes(key, desc, sizeof(*desc));
  hash_bytes(key, desc->layers, sizeof(rgnn_conv_desc) * desc->n_layers);
  if (desc->bn_weight != nullptr) hash_bytes(key, desc->bn_weight, sizeof(float*) * desc->n_layers);
  if (desc->bn_bias != nullptr) hash_bytes(key, desc->bn_bias, sizeof(float*) * desc->n_layers);
  if (desc->bn_running_mean != nullptr) hash_bytes(key, desc->bn_running_mean, sizeof(float*) * desc->n_layers);
  if (desc->bn_running_var != nullptr) hash_bytes(key, desc->bn_running_var, sizeof(float*) * desc->n_layers);
  hash_bytes(key, frame_ptr_host, sizeof(int64_t) * (n_frames + 1));
  const void* ptrs[] = {pos_host, vel_host, x0_host, edge_index_host, edge_attr_host, h_host, workspace};
  hash_bytes(key, ptrs, sizeof(ptrs));
  const int64_t nums[] = {c0, n_frames, n_edges, static_cast<int64_t>(workspace_bytes)};
  hash_bytes(key, nums, sizeof(nums));
  if (key == 0) key = 1;

  bool launched = false;
  if (graph_enabled && n > 0 && desc->search == 0) {
    if (hs->graph_exec != nullptr && hs->graph_key == key) {
      launched = cudaGraphLaunch(hs->graph_exec, ms) == cudaSuccess;
    } else {
      if (hs->graph_exec != nullptr) { cudaGraphExecDestroy(hs->graph_exec); hs->graph_exec = nullptr; hs->graph_key = 0; }
      if (cudaStreamBeginCapture(ms, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int st = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ms, &graph);
        if (st == RGNN_OK && ce == cudaSuccess && graph != nullptr &&
            cudaGraphInstantiate(&hs->graph_exec, graph, 0) == cudaSuccess) {
          hs->graph_key = key;
          launched = cudaGraphLaunch(hs->graph_exec, ms) == cudaSuccess;
        } else {
          hs->graph_exec = nullptr;
          (void)cudaGetLastError();   // a failed capture falls back to eager launches below
        }
        if (graph != nullptr) cudaGraphDestroy(graph);
        if (st != RGNN_OK && st != RGNN_ERR_CUDA) return st;   // argument errors are the caller's
      }
    }
  }
  if (!launched) RGNN_RETURN_IF_ERROR(enqueue());
  RGNN_CUDA_CHECK(cudaStreamSynchronize(ms));
  const int32_t flag_host = *hs->flag_pinned;
  return flag_host != 0 ? flag_host : RGNN_OK;
}

}  // extern "C"
