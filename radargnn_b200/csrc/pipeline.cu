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

// The pipelined host-buffer path (rgnn_pipeline_submit_host / rgnn_pipeline_wait_host): one upload, one compute
// and one download stream per device, shared by every slot, so that consecutive calls form a software pipeline --
// the upload of call i + 1 and the download of call i - 1 travel while call i computes (PCIe is full duplex), and
// the compute of consecutive calls stays in submission order (BatchNorm running statistics are updated in that
// order).  The two halves of the compute are replayed from per-slot CUDA graphs; the copies and the events that
// order them are enqueued directly.
constexpr int kHostSlots = RGNN_HOST_SLOTS;
struct HostSlot {
  cudaEvent_t pos_up, x0_up, graph_done, layers_done, done;
  int32_t* flag_pinned;
  uint64_t graph_key;
  cudaGraphExec_t build_exec, layers_exec;
  bool busy, has_work;
};
struct HostPipelineState {
  cudaStream_t upload, compute, download;
  cudaEvent_t start;
  HostSlot slot[kHostSlots];
  bool ok;
};
HostPipelineState* host_pipeline_state(bool create = true) {
  static HostPipelineState per_device[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostPipelineState& h = per_device[dev];
  if (!h.ok && !create) return nullptr;
  if (!h.ok) {
    if (cudaStreamCreateWithFlags(&h.upload, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h.compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h.download, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (HostSlot& s : h.slot) {
      cudaEvent_t* evs[] = {&s.pos_up, &s.x0_up, &s.graph_done, &s.layers_done, &s.done};
      for (cudaEvent_t* e : evs)
        if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaMallocHost(reinterpret_cast<void**>(&s.flag_pinned), 64) != cudaSuccess) return nullptr;
      s.graph_key = 0;
      s.build_exec = s.layers_exec = nullptr;
      s.busy = s.has_work = false;
    }
    h.ok = true;
  }
  return &h;
}
std::mutex& host_path_mutex() {
  static std::mutex m;
  return m;
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
// recorded once edge_index / edge_attr are final -- the host entry point overlaps its copies with them.
// phases: kPhaseBuild (neighbour search .. CSC view) | kPhaseLayers (the conv stack); the pipelined host
// entry point enqueues the two halves separately so that its copies can be ordered between them.
enum { kPhaseBuild = 1, kPhaseLayers = 2, kPhaseAll = 3 };
static int pipeline_forward_impl(const rgnn_pipeline_desc* desc, const float* pos, const float* vel, const float* x0,
                                 const int64_t* frame_ptr_host, int32_t n_frames, int64_t* edge_index, int64_t n_edges,
                                 float* edge_attr, float* h, int32_t* error_flag, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream, cudaEvent_t x0_ready, cudaEvent_t graph_done,
                                 int phases = kPhaseAll) {
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
  if (phases & kPhaseBuild) RGNN_CUDA_CHECK(cudaMemsetAsync(error_flag, 0, sizeof(int32_t), stream));
  if (n == 0) return RGNN_OK;

  if (phases & kPhaseBuild) {
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
  }   // kPhaseBuild

  if (graph_done != nullptr) RGNN_CUDA_CHECK(cudaEventRecord(graph_done, stream));
  if (x0_ready != nullptr) RGNN_CUDA_CHECK(cudaStreamWaitEvent(stream, x0_ready, 0));
  if (!(phases & kPhaseLayers)) return RGNN_OK;

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

}  // extern "C"

namespace {

// Device staging area of one host-buffer call, carved from the caller's workspace.
struct HostCall {
  int32_t de, c_last;
  int64_t n;
  float *pos, *vel, *x0;
  int64_t* edge_index;
  float *edge_attr, *h;
  int32_t* flag;
  char* inner_ws;
  size_t inner;
};

int carve_host_call(const rgnn_pipeline_desc* desc, const float* pos_host, const float* vel_host, const float* x0_host,
                    int32_t c0, const int64_t* frame_ptr_host, int32_t n_frames, int64_t n_edges, void* workspace,
                    size_t workspace_bytes, HostCall* c) {
  int32_t c_max = 0;
  RGNN_RETURN_IF_ERROR(validate(desc, &c->de, &c_max));
  if (frame_ptr_host == nullptr || n_frames < 1 || n_edges < 0 || c0 != desc->layers[0].in_channels) return RGNN_ERR_INVALID_ARGUMENT;
  const int64_t n = frame_ptr_host[n_frames];
  if (n < 0) return RGNN_ERR_INVALID_ARGUMENT;
  if (n > 0 && (pos_host == nullptr || vel_host == nullptr || x0_host == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  const size_t need = rgnn_pipeline_host_workspace_bytes(desc, n, n_frames, n_edges, c0);
  if (workspace == nullptr || need == 0 || workspace_bytes < need) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  c->n = n;
  c->c_last = desc->layers[desc->n_layers - 1].out_channels;
  Arena a(workspace, workspace_bytes);
  c->pos = a.take<float>(static_cast<size_t>(n) * 2);
  c->vel = a.take<float>(static_cast<size_t>(n) * 2);
  c->x0 = a.take<float>(static_cast<size_t>(n) * c0);
  c->edge_index = a.take<int64_t>(static_cast<size_t>(n_edges) * 2);
  c->edge_attr = a.take<float>(static_cast<size_t>(n_edges) * c->de);
  c->h = a.take<float>(static_cast<size_t>(n) * c->c_last);
  c->flag = a.take<int32_t>(64);
  c->inner = rgnn_pipeline_workspace_bytes(desc, n, n_frames, n_edges);
  c->inner_ws = a.take<char>(c->inner);
  return a.overflow ? RGNN_ERR_WORKSPACE_TOO_SMALL : RGNN_OK;
}

// replay key: every argument the enqueued work depends on (pointers are compared, not their targets --
// buffers are read when the graph runs)
uint64_t host_call_key(const rgnn_pipeline_desc* desc, const void* pos_host, const void* vel_host, const void* x0_host,
                       int32_t c0, const int64_t* frame_ptr_host, int32_t n_frames, const void* edge_index_host,
                       int64_t n_edges, const void* edge_attr_host, const void* h_host, const void* workspace,
                       size_t workspace_bytes) {
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
  return key == 0 ? 1 : key;
}

bool host_graphs_enabled() {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("RGNN_HOST_GRAPH"); enabled = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return enabled != 0;
}

// Captures what `enqueue` puts on `stream` into *exec (replacing what it held).  Returns RGNN_OK with *exec == nullptr
// when the capture itself failed (the caller then launches eagerly), the status of `enqueue` when that is the
// caller's error.
template <typename F>
int capture_graph(cudaStream_t stream, F&& enqueue, cudaGraphExec_t* exec) {
  if (*exec != nullptr) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
  if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { (void)cudaGetLastError(); return RGNN_OK; }
  const int st = enqueue();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
  if (!(st == RGNN_OK && ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(exec, graph, 0) == cudaSuccess)) {
    *exec = nullptr;
    (void)cudaGetLastError();
  }
  if (graph != nullptr) cudaGraphDestroy(graph);
  return (st != RGNN_OK && st != RGNN_ERR_CUDA) ? st : RGNN_OK;
}

}  // namespace

extern "C" {

int rgnn_pipeline_forward_host(const rgnn_pipeline_desc* desc, const float* pos_host, const float* vel_host,
                               const float* x0_host, int32_t c0, const int64_t* frame_ptr_host, int32_t n_frames,
                               int64_t* edge_index_host, int64_t n_edges, float* edge_attr_host, float* h_host,
                               void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HostCall c;
  RGNN_RETURN_IF_ERROR(carve_host_call(desc, pos_host, vel_host, x0_host, c0, frame_ptr_host, n_frames, n_edges, workspace,
                                       workspace_bytes, &c));
  const int64_t n = c.n;
  const int32_t de = c.de;
  // Copies overlap the compute: x0 (the bulk of the input) travels on a side stream while the main stream
  // builds the graph, and edge_index / edge_attr travel back while the layers run.  The whole sequence
  // (copies, ~40 kernels, cross-stream events) is captured into a CUDA graph the first time it is seen and
  // replayed while the arguments stay the same: the eager launches cost ~0.1 ms of host time per call.
  // one call at a time per process: the per-device streams, events and the cached graph are shared state
  std::lock_guard<std::mutex> host_path_lock(host_path_mutex());
  HostPathStreams* hs = host_path_streams();
  if (hs == nullptr) return RGNN_ERR_CUDA;
  cudaStream_t ms = hs->main;
  RGNN_CUDA_CHECK(cudaEventRecord(hs->start, stream));   // order our stream after the caller's
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ms, hs->start, 0));
  {
    // ... and after the calls still in flight on the pipelined path (running statistics stay in call order)
    HostPipelineState* ps = host_pipeline_state(false);
    if (ps != nullptr) {
      for (HostSlot& s : ps->slot)
        if (s.busy && s.has_work) RGNN_CUDA_CHECK(cudaStreamWaitEvent(ms, s.done, 0));
    }
  }

  auto enqueue = [&]() -> int {
    if (n > 0) {
      RGNN_CUDA_CHECK(cudaMemcpyAsync(c.pos, pos_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ms));
      RGNN_CUDA_CHECK(cudaMemcpyAsync(c.vel, vel_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ms));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->x0_ready, ms));                 // fork the side stream
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(hs->copy, hs->x0_ready, 0));
      RGNN_CUDA_CHECK(cudaMemcpyAsync(c.x0, x0_host, sizeof(float) * n * c0, cudaMemcpyHostToDevice, hs->copy));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->x0_ready, hs->copy));
    }
    RGNN_RETURN_IF_ERROR(pipeline_forward_impl(desc, c.pos, c.vel, c.x0, frame_ptr_host, n_frames, c.edge_index, n_edges,
                                               c.edge_attr, c.h, c.flag, c.inner_ws, c.inner, ms, n > 0 ? hs->x0_ready : nullptr,
                                               n > 0 ? hs->graph_done : nullptr));
    if (n > 0) {
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(hs->copy, hs->graph_done, 0));
      if (edge_index_host != nullptr && n_edges > 0)
        RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_index_host, c.edge_index, sizeof(int64_t) * n_edges * 2, cudaMemcpyDeviceToHost, hs->copy));
      if (edge_attr_host != nullptr && n_edges > 0 && de > 0)
        RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_attr_host, c.edge_attr, sizeof(float) * n_edges * de, cudaMemcpyDeviceToHost, hs->copy));
      RGNN_CUDA_CHECK(cudaEventRecord(hs->copies_done, hs->copy));
      RGNN_CUDA_CHECK(cudaStreamWaitEvent(ms, hs->copies_done, 0));        // join
    }
    if (h_host != nullptr && n > 0)
      RGNN_CUDA_CHECK(cudaMemcpyAsync(h_host, c.h, sizeof(float) * n * c.c_last, cudaMemcpyDeviceToHost, ms));
    RGNN_CUDA_CHECK(cudaMemcpyAsync(hs->flag_pinned, c.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, ms));
    return RGNN_OK;
  };

  const uint64_t key = host_call_key(desc, pos_host, vel_host, x0_host, c0, frame_ptr_host, n_frames, edge_index_host, n_edges,
                                     edge_attr_host, h_host, workspace, workspace_bytes);
  bool launched = false;
  if (host_graphs_enabled() && n > 0 && desc->search == 0) {
    if (hs->graph_exec == nullptr || hs->graph_key != key) {
      hs->graph_key = 0;
      RGNN_RETURN_IF_ERROR(capture_graph(ms, enqueue, &hs->graph_exec));
      if (hs->graph_exec != nullptr) hs->graph_key = key;
    }
    if (hs->graph_exec != nullptr) launched = cudaGraphLaunch(hs->graph_exec, ms) == cudaSuccess;
  }
  if (!launched) RGNN_RETURN_IF_ERROR(enqueue());
  RGNN_CUDA_CHECK(cudaStreamSynchronize(ms));
  const int32_t flag_host = *hs->flag_pinned;
  return flag_host != 0 ? flag_host : RGNN_OK;
}

int rgnn_pipeline_submit_host(int32_t slot, const rgnn_pipeline_desc* desc, const float* pos_host, const float* vel_host,
                              const float* x0_host, int32_t c0, const int64_t* frame_ptr_host, int32_t n_frames,
                              int64_t* edge_index_host, int64_t n_edges, float* edge_attr_host, float* h_host,
                              void* workspace, size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (slot < 0 || slot >= kHostSlots) return RGNN_ERR_INVALID_ARGUMENT;
  HostCall c;
  RGNN_RETURN_IF_ERROR(carve_host_call(desc, pos_host, vel_host, x0_host, c0, frame_ptr_host, n_frames, n_edges, workspace,
                                       workspace_bytes, &c));
  const int64_t n = c.n;
  const int32_t de = c.de;
  std::lock_guard<std::mutex> host_path_lock(host_path_mutex());
  HostPipelineState* ps = host_pipeline_state();
  if (ps == nullptr) return RGNN_ERR_CUDA;
  HostSlot& s = ps->slot[slot];
  if (s.busy) return RGNN_ERR_INVALID_ARGUMENT;   // rgnn_pipeline_wait_host(slot) first
  *s.flag_pinned = 0;
  if (n == 0) { s.busy = true; s.has_work = false; return RGNN_OK; }

  auto build = [&]() -> int {
    return pipeline_forward_impl(desc, c.pos, c.vel, c.x0, frame_ptr_host, n_frames, c.edge_index, n_edges, c.edge_attr, c.h,
                                 c.flag, c.inner_ws, c.inner, ps->compute, nullptr, nullptr, kPhaseBuild);
  };
  auto layers = [&]() -> int {
    return pipeline_forward_impl(desc, c.pos, c.vel, c.x0, frame_ptr_host, n_frames, c.edge_index, n_edges, c.edge_attr, c.h,
                                 c.flag, c.inner_ws, c.inner, ps->compute, nullptr, nullptr, kPhaseLayers);
  };
  const bool use_graphs = host_graphs_enabled() && desc->search == 0;
  if (use_graphs) {
    const uint64_t key = host_call_key(desc, pos_host, vel_host, x0_host, c0, frame_ptr_host, n_frames, edge_index_host,
                                       n_edges, edge_attr_host, h_host, workspace, workspace_bytes);
    if (s.graph_key != key || s.build_exec == nullptr || s.layers_exec == nullptr) {
      s.graph_key = 0;
      RGNN_RETURN_IF_ERROR(capture_graph(ps->compute, build, &s.build_exec));
      RGNN_RETURN_IF_ERROR(capture_graph(ps->compute, layers, &s.layers_exec));
      if (s.build_exec != nullptr && s.layers_exec != nullptr) s.graph_key = key;
    }
  }
  const bool replay = use_graphs && s.graph_key != 0;

  // upload: positions first (the neighbour search needs only them), the node features behind
  RGNN_CUDA_CHECK(cudaEventRecord(ps->start, stream));   // order the call after the caller's stream
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ps->upload, ps->start, 0));
  RGNN_CUDA_CHECK(cudaMemcpyAsync(c.pos, pos_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ps->upload));
  RGNN_CUDA_CHECK(cudaMemcpyAsync(c.vel, vel_host, sizeof(float) * n * 2, cudaMemcpyHostToDevice, ps->upload));
  RGNN_CUDA_CHECK(cudaEventRecord(s.pos_up, ps->upload));
  RGNN_CUDA_CHECK(cudaMemcpyAsync(c.x0, x0_host, sizeof(float) * n * c0, cudaMemcpyHostToDevice, ps->upload));
  RGNN_CUDA_CHECK(cudaEventRecord(s.x0_up, ps->upload));
  // compute: graph build, then the layers once the node features have arrived
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ps->compute, s.pos_up, 0));
  if (replay) RGNN_CUDA_CHECK(cudaGraphLaunch(s.build_exec, ps->compute));
  else RGNN_RETURN_IF_ERROR(build());
  RGNN_CUDA_CHECK(cudaEventRecord(s.graph_done, ps->compute));
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ps->compute, s.x0_up, 0));
  if (replay) RGNN_CUDA_CHECK(cudaGraphLaunch(s.layers_exec, ps->compute));
  else RGNN_RETURN_IF_ERROR(layers());
  RGNN_CUDA_CHECK(cudaEventRecord(s.layers_done, ps->compute));
  // download: the graph while the layers run, the embeddings and the error flag behind them
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ps->download, s.graph_done, 0));
  if (edge_index_host != nullptr && n_edges > 0)
    RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_index_host, c.edge_index, sizeof(int64_t) * n_edges * 2, cudaMemcpyDeviceToHost, ps->download));
  if (edge_attr_host != nullptr && n_edges > 0 && de > 0)
    RGNN_CUDA_CHECK(cudaMemcpyAsync(edge_attr_host, c.edge_attr, sizeof(float) * n_edges * de, cudaMemcpyDeviceToHost, ps->download));
  RGNN_CUDA_CHECK(cudaStreamWaitEvent(ps->download, s.layers_done, 0));
  if (h_host != nullptr)
    RGNN_CUDA_CHECK(cudaMemcpyAsync(h_host, c.h, sizeof(float) * n * c.c_last, cudaMemcpyDeviceToHost, ps->download));
  RGNN_CUDA_CHECK(cudaMemcpyAsync(s.flag_pinned, c.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, ps->download));
  RGNN_CUDA_CHECK(cudaEventRecord(s.done, ps->download));
  s.busy = true;
  s.has_work = true;
  return RGNN_OK;
}

int rgnn_pipeline_wait_host(int32_t slot) {
  if (slot < 0 || slot >= kHostSlots) return RGNN_ERR_INVALID_ARGUMENT;
  cudaEvent_t done = nullptr;
  HostSlot* s = nullptr;
  {
    std::lock_guard<std::mutex> host_path_lock(host_path_mutex());
    HostPipelineState* ps = host_pipeline_state();
    if (ps == nullptr) return RGNN_ERR_CUDA;
    s = &ps->slot[slot];
    if (!s->busy) return RGNN_ERR_INVALID_ARGUMENT;   // nothing submitted on this slot
    if (s->has_work) done = s->done;
  }
  if (done != nullptr) RGNN_CUDA_CHECK(cudaEventSynchronize(done));   // outside the lock: other slots stay usable
  std::lock_guard<std::mutex> host_path_lock(host_path_mutex());
  const int32_t flag_host = *s->flag_pinned;
  s->busy = s->has_work = false;
  return flag_host != 0 ? flag_host : RGNN_OK;
}

}  // extern "C"
