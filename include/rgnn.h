/* rgnn.h -- C ABI of librgnn_b200.so: B200 (sm_100a) kernels for RadarGNN's hot path.
 *
 * The reference (TUMFTM/RadarGNN) has no FFI of its own: its boundary for this path is
 * a Python API backed by scikit-learn (KD-tree) and PyG / torch_scatter / cuBLAS.  The
 * entry points below are what a binding for that path attaches to; each one names the
 * reference interface it replaces (paths relative to the reference's
 * src/gnnradarobjectdetection/).  radargnn_b200/_lib.py is the ctypes binding used by
 * the Python mirror of the reference classes; INTEGRATION.md shows the stub a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - pointers are DEVICE pointers unless the parameter name ends in `_host`;
 *   - every function returns an rgnn_status (0 = ok), never throws, never allocates
 *     device memory: scratch comes from the caller through `workspace`
 *     (size from the matching *_workspace_bytes query; 256-byte aligned);
 *   - all work is enqueued on `stream` (a cudaStream_t); functions that must hand a
 *     count back to the host synchronise that stream and say so;
 *   - matrices are dense row-major; edge_index is int64 [2, E] exactly as PyG wants it:
 *     row 0 = the query point i (reference E[:,0]), row 1 = its neighbour j (E[:,1]).
 */
#ifndef RGNN_H_
#define RGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGNN_ABI_VERSION 4

typedef void* rgnn_stream_t; /* cudaStream_t */

typedef enum rgnn_status {
  RGNN_OK = 0,
  RGNN_ERR_INVALID_ARGUMENT = 1,
  RGNN_ERR_K_NOT_SMALLER_THAN_N = 2, /* sklearn: ValueError "Expected n_neighbors < n_samples_fit" */
  RGNN_ERR_WORKSPACE_TOO_SMALL = 3,
  RGNN_ERR_CUDA = 4,
  RGNN_ERR_DOT_PRODUCT = 5,     /* graph_constructor/features.py:56 "Error in dot product calculation" */
  RGNN_ERR_INVALID_FEATURE = 6, /* graph_constructor/graph.py:220 "Invalid feature specified" */
  RGNN_ERR_UNSUPPORTED = 7,
  RGNN_ERR_NO_DEVICE = 8,
  RGNN_ERR_NON_FINITE_INPUT = 9,   /* sklearn check_array: ValueError "Input contains NaN" / "infinity" */
  RGNN_ERR_INDEX_OUT_OF_RANGE = 10 /* edge_index entry outside [0, N): PyG's gather raises an index error */
} rgnn_status;

typedef enum rgnn_dtype { RGNN_F32 = 0, RGNN_F64 = 1 } rgnn_dtype;
typedef enum rgnn_edge_mode { RGNN_DIRECTED = 0, RGNN_UNDIRECTED = 1 } rgnn_edge_mode;

/* graph.py:139-223 feature names, in the order the caller lists them */
typedef enum rgnn_edge_feature {
  RGNN_EF_POINT_PAIR_FEATURES = 0,         /* 4 columns */
  RGNN_EF_SPATIAL_EUCLIDEAN_DISTANCE = 1,  /* 1 */
  RGNN_EF_VELOCITY_EUCLIDEAN_DISTANCE = 2, /* 1 */
  RGNN_EF_RELATIVE_POSITION = 3,           /* 2 */
  RGNN_EF_RELATIVE_VELOCITY = 4            /* 2 */
} rgnn_edge_feature;

/* graph.py:225-275 feature names */
typedef enum rgnn_node_feature {
  RGNN_NF_RCS = 0,
  RGNN_NF_TIME_INDEX = 1,
  RGNN_NF_DEGREE = 2,
  RGNN_NF_VELOCITY_VECTOR_LENGTH = 3,
  RGNN_NF_VELOCITY_VECTOR = 4,     /* 2 columns */
  RGNN_NF_SPATIAL_COORDINATES = 5  /* 2 columns */
} rgnn_node_feature;

typedef enum rgnn_aggr { RGNN_AGGR_MAX = 0, RGNN_AGGR_ADD = 1, RGNN_AGGR_MEAN = 2, RGNN_AGGR_MIN = 3 } rgnn_aggr;
typedef enum rgnn_conv_type { RGNN_CONV_MPNN = 0, RGNN_CONV_RADAR_POINT_GNN = 1 } rgnn_conv_type;

#define RGNN_MAX_MLP_LAYERS 8
#define RGNN_MAX_EDGE_FEATURES 8
#define RGNN_MAX_NODE_FEATURES 8
#define RGNN_MAX_K 64

/* ------------------------------------------------------------------------------------ */
/* library                                                                              */
/* ------------------------------------------------------------------------------------ */
int rgnn_abi_version(void);
const char* rgnn_status_string(int status);
/* Last CUDA error text seen by this thread ("" if none). */
const char* rgnn_last_cuda_error(void);
/* SM count / compute capability of the current device; RGNN_ERR_NO_DEVICE without a GPU. */
int rgnn_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t rgnn_kernel_launch_count(void);

/* Per-kernel device timing for bench.py's roofline line.  While enabled, every kernel family the
 * library launches is bracketed by CUDA events on its launch stream.  collect() synchronises the
 * recorded events, adds their durations to per-name totals and returns the number of names;
 * entry(i) reads one total (name is owned by the library, valid until reset). */
void rgnn_profile_enable(int32_t on);
int32_t rgnn_profile_collect(void);
int rgnn_profile_entry(int32_t index, const char** name, double* total_ms, int64_t* count);
void rgnn_profile_reset(void);

/* ------------------------------------------------------------------------------------ */
/* neighbour search  (replaces Graph.build, graph_constructor/graph.py:32-82, i.e.       */
/* sklearn kneighbors_graph / radius_neighbors_graph + nonzero())                        */
/* ------------------------------------------------------------------------------------ */
/* A batch of frames is one array of points plus frame_ptr_host[n_frames+1]; edges never
 * cross frames (disjoint union, utils/data_handling.py:30); node ids are global.
 * `basis` is the search space [n_points, dims] (dims = 2 for "X", 4 for "XV";
 * preprocessor/radarscenes/dataset_creation.py:203-206), f32 or f64; distances are
 * always evaluated in fp64 exactly as sklearn does (sum over dims of (a-b)*(a-b)). */
size_t rgnn_graph_workspace_bytes(int64_t n_points, int32_t n_frames);

/* Number of edges a k-NN build emits: sum over frames with n_f > 1 of n_f * k.
 * Returns -1 and sets *status = RGNN_ERR_K_NOT_SMALLER_THAN_N if some frame has 1 < n_f <= k. */
int64_t rgnn_knn_edge_count(const int64_t* frame_ptr_host, int32_t n_frames, int32_t k, int* status);

/* k-NN graph (graph.py:52-66).  Row i*k..i*k+k-1 (per frame) lists the k nearest
 * neighbours of point i by ascending (fp64 squared distance, index); self excluded by
 * index; frames with n_f <= 1 emit nothing (graph.py:45).
 * *error_flag (device int32, may be NULL; zeroed by the call) becomes RGNN_ERR_NON_FINITE_INPUT when a
 * coordinate is NaN or infinite (sklearn's check_array raises "Input contains NaN"); the rows of such
 * points are then meaningless.  The radius count call reports the same condition as its return value. */
int rgnn_graph_build_knn(const void* basis, int32_t basis_dtype, int32_t dims,
                         const int64_t* frame_ptr_host, int32_t n_frames, int32_t k,
                         int64_t* edge_index, int64_t n_edges, int32_t* error_flag,
                         void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* Radius graph (graph.py:68-82), two calls sharing the workspace:
 *   count: bins the points, counts  j != i with sum_d (a_d-b_d)^2 <= r*r  (fp64,
 *          inclusive), synchronises `stream`, returns E in *n_edges_host;
 *   fill:  writes edge_index [2, E]; rows ascend in i, columns ascend in j (canonical
 *          order; sklearn's KD-tree order inside a row is not reproduced). */
int rgnn_graph_build_radius_count(const void* basis, int32_t basis_dtype, int32_t dims,
                                  const int64_t* frame_ptr_host, int32_t n_frames, double r,
                                  int64_t* n_edges_host,
                                  void* workspace, size_t workspace_bytes, rgnn_stream_t stream);
int rgnn_graph_build_radius_fill(const void* basis, int32_t basis_dtype, int32_t dims,
                                 const int64_t* frame_ptr_host, int32_t n_frames, double r,
                                 int64_t* edge_index, int64_t n_edges,
                                 void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* ------------------------------------------------------------------------------------ */
/* edge / node features                                                                  */
/* ------------------------------------------------------------------------------------ */
/* Columns a feature list produces (graph.py:157-166); -1 for an unknown feature id. */
int32_t rgnn_edge_feature_width(const int32_t* features_host, int32_t n_features);

/* GeometricGraph.extract_node_pair_features (graph.py:139-223) +
 * get_En_equivariant_point_pair_metrics (features.py:6-122): fp64 arithmetic per edge,
 * written as `out_dtype` [E, De].  pos [N, pos_dims], vel [N, vel_dims] of `in_dtype`.
 * *error_flag (device int32, zeroed by the call) becomes RGNN_ERR_DOT_PRODUCT if a
 * normalised dot product exceeds 1 by 1e-3 or more (features.py:49-56). */
int rgnn_edge_features(const void* pos, const void* vel, int32_t in_dtype,
                       int32_t pos_dims, int32_t vel_dims, int64_t n_points,
                       const int64_t* edge_index, int64_t n_edges,
                       const int32_t* features_host, int32_t n_features, int32_t edge_mode,
                       void* edge_attr, int32_t out_dtype, int32_t* error_flag,
                       rgnn_stream_t stream);

/* Graph.get_degree (graph.py:93-96): degree of the *undirected* graph networkx builds
 * from the adjacency matrix, |N_out(i) U N_in(i)|.  edge_index rows must be grouped by
 * ascending source (what the builders above emit). */
int rgnn_undirected_degree(const int64_t* edge_index, int64_t n_edges, int64_t n_points,
                           int32_t* degree, rgnn_stream_t stream);

/* GeometricGraph.extract_single_node_features (graph.py:225-275): concatenates the
 * listed columns into out [N, Fn] (`out_dtype`).  rcs / time_index [N] f64, degree [N]
 * int32, pos / vel [N, 2] f64; unused inputs may be NULL. */
int32_t rgnn_node_feature_width(const int32_t* features_host, int32_t n_features);
int rgnn_node_features(const double* rcs, const double* time_index, const int32_t* degree,
                       const double* pos, const double* vel, int64_t n_points,
                       const int32_t* features_host, int32_t n_features,
                       void* out, int32_t out_dtype, rgnn_stream_t stream);

/* ------------------------------------------------------------------------------------ */
/* message passing  (replaces PyG MessagePassing.propagate + torch_scatter + addmm under */
/* MPNNConv / RadarPointGNNConv, gnn/mpnn_layers.py:86-101, 171-184)                     */
/* ------------------------------------------------------------------------------------ */
/* Target-major (CSC) view of edge_index: csc_ptr [N+1], and for every slot the source
 * node (csc_src) and the position of that edge in edge_index / edge_attr (csc_eid).
 * Messages are reduced at edge_index[1] (flow = source_to_target). */
size_t rgnn_csc_workspace_bytes(int64_t n_nodes, int64_t n_edges);
/* *error_flag (device int32, may be NULL; zeroed by the call) becomes RGNN_ERR_INDEX_OUT_OF_RANGE when
 * edge_index holds an id outside [0, n_nodes) -- PyG's gather raises an index error there; such edges
 * are left out of the view instead of being used as addresses. */
int rgnn_csc_build(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes,
                   int32_t* csc_ptr, int32_t* csc_src, int32_t* csc_eid, int32_t* error_flag,
                   void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* One graph-convolution layer.  Weights are PyG `Linear` parameters: [out, in] row-major
 * fp32, read at call time (the reference's tests swap them after construction).
 *   MPNN:            m_e = pre_mlp([x_t ; x_s ; e'])     P = 2C + De  (3C with edge encoder)
 *                    h_n = post_mlp([x_n ; aggr_e m_e])
 *   RADAR_POINT_GNN: m_e = pre_mlp([x_s ; e])            P = C + De
 *                    h_n = post_mlp([x_n ; aggr_e m_e]) + x_n
 * pre_mlp  = Linear(P,P) [ReLU Linear(P,P)]*(pre_layers-1)
 * post_mlp = Linear(P+C,C_out) [ReLU Linear(C_out,C_out)]*(post_layers-1)
 * Nodes without incoming edge aggregate to 0 for every `aggr`. */
typedef struct rgnn_conv_desc {
  int32_t conv_type;   /* rgnn_conv_type */
  int32_t aggr;        /* rgnn_aggr */
  int32_t in_channels; /* C */
  int32_t out_channels;
  int32_t edge_dim;    /* De */
  int32_t pre_layers;
  int32_t post_layers;
  int32_t use_edge_encoder;
  const float* edge_encoder_weight; /* [C, De] or NULL */
  const float* edge_encoder_bias;   /* [C] or NULL */
  const float* pre_weight[RGNN_MAX_MLP_LAYERS];
  const float* pre_bias[RGNN_MAX_MLP_LAYERS];
  const float* post_weight[RGNN_MAX_MLP_LAYERS];
  const float* post_bias[RGNN_MAX_MLP_LAYERS];
  /* Optional (may be NULL): tensor-core weight images made by rgnn_conv_pack_weights for exactly these
   * weights.  NULL = the images are rebuilt inside every forward call (weights are then read at call
   * time, as the reference's tests require); callers whose weights are unchanged between calls pack once. */
  const void* packed_weights;
} rgnn_conv_desc;

/* Bytes of the packed images of a layer (0 when the layer does not use the tensor-core path), and the
 * packing itself: hi/lo TF32 split of W_s and of the update weights with W_m W_t folded in, laid out
 * as 128-byte-swizzled K-major panels (radargnn_b200/csrc/node_gemm.cu). */
size_t rgnn_conv_packed_bytes(const rgnn_conv_desc* desc);
int rgnn_conv_pack_weights(const rgnn_conv_desc* desc, void* packed, size_t packed_bytes, rgnn_stream_t stream);

size_t rgnn_conv_workspace_bytes(const rgnn_conv_desc* desc, int64_t n_nodes, int64_t n_edges);
int rgnn_conv_forward(const rgnn_conv_desc* desc, const float* x, int64_t n_nodes,
                      const int32_t* csc_ptr, const int32_t* csc_src, const int32_t* csc_eid,
                      const float* edge_attr, int64_t n_edges, float* out,
                      void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* PyG BatchNorm (= BatchNorm1d) in TRAINING mode followed by optional ReLU
 * (gnn/gnn_models.py:126-128): batch statistics over all n rows, biased variance for
 * the normalisation; running_mean / running_var (unbiased) updated with `momentum`
 * when non-NULL.  weight / bias may be NULL (affine=False). */
size_t rgnn_batchnorm_workspace_bytes(int64_t n, int32_t channels);
int rgnn_batchnorm_relu_forward(const float* x, int64_t n, int32_t channels,
                                const float* weight, const float* bias, float eps, float momentum,
                                float* running_mean, float* running_var, int32_t apply_relu,
                                float* out, void* workspace, size_t workspace_bytes,
                                rgnn_stream_t stream);

/* Per-channel affine map (+ optional ReLU): out = relu?((x - mean[c]) * scale[c] + beta[c]).
 * This is BatchNorm in EVAL mode with mean = running_mean, scale = weight / sqrt(running_var + eps),
 * beta = bias (the reference itself never leaves training mode). */
int rgnn_affine_relu_forward(const float* x, int64_t n, int32_t channels, const float* mean,
                             const float* scale, const float* beta, int32_t apply_relu, float* out,
                             rgnn_stream_t stream);

/* Deterministic sum of `count` floats into *result (DEVICE double): the per-rank loss partial that
 * the data-parallel step all-reduces.  workspace: rgnn_sum_workspace_bytes(). */
size_t rgnn_sum_workspace_bytes(void);
int rgnn_sum_f32(const float* x, int64_t count, double* result, void* workspace, size_t workspace_bytes,
                 rgnn_stream_t stream);

/* y[n, out] = act(x)[n, in] . W[out, in]^T + b  -- PyG `Linear` (embedding MLPs and heads,
 * gnn/gnn_models.py:137-178).  relu_input applies ReLU to x on load. */
int rgnn_linear_forward(const float* x, int64_t n, int32_t in_features, const float* weight,
                        const float* bias, int32_t out_features, int32_t relu_input, float* y,
                        rgnn_stream_t stream);

/* ------------------------------------------------------------------------------------ */
/* callers either side of the path                               */
/* ------------------------------------------------------------------------------------ */
/* Loss of the detection heads (gnn/trainer.py:184-206 training, 276-298 validation):
 *   loss_cls = torch.nn.CrossEntropyLoss(weight = class_weight)(cls, y[:, 0].long())  (weighted mean; label -100 ignored)
 *   loss_bb  = mean over the nodes with y[:, 0] != bg_index of torch.nn.HuberLoss(delta)(y[i, 1:], bb[i])
 *              (0 without such a node; nan_to_zero: a NaN box loss counts as 0, trainer.py:206-216)
 *   loss     = cls_loss_weight * loss_cls + bb_loss_weight * loss_bb
 * cls [N, n_classes], bb [N, n_box], y [N, ldy >= 1 + n_box] (label stored as float, like graph_batch.y).
 * out (DEVICE double[5]) = loss, loss_cls, loss_bb, number of foreground nodes, number of labels outside
 * [0, n_classes) (torch raises an index error for those; they are left out of the sums).  Replaces the
 * reference's per-node Python loop by one deterministic reduction (fp64, fixed order). */
size_t rgnn_detection_loss_workspace_bytes(void);
int rgnn_detection_loss(const float* cls, int32_t n_classes, const float* bb, int32_t n_box, const float* y, int64_t ldy,
                        int64_t n_nodes, const float* class_weight, int32_t bg_index, float cls_loss_weight,
                        float bb_loss_weight, float huber_delta, int32_t nan_to_zero, double* out, void* workspace,
                        size_t workspace_bytes, rgnn_stream_t stream);

/* Greedy non-maximum suppression (postprocessor/postprocessing.py:336-435).
 *   rotated == 0: boxes f32 [n, 4] (x1, y1, x2, y2), scores f32 [n]: torchvision.ops.nms (:408);
 *   rotated == 1: boxes f64 [n, 5] (cx, cy, w, h, angle in degrees), scores f64 [n]: detectron2 nms_rotated (:370).
 * A box is dropped when its IoU with a kept box of higher score (ties: lower index first) exceeds
 * iou_threshold.  shift_negative reproduces the reference moving all boxes by |min| + 100 when a coordinate
 * is negative (:361-365, :400-404).  box_frame (int32 [n], may be NULL): boxes of different frames never
 * suppress each other.  keep (int64 [n]) receives the kept indices ordered by (frame, descending score),
 * *keep_count (device int32) their number, keep_flag (uint8 [n]) one flag per box.  n <= 65536. */
size_t rgnn_nms_workspace_bytes(int64_t n_boxes);
int rgnn_nms(const void* boxes, int32_t rotated, const void* scores, const int32_t* box_frame, int64_t n_boxes,
             double iou_threshold, int32_t shift_negative, int64_t* keep, int32_t* keep_count, uint8_t* keep_flag,
             void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* Nearest neighbour of every point inside its frame -- kneighbors_graph(X, 1, include_self=False) followed by
 * X[np.where(A == 1)[1]] (preprocessor/radarscenes/dataset_creation.py:314-318, nuscenes/conversion.py:133-137,
 * postprocessor/postprocessing.py:233-237, 468-472).  nn_index int64 [N] (-1 for the point of a one-point
 * frame), nn_points [N, dims] of the basis dtype (row untouched where there is no neighbour); either may be NULL. */
size_t rgnn_nearest_neighbor_workspace_bytes(int64_t n_points, int32_t n_frames);
int rgnn_nearest_neighbor(const void* basis, int32_t basis_dtype, int32_t dims, const int64_t* frame_ptr_host, int32_t n_frames,
                          int64_t* nn_index, void* nn_points, void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* time_index node feature (dataset_creation.py:214-223): position of the point's timestamp among the sorted
 * distinct timestamps of its frame, as double.  frame_ptr (device) and frame_ptr_host hold the same
 * [n_frames + 1] offsets.  Frames of up to 8192 points are ranked in shared memory (workspace may be NULL);
 * a batch with a longer frame takes the global-memory path and needs rgnn_time_index_workspace_bytes(N). */
size_t rgnn_time_index_workspace_bytes(int64_t n_points);
int rgnn_time_index(const double* timestamp, const int64_t* frame_ptr, const int64_t* frame_ptr_host, int32_t n_frames,
                    double* time_index, void* workspace, size_t workspace_bytes, rgnn_stream_t stream);

/* Node-id offsets of the disjoint-union collate (utils/data_handling.py:30, PyG Batch.from_data_list):
 * edge_index [2, E] holds frame-local ids, columns edge_ptr[f] .. edge_ptr[f+1] belong to frame f (device
 * int64 [n_frames + 1] tables); both rows get node_ptr[f] added in place.  batch_of_node (int64 [n_nodes], may be
 * NULL) receives PyG's `batch` vector. */
int rgnn_collate_offsets(int64_t* edge_index, int64_t n_edges, const int64_t* edge_ptr, const int64_t* node_ptr, int32_t n_frames,
                         int64_t n_nodes, int64_t* batch_of_node, rgnn_stream_t stream);

/* Graph-specific half of the backward of MPNNConv / RadarPointGNNConv with one Linear in pre_mlp (autograd through
 * gnn/mpnn_layers.py:86-101, 171-184; gnn/trainer.py:228-231).  With m_e = A[t_e] + B[s_e] + W_e e_e + bias
 * (A = x W_t^T or NULL for RadarPointGNNConv, B = x W_s^T, both [N, pp], pp = p rounded up to 4) and grad_m [N, ld_grad]
 * the gradient of the aggregated messages, it returns
 *   m_out [N, pp]     the aggregated messages themselves (0 for nodes without incoming edge),
 *   grad_a [N, pp]    sum over a node's incoming edges of the per-edge message gradient (also the bias gradient, summed),
 *   grad_b [N, pp]    the same sum over the node's outgoing edges (scatter by source),
 *   grad_we [p, de]   gradient of W_e,   grad_ea_csc [E, de] gradient of the edge attributes in CSC slot order (may be NULL).
 * max / min route every channel's gradient to the winning edge (first slot on exact ties), add to every edge, mean to
 * every edge / deg.  grad_b, grad_we, grad_ea_csc are zeroed by the call and accumulated with fp32 atomics. */
int rgnn_conv_backward_route(int32_t aggr, const float* a, const float* b, int32_t p, const float* bias, const float* w_e,
                             int64_t ldwe, int32_t de, const float* ea_csc, const int32_t* csc_ptr, const int32_t* csc_src,
                             int64_t n_nodes, int64_t n_edges, const float* grad_m, int64_t ld_grad, float* m_out, float* grad_a,
                             float* grad_b, float* grad_we, float* grad_ea_csc, rgnn_stream_t stream);

/* ------------------------------------------------------------------------------------ */
/* fused path: graph build + L-layer MPNN forward (the north-star hot path)               */
/* ------------------------------------------------------------------------------------ */
typedef struct rgnn_pipeline_desc {
  int32_t search;  /* 0 = knn, 1 = radius (radius: n_edges must come from a prior count) */
  int32_t k;
  double r;
  int32_t distance_dims;  /* 2 = "X", 4 = "XV" */
  int32_t edge_mode;
  int32_t n_edge_features;
  int32_t edge_features[RGNN_MAX_EDGE_FEATURES];
  int32_t n_layers;
  const rgnn_conv_desc* layers;   /* HOST array [n_layers] */
  const float* const* bn_weight;  /* HOST array of device pointers [n_layers], entries may be NULL */
  const float* const* bn_bias;
  float bn_eps;
  /* Optional running statistics (torch BatchNorm1d semantics: running = (1 - momentum) * running +
   * momentum * batch, unbiased variance), updated in place by every forward when non-NULL: HOST arrays of
   * device pointers [n_layers], the arrays themselves or single entries may be NULL. */
  float bn_momentum;
  float* const* bn_running_mean;
  float* const* bn_running_var;
} rgnn_pipeline_desc;

size_t rgnn_pipeline_workspace_bytes(const rgnn_pipeline_desc* desc, int64_t n_points,
                                     int32_t n_frames, int64_t n_edges);

/* Everything resident in HBM.  pos / vel f32 [N, 2]; x0 f32 [N, C0].  Outputs:
 * edge_index int64 [2, E], edge_attr f32 [E, De], h f32 [N, C_last] with
 * h = relu(bn(conv_L(... relu(bn(conv_1(x0)))))).  *error_flag as in rgnn_edge_features. */
int rgnn_pipeline_forward(const rgnn_pipeline_desc* desc, const float* pos, const float* vel,
                          const float* x0, const int64_t* frame_ptr_host, int32_t n_frames,
                          int64_t* edge_index, int64_t n_edges, float* edge_attr, float* h,
                          int32_t* error_flag, void* workspace, size_t workspace_bytes,
                          rgnn_stream_t stream);

/* Same path with HOST buffers (pinned or pageable): copies pos / vel / x0 to the device
 * staging area inside `workspace`, runs the path, copies edge_index / edge_attr / h back
 * and synchronises `stream`.  Output pointers may be NULL to skip that copy.
 * workspace must hold rgnn_pipeline_host_workspace_bytes(). */
size_t rgnn_pipeline_host_workspace_bytes(const rgnn_pipeline_desc* desc, int64_t n_points,
                                          int32_t n_frames, int64_t n_edges, int32_t c0);
int rgnn_pipeline_forward_host(const rgnn_pipeline_desc* desc, const float* pos_host,
                               const float* vel_host, const float* x0_host, int32_t c0,
                               const int64_t* frame_ptr_host, int32_t n_frames,
                               int64_t* edge_index_host, int64_t n_edges, float* edge_attr_host,
                               float* h_host, void* workspace, size_t workspace_bytes,
                               rgnn_stream_t stream);

/* The same call split into submit and wait, so that consecutive batches form a pipeline (the role of the
 * prefetching DataLoader in front of the reference's loop, gnn/trainer.py:210-233): while batch i computes,
 * batch i + 1 uploads and batch i - 1 downloads.  rgnn_pipeline_submit_host enqueues the work of one batch
 * on `slot` (0 .. RGNN_HOST_SLOTS - 1) and returns without waiting; rgnn_pipeline_wait_host blocks until
 * that batch's outputs are in the host buffers and returns its status (the error codes of
 * rgnn_pipeline_forward_host).  Every slot in flight needs its own workspace and its own output buffers;
 * the input buffers must stay untouched until the wait.  Batches compute in submission order (BatchNorm
 * running statistics see them in that order).  Submitting to a slot that has not been waited for returns
 * RGNN_ERR_INVALID_ARGUMENT, and so does waiting on an idle slot.  Pinned host buffers are needed for the
 * copies to overlap. */
#define RGNN_HOST_SLOTS 4
int rgnn_pipeline_submit_host(int32_t slot, const rgnn_pipeline_desc* desc, const float* pos_host,
                              const float* vel_host, const float* x0_host, int32_t c0,
                              const int64_t* frame_ptr_host, int32_t n_frames,
                              int64_t* edge_index_host, int64_t n_edges, float* edge_attr_host,
                              float* h_host, void* workspace, size_t workspace_bytes,
                              rgnn_stream_t stream);
int rgnn_pipeline_wait_host(int32_t slot);

#ifdef __cplusplus
}
#endif
#endif /* RGNN_H_ */
