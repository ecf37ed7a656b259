"""Tensor-level wrappers over the C ABI (include/rgnn.h): torch CUDA tensors in, torch CUDA
tensors out, all work enqueued on torch's current stream.  torch is used for device memory
and streams only -- every computation below happens in librgnn_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

__all__ = [
    "knn_graph", "radius_graph", "edge_features", "undirected_degree", "node_features", "csc_build",
    "CscGraph", "ConvParams", "conv_forward", "batchnorm_relu", "affine_relu", "sum_f32", "linear", "PipelineConfig",
    "pipeline_forward", "pipeline_forward_host", "HostPipeline", "knn_edge_count",
]


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float64:
        return _lib.F64
    raise TypeError(f"expected a float32 or float64 tensor, got {t.dtype}")


def _frame_ptr(frame_ptr, n: int) -> np.ndarray:
    if frame_ptr is None:
        return np.array([0, n], dtype=np.int64)
    fp = np.ascontiguousarray(np.asarray(frame_ptr, dtype=np.int64))
    if fp.ndim != 1 or fp.shape[0] < 2 or fp[0] != 0 or fp[-1] != n:
        raise ValueError("frame_ptr must be [F+1] int64 with frame_ptr[0] = 0 and frame_ptr[-1] = N")
    return fp


def _version_of(t: torch.Tensor) -> int:
    """torch's in-place version counter; tensors made under inference_mode have none (and cannot change)."""
    try:
        return t._version
    except RuntimeError:
        return -1


def _cuda_contig(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: radargnn_b200 has no CPU fallback")
    return t.contiguous()


def knn_edge_count(frame_ptr: np.ndarray, k: int) -> int:
    lib = _lib.load()
    st = C.c_int(0)
    e = lib.rgnn_knn_edge_count(frame_ptr.ctypes.data, len(frame_ptr) - 1, int(k), C.byref(st))
    _lib.check(st.value)
    return int(e)


def knn_graph(basis: torch.Tensor, k: int, frame_ptr=None, check_input: bool = True) -> torch.Tensor:
    """k-NN graph of every frame (graph.py:52-66): edge_index int64 [2, E], row 0 the query point,
    row 1 its neighbours by ascending (fp64 squared distance, index)."""
    _lib.require_device()
    lib = _lib.load()
    basis = _cuda_contig(basis, "basis")
    n, dims = basis.shape
    fp = _frame_ptr(frame_ptr, n)
    n_edges = knn_edge_count(fp, k)
    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=basis.device)
    flag = torch.zeros(1, dtype=torch.int32, device=basis.device)
    with torch.cuda.device(basis.device):
        ws = _lib.workspace(lib.rgnn_graph_workspace_bytes(n, len(fp) - 1), basis.device)
        _lib.check(lib.rgnn_graph_build_knn(basis.data_ptr(), _dtype_code(basis), dims, fp.ctypes.data,
                                            len(fp) - 1, int(k), edge_index.data_ptr(), n_edges, flag.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    if check_input:
        _lib.check(int(flag.item()))   # sklearn: ValueError("Input contains NaN")
    return edge_index


def radius_graph(basis: torch.Tensor, r: float, frame_ptr=None) -> torch.Tensor:
    """Radius graph (graph.py:68-82): all j != i with squared distance <= r*r (fp64, inclusive);
    rows ascend in i, columns ascend in j (canonical order)."""
    _lib.require_device()
    lib = _lib.load()
    basis = _cuda_contig(basis, "basis")
    n, dims = basis.shape
    fp = _frame_ptr(frame_ptr, n)
    with torch.cuda.device(basis.device):
        ws = _lib.workspace(lib.rgnn_graph_workspace_bytes(n, len(fp) - 1), basis.device)
        count = C.c_int64(0)
        _lib.check(lib.rgnn_graph_build_radius_count(basis.data_ptr(), _dtype_code(basis), dims, fp.ctypes.data,
                                                     len(fp) - 1, float(r), C.byref(count), ws.data_ptr(),
                                                     ws.numel(), _lib.stream_ptr()))
        n_edges = int(count.value)
        edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=basis.device)
        _lib.check(lib.rgnn_graph_build_radius_fill(basis.data_ptr(), _dtype_code(basis), dims, fp.ctypes.data,
                                                    len(fp) - 1, float(r), edge_index.data_ptr(), n_edges,
                                                    ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return edge_index


def _edge_feature_ids(features: Sequence[str]) -> List[int]:
    ids = []
    for f in features:
        if f not in _lib.EDGE_FEATURES:
            raise Exception("Invalid feature specified")  # graph.py:220
        ids.append(_lib.EDGE_FEATURES[f])
    return ids


def _edge_mode(edge_mode: str) -> int:
    if edge_mode == "directed":
        return _lib.DIRECTED
    if edge_mode == "undirected":
        return _lib.UNDIRECTED
    raise Exception("Invalid edge mode specified")


def edge_features(pos: torch.Tensor, vel: torch.Tensor, edge_index: torch.Tensor, features: Sequence[str],
                  edge_mode: str, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """graph.py:139-223 + features.py:6-122: fp64 arithmetic per edge, [E, De] of ``out_dtype``."""
    _lib.require_device()
    lib = _lib.load()
    pos, vel = _cuda_contig(pos, "pos"), _cuda_contig(vel, "vel")
    if pos.dtype != vel.dtype:
        raise TypeError("pos and vel must have the same dtype")
    edge_index = _cuda_contig(edge_index, "edge_index")
    ids = _edge_feature_ids(features)
    arr = _lib.int32_array(ids)
    width = lib.rgnn_edge_feature_width(arr, len(ids))
    n_edges = edge_index.shape[1]
    out = torch.empty((n_edges, width), dtype=out_dtype, device=pos.device)
    flag = torch.zeros(1, dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        _lib.check(lib.rgnn_edge_features(pos.data_ptr(), vel.data_ptr(), _dtype_code(pos), pos.shape[1],
                                          vel.shape[1], pos.shape[0], edge_index.data_ptr(), n_edges, arr,
                                          len(ids), _edge_mode(edge_mode), out.data_ptr(), _dtype_code(out),
                                          flag.data_ptr(), _lib.stream_ptr()))
    _lib.check(int(flag.item()))
    return out


def undirected_degree(edge_index: torch.Tensor, n: int) -> torch.Tensor:
    """graph.py:93-96: degree of the undirected graph, int32 [N]."""
    _lib.require_device()
    lib = _lib.load()
    edge_index = _cuda_contig(edge_index, "edge_index")
    deg = torch.empty(n, dtype=torch.int32, device=edge_index.device)
    with torch.cuda.device(edge_index.device):
        _lib.check(lib.rgnn_undirected_degree(edge_index.data_ptr(), edge_index.shape[1], n, deg.data_ptr(),
                                              _lib.stream_ptr()))
    return deg


def node_features(features: Sequence[str], n: int, device, *, rcs=None, time_index=None, degree=None,
                  pos=None, vel=None, out_dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """graph.py:225-275: concatenate the listed node features into [N, Fn]."""
    _lib.require_device()
    lib = _lib.load()
    ids = [_lib.NODE_FEATURES[f] for f in features]
    arr = _lib.int32_array(ids)
    width = lib.rgnn_node_feature_width(arr, len(ids))

    def f64(t):
        return None if t is None else _cuda_contig(t.to(torch.float64), "node feature")

    rcs, time_index, pos, vel = f64(rcs), f64(time_index), f64(pos), f64(vel)
    degree = None if degree is None else _cuda_contig(degree.to(torch.int32), "degree")
    out = torch.empty((n, width), dtype=out_dtype, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.rgnn_node_features(_lib.ptr(rcs), _lib.ptr(time_index), _lib.ptr(degree), _lib.ptr(pos),
                                          _lib.ptr(vel), n, arr, len(ids), out.data_ptr(), _dtype_code(out),
                                          _lib.stream_ptr()))
    return out


@dataclass
class CscGraph:
    """Target-major view of edge_index (messages are reduced at edge_index[1])."""
    ptr: torch.Tensor   # int32 [N + 1]
    src: torch.Tensor   # int32 [E]
    eid: torch.Tensor   # int32 [E]
    n_nodes: int
    n_edges: int


def csc_build(edge_index: torch.Tensor, n_nodes: int) -> CscGraph:
    _lib.require_device()
    lib = _lib.load()
    edge_index = _cuda_contig(edge_index, "edge_index")
    if edge_index.dtype != torch.int64:
        raise TypeError("edge_index must be int64 [2, E]")
    n_edges = edge_index.shape[1]
    dev = edge_index.device
    ptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    src = torch.empty(n_edges, dtype=torch.int32, device=dev)
    eid = torch.empty(n_edges, dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _lib.workspace(lib.rgnn_csc_workspace_bytes(n_nodes, n_edges), dev)
        _lib.check(lib.rgnn_csc_build(edge_index.data_ptr(), n_edges, n_nodes, ptr.data_ptr(), src.data_ptr(),
                                      eid.data_ptr(), flag.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    _lib.check(int(flag.item()))   # an id outside [0, N) raises like PyG's gather does
    return CscGraph(ptr, src, eid, n_nodes, n_edges)


@dataclass
class ConvParams:
    """Parameters of one MPNNConv / RadarPointGNNConv, PyG `Linear` layout ([out, in] fp32)."""
    conv_type: str                      # "MPNNConv" | "RadarPointGNNConv"
    in_channels: int
    out_channels: int
    edge_dim: int
    aggr: str
    pre: List[Tuple[torch.Tensor, torch.Tensor]]    # (weight, bias) of every Linear in pre_mlp
    post: List[Tuple[torch.Tensor, torch.Tensor]]
    edge_encoder: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
    # tensor-core weight images, rebuilt only when a weight tensor was replaced or modified in place
    # (the key holds every tensor's address and torch version counter).  Edits through ``.data`` bypass the
    # version counter: call invalidate() after them (a held PipelineConfig / ConvParams keeps its images).
    _packed: Optional[torch.Tensor] = None
    _packed_key: Optional[tuple] = None

    def invalidate(self) -> None:
        self._packed, self._packed_key = None, None

    def _tensors(self):
        out = [t for pair in self.pre + self.post for t in pair]
        if self.edge_encoder is not None:
            out += list(self.edge_encoder)
        return out

    def desc(self, keep: list, pack: bool = True) -> _lib.ConvDesc:
        if self.aggr not in _lib.AGGR:
            raise ValueError(f"unsupported aggregation {self.aggr!r}")
        if len(self.pre) > _lib.MAX_MLP_LAYERS or len(self.post) > _lib.MAX_MLP_LAYERS:
            raise ValueError("too many MLP layers")
        d = _lib.ConvDesc()
        d.conv_type = _lib.CONV_MPNN if self.conv_type == "MPNNConv" else _lib.CONV_RADAR_POINT_GNN
        d.aggr = _lib.AGGR[self.aggr]
        d.in_channels, d.out_channels, d.edge_dim = self.in_channels, self.out_channels, self.edge_dim
        d.pre_layers, d.post_layers = len(self.pre), len(self.post)
        d.use_edge_encoder = 1 if self.edge_encoder is not None else 0

        def dev(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                t = t.to(dtype=torch.float32).contiguous()
                if not t.is_cuda:
                    raise RuntimeError("layer parameters must live on the CUDA device")
            keep.append(t)
            return t.data_ptr()

        for i, (w, b) in enumerate(self.pre):
            d.pre_weight[i], d.pre_bias[i] = dev(w), dev(b)
        for i, (w, b) in enumerate(self.post):
            d.post_weight[i], d.post_bias[i] = dev(w), dev(b)
        if self.edge_encoder is not None:
            d.edge_encoder_weight, d.edge_encoder_bias = dev(self.edge_encoder[0]), dev(self.edge_encoder[1])
        if pack:
            lib = _lib.load()
            nbytes = lib.rgnn_conv_packed_bytes(C.byref(d))
            if nbytes > 0:
                key = tuple((t.data_ptr(), _version_of(t), tuple(t.shape)) for t in self._tensors()) + (self.aggr,)
                if self._packed is None or self._packed_key != key or self._packed.numel() < nbytes:
                    device = self.pre[0][0].device
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
                    with torch.cuda.device(device):
                        _lib.check(lib.rgnn_conv_pack_weights(C.byref(d), buf.data_ptr(), nbytes, _lib.stream_ptr()))
                    self._packed, self._packed_key = buf, key
                keep.append(self._packed)
                d.packed_weights = self._packed.data_ptr()
        return d


def conv_forward(params: ConvParams, x: torch.Tensor, csc: CscGraph, edge_attr: torch.Tensor) -> torch.Tensor:
    """One graph convolution (mpnn_layers.py:86-101 / 171-184) -> [N, out_channels] fp32."""
    _lib.require_device()
    lib = _lib.load()
    x = _cuda_contig(x.to(torch.float32), "x")
    edge_attr = _cuda_contig(edge_attr.to(torch.float32), "edge_attr")
    if x.shape[1] != params.in_channels:
        raise ValueError(f"x has {x.shape[1]} channels, layer expects {params.in_channels}")
    if edge_attr.shape[0] != csc.n_edges or (csc.n_edges and edge_attr.shape[1] != params.edge_dim):
        raise ValueError("edge_attr shape does not match the graph / layer")
    keep: list = []
    d = params.desc(keep)
    out = torch.empty((csc.n_nodes, params.out_channels), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _lib.workspace(lib.rgnn_conv_workspace_bytes(C.byref(d), csc.n_nodes, csc.n_edges), x.device)
        _lib.check(lib.rgnn_conv_forward(C.byref(d), x.data_ptr(), csc.n_nodes, csc.ptr.data_ptr(),
                                         csc.src.data_ptr(), csc.eid.data_ptr(), edge_attr.data_ptr(),
                                         csc.n_edges, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr()))
    return out


def batchnorm_relu(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                   eps: float = 1e-5, momentum: float = 0.1, running_mean: Optional[torch.Tensor] = None,
                   running_var: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """Training-mode BatchNorm1d (+ReLU), gnn_models.py:126-128.  Running statistics are updated
    in place when given."""
    _lib.require_device()
    lib = _lib.load()
    x = _cuda_contig(x.to(torch.float32), "x")
    n, c = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _lib.workspace(lib.rgnn_batchnorm_workspace_bytes(n, c), x.device)
        _lib.check(lib.rgnn_batchnorm_relu_forward(
            x.data_ptr(), n, c, _lib.ptr(None if weight is None else weight.detach()),
            _lib.ptr(None if bias is None else bias.detach()), float(eps), float(momentum),
            _lib.ptr(running_mean), _lib.ptr(running_var), 1 if relu else 0, out.data_ptr(), ws.data_ptr(),
            ws.numel(), _lib.stream_ptr()))
    return out


def affine_relu(x: torch.Tensor, mean: torch.Tensor, scale: torch.Tensor, beta: torch.Tensor, relu: bool) -> torch.Tensor:
    """out = relu?((x - mean) * scale + beta) per channel (eval-mode BatchNorm)."""
    _lib.require_device()
    lib = _lib.load()
    x = _cuda_contig(x.to(torch.float32), "x")
    out = torch.empty_like(x)
    mean, scale, beta = (t.detach().to(torch.float32).contiguous() for t in (mean, scale, beta))
    with torch.cuda.device(x.device):
        _lib.check(lib.rgnn_affine_relu_forward(x.data_ptr(), x.shape[0], x.shape[1], mean.data_ptr(),
                                                scale.data_ptr(), beta.data_ptr(), 1 if relu else 0,
                                                out.data_ptr(), _lib.stream_ptr()))
    return out


def sum_f32(x: torch.Tensor) -> torch.Tensor:
    """Deterministic fp64 sum of a float32 tensor, left on the device (double [1])."""
    _lib.require_device()
    lib = _lib.load()
    x = _cuda_contig(x.to(torch.float32), "x")
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        ws = _lib.workspace(lib.rgnn_sum_workspace_bytes(), x.device)
        _lib.check(lib.rgnn_sum_f32(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _lib.stream_ptr()))
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu_input: bool = False) -> torch.Tensor:
    """PyG `Linear`: y = act(x) W^T + b with W [out, in]."""
    _lib.require_device()
    lib = _lib.load()
    x = _cuda_contig(x.to(torch.float32), "x")
    w = _cuda_contig(weight.detach().to(torch.float32), "weight")
    b = None if bias is None else _cuda_contig(bias.detach().to(torch.float32), "bias")
    y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.rgnn_linear_forward(x.data_ptr(), x.shape[0], x.shape[1], w.data_ptr(), _lib.ptr(b),
                                           w.shape[0], 1 if relu_input else 0, y.data_ptr(), _lib.stream_ptr()))
    return y


@dataclass
class PipelineConfig:
    """Graph construction + conv stack of the fused hot path (rgnn_pipeline_desc)."""
    layers: List[ConvParams]
    bn: List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]   # (weight, bias) per layer
    algorithm: str = "knn"          # "knn" | "radius"
    k: int = 16
    r: float = 1.0
    distance_definition: str = "X"  # "X" (2-D) | "XV" (4-D)
    edge_features: Sequence[str] = ("relative_position",)
    edge_mode: str = "directed"
    bn_eps: float = 1e-5
    # optional running statistics per layer, updated in place by every forward (torch BatchNorm1d semantics)
    bn_running: Optional[List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]] = None
    bn_momentum: float = 0.1


class _PipelineHandle:
    """Keeps the ctypes structures (and the tensors they point to) alive for a call."""

    def __init__(self, cfg: PipelineConfig):
        self.keep: list = []
        n_layers = len(cfg.layers)
        self.layer_array = (_lib.ConvDesc * n_layers)(*[p.desc(self.keep) for p in cfg.layers])
        self.bn_w = (C.c_void_p * n_layers)()
        self.bn_b = (C.c_void_p * n_layers)()
        for i, (w, b) in enumerate(cfg.bn):
            for arr, t in ((self.bn_w, w), (self.bn_b, b)):
                if t is not None:
                    t = t.detach().to(torch.float32).contiguous()
                    self.keep.append(t)
                    arr[i] = t.data_ptr()
        d = _lib.PipelineDesc()
        if cfg.algorithm not in ("knn", "radius"):
            raise Exception("Invalid graph construction algorithm selected")  # preprocessor/configs.py:26
        d.search = 0 if cfg.algorithm == "knn" else 1
        d.k, d.r = int(cfg.k), float(cfg.r)
        if cfg.distance_definition not in ("X", "XV"):
            raise Exception("Invalid distance definition")
        d.distance_dims = 2 if cfg.distance_definition == "X" else 4
        d.edge_mode = _edge_mode(cfg.edge_mode)
        ids = _edge_feature_ids(cfg.edge_features)
        d.n_edge_features = len(ids)
        for i, v in enumerate(ids):
            d.edge_features[i] = v
        d.n_layers = n_layers
        d.layers = C.cast(self.layer_array, C.POINTER(_lib.ConvDesc))
        d.bn_weight = C.cast(self.bn_w, C.POINTER(C.c_void_p))
        d.bn_bias = C.cast(self.bn_b, C.POINTER(C.c_void_p))
        d.bn_eps = float(cfg.bn_eps)
        d.bn_momentum = float(cfg.bn_momentum)
        if cfg.bn_running is not None:
            self.bn_rm = (C.c_void_p * n_layers)()
            self.bn_rv = (C.c_void_p * n_layers)()
            for i, (rm, rv) in enumerate(cfg.bn_running):
                if rm is None or rv is None:
                    continue
                if rm.dtype != torch.float32 or rv.dtype != torch.float32 or not rm.is_cuda or not rv.is_cuda \
                        or not rm.is_contiguous() or not rv.is_contiguous():
                    raise ValueError("running statistics must be contiguous float32 CUDA tensors (they are updated in place)")
                self.keep += [rm, rv]
                self.bn_rm[i], self.bn_rv[i] = rm.data_ptr(), rv.data_ptr()
            d.bn_running_mean = C.cast(self.bn_rm, C.POINTER(C.c_void_p))
            d.bn_running_var = C.cast(self.bn_rv, C.POINTER(C.c_void_p))
        self.desc = d
        self.edge_dim = _lib.load().rgnn_edge_feature_width(_lib.int32_array(ids), len(ids))


def _pipeline_edge_count(cfg: PipelineConfig, pos: torch.Tensor, vel: torch.Tensor, fp: np.ndarray) -> int:
    if cfg.algorithm == "knn":
        return knn_edge_count(fp, cfg.k)
    lib = _lib.load()
    basis = pos if cfg.distance_definition == "X" else torch.cat([pos, vel], dim=1).contiguous()
    ws = _lib.workspace(lib.rgnn_graph_workspace_bytes(pos.shape[0], len(fp) - 1), pos.device)
    count = C.c_int64(0)
    _lib.check(lib.rgnn_graph_build_radius_count(basis.data_ptr(), _lib.F32, basis.shape[1], fp.ctypes.data,
                                                 len(fp) - 1, float(cfg.r), C.byref(count), ws.data_ptr(),
                                                 ws.numel(), _lib.stream_ptr()))
    return int(count.value)


def pipeline_forward(cfg: PipelineConfig, pos: torch.Tensor, vel: torch.Tensor, x0: torch.Tensor,
                     frame_ptr=None, workspace: Optional[torch.Tensor] = None,
                     out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                     check_errors: bool = True):
    """Graph build + L x (conv, BatchNorm(train), ReLU) with everything resident in HBM.
    Returns (edge_index int64 [2, E], edge_attr f32 [E, De], h f32 [N, C_last])."""
    _lib.require_device()
    lib = _lib.load()
    pos = _cuda_contig(pos.to(torch.float32), "pos")
    vel = _cuda_contig(vel.to(torch.float32), "vel")
    x0 = _cuda_contig(x0.to(torch.float32), "x0")
    n = pos.shape[0]
    fp = _frame_ptr(frame_ptr, n)
    handle = _PipelineHandle(cfg)
    dev = pos.device
    with torch.cuda.device(dev):
        n_edges = _pipeline_edge_count(cfg, pos, vel, fp)
        if out is None:
            edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
            edge_attr = torch.empty((n_edges, handle.edge_dim), dtype=torch.float32, device=dev)
            h = torch.empty((n, cfg.layers[-1].out_channels), dtype=torch.float32, device=dev)
        else:
            edge_index, edge_attr, h = out
        need = lib.rgnn_pipeline_workspace_bytes(C.byref(handle.desc), n, len(fp) - 1, n_edges)
        if need == 0:
            raise ValueError("invalid pipeline configuration")
        if workspace is None or workspace.numel() < need:
            workspace = _lib.workspace(need, dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.rgnn_pipeline_forward(C.byref(handle.desc), pos.data_ptr(), vel.data_ptr(), x0.data_ptr(),
                                             fp.ctypes.data, len(fp) - 1, edge_index.data_ptr(), n_edges,
                                             edge_attr.data_ptr(), h.data_ptr(), flag.data_ptr(),
                                             workspace.data_ptr(), workspace.numel(), _lib.stream_ptr()))
    if check_errors:
        _lib.check(int(flag.item()))
    return edge_index, edge_attr, h


def pipeline_forward_host(cfg: PipelineConfig, pos: np.ndarray, vel: np.ndarray, x0: np.ndarray,
                          frame_ptr=None, device="cuda:0", want_graph: bool = True):
    """Same path through the HOST-buffer entry point (rgnn_pipeline_forward_host): numpy in,
    numpy out, host<->device copies inside the call.  Radius graphs need a device pass to learn E."""
    _lib.require_device()
    lib = _lib.load()
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    vel = np.ascontiguousarray(vel, dtype=np.float32)
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    n = pos.shape[0]
    fp = _frame_ptr(frame_ptr, n)
    handle = _PipelineHandle(cfg)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        if cfg.algorithm == "knn":
            n_edges = knn_edge_count(fp, cfg.k)
        else:
            n_edges = _pipeline_edge_count(cfg, torch.from_numpy(pos).to(dev), torch.from_numpy(vel).to(dev), fp)
        edge_index = np.empty((2, n_edges), dtype=np.int64) if want_graph else None
        edge_attr = np.empty((n_edges, handle.edge_dim), dtype=np.float32) if want_graph else None
        h = np.empty((n, cfg.layers[-1].out_channels), dtype=np.float32)
        need = lib.rgnn_pipeline_host_workspace_bytes(C.byref(handle.desc), n, len(fp) - 1, n_edges, x0.shape[1])
        if need == 0:
            raise ValueError("invalid pipeline configuration")
        ws = _lib.workspace(need, dev)
        _lib.check(lib.rgnn_pipeline_forward_host(
            C.byref(handle.desc), pos.ctypes.data, vel.ctypes.data, x0.ctypes.data, x0.shape[1], fp.ctypes.data,
            len(fp) - 1, None if edge_index is None else edge_index.ctypes.data, n_edges,
            None if edge_attr is None else edge_attr.ctypes.data, h.ctypes.data, ws.data_ptr(), ws.numel(),
            _lib.stream_ptr()))
    return edge_index, edge_attr, h


class HostPipeline:
    """Batches through the host-buffer path with several calls in flight (rgnn_pipeline_submit_host /
    rgnn_pipeline_wait_host): while batch i computes, batch i + 1 uploads and batch i - 1 downloads -- the role the
    prefetching ``DataLoader`` plays in front of the reference's loop (gnn/trainer.py:210-233).

        pipe = HostPipeline(cfg, depth=2)
        for batch in batches:
            if pipe.full:
                edge_index, edge_attr, h = pipe.wait()      # oldest batch, numpy views of pinned memory
            pipe.submit(batch.pos, batch.vel, batch.x0, batch.frame_ptr)
        while pipe.in_flight:
            edge_index, edge_attr, h = pipe.wait()

    The arrays ``wait`` returns are views of the slot's pinned buffers: valid until ``depth`` further submits
    (pass ``copy=True`` to own them).  k-NN graphs only: the edge count of a radius graph needs a device pass."""

    def __init__(self, cfg: PipelineConfig, depth: int = 2, device="cuda:0", want_graph: bool = True):
        _lib.require_device()
        if not 1 <= depth <= _lib.HOST_SLOTS:
            raise ValueError(f"depth must be in 1 .. {_lib.HOST_SLOTS}")
        if cfg.algorithm != "knn":
            raise ValueError("HostPipeline handles k-NN graphs (a radius graph's edge count needs a device pass)")
        self.cfg, self.depth, self.want_graph = cfg, depth, want_graph
        self.device = torch.device(device)
        self._lib = _lib.load()
        self._handle = _PipelineHandle(cfg)
        self._slots = [{} for _ in range(depth)]
        self._order: list = []      # slots in flight, oldest first
        self._next = 0

    @property
    def in_flight(self) -> int:
        return len(self._order)

    @property
    def full(self) -> bool:
        return len(self._order) == self.depth

    @staticmethod
    def _pinned(slot: dict, name: str, shape, dtype) -> torch.Tensor:
        """Pinned buffer of the slot, reallocated only when the batch outgrows it (the replay cache of the
        library keys on the addresses)."""
        need = int(np.prod(shape))
        buf = slot.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < need:
            buf = torch.empty(max(need, 1), dtype=dtype).pin_memory()
            slot[name] = buf
        return buf[:need].view(*shape)

    def submit(self, pos: np.ndarray, vel: np.ndarray, x0: np.ndarray, frame_ptr=None) -> None:
        if self.full:
            raise RuntimeError("every slot is in flight: wait() first")
        lib, handle = self._lib, self._handle
        n = int(pos.shape[0])
        fp = _frame_ptr(frame_ptr, n)
        n_edges = knn_edge_count(fp, self.cfg.k)
        c0 = int(x0.shape[1])
        idx = self._next
        slot = self._slots[idx]
        pos_p = self._pinned(slot, "pos", (n, 2), torch.float32)
        vel_p = self._pinned(slot, "vel", (n, 2), torch.float32)
        x0_p = self._pinned(slot, "x0", (n, c0), torch.float32)
        pos_p.copy_(torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float32)))
        vel_p.copy_(torch.from_numpy(np.ascontiguousarray(vel, dtype=np.float32)))
        x0_p.copy_(torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32)))
        h = self._pinned(slot, "h", (n, self.cfg.layers[-1].out_channels), torch.float32)
        ei = self._pinned(slot, "edge_index", (2, n_edges), torch.int64) if self.want_graph else None
        ea = self._pinned(slot, "edge_attr", (n_edges, handle.edge_dim), torch.float32) if self.want_graph else None
        with torch.cuda.device(self.device):
            need = lib.rgnn_pipeline_host_workspace_bytes(C.byref(handle.desc), n, len(fp) - 1, n_edges, c0)
            if need == 0:
                raise ValueError("invalid pipeline configuration")
            ws = slot.get("ws")
            if ws is None or ws.numel() < need:
                ws = slot["ws"] = _lib.workspace(need, self.device)
            slot["fp"] = fp   # read by the call: keep it alive
            _lib.check(lib.rgnn_pipeline_submit_host(
                idx, C.byref(handle.desc), pos_p.data_ptr(), vel_p.data_ptr(), x0_p.data_ptr(), c0, fp.ctypes.data,
                len(fp) - 1, None if ei is None else ei.data_ptr(), n_edges, None if ea is None else ea.data_ptr(),
                h.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        slot["out"] = (ei, ea, h)
        self._order.append(idx)
        self._next = (idx + 1) % self.depth

    def wait(self, copy: bool = False):
        """Outputs of the oldest batch in flight: (edge_index, edge_attr, h) as numpy arrays."""
        if not self._order:
            raise RuntimeError("nothing in flight")
        idx = self._order.pop(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rgnn_pipeline_wait_host(idx))
        out = tuple(None if t is None else (t.numpy().copy() if copy else t.numpy()) for t in self._slots[idx]["out"])
        return out

    def drain(self) -> None:
        while self._order:
            idx = self._order.pop(0)
            with torch.cuda.device(self.device):
                self._lib.rgnn_pipeline_wait_host(idx)

    def __del__(self):
        try:
            self.drain()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# callers either side of the path: loss, NMS, nearest neighbour, time index, collate
# ---------------------------------------------------------------------------------------------
def detection_loss(cls: torch.Tensor, bb: torch.Tensor, y: torch.Tensor, class_weight: Optional[torch.Tensor],
                   bg_index: int, cls_loss_weight: float = 1.0, bb_loss_weight: float = 1.0, huber_delta: float = 1.0,
                   nan_to_zero: bool = True) -> torch.Tensor:
    """Weighted cross entropy + Huber box loss over the foreground nodes (reference gnn/trainer.py:184-231).
    Returns a DEVICE double[5]: loss, loss_cls, loss_bb, foreground nodes, labels outside [0, n_classes)."""
    _lib.require_device()
    lib = _lib.load()
    cls = _cuda_contig(cls.to(torch.float32), "cls")
    bb = _cuda_contig(bb.to(torch.float32), "bb")
    y = _cuda_contig(y.to(torch.float32), "y")
    n, k = cls.shape
    nb = bb.shape[1] if bb.dim() == 2 else 0
    if y.shape[0] != n or bb.shape[0] != n or y.shape[1] < 1 + nb:
        raise ValueError("cls [N, K], bb [N, B] and y [N, 1 + B] must agree")
    w = None if class_weight is None else _cuda_contig(class_weight.to(torch.float32), "class_weight")
    if w is not None and w.numel() != k:
        raise ValueError("class_weight needs one entry per class")
    out = torch.empty(5, dtype=torch.float64, device=cls.device)
    with torch.cuda.device(cls.device):
        ws = _lib.workspace(lib.rgnn_detection_loss_workspace_bytes(), cls.device)
        _lib.check(lib.rgnn_detection_loss(cls.data_ptr(), k, bb.data_ptr(), nb, y.data_ptr(), y.shape[1], n, _lib.ptr(w),
                                           int(bg_index), float(cls_loss_weight), float(bb_loss_weight), float(huber_delta),
                                           1 if nan_to_zero else 0, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _lib.stream_ptr()))
    return out


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float, rotated: bool = False,
        box_frame: Optional[torch.Tensor] = None, shift_negative: bool = True) -> torch.Tensor:
    """Greedy NMS (postprocessor/postprocessing.py:336-435): kept indices (int64) by (frame, descending score).
    aligned: boxes [n, 4] (x1, y1, x2, y2) fp32 (torchvision.ops.nms); rotated: [n, 5] (cx, cy, w, h, degrees) fp64."""
    _lib.require_device()
    lib = _lib.load()
    dt = torch.float64 if rotated else torch.float32
    boxes = _cuda_contig(boxes.to(dt), "boxes")
    scores = _cuda_contig(scores.to(dt).reshape(-1), "scores")
    n = boxes.shape[0]
    if boxes.dim() != 2 or boxes.shape[1] != (5 if rotated else 4) or scores.numel() != n:
        raise ValueError("boxes must be [n, 4] (aligned) or [n, 5] (rotated) with one score per box")
    fr = None if box_frame is None else _cuda_contig(box_frame.to(torch.int32), "box_frame")
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    count = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    flag = torch.empty(n, dtype=torch.uint8, device=boxes.device)
    with torch.cuda.device(boxes.device):
        nbytes = lib.rgnn_nms_workspace_bytes(n)
        if n > 0 and nbytes == 0:
            raise ValueError("nms handles at most 65536 boxes per call")
        ws = _lib.workspace(nbytes, boxes.device)
        _lib.check(lib.rgnn_nms(boxes.data_ptr(), 1 if rotated else 0, scores.data_ptr(), _lib.ptr(fr), n, float(iou_threshold),
                                1 if shift_negative else 0, keep.data_ptr(), count.data_ptr(), flag.data_ptr(),
                                ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return keep[: int(count.item())]


def nearest_neighbor(basis: torch.Tensor, frame_ptr=None):
    """(nn_index int64 [N], nn_points [N, D]): kneighbors_graph(X, 1) + X[np.where(A == 1)[1]]
    (dataset_creation.py:314-318, postprocessing.py:233-237)."""
    _lib.require_device()
    lib = _lib.load()
    basis = _cuda_contig(basis, "basis")
    if basis.dtype not in (torch.float32, torch.float64):
        basis = basis.to(torch.float64)
    n, d = basis.shape
    fp = _frame_ptr(frame_ptr, n)
    idx = torch.empty(n, dtype=torch.int64, device=basis.device)
    pts = torch.full_like(basis, float("nan"))
    with torch.cuda.device(basis.device):
        ws = _lib.workspace(lib.rgnn_nearest_neighbor_workspace_bytes(n, len(fp) - 1), basis.device)
        _lib.check(lib.rgnn_nearest_neighbor(basis.data_ptr(), _dtype_code(basis), d, fp.ctypes.data, len(fp) - 1,
                                             idx.data_ptr(), pts.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return idx, pts


def time_index(timestamp: torch.Tensor, frame_ptr=None) -> torch.Tensor:
    """Dense rank of every point's timestamp inside its frame (dataset_creation.py:214-223), float64."""
    _lib.require_device()
    lib = _lib.load()
    ts = _cuda_contig(timestamp.to(torch.float64).reshape(-1), "timestamp")
    n = ts.numel()
    fp = _frame_ptr(frame_ptr, n)
    fp_dev = torch.from_numpy(fp).to(ts.device)
    out = torch.zeros(n, dtype=torch.float64, device=ts.device)
    with torch.cuda.device(ts.device):
        # frames above 8192 points are ranked in global memory and need scratch
        ws = _lib.workspace(lib.rgnn_time_index_workspace_bytes(n), ts.device) if int(np.diff(fp).max(initial=0)) > 8192 else None
        _lib.check(lib.rgnn_time_index(ts.data_ptr(), fp_dev.data_ptr(), fp.ctypes.data, len(fp) - 1, out.data_ptr(),
                                       None if ws is None else ws.data_ptr(), 0 if ws is None else ws.numel(),
                                       _lib.stream_ptr()))
    return out


def collate_offsets(edge_index: torch.Tensor, edge_ptr, node_ptr):
    """PyG's disjoint-union collate of edge_index (utils/data_handling.py:30) on the device: edge_index [2, E] with
    frame-local ids (frames concatenated) gets each frame's node offset added IN PLACE; returns (edge_index, batch)."""
    _lib.require_device()
    lib = _lib.load()
    if edge_index.dtype != torch.int64 or not edge_index.is_contiguous() or not edge_index.is_cuda:
        raise ValueError("edge_index must be a contiguous CUDA int64 [2, E] tensor")
    ep = np.ascontiguousarray(np.asarray(edge_ptr, dtype=np.int64))
    npt = np.ascontiguousarray(np.asarray(node_ptr, dtype=np.int64))
    if ep.shape != npt.shape or ep[0] != 0 or npt[0] != 0 or ep[-1] != edge_index.shape[1]:
        raise ValueError("edge_ptr / node_ptr must be [F + 1] offset tables matching edge_index")
    dev = edge_index.device
    ep_d, np_d = torch.from_numpy(ep).to(dev), torch.from_numpy(npt).to(dev)
    batch = torch.empty(int(npt[-1]), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rgnn_collate_offsets(edge_index.data_ptr(), edge_index.shape[1], ep_d.data_ptr(), np_d.data_ptr(),
                                            len(ep) - 1, int(npt[-1]), batch.data_ptr(), _lib.stream_ptr()))
    return edge_index, batch
