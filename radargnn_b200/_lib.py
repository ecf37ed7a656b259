"""ctypes binding of librgnn_b200.so (C ABI declared in include/rgnn.h).

This is the only place that touches the shared library.  There is no CPU fallback: loading
fails loudly when the library has not been built (``python -m radargnn_b200.build``), and
every compute entry point raises when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librgnn_b200.so")

# ---- enums of include/rgnn.h -------------------------------------------------------------
OK, ERR_INVALID_ARGUMENT, ERR_K_NOT_SMALLER_THAN_N, ERR_WORKSPACE_TOO_SMALL, ERR_CUDA, \
    ERR_DOT_PRODUCT, ERR_INVALID_FEATURE, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_NON_FINITE_INPUT, \
    ERR_INDEX_OUT_OF_RANGE = range(11)
ABI_VERSION = 4
HOST_SLOTS = 4   # RGNN_HOST_SLOTS
F32, F64 = 0, 1
DIRECTED, UNDIRECTED = 0, 1
EDGE_FEATURES = {
    "point_pair_features": 0, "spatial_euclidean_distance": 1, "velocity_euclidean_distance": 2,
    "relative_position": 3, "relative_velocity": 4,
}
NODE_FEATURES = {
    "rcs": 0, "time_index": 1, "degree": 2, "velocity_vector_length": 3, "velocity_vector": 4,
    "spatial_coordinates": 5,
}
AGGR = {"max": 0, "add": 1, "sum": 1, "mean": 2, "min": 3}
CONV_MPNN, CONV_RADAR_POINT_GNN = 0, 1
MAX_MLP_LAYERS = 8
MAX_EDGE_FEATURES = 8
MAX_K = 64

_f32p = C.POINTER(C.c_float)


class ConvDesc(C.Structure):
    _fields_ = [
        ("conv_type", C.c_int32), ("aggr", C.c_int32), ("in_channels", C.c_int32),
        ("out_channels", C.c_int32), ("edge_dim", C.c_int32), ("pre_layers", C.c_int32),
        ("post_layers", C.c_int32), ("use_edge_encoder", C.c_int32),
        ("edge_encoder_weight", C.c_void_p), ("edge_encoder_bias", C.c_void_p),
        ("pre_weight", C.c_void_p * MAX_MLP_LAYERS), ("pre_bias", C.c_void_p * MAX_MLP_LAYERS),
        ("post_weight", C.c_void_p * MAX_MLP_LAYERS), ("post_bias", C.c_void_p * MAX_MLP_LAYERS),
        ("packed_weights", C.c_void_p),
    ]


class PipelineDesc(C.Structure):
    _fields_ = [
        ("search", C.c_int32), ("k", C.c_int32), ("r", C.c_double), ("distance_dims", C.c_int32),
        ("edge_mode", C.c_int32), ("n_edge_features", C.c_int32),
        ("edge_features", C.c_int32 * MAX_EDGE_FEATURES), ("n_layers", C.c_int32),
        ("layers", C.POINTER(ConvDesc)), ("bn_weight", C.POINTER(C.c_void_p)),
        ("bn_bias", C.POINTER(C.c_void_p)), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("bn_running_mean", C.POINTER(C.c_void_p)), ("bn_running_var", C.POINTER(C.c_void_p)),
    ]


class RgnnError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_lib: Optional[C.CDLL] = None

_PROTOTYPES = {
    "rgnn_abi_version": (C.c_int, []),
    "rgnn_status_string": (C.c_char_p, [C.c_int]),
    "rgnn_last_cuda_error": (C.c_char_p, []),
    "rgnn_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "rgnn_kernel_launch_count": (C.c_int64, []),
    "rgnn_profile_enable": (None, [C.c_int32]),
    "rgnn_profile_collect": (C.c_int32, []),
    "rgnn_profile_entry": (C.c_int, [C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rgnn_profile_reset": (None, []),
    "rgnn_graph_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "rgnn_knn_edge_count": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int)]),
    "rgnn_graph_build_knn": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_graph_build_radius_count": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                                C.c_double, C.POINTER(C.c_int64), C.c_void_p, C.c_size_t,
                                                C.c_void_p]),
    "rgnn_graph_build_radius_fill": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                               C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
    "rgnn_edge_feature_width": (C.c_int32, [C.c_void_p, C.c_int32]),
    "rgnn_edge_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_void_p]),
    "rgnn_undirected_degree": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "rgnn_node_feature_width": (C.c_int32, [C.c_void_p, C.c_int32]),
    "rgnn_node_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "rgnn_csc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "rgnn_csc_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_conv_packed_bytes": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "rgnn_conv_pack_weights": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_conv_workspace_bytes": (C.c_size_t, [C.POINTER(ConvDesc), C.c_int64, C.c_int64]),
    "rgnn_conv_forward": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "rgnn_batchnorm_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "rgnn_batchnorm_relu_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_affine_relu_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.c_void_p]),
    "rgnn_sum_workspace_bytes": (C.c_size_t, []),
    "rgnn_sum_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_linear_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_void_p]),
    "rgnn_detection_loss_workspace_bytes": (C.c_size_t, []),
    "rgnn_detection_loss": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_nms_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "rgnn_nms": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int32, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_nearest_neighbor_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "rgnn_nearest_neighbor": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_time_index_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "rgnn_time_index": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "rgnn_collate_offsets": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                       C.c_void_p]),
    "rgnn_conv_backward_route": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgnn_pipeline_workspace_bytes": (C.c_size_t, [C.POINTER(PipelineDesc), C.c_int64, C.c_int32, C.c_int64]),
    "rgnn_pipeline_forward": (C.c_int, [C.POINTER(PipelineDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_pipeline_host_workspace_bytes": (C.c_size_t, [C.POINTER(PipelineDesc), C.c_int64, C.c_int32,
                                                        C.c_int64, C.c_int32]),
    "rgnn_pipeline_forward_host": (C.c_int, [C.POINTER(PipelineDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_pipeline_submit_host": (C.c_int, [C.c_int32, C.POINTER(PipelineDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "rgnn_pipeline_wait_host": (C.c_int, [C.c_int32]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def load() -> C.CDLL:
    """dlopen the library (once) and attach the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m radargnn_b200.build` "
            "(nvcc, sm_100a).  radargnn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.rgnn_abi_version() != ABI_VERSION:
        raise ImportError("librgnn_b200.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


def require_device() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("radargnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def check(status: int) -> None:
    """Map an rgnn_status to the exception the reference raises for the same condition."""
    if status == OK:
        return
    lib = load()
    msg = lib.rgnn_status_string(status).decode()
    if status in (ERR_K_NOT_SMALLER_THAN_N, ERR_NON_FINITE_INPUT):
        raise ValueError(msg)  # sklearn's ValueError through graph.py:57 (n_neighbors / check_array)
    if status == ERR_INDEX_OUT_OF_RANGE:
        raise IndexError(msg)  # PyG: index_select on edge_index raises an index error
    if status in (ERR_DOT_PRODUCT, ERR_INVALID_FEATURE):
        raise Exception(msg)  # features.py:56, graph.py:220
    if status == ERR_CUDA:
        msg = f"{msg}: {lib.rgnn_last_cuda_error().decode()}"
    raise RgnnError(status, msg)


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def int32_array(values: Sequence[int]):
    return (C.c_int32 * max(len(values), 1))(*values)


def launch_count() -> int:
    return int(load().rgnn_kernel_launch_count())


def profile_enable(on: bool) -> None:
    load().rgnn_profile_enable(1 if on else 0)


def profile_reset() -> None:
    load().rgnn_profile_reset()


def profile_totals() -> dict:
    """{kernel family: (total device ms, launches)} accumulated since the last reset."""
    lib = load()
    out = {}
    for i in range(lib.rgnn_profile_collect()):
        name, ms, count = C.c_char_p(), C.c_double(), C.c_int64()
        check(lib.rgnn_profile_entry(i, C.byref(name), C.byref(ms), C.byref(count)))
        out[name.value.decode()] = (ms.value, count.value)
    return out
