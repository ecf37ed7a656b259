"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rgnn.h declares,
and its host-only entry points (sizes, counts, status text) behave.  No compute call needs a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from radargnn_b200 import _lib, build
    build.build_library()          # nvcc cross-compiles for sm_100a without a GPU
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rgnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rgnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from radargnn_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/rgnn.h but not exported"
    # the ctypes binding covers the whole header, nothing more
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "rgnn.h")).read()
    for needle in ("graph.py:32-82", "graph.py:139-223", "features.py:6-122", "graph.py:93-96", "graph.py:225-275",
                   "mpnn_layers.py:86-101", "gnn_models.py:126-128"):
        assert needle in text


def test_status_strings_match_reference_messages(lib):
    assert lib.rgnn_abi_version() == 4
    assert lib.rgnn_status_string(0) == b"ok"
    assert lib.rgnn_status_string(2) == b"Expected n_neighbors < n_samples_fit"   # sklearn's ValueError text
    assert lib.rgnn_status_string(5) == b"Error in dot product calculation"        # features.py:56
    assert lib.rgnn_status_string(6) == b"Invalid feature specified"               # graph.py:220
    assert lib.rgnn_status_string(9) == b"Input contains NaN or infinity"          # sklearn check_array
    assert b"outside [0, N)" in lib.rgnn_status_string(10)


def test_knn_edge_count_and_k_check(lib):
    ptr = np.array([0, 300, 301, 301, 421], dtype=np.int64)   # frames of 300, 1, 0, 120 points
    st = C.c_int(0)
    assert lib.rgnn_knn_edge_count(ptr.ctypes.data, 4, 16, C.byref(st)) == (300 + 120) * 16 and st.value == 0
    assert lib.rgnn_knn_edge_count(ptr.ctypes.data, 4, 64, C.byref(st)) == (300 + 120) * 64
    ptr2 = np.array([0, 17, 34], dtype=np.int64)
    assert lib.rgnn_knn_edge_count(ptr2.ctypes.data, 2, 17, C.byref(st)) == -1 and st.value == 2
    assert lib.rgnn_knn_edge_count(ptr2.ctypes.data, 2, 0, C.byref(st)) == -1 and st.value == 1


def test_feature_widths(lib):
    from radargnn_b200 import _lib
    ids = _lib.int32_array([0, 1, 2, 3, 4])
    assert lib.rgnn_edge_feature_width(ids, 5) == 10          # graph.py:157-166
    assert lib.rgnn_edge_feature_width(_lib.int32_array([3]), 1) == 2
    assert lib.rgnn_edge_feature_width(_lib.int32_array([9]), 1) == -1
    assert lib.rgnn_node_feature_width(_lib.int32_array([0, 1, 2, 3, 4, 5]), 6) == 8
    assert lib.rgnn_node_feature_width(_lib.int32_array([7]), 1) == -1


def test_workspace_sizes_are_monotone_and_aligned(lib):
    from radargnn_b200 import _lib
    small = lib.rgnn_graph_workspace_bytes(1000, 1)
    big = lib.rgnn_graph_workspace_bytes(100_000, 1)
    assert 0 < small < big and big < 64 * 100_000 + (1 << 20)
    assert lib.rgnn_csc_workspace_bytes(1000, 16000) > 0
    d = _lib.ConvDesc()
    d.conv_type, d.aggr, d.in_channels, d.out_channels, d.edge_dim = 0, 0, 64, 64, 2
    d.pre_layers = d.post_layers = 1
    w1 = lib.rgnn_conv_workspace_bytes(C.byref(d), 100_000, 1_600_000)
    # B, M in the split layout ([N, 128] main + [N, 4] tail, fp32) + edge attributes in slot order
    assert w1 >= 2 * 100_000 * 132 * 4 + 1_600_000 * 2 * 4
    # tensor-core weight images of the factored path: W_s [144 x 64] and update [64 x 224], hi + lo
    assert lib.rgnn_conv_packed_bytes(C.byref(d)) >= 2 * 4 * (144 * 64 + 64 * 224)
    d.pre_layers = 2   # general path: two [E, 132] per-edge activation buffers, no tensor-core images
    assert lib.rgnn_conv_packed_bytes(C.byref(d)) == 0
    assert lib.rgnn_conv_workspace_bytes(C.byref(d), 100_000, 1_600_000) > w1 + 2 * 1_600_000 * 132 * 4 - (4 << 20)
    d.aggr = 17
    assert lib.rgnn_conv_workspace_bytes(C.byref(d), 10, 10) == 0
    assert lib.rgnn_batchnorm_workspace_bytes(100_000, 64) > 0
    assert lib.rgnn_sum_workspace_bytes() > 0


def test_compute_entry_points_fail_loudly_without_a_device(lib):
    import torch
    from radargnn_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sm = C.c_int32(0)
    assert lib.rgnn_device_info(C.byref(sm), None, None) == 8          # RGNN_ERR_NO_DEVICE
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.knn_graph(torch.zeros(4, 2), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.csc_build(torch.zeros(2, 3, dtype=torch.int64), 4)
    from radargnn_b200.graph_constructor.graph import Graph
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Graph().build(np.zeros((3, 2)), "knn", k=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.HostPipeline(None)
    assert lib.rgnn_pipeline_wait_host(0) != 0 and lib.rgnn_pipeline_wait_host(99) == 1   # no device / slot out of range


def test_product_never_imports_the_oracle():
    for pkg in (os.path.join(ROOT, "radargnn_b200"), os.path.join(ROOT, "src")):
        for dirpath, _, files in os.walk(pkg):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f"{f} mentions the oracle"
