"""radargnn_b200 -- B200-native (sm_100a) graph construction + MPNN forward for RadarGNN.

Drop-in for ONE hot path of TUMFTM/RadarGNN: the radius / k-NN neighbour search that
emits ``edge_index`` + edge attributes, and the stacked MPNN layers' forward.  The
compute lives in hand-written CUDA kernels behind a C ABI (``include/rgnn.h``,
``radargnn_b200/csrc``); the Python modules mirror the reference's own interface
(``graph_constructor``, ``gnn``, ``preprocessor``).  There is no CPU fallback: ops
raise if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
