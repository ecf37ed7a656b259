"""Reference path ``gnnradarobjectdetection.gnn`` -> CUDA-backed mirror (radargnn_b200.gnn)."""
