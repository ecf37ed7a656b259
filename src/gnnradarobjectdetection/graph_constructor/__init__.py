"""Reference path ``gnnradarobjectdetection.graph_constructor`` -> CUDA-backed mirror."""
