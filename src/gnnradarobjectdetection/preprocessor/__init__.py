"""Reference path ``gnnradarobjectdetection.preprocessor`` -> CUDA-backed mirror (hot-path part)."""
