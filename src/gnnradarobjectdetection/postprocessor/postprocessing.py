"""Reference postprocessor/postprocessing.py:336-435 (BoxSuppressor) and the nearest-neighbour lookup of :233-237 /
:468-472 -> radargnn_b200.postprocessor (box matrices in, rgnn_nms / rgnn_nearest_neighbor on the GPU)."""
from radargnn_b200.postprocessor.postprocessing import BoxSuppressor, nearest_neighbor_positions  # noqa: F401

__all__ = ["BoxSuppressor", "nearest_neighbor_positions"]
