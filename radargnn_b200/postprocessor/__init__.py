"""Mirror of the reference's postprocessor package for the part that sits next to the hot path: the box
suppression of ``postprocessor/postprocessing.py:336-435`` and the nearest-neighbour lookup of the "en" box
representation (:233-237, :468-472), both on the device."""
from .postprocessing import BoxSuppressor, nearest_neighbor_positions  # noqa: F401
