"""Reference preprocessor/radar_point_cloud.py -> radargnn_b200.preprocessor.radar_point_cloud."""
from radargnn_b200.preprocessor.radar_point_cloud import RadarPointCloud  # noqa: F401

__all__ = ["RadarPointCloud"]
