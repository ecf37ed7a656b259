"""Reference preprocessor/configs.py:7-27 -> radargnn_b200.preprocessor.configs."""
from radargnn_b200.preprocessor.configs import GraphConstructionConfiguration  # noqa: F401

__all__ = ["GraphConstructionConfiguration"]
