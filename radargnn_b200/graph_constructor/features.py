"""Mirror of reference graph_constructor/features.py: the point-pair-feature function, computed
by the CUDA edge-feature kernel (fp64) instead of numpy."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import ops


def get_En_equivariant_point_pair_metrics(p1: np.ndarray, p2: np.ndarray, v1: np.ndarray, v2: np.ndarray,
                                          mode: str) -> Tuple[float]:
    """Point-pair features of two radar points (reference features.py:6-122).

    ``p1, p2, v1, v2`` are ``[D, 1]`` column vectors as in the reference.  Returns
    ``(d, theta_v1_v2, theta_d_v_min, theta_d_v_max)`` in degrees.  Raises
    ``Exception("Error in dot product calculation")`` like features.py:56.
    """
    if mode not in ("directed", "undirected"):
        raise UnboundLocalError("mode must be 'directed' or 'undirected'")  # the reference falls through
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    pos = torch.tensor(np.stack([np.asarray(p1, dtype=np.float64).reshape(-1),
                                 np.asarray(p2, dtype=np.float64).reshape(-1)]), device=dev)
    vel = torch.tensor(np.stack([np.asarray(v1, dtype=np.float64).reshape(-1),
                                 np.asarray(v2, dtype=np.float64).reshape(-1)]), device=dev)
    edge = torch.tensor([[0], [1]], dtype=torch.int64, device=dev)
    out = ops.edge_features(pos, vel, edge, ["point_pair_features"], mode, out_dtype=torch.float64)
    d, t12, tmin, tmax = out[0].tolist()
    return d, t12, tmin, tmax
