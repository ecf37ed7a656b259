"""Synthetic radar frames of the shapes BASELINE.json / SURVEY.md section 8(d) name.

There is no dataset in this environment (RadarScenes / nuScenes need downloads),
so benchmarks and parity tests run on seeded synthetic point clouds.  All
coordinates are float32-representable values (stored as float64 where the
reference's API wants fp64) so that the fp64 neighbour search sees identical
inputs on CPU and GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass
class SyntheticFrame:
    """One radar frame with the fields of the reference's ``RadarPointCloud``
    that the graph constructor reads (preprocessor/radar_point_cloud.py:18-37)."""
    X_cc: np.ndarray              # [N, 2] float64 (float32-representable)
    V_cc_compensated: np.ndarray  # [N, 2] float64
    rcs: np.ndarray               # [N, 1] float64
    timestamp: np.ndarray         # [N, 1] float64

    @property
    def n(self) -> int:
        return self.X_cc.shape[0]


def _f32(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float32).astype(np.float64)


def radar_frame(n: int = 300, seed: int = 0, extent: Tuple[float, float] = (100.0, 100.0),
                cluster_fraction: float = 0.3, static_fraction: float = 0.6,
                n_timestamps: int = 8) -> SyntheticFrame:
    """RadarScenes-like frame: x in [0, extent_x),
    y in [-extent_y/2, extent_y/2), ``cluster_fraction`` of the points in Gaussian
    object-like clusters (sigma = 1 m), exactly-zero velocity for ``static_fraction``."""
    rng = np.random.default_rng(seed)
    n_cl = int(round(n * cluster_fraction))
    n_bg = n - n_cl
    bg = np.stack([rng.uniform(0, extent[0], n_bg),
                   rng.uniform(-extent[1] / 2, extent[1] / 2, n_bg)], axis=1)
    n_centres = max(1, n_cl // 12)
    centres = np.stack([rng.uniform(5, extent[0] - 5, n_centres),
                        rng.uniform(-extent[1] / 2 + 5, extent[1] / 2 - 5, n_centres)], axis=1)
    which = rng.integers(0, n_centres, n_cl)
    cl = centres[which] + rng.normal(0, 1.0, (n_cl, 2))
    X = np.concatenate([bg, cl], axis=0)
    V = rng.normal(0, 5.0, (n, 2))
    V[rng.random(n) < static_fraction] = 0.0
    order = rng.permutation(n)
    X, V = X[order], V[order]
    rcs = rng.normal(-5, 10, (n, 1))
    ts = rng.integers(0, n_timestamps, (n, 1)).astype(np.float64) * 17000.0 + 1.0e9
    return SyntheticFrame(_f32(X), _f32(V), _f32(rcs), ts)


def uniform_square(n: int, seed: int = 0, density: float = 1.0) -> SyntheticFrame:
    """``n`` points uniform on a square of side sqrt(n / density) metres
    (config 2 / headline: 10 k points on 100 m x 100 m, 100 k on 316 m x 316 m)."""
    rng = np.random.default_rng(seed)
    side = float(np.sqrt(n / density))
    X = rng.uniform(0, side, (n, 2))
    V = rng.normal(0, 5.0, (n, 2))
    rcs = rng.normal(-5, 10, (n, 1))
    ts = np.zeros((n, 1))
    return SyntheticFrame(_f32(X), _f32(V), _f32(rcs), ts)


def nuscenes_frame(n: int = 2000, seed: int = 0) -> SyntheticFrame:
    """nuScenes-like multi-sweep frame (config 4): x, y uniform in [-100, 100) m,
    30 distinct sweep timestamps (5 radars x 6 sweeps)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-100, 100, (n, 2))
    V = rng.normal(0, 4.0, (n, 2))
    V[rng.random(n) < 0.5] = 0.0
    rcs = rng.normal(5, 8, (n, 1))
    ts = rng.integers(0, 30, (n, 1)).astype(np.float64) * 77000.0 + 1.5e9
    return SyntheticFrame(_f32(X), _f32(V), _f32(rcs), ts)


def frame_batch(frames: List[SyntheticFrame]):
    """Concatenate frames into the disjoint-union layout the kernels take:
    positions/velocities stacked on dim 0 plus ``frame_ptr`` [F+1] (int64)."""
    X = np.concatenate([f.X_cc for f in frames], axis=0)
    V = np.concatenate([f.V_cc_compensated for f in frames], axis=0)
    ptr = np.zeros(len(frames) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([f.n for f in frames])
    return X, V, ptr


def node_embeddings(n: int, channels: int, seed: int = 0) -> np.ndarray:
    """x0 ~ N(0, 1), float32 [n, channels] (config 2: channels = 64)."""
    rng = np.random.default_rng(seed + 7919)
    return rng.standard_normal((n, channels), dtype=np.float32)
