"""Per-edge restatement of the reference's CPU path, for the *timed* CPU baseline.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is Python and
cannot travel to the GPU box (/root/reference does not exist there), so
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs time this port of
the same algorithm with the same cost structure:

* neighbour search: the very sklearn calls of graph.py:57-58 / 73-74 (1 thread,
  ``n_jobs=None``), including the dense ``toarray()`` (graph.py:59,75) when the
  frame is small enough for it to be feasible;
* edge attributes: one Python iteration per edge on [D,1] column vectors with
  ``np.linalg.norm(..., ord=2)`` / ``np.dot`` / ``np.arccos`` exactly as
  graph.py:172-223 and features.py:24-122 structure the work;
* node ``degree`` through networkx on the dense adjacency (graph.py:93-96);
* MPNN forward: ``oracle.mpnn_oracle`` on all host cores (the PyG CPU ops).

It is also checked against ``graph_oracle.edge_features`` in the tests, which
makes it a second, independent statement of the feature arithmetic.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def _unit(v: np.ndarray) -> np.ndarray:
    if np.count_nonzero(v) == 0:
        return np.zeros_like(v, dtype=np.float64)
    return v / np.linalg.norm(v, ord=2)


def _snap(dot: float) -> float:
    if abs(dot) > 1:
        if abs(dot) - 1 < 1e-3:
            return 1.0 if dot > 0 else -1.0
        raise Exception("Error in dot product calculation")
    return dot


def _angle(dot: float) -> float:
    return float(np.arccos(dot)) * 180 / np.pi


def pair_metrics(p1, p2, v1, v2, mode: str):
    """features.py:6-122 for one point pair of [D,1] column vectors."""
    u1, u2 = _unit(v1), _unit(v2)
    dist = np.linalg.norm(p1 - p2, ord=2)
    theta_v = _angle(_snap(float(np.dot(u1.T, u2)[0, 0])))
    if mode == "directed":
        span = np.linalg.norm(p2 - p1, ord=2)
        dvec = np.zeros_like(p2, dtype=np.float64) if span == 0 else (p2 - p1) / span
        t1 = _angle(_snap(float(np.dot(u1.T, dvec)[0, 0])))
        t2 = _angle(_snap(float(np.dot(u2.T, dvec)[0, 0])))
        return dist, theta_v, t1, t2
    if mode == "undirected":
        span = np.linalg.norm(p1 - p2, ord=2)
        fwd = np.zeros_like(p2, dtype=np.float64) if span == 0 else (p1 - p2) / span
        span = np.linalg.norm(p2 - p1, ord=2)
        bwd = np.zeros_like(p2, dtype=np.float64) if span == 0 else (p2 - p1) / span
        with np.errstate(invalid="ignore"):
            a1 = min(_angle(float(np.dot(u1.T, fwd)[0, 0])), _angle(float(np.dot(u1.T, bwd)[0, 0])))
            a2 = min(_angle(float(np.dot(u2.T, fwd)[0, 0])), _angle(float(np.dot(u2.T, bwd)[0, 0])))
        return dist, theta_v, min(a1, a2), max(a1, a2)
    raise Exception("Invalid edge mode specified")


def edge_feature_loop(X: np.ndarray, V: np.ndarray, E: np.ndarray,
                      features: Sequence[str], edge_mode: str) -> np.ndarray:
    """graph.py:139-223 -- the per-edge loop."""
    width = sum(4 if f == "point_pair_features" else
                2 if f in ("relative_position", "relative_velocity") else 1 for f in features)
    out = np.empty([E.shape[0], width])
    dx, dv = X.shape[1], V.shape[1]
    for row, (a, b) in enumerate(E):
        xa, xb = X[int(a), :].reshape(dx, 1), X[int(b), :].reshape(dx, 1)
        va, vb = V[int(a), :].reshape(dv, 1), V[int(b), :].reshape(dv, 1)
        vals: List[float] = []
        for name in features:
            if name == "point_pair_features":
                vals.extend(pair_metrics(xa, xb, va, vb, edge_mode))
            elif name == "spatial_euclidean_distance":
                vals.append(np.linalg.norm(xa - xb, ord=2))
            elif name == "velocity_euclidean_distance":
                vals.append(np.linalg.norm(va - vb, ord=2))
            elif name in ("relative_position", "relative_velocity"):
                pa, pb = (xa, xb) if name == "relative_position" else (va, vb)
                d0, d1 = pa[0, 0] - pb[0, 0], pa[1, 0] - pb[1, 0]
                if edge_mode == "undirected":
                    d0, d1 = abs(d0), abs(d1)
                vals.extend([d0, d1])
            else:
                raise Exception("Invalid feature specified")
        out[row, :] = np.array(vals).reshape(1, len(vals))
    return out


def build_edges_like_reference(X: np.ndarray, routine: str, k: int, r: float, dense: bool):
    """graph.py:52-82: sklearn call, optional dense ``toarray`` and ``nonzero``."""
    from sklearn.neighbors import kneighbors_graph, radius_neighbors_graph
    if routine == "knn":
        a_sparse = kneighbors_graph(X, k, mode="connectivity", include_self=False)
    else:
        a_sparse = radius_neighbors_graph(X, r, mode="connectivity", include_self=False)
    A = a_sparse.toarray() if dense else None
    e1 = a_sparse.nonzero()[0].reshape(-1, 1)
    e2 = a_sparse.nonzero()[1].reshape(-1, 1)
    return np.concatenate((e1, e2), axis=1), A


def degree_like_reference(A: np.ndarray) -> np.ndarray:
    """graph.py:93-96 via networkx (``from_numpy_array`` is the 3.x name of
    ``from_numpy_matrix``)."""
    import networkx as nx
    G = nx.from_numpy_array(A)
    return np.array([val for (_, val) in G.degree()]).reshape(-1, 1)
