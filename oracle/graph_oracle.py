"""Graph-construction oracle (CPU, numpy): neighbour search, edge and node features.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, in vectorised
numpy, what the reference computes with sklearn + a per-edge Python loop:

* ``Graph.build`` / ``__build_knn`` / ``__build_rn``
  (reference src/gnnradarobjectdetection/graph_constructor/graph.py:32-82)
* ``GeometricGraph.extract_node_pair_features`` (graph.py:139-223) with
  ``get_En_equivariant_point_pair_metrics`` (graph_constructor/features.py:6-122)
* ``GeometricGraph.extract_single_node_features`` / ``get_degree``
  (graph.py:225-275, 93-96)
* ``GraphConstructor.build_geometric_graph``
  (preprocessor/radarscenes/dataset_creation.py:187-229; nuScenes twin
  preprocessor/nuscenes/conversion.py:70-109)

The neighbour search itself lives in a third-party dependency that is NOT under
/root/reference: scikit-learn (undeclared and unpinned by the reference; 1.9.0
in this image).  Its published algorithm, restated here from
sklearn/neighbors/_binary_tree.pxi.tp and _base.py:

* distances are *reduced* (squared) Euclidean distances in fp64, accumulated
  left to right over the dimensions: ``d = 0; d += (a_j - b_j) * (a_j - b_j)``
  (no fused multiply-add on the x86-64 wheels);
* k-NN asks for k+1 neighbours of every training point, drops the point's own
  index and lists the remaining k by ascending reduced distance;
* radius: j is a neighbour of i iff j != i and ``rdist(i, j) <= r * r``
  (inclusive).  Within a row sklearn lists them in KD-tree traversal order; the
  canonical order used for parity is ascending column (``canonicalise_rows``);
* the sparse matrix' ``nonzero()`` lists rows ascending, columns in stored order.

Exact distance ties have no specified order in sklearn; the documented tie rule
of this project is ascending (reduced distance, neighbour index).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

EDGE_FEATURE_WIDTH = {
    "point_pair_features": 4,
    "spatial_euclidean_distance": 1,
    "velocity_euclidean_distance": 1,
    "relative_position": 2,
    "relative_velocity": 2,
}

NODE_FEATURE_NAMES = (
    "rcs", "time_index", "degree", "velocity_vector_length",
    "velocity_vector", "spatial_coordinates",
)


# --------------------------------------------------------------------------- #
# neighbour search
# --------------------------------------------------------------------------- #
def reduced_distance_matrix(X: np.ndarray) -> np.ndarray:
    """fp64 squared distances, accumulated dimension by dimension like
    sklearn's ``rdist`` (metrics/_dist_metrics: euclidean_rdist)."""
    X = np.asarray(X, dtype=np.float64)
    n, dims = X.shape
    acc = np.zeros((n, n), dtype=np.float64)
    for j in range(dims):
        diff = X[:, j][:, None] - X[:, j][None, :]
        acc = acc + diff * diff
    return acc


def knn_edges_bruteforce(X: np.ndarray, k: int) -> np.ndarray:
    """All-pairs restatement of graph.py:52-66 (sklearn ``kneighbors_graph``).

    Returns ``E`` of shape [N*k, 2], int64: column 0 the query point, column 1
    its neighbour; rows grouped by ascending query, neighbours by ascending
    (reduced distance, index).  ``k >= N`` raises ``ValueError`` like sklearn.
    """
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if n <= 1:  # graph.py:45 -- nothing is built
        return np.zeros((0, 2), dtype=np.int64)
    if k >= n:
        raise ValueError(
            f"Expected n_neighbors < n_samples_fit, but n_neighbors = {k + 1}, "
            f"n_samples_fit = {n}, n_samples = {n}")
    rd = reduced_distance_matrix(X)
    idx = np.arange(n)
    rd[idx, idx] = np.inf  # self is excluded by index, never by distance
    # lexsort: last key is primary -> (distance, then index)
    cols = np.empty((n, k), dtype=np.int64)
    for i in range(n):
        order = np.lexsort((idx, rd[i]))
        cols[i] = order[:k]
    rows = np.repeat(idx.astype(np.int64), k)
    return np.stack([rows, cols.reshape(-1)], axis=1)


def radius_edges_bruteforce(X: np.ndarray, r: float) -> np.ndarray:
    """All-pairs restatement of graph.py:68-82 (sklearn ``radius_neighbors_graph``),
    in canonical (row, then column ascending) order."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if n <= 1:
        return np.zeros((0, 2), dtype=np.int64)
    rd = reduced_distance_matrix(X)
    reduced_r = np.float64(r) * np.float64(r)
    mask = rd <= reduced_r
    mask[np.arange(n), np.arange(n)] = False
    rows, cols = np.nonzero(mask)
    return np.stack([rows.astype(np.int64), cols.astype(np.int64)], axis=1)


def knn_edges_sklearn(X: np.ndarray, k: int) -> np.ndarray:
    """The reference's own call (graph.py:57-63) without the dense ``toarray``."""
    from sklearn.neighbors import kneighbors_graph
    X = np.asarray(X, dtype=np.float64)
    if X.shape[0] <= 1:
        return np.zeros((0, 2), dtype=np.int64)
    a_sparse = kneighbors_graph(X, k, mode="connectivity", include_self=False)
    rows, cols = a_sparse.nonzero()
    return np.stack([rows.astype(np.int64), cols.astype(np.int64)], axis=1)


def radius_edges_sklearn(X: np.ndarray, r: float) -> np.ndarray:
    """The reference's own call (graph.py:73-79) without the dense ``toarray``.
    Row-internal order is sklearn's KD-tree order (not canonical)."""
    from sklearn.neighbors import radius_neighbors_graph
    X = np.asarray(X, dtype=np.float64)
    if X.shape[0] <= 1:
        return np.zeros((0, 2), dtype=np.int64)
    a_sparse = radius_neighbors_graph(X, r, mode="connectivity", include_self=False)
    rows, cols = a_sparse.nonzero()
    return np.stack([rows.astype(np.int64), cols.astype(np.int64)], axis=1)


def canonicalise_rows(E: np.ndarray, *per_edge: np.ndarray):
    """Sort edges by (row, column) and permute per-edge arrays in lock step.
    Used to compare radius graphs whose row-internal order is a KD-tree artefact."""
    E = np.asarray(E)
    order = np.lexsort((E[:, 1], E[:, 0]))
    out = [E[order]] + [np.asarray(a)[order] for a in per_edge]
    return out[0] if not per_edge else tuple(out)


def kth_gap_is_tie_free(X: np.ndarray, E: np.ndarray, k: int) -> bool:
    """True when, for every query, the k listed reduced distances are strictly
    increasing -- the precondition for "bit-exact edge order" to be well defined
    (checked on the emitted edges; the k/k+1 boundary is covered by comparing two
    independent implementations)."""
    X = np.asarray(X, dtype=np.float64)
    d = np.zeros(E.shape[0])
    for j in range(X.shape[1]):
        t = X[E[:, 0], j] - X[E[:, 1], j]
        d = d + t * t
    d = d.reshape(-1, k)
    return bool(np.all(d[:, 1:] > d[:, :-1])) if k > 1 else True


def batched_edges(frames: Sequence[np.ndarray], routine: str, k: int = 6, r: float = 1.0,
                  backend: str = "bruteforce") -> np.ndarray:
    """Disjoint union of per-frame graphs with node indices offset by the
    cumulative point count (PyG collate, utils/data_handling.py:30)."""
    out, offset = [], 0
    for X in frames:
        n = X.shape[0]
        if n > 1:
            if routine == "knn":
                e = (knn_edges_sklearn if backend == "sklearn" else knn_edges_bruteforce)(X, k)
            elif routine == "radius":
                e = (radius_edges_sklearn if backend == "sklearn" else radius_edges_bruteforce)(X, r)
                e = canonicalise_rows(e)
            else:
                e = np.zeros((0, 2), dtype=np.int64)
            out.append(e + offset)
        offset += n
    return np.concatenate(out, axis=0) if out else np.zeros((0, 2), dtype=np.int64)


# --------------------------------------------------------------------------- #
# edge features
# --------------------------------------------------------------------------- #
def _row_norm(a: np.ndarray) -> np.ndarray:
    acc = np.zeros(a.shape[0], dtype=np.float64)
    for j in range(a.shape[1]):
        acc = acc + a[:, j] * a[:, j]
    return np.sqrt(acc)


def _unit_or_zero(a: np.ndarray) -> np.ndarray:
    """features.py:24-40 / 62-65: exact-zero vectors stay zero, everything else
    is divided by its 2-norm."""
    norm = _row_norm(a)
    is_zero = np.all(a == 0.0, axis=1)
    safe = np.where(is_zero, 1.0, norm)
    out = a / safe[:, None]
    out[is_zero] = 0.0
    return out


def _clamped_dot(u: np.ndarray, w: np.ndarray) -> np.ndarray:
    """features.py:46-56: |dot| in (1, 1+1e-3) snaps to +-1, anything further raises."""
    dot = np.zeros(u.shape[0], dtype=np.float64)
    for j in range(u.shape[1]):
        dot = dot + u[:, j] * w[:, j]
    over = np.abs(dot) > 1.0
    if np.any(over & ~((np.abs(dot) - 1.0) < 1e-3)):
        raise Exception("Error in dot product calculation")
    return np.where(over, np.sign(dot), dot)


def _plain_dot(u: np.ndarray, w: np.ndarray) -> np.ndarray:
    dot = np.zeros(u.shape[0], dtype=np.float64)
    for j in range(u.shape[1]):
        dot = dot + u[:, j] * w[:, j]
    return dot


def _deg(cosine: np.ndarray) -> np.ndarray:
    with np.errstate(invalid="ignore"):
        return np.arccos(cosine) * 180 / np.pi


def _py_min(a, b):  # Python's min(a, b): b only if b < a
    return np.where(b < a, b, a)


def _py_max(a, b):  # Python's max(a, b): b only if b > a
    return np.where(b > a, b, a)


def point_pair_features(p1, p2, v1, v2, mode: str) -> np.ndarray:
    """Vectorised features.py:6-122 for arrays of point pairs ([E, D] each).
    Returns [E, 4]: distance, angle(v1,v2), angle(v1,d), angle(v2,d) in degrees."""
    p1, p2, v1, v2 = (np.asarray(a, dtype=np.float64) for a in (p1, p2, v1, v2))
    v1n, v2n = _unit_or_zero(v1), _unit_or_zero(v2)
    d = _row_norm(p1 - p2)
    theta_v = _deg(_clamped_dot(v1n, v2n))
    if mode == "directed":
        dvec = _unit_or_zero_by_norm(p2 - p1)
        t1 = _deg(_clamped_dot(v1n, dvec))
        t2 = _deg(_clamped_dot(v2n, dvec))
        return np.stack([d, theta_v, t1, t2], axis=1)
    if mode == "undirected":
        d1 = _unit_or_zero_by_norm(p1 - p2)
        d2 = _unit_or_zero_by_norm(p2 - p1)
        t_d1_v1, t_d1_v2 = _deg(_plain_dot(v1n, d1)), _deg(_plain_dot(v2n, d1))
        t_d2_v1, t_d2_v2 = _deg(_plain_dot(v1n, d2)), _deg(_plain_dot(v2n, d2))
        t1 = _py_min(t_d1_v1, t_d2_v1)
        t2 = _py_min(t_d1_v2, t_d2_v2)
        return np.stack([d, theta_v, _py_min(t1, t2), _py_max(t1, t2)], axis=1)
    raise Exception("Invalid edge mode specified")


def _unit_or_zero_by_norm(a: np.ndarray) -> np.ndarray:
    """features.py:62-65: the connection vector is zeroed when its *norm* is 0."""
    norm = _row_norm(a)
    is_zero = norm == 0.0
    out = a / np.where(is_zero, 1.0, norm)[:, None]
    out[is_zero] = 0.0
    return out


def edge_feature_width(features: Sequence[str]) -> int:
    # graph.py:157-166 -- unknown names count one column and fail later in the loop
    return sum(EDGE_FEATURE_WIDTH.get(f, 1) for f in features)


def edge_features(X: np.ndarray, V: np.ndarray, E: np.ndarray,
                  features: Sequence[str], edge_mode: str) -> np.ndarray:
    """Vectorised graph.py:139-223: ``E_feat`` [E, De] fp64, columns in list order.
    Point i is ``E[:, 0]`` (the query), point j is ``E[:, 1]`` (its neighbour)."""
    X = np.asarray(X, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    Xi, Xj = X[E[:, 0]], X[E[:, 1]]
    Vi, Vj = V[E[:, 0]], V[E[:, 1]]
    cols: List[np.ndarray] = []
    for name in features:
        if name == "point_pair_features":
            cols.append(point_pair_features(Xi, Xj, Vi, Vj, edge_mode))
        elif name == "spatial_euclidean_distance":
            cols.append(_row_norm(Xi - Xj)[:, None])
        elif name == "velocity_euclidean_distance":
            cols.append(_row_norm(Vi - Vj)[:, None])
        elif name == "relative_position":
            rel = Xi[:, 0:2] - Xj[:, 0:2]
            cols.append(np.abs(rel) if edge_mode == "undirected" else rel)
        elif name == "relative_velocity":
            rel = Vi[:, 0:2] - Vj[:, 0:2]
            cols.append(np.abs(rel) if edge_mode == "undirected" else rel)
        else:
            raise Exception("Invalid feature specified")
    if not cols:
        return np.empty((E.shape[0], 0), dtype=np.float64)
    return np.concatenate(cols, axis=1)


# --------------------------------------------------------------------------- #
# node features
# --------------------------------------------------------------------------- #
def undirected_degree(E: np.ndarray, n: int) -> np.ndarray:
    """graph.py:93-96 -- networkx builds an *undirected* graph from the
    (asymmetric) adjacency matrix: degree = |out-neighbours U in-neighbours|."""
    if E.shape[0] == 0:
        return np.zeros(n, dtype=np.int64)
    a = np.minimum(E[:, 0], E[:, 1])
    b = np.maximum(E[:, 0], E[:, 1])
    und = np.unique(a * n + b)  # one entry per unordered pair
    lo, hi = und // n, und % n
    return np.bincount(lo, minlength=n) + np.bincount(hi, minlength=n)


def time_index(timestamp: np.ndarray) -> np.ndarray:
    """dataset_creation.py:214-223 -- rank of every timestamp among the sorted
    distinct values, same shape/dtype as the input."""
    ts = np.asarray(timestamp)
    _, inverse = np.unique(ts, return_inverse=True)
    return inverse.reshape(ts.shape).astype(ts.dtype)


def node_features(X: np.ndarray, V: np.ndarray, F: Dict[str, np.ndarray], E: Optional[np.ndarray],
                  features: Sequence[str]) -> np.ndarray:
    """Vectorised graph.py:225-275: ``X_feat`` [N, Fn], columns in list order."""
    n = X.shape[0]
    blocks = []
    for name in features:
        if name == "rcs":
            blocks.append(np.asarray(F["rcs"]).reshape(n, -1))
        elif name == "time_index":
            blocks.append(np.asarray(F["time_index"]).reshape(n, -1))
        elif name == "degree":
            blocks.append(undirected_degree(E, n).reshape(n, 1))
        elif name == "velocity_vector_length":
            blocks.append(_row_norm(np.asarray(V, dtype=np.float64)).reshape(n, 1))
        elif name == "velocity_vector":
            blocks.append(np.asarray(V))
        elif name == "spatial_coordinates":
            blocks.append(np.asarray(X))
        # graph.py:252-275 has no else branch: an unknown name silently re-appends
        # the previous block; callers never do that, so it is not restated.
    return np.concatenate(blocks, axis=1) if blocks else np.empty((n, 0))


# --------------------------------------------------------------------------- #
# GraphConstructor.build_geometric_graph (the boundary function)
# --------------------------------------------------------------------------- #
def build_geometric_graph(*, X_cc: np.ndarray, V_cc: np.ndarray, rcs: Optional[np.ndarray],
                          timestamp: Optional[np.ndarray], algorithm: str, k: Optional[int],
                          r: Optional[float], node_feature_names: Sequence[str],
                          edge_feature_names: Sequence[str], edge_mode: str,
                          distance_definition: str, backend: str = "bruteforce"
                          ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """dataset_creation.py:187-229 on plain arrays.  Returns (E, E_feat, X_feat);
    radius graphs come back in canonical row order."""
    if distance_definition == "X":
        basis = X_cc
    elif distance_definition == "XV":
        basis = np.concatenate((X_cc, V_cc), axis=1)
    else:
        raise Exception("Invalid distance definition")
    F = {"rcs": rcs}
    if "time_index" in node_feature_names:
        F["time_index"] = time_index(timestamp)
    E = batched_edges([basis], algorithm, k=k if k is not None else 6,
                      r=r if r is not None else 1.0, backend=backend)
    E_feat = edge_features(X_cc, V_cc, E, edge_feature_names, edge_mode)
    X_feat = node_features(X_cc, V_cc, F, E, node_feature_names)
    return E, E_feat, X_feat
