"""Mirror of reference graph_constructor/graph.py: ``Graph`` and ``GeometricGraph`` with the same
attributes and methods, backed by the CUDA neighbour search and feature kernels.

Differences that are deliberate:
* ``A`` (dense N x N adjacency) is materialised lazily on attribute access -- the reference
  builds it eagerly with ``toarray()`` (graph.py:59,75), O(N^2) memory;
* radius graphs list a row's neighbours in ascending column order; the reference inherits
  sklearn's KD-tree traversal order (same edge set, different order inside a row).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .. import ops


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("radargnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Graph():
    """General graph (reference graph.py:10-101).

    Attributes: ``X_feat`` node features, ``E_feat`` edge features, ``A`` adjacency matrix,
    ``E`` edge connection matrix ``[E, 2]`` (column 0 = query point, column 1 = neighbour).
    """

    def __init__(self):
        self.X_feat = None
        self.E_feat = None
        self._A = None
        self._E = None
        self._n = 0
        self._edge_index_dev: Optional[torch.Tensor] = None  # int64 [2, E] on the device, mirrors _E
        self._dev_in_E_order = False   # the device copy lists the edges in E's own row order (set by build())

    # -- edge list: assigning E by hand drops the device copy and the cached adjacency matrix -----
    @property
    def E(self):
        if self._E is None and self._edge_index_dev is not None and self._dev_in_E_order:
            self._E = self._edge_index_dev.t().contiguous().cpu().numpy()   # adopted device edges: copied on first use
        return self._E

    @E.setter
    def E(self, value):
        self._E = value
        self._edge_index_dev = None
        self._dev_in_E_order = False
        self._A = None

    # -- adjacency matrix, built only when somebody looks at it ----------------------------------
    @property
    def A(self):
        if self._A is None and self.E is not None:
            A = np.zeros((self._n, self._n))
            A[self.E[:, 0], self.E[:, 1]] = 1.0
            self._A = A
        return self._A

    @A.setter
    def A(self, value):
        self._A = value

    def build(self, X: np.ndarray, routine: str, k: int = 6, r: float = 1) -> None:
        """Creates the edges between the points ``X`` (all dimensions of ``X`` enter the distance).
        Nothing is built for fewer than two points or an unknown routine (reference graph.py:45-50)."""
        X = np.asarray(X)
        if X.shape[0] > 1:
            if routine == "knn":
                self.__build_knn(X, k)
            elif routine == "radius":
                self.__build_rn(X, r)

    def __set_edges(self, X: np.ndarray, edge_index: torch.Tensor) -> None:
        self._n = X.shape[0]
        self.E = edge_index.t().contiguous().cpu().numpy()
        self._edge_index_dev = edge_index   # after the setter (which drops any stale device copy)
        self._dev_in_E_order = True

    def adopt_edges(self, n_nodes: int, edge_index: Optional[torch.Tensor], E: Optional[np.ndarray] = None) -> None:
        """Takes over an edge list built elsewhere (the batched constructor): a device ``edge_index`` [2, E] whose
        rows are grouped by query point, or a host ``E`` [E, 2]."""
        self._n = int(n_nodes)
        if edge_index is not None:
            self.E = None                      # the host copy is made when somebody reads E
            self._edge_index_dev = edge_index
            self._dev_in_E_order = True
        else:
            self.E = np.ascontiguousarray(E, dtype=np.int64)

    def __build_knn(self, X: np.ndarray, k: int) -> None:
        basis = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64), device=_device())
        self.__set_edges(X, ops.knn_graph(basis, k))

    def __build_rn(self, X: np.ndarray, r: float) -> None:
        basis = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64), device=_device())
        self.__set_edges(X, ops.radius_graph(basis, r))

    def _built_edges_current(self) -> bool:
        """The device edge_index of the last build() still describes ``E`` (rows in E's order)."""
        return self._edge_index_dev is not None and self._dev_in_E_order

    def add_node_features(self, feat: np.ndarray) -> None:
        if self.X_feat is None:
            self.X_feat = feat
        else:
            if feat.shape[0] == self.X_feat.shape[0]:
                self.X_feat = np.concatenate((self.X_feat, feat), axis=1)
            else:
                raise Exception("Feature dimension not compatible")

    def _edge_index(self) -> torch.Tensor:
        if self._edge_index_dev is None:
            # E was assigned by hand (or edited in place: callers must re-assign it): group rows by source
            # for the degree kernel
            E = np.asarray(self.E, dtype=np.int64)
            order = np.lexsort((E[:, 1], E[:, 0]))
            self._edge_index_dev = torch.as_tensor(np.ascontiguousarray(E[order].T), device=_device())
            self._n = max(self._n, int(E.max()) + 1 if E.size else 0)
        return self._edge_index_dev

    def _ensure_edges(self) -> None:
        """Adopts the edges of a hand-assigned adjacency matrix when no edge list exists."""
        if self._E is None and self._edge_index_dev is None:
            A = self._A
            if A is None:
                raise AttributeError("graph has not been built")
            rows, cols = np.nonzero(A)
            self.E = np.stack([rows, cols], axis=1).astype(np.int64)
            self._A = A       # the E setter dropped it
            self._n = A.shape[0]

    def get_degree(self) -> list:
        """Degree of every node of the undirected graph over ``A`` (reference graph.py:93-96)."""
        return self._degree_array().tolist()

    def _degree_array(self) -> np.ndarray:
        """``get_degree`` as an int64 array (no Python list in between)."""
        self._ensure_edges()
        n = self._n if self._A is None else self._A.shape[0]
        return ops.undirected_degree(self._edge_index(), n).cpu().numpy().astype(np.int64, copy=False)

    def show(self, node_size: float = 60) -> None:  # pragma: no cover - plotting only
        import matplotlib.pyplot as plt
        import networkx as nx
        G = nx.from_numpy_array(self.A)
        fig, ax = plt.subplots()
        nx.draw(G, ax=ax, node_size=node_size)


class GeometricGraph(Graph):
    """Geometric graph separating spatial (``X``), velocity (``V``) and invariant (``F``) node
    data (reference graph.py:104-302)."""

    def __init__(self):
        super().__init__()
        self.X = None
        self.V = None
        self.F = None

    def add_invariant_feature(self, name: str, F_add: np.ndarray) -> None:
        if self.F is None:
            self.F = {name: F_add}
        else:
            self.F[name] = F_add

    def add_degree_to_inv_features(self) -> None:
        deg = self._degree_array()
        self.add_invariant_feature("degree", deg.reshape(len(deg), 1))

    def extract_node_pair_features(self, features: List[str], edge_mode: str) -> None:
        """Edge feature matrix ``E_feat`` [E, De] (fp64), columns in list order
        (reference graph.py:139-223)."""
        dev = _device()
        pos = torch.as_tensor(np.ascontiguousarray(self.X, dtype=np.float64), device=dev)
        vel = torch.as_tensor(np.ascontiguousarray(self.V, dtype=np.float64), device=dev)
        if self._built_edges_current():
            ei = self._edge_index_dev
        else:
            ei = torch.as_tensor(np.ascontiguousarray(np.asarray(self.E, dtype=np.int64).T), device=dev)
        feat = ops.edge_features(pos, vel, ei, features, edge_mode, out_dtype=torch.float64)
        feat = feat.cpu().numpy()
        if self.E_feat is None:
            self.E_feat = feat
        else:
            self.E_feat[:, :] = feat

    def extract_single_node_features(self, features: List[str]) -> None:
        """Node feature matrix ``X_feat`` [N, Fn], columns in list order (reference graph.py:225-275)."""
        dev = _device()
        n = np.asarray(self.X).shape[0]
        # the node-feature kernel reads exactly two columns of X / V: check the ones the request needs
        for arr, label, users in ((self.X, "X", ("spatial_coordinates",)),
                                  (self.V, "V", ("velocity_vector", "velocity_vector_length"))):
            if any(f in users for f in features) and (arr is None or np.asarray(arr).ndim != 2 or np.asarray(arr).shape != (n, 2)):
                raise ValueError(f"{label} must be an [N, 2] array for the node features {users}")
        names = ("rcs", "time_index", "degree", "velocity_vector_length", "velocity_vector", "spatial_coordinates")
        known = []
        for f in features:
            if f in names:
                known.append(f)
            elif known:
                known.append(known[-1])   # reference quirk (graph.py:253-275): an unknown name re-appends the previous feature
            else:
                raise UnboundLocalError("local variable 'feat' referenced before assignment")  # what the reference raises
        for f in features:
            if f == "degree":
                self.add_degree_to_inv_features()   # once per occurrence, like the reference (graph.py:246-248)
        F = self.F or {}

        def dev_f64(a):
            return None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(n, -1), device=dev)

        deg = F.get("degree")
        out = ops.node_features(
            known, n, dev, rcs=dev_f64(F.get("rcs")) if "rcs" in known else None,
            time_index=dev_f64(F.get("time_index")) if "time_index" in known else None,
            degree=None if deg is None or "degree" not in known else torch.as_tensor(np.asarray(deg).reshape(-1), device=dev),
            pos=dev_f64(self.X), vel=dev_f64(self.V), out_dtype=torch.float64)
        feat = out.cpu().numpy()
        if feat.shape[1]:
            self.X_feat = feat if self.X_feat is None else np.concatenate((self.X_feat, feat), axis=1)

    def show(self, *args, **kwargs) -> None:  # pragma: no cover - plotting only
        super().show(*args, **kwargs)
