"""Import the *reference's own* hot-path modules from /root/reference (this
container only -- the GPU box has no /root/reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the local
(``-m "not gpu"``) tests that validate the numpy/torch restatements against the
real reference code when it is present.  Nothing is copied: the modules are
imported in place with three environment shims,

* ``matplotlib`` / ``matplotlib.pyplot`` stubbed (graph.py:5 imports pyplot, only
  ``show`` uses it),
* ``networkx.from_numpy_matrix`` aliased to ``from_numpy_array`` (graph.py:94;
  removed in networkx 3.x),
* ``torch_geometric`` replaced by ``oracle/pyg_shim.py`` (not installable here).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "gnnradarobjectdetection"))


def _install_env_shims() -> None:
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    import networkx as nx
    if not hasattr(nx, "from_numpy_matrix"):
        nx.from_numpy_matrix = nx.from_numpy_array
    from . import pyg_shim
    pyg_shim.install()


def load(module: str):
    """``load("graph_constructor.graph")`` -> the reference module object."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    _install_env_shims()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    return importlib.import_module("gnnradarobjectdetection." + module)
