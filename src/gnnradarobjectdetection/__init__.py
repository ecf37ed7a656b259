"""``gnnradarobjectdetection`` -- the reference's import path, backed by the B200 kernels.

BASELINE.json's north star asks that the kernels "drop in under src/gnnradarobjectdetection":
putting this ``src`` directory on ``sys.path`` (or ``pip install -e`` it in place of the reference's
``src``) makes the reference's own imports for the hot path

    from gnnradarobjectdetection.graph_constructor.graph import GeometricGraph
    from gnnradarobjectdetection.gnn.mpnn_layers import MPNNConv, RadarPointGNNConv
    from gnnradarobjectdetection.gnn.gnn_models import DetNetBasic, get_mlp
    from gnnradarobjectdetection.preprocessor.radarscenes.dataset_creation import GraphConstructor

resolve to the CUDA-backed classes of ``radargnn_b200`` (same names, signatures, attributes and
parameter names as reference src/gnnradarobjectdetection/...).  Only the modules of the hot path
 exist here; everything else of the reference (trainer, dataset I/O, evaluation)
keeps coming from the reference itself.
"""
import os as _os
import sys as _sys

# the kernels' package lives beside `src/` in this repository
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path and _os.path.isdir(_os.path.join(_ROOT, "radargnn_b200")):
    _sys.path.insert(0, _ROOT)

import radargnn_b200 as _impl  # noqa: E402

__version__ = _impl.__version__
BACKEND = "radargnn_b200 (sm_100a CUDA kernels behind librgnn_b200.so)"
