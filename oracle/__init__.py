"""CPU oracle for the RadarGNN hot path (graph build + MPNN forward).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline -- never as a code path of
``radargnn_b200``.  The product fails loudly when its CUDA library is missing.

Parity status (details in DESIGN.md, section "Oracle"):

* graph half  -- PINNED.  ``graph_oracle`` is checked against (a) the known
  answers of the reference's own tests (test/test_graph_constructor.py,
  test/test_preprocessor.py:207-257) and (b) golden vectors produced by
  importing the reference's ``graph_constructor/graph.py`` + ``features.py``
  from /root/reference (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* MPNN half   -- PINNED on the reference's known answers
  (test/test_gnn.py:9-25,79-116,119-172,175-221) and on golden vectors
  produced by running the reference's own ``mpnn_layers.py`` /
  ``gnn_models.py`` classes on top of a minimal stand-in for the missing
  torch_geometric package (``oracle/pyg_shim.py``; PyG 2.1.0 itself is not
  installable here -- no network).  Aggregation semantics of PyG's
  ``propagate`` (torch_scatter) for ``add``/``mean``/zero in-degree are
  therefore restated from the published behaviour, not executed:
  "parity unpinned by reference tests" for those cases.
"""
