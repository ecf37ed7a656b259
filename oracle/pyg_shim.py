"""Minimal stand-in for the parts of torch_geometric 2.1.0 the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch_geometric /
torch_scatter are pinned by the reference (Dockerfile:21-26: torch-geometric
2.1.0.post1, torch-scatter 2.0.9) but are absent from /root/reference and not
installable here (no network).  This module restates their *published*
behaviour for exactly the symbols the reference touches, so that the
reference's own ``gnn/mpnn_layers.py`` and ``gnn/gnn_models.py`` can be executed
unmodified to produce golden vectors (oracle/make_golden.py):

* ``MessagePassing(aggr=...)`` with ``flow="source_to_target"``:
  ``x_j = x[edge_index[0]]``, ``x_i = x[edge_index[1]]``; messages are reduced
  at ``edge_index[1]`` with ``dim_size = x.size(0)`` (mpnn_layers.py:48,88,137,173).
* ``torch_scatter.scatter(reduce=...)``: ``sum``/``add`` start from zeros;
  ``mean`` divides the sum by ``max(count, 1)``; ``max``/``min`` leave rows
  without any message at 0.
* ``nn.dense.linear.Linear``: ``F.linear`` on a ``[out, in]`` weight with
  Kaiming-uniform(a=sqrt 5) weight and uniform(+-1/sqrt(in)) bias -- the same
  initialisation as ``torch.nn.Linear``.
* ``nn.BatchNorm``: ``torch.nn.BatchNorm1d(C, eps=1e-5, momentum=0.1)`` held as
  ``.module``.
* ``nn.inits.reset``: calls ``reset_parameters`` where present.

The reference's known answers (test/test_gnn.py) pin the flow direction, the
concat orders and max-aggregation over parallel edges against this shim.
"""
from __future__ import annotations

import inspect
import sys
import types

import torch


def scatter(src: torch.Tensor, index: torch.Tensor, dim_size: int, reduce: str) -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce=reduce)."""
    width = src.shape[1:]
    out = torch.zeros((dim_size,) + tuple(width), dtype=src.dtype, device=src.device)
    if src.shape[0] == 0:
        return out
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return out.scatter_add_(0, idx, src)
    if reduce == "mean":
        out.scatter_add_(0, idx, src)
        count = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        count.scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / count.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
    if reduce in ("max", "min"):
        return out.scatter_reduce_(0, idx, src, "amax" if reduce == "max" else "amin",
                                   include_self=False)
    raise ValueError(f"unsupported aggregation {reduce!r}")


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr: str = "add", flow: str = "source_to_target", node_dim: int = -2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr

    def propagate(self, edge_index: torch.Tensor, size=None, **kwargs) -> torch.Tensor:
        x = kwargs.get("x")
        src, dst = edge_index[0], edge_index[1]
        wanted = inspect.signature(self.message).parameters
        args = {}
        for name in wanted:
            if name.endswith("_j"):
                args[name] = kwargs[name[:-2]].index_select(0, src)
            elif name.endswith("_i"):
                args[name] = kwargs[name[:-2]].index_select(0, dst)
            else:
                args[name] = kwargs[name]
        messages = self.message(**args)
        return scatter(messages, dst, x.size(0), self.aggr)

    def message(self, x_j):  # pragma: no cover - overridden by the reference layers
        return x_j


class Linear(torch.nn.Linear):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, **_):
        super().__init__(in_channels, out_channels, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels


class BatchNorm(torch.nn.Module):
    def __init__(self, in_channels: int, eps: float = 1e-5, momentum: float = 0.1,
                 affine: bool = True, track_running_stats: bool = True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.module(x)


def reset(value) -> None:
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)


def install() -> None:
    """Register the stand-in under the module names the reference imports."""
    if "torch_geometric" in sys.modules and not getattr(sys.modules["torch_geometric"], "_rgnn_shim", False):
        return  # a real torch_geometric is importable: use it
    def mod(name):
        m = types.ModuleType(name)
        m._rgnn_shim = True
        sys.modules[name] = m
        return m
    tg = mod("torch_geometric")
    nn = mod("torch_geometric.nn")
    conv = mod("torch_geometric.nn.conv")
    dense = mod("torch_geometric.nn.dense")
    linear = mod("torch_geometric.nn.dense.linear")
    inits = mod("torch_geometric.nn.inits")
    typing_mod = mod("torch_geometric.typing")
    tg.nn, nn.conv, nn.dense, dense.linear, nn.inits, tg.typing = nn, conv, dense, linear, inits, typing_mod
    conv.MessagePassing = MessagePassing
    nn.MessagePassing = MessagePassing
    linear.Linear = Linear
    nn.Linear = Linear
    nn.BatchNorm = BatchNorm
    inits.reset = reset
    typing_mod.Adj = torch.Tensor
    typing_mod.OptTensor = torch.Tensor
    tg.seed_everything = lambda seed: torch.manual_seed(seed)
