"""Building blocks standing in for the PyG classes the reference subclasses / instantiates
(``MessagePassing``, ``nn.dense.linear.Linear``, ``nn.BatchNorm``), with the same parameter
names so that the reference's ``state_dict``s load, and with every forward routed to the CUDA
kernels of librgnn_b200.so; gradients flow through the ``torch.autograd.Function``s of ``_autograd.py``."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from .. import ops


class Linear(torch.nn.Linear):
    """PyG ``Linear``: ``F.linear`` on a ``[out, in]`` weight, torch.nn.Linear initialisation."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, **_):
        super().__init__(in_channels, out_channels, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels

    def forward(self, x: torch.Tensor, relu_input: bool = False) -> torch.Tensor:
        from ._autograd import LinearFunction, wants_grad
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if wants_grad(x2, self.weight, self.bias):
            y = LinearFunction.apply(x2, self.weight, self.bias, relu_input)
        else:
            y = ops.linear(x2, self.weight, self.bias, relu_input=relu_input)
        return y.reshape(*lead, y.shape[-1])


class BatchNorm(torch.nn.Module):
    """PyG ``BatchNorm``: a ``torch.nn.BatchNorm1d`` held as ``.module`` (parameter names
    ``module.weight`` ... ``module.running_var``).  Training mode normalises with the batch
    statistics and updates the running ones; eval mode uses the running statistics."""

    def __init__(self, in_channels: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
        m = self.module
        if m.training or m.running_mean is None:
            momentum = 0.1 if m.momentum is None else m.momentum
            if m.num_batches_tracked is not None and m.training:
                m.num_batches_tracked += 1
            from ._autograd import BatchNormReluFunction, wants_grad
            rm, rv = (m.running_mean, m.running_var) if m.training else (None, None)
            if wants_grad(x, m.weight, m.bias):
                return BatchNormReluFunction.apply(x, m.weight, m.bias, m.eps, momentum, rm, rv, relu)
            return ops.batchnorm_relu(x, m.weight, m.bias, m.eps, momentum, rm, rv, relu=relu)
        scale = torch.rsqrt(m.running_var + m.eps)
        beta = torch.zeros_like(scale)
        if m.weight is not None:
            scale = scale * m.weight.detach()
            beta = m.bias.detach()
        return ops.affine_relu(x, m.running_mean, scale, beta, relu)


def _tensor_version(t: torch.Tensor) -> int:
    """Version counter of a tensor; inference-mode tensors have none (they cannot be modified in place)."""
    try:
        return t._version
    except RuntimeError:
        return -1


def reset(value) -> None:
    """PyG ``nn.inits.reset``."""
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)


class _CscCache:
    """edge_index -> CSC view, kept for the last few graphs: every conv of a stack receives the
    same edge_index tensor (gnn_models.py:124-125), so the counting sort runs once per batch."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self.items: "OrderedDict[tuple, ops.CscGraph]" = OrderedDict()

    def get(self, edge_index: torch.Tensor, n_nodes: int) -> ops.CscGraph:
        key = (edge_index.data_ptr(), tuple(edge_index.shape), _tensor_version(edge_index), n_nodes, str(edge_index.device))
        hit = self.items.get(key)
        if hit is not None:
            self.items.move_to_end(key)
            return hit[1]
        csc = ops.csc_build(edge_index, n_nodes)
        # keep the tensor alive so that its address cannot be recycled while the entry exists
        self.items[key] = (edge_index, csc)
        while len(self.items) > self.capacity:
            self.items.popitem(last=False)
        return csc


_csc_cache = _CscCache()


def _pyg_message_passing():
    """``torch_geometric.nn.MessagePassing`` when PyG is importable (the reference's own base class,
    gnn/mpnn_layers.py:4), else None.  The layers then ARE PyG MessagePassing modules -- ``isinstance``
    checks, ``jittable`` / hooks and ``aggr`` bookkeeping of a PyG code base keep working -- while
    ``forward`` still runs the CUDA kernels instead of ``propagate``."""
    try:
        from torch_geometric.nn import MessagePassing as base   # noqa: WPS433 (optional dependency)
        return base
    except Exception:
        return None


_PygBase = _pyg_message_passing()
USES_PYG_BASE = _PygBase is not None


class MessagePassing(_PygBase if USES_PYG_BASE else torch.nn.Module):
    """The part of PyG's ``MessagePassing`` the reference layers rely on: ``aggr``,
    ``flow = source_to_target`` (``x_j = x[edge_index[0]]``, ``x_i = x[edge_index[1]]``,
    messages reduced at ``edge_index[1]`` with ``dim_size = x.size(0)``).  Subclasses the real PyG class
    when ``torch_geometric`` is importable, a plain ``torch.nn.Module`` otherwise."""

    def __init__(self, aggr: Optional[str] = "add", flow: str = "source_to_target", node_dim: int = -2):
        if flow != "source_to_target":
            raise ValueError("only flow='source_to_target' is implemented (the reference's setting)")
        if USES_PYG_BASE:
            super().__init__(aggr=aggr, flow=flow, node_dim=node_dim)
        else:
            super().__init__()
            self.aggr = aggr
            self.flow = flow
            self.node_dim = node_dim
        self.aggr_name = aggr   # PyG >= 2.2 may turn .aggr into an Aggregation module: the kernels take the name

    @staticmethod
    def _check_inputs(x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor) -> None:
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
        if not (x.is_cuda and edge_index.is_cuda and edge_attr.is_cuda):
            raise RuntimeError("radargnn_b200 layers need CUDA tensors: there is no CPU fallback")

    @staticmethod
    def csc(edge_index: torch.Tensor, n_nodes: int) -> ops.CscGraph:
        return _csc_cache.get(edge_index, n_nodes)
