"""MPNN-forward oracle (CPU, pure torch): message MLP -> scatter aggregate -> update MLP.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Functional restatement of

* ``MPNNConv.forward/message``          reference gnn/mpnn_layers.py:86-101
* ``RadarPointGNNConv.forward/message`` reference gnn/mpnn_layers.py:171-184
* the conv stack of ``DetNetBasic.forward``  reference gnn/gnn_models.py:117-134
* ``get_mlp``-built Sequentials         reference gnn/gnn_models.py:137-178

expressed with the primitive tensor ops PyG 2.1 / torch_scatter 2.0.9 execute
(``index_select`` x2 -> ``cat`` -> ``linear`` -> scatter reduce -> ``cat`` ->
``linear``).  Parameters are read from a plain ``state_dict`` using the
reference's key names (``pre_mlp.0.weight`` ... ``post_mlp.2.bias``,
``edge_encoder.weight``), so the same dict drives the oracle and the CUDA path.

``dtype=torch.float64`` gives the "truth" mode used to judge fp32 error budgets.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Mapping[str, torch.Tensor]


def _mlp_keys(params: Params, prefix: str) -> List[int]:
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in params
                  if k.startswith(prefix + ".") and k.endswith(".weight")})
    return idx


def run_sequential(params: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``Linear [, ReLU, Linear]*`` as built at mpnn_layers.py:64-74 -- the Linear
    modules sit at even positions, a ReLU between consecutive ones."""
    keys = _mlp_keys(params, prefix)
    for n, i in enumerate(keys):
        if n > 0:
            x = F.relu(x)
        x = F.linear(x, params[f"{prefix}.{i}.weight"].to(x.dtype),
                     params[f"{prefix}.{i}.bias"].to(x.dtype))
    return x


def scatter_aggregate(messages: torch.Tensor, target: torch.Tensor, n_nodes: int, aggr: str) -> torch.Tensor:
    """torch_scatter.scatter(messages, target, dim=0, dim_size=n_nodes, reduce=aggr):
    nodes without an incoming edge end at 0 for every reduction."""
    out = torch.zeros((n_nodes, messages.shape[1]), dtype=messages.dtype)
    if messages.shape[0] == 0:
        return out
    idx = target.view(-1, 1).expand_as(messages)
    if aggr in ("add", "sum"):
        return out.scatter_add_(0, idx, messages)
    if aggr == "mean":
        out.scatter_add_(0, idx, messages)
        count = torch.bincount(target, minlength=n_nodes).clamp(min=1).to(messages.dtype)
        return out / count.view(-1, 1)
    if aggr == "max":
        return out.scatter_reduce_(0, idx, messages, "amax", include_self=False)
    if aggr == "min":
        return out.scatter_reduce_(0, idx, messages, "amin", include_self=False)
    raise ValueError(f"unsupported aggregation {aggr!r}")


def mpnn_conv_forward(params: Params, x: torch.Tensor, edge_index: torch.Tensor,
                      edge_attr: torch.Tensor, aggr: str = "max",
                      use_edge_encoder: bool = False,
                      dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """mpnn_layers.py:86-101.  ``edge_index[0]`` is the source j, ``edge_index[1]``
    the target i; the message is ``pre_mlp([x_i ; x_j ; e])`` reduced at i, the
    update ``post_mlp([x ; m])``."""
    x = x.to(dtype)
    edge_attr = edge_attr.to(dtype)
    src, dst = edge_index[0], edge_index[1]
    x_j, x_i = x.index_select(0, src), x.index_select(0, dst)
    if use_edge_encoder:
        edge_attr = F.linear(edge_attr, params["edge_encoder.weight"].to(dtype),
                             params["edge_encoder.bias"].to(dtype))
    m = run_sequential(params, "pre_mlp", torch.cat([x_i, x_j, edge_attr], dim=-1))
    m_emb = scatter_aggregate(m, dst, x.shape[0], aggr)
    return run_sequential(params, "post_mlp", torch.cat([x, m_emb], dim=-1))


def radar_point_gnn_conv_forward(params: Params, x: torch.Tensor, edge_index: torch.Tensor,
                                 edge_attr: torch.Tensor, aggr: str = "max",
                                 dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """mpnn_layers.py:171-184: message ``pre_mlp([x_j ; e])``, residual update."""
    x = x.to(dtype)
    edge_attr = edge_attr.to(dtype)
    src, dst = edge_index[0], edge_index[1]
    m = run_sequential(params, "pre_mlp", torch.cat([x.index_select(0, src), edge_attr], dim=-1))
    m_emb = scatter_aggregate(m, dst, x.shape[0], aggr)
    return run_sequential(params, "post_mlp", torch.cat([x, m_emb], dim=-1)) + x


def batch_norm_train(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                     eps: float = 1e-5) -> torch.Tensor:
    """PyG ``BatchNorm`` == ``BatchNorm1d`` in training mode (the reference never
    leaves it): batch statistics over all nodes, biased variance."""
    mean = x.mean(dim=0)
    var = x.var(dim=0, unbiased=False)
    y = (x - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight.to(x.dtype) + bias.to(x.dtype)
    return y


def sub_params(params: Params, prefix: str) -> Dict[str, torch.Tensor]:
    p = prefix if prefix.endswith(".") else prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


def conv_stack_forward(params: Params, x: torch.Tensor, edge_index: torch.Tensor,
                       edge_attr: torch.Tensor, n_layers: int, conv_type: str = "MPNNConv",
                       aggr: str = "max", use_edge_encoder: bool = False,
                       dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """gnn_models.py:124-128: ``x = relu(batch_norm(conv(x, edge_index, edge_attr)))``
    per layer; keys ``convs.{i}.*`` and ``batch_norms.{i}.module.*``."""
    x = x.to(dtype)
    for i in range(n_layers):
        cp = sub_params(params, f"convs.{i}")
        if conv_type == "MPNNConv":
            x = mpnn_conv_forward(cp, x, edge_index, edge_attr, aggr, use_edge_encoder, dtype)
        elif conv_type == "RadarPointGNNConv":
            x = radar_point_gnn_conv_forward(cp, x, edge_index, edge_attr, aggr, dtype)
        else:
            raise Exception(f"{conv_type} is invalid GNN conv layer type. Chose either MPNNConv or RadarPointGNNConv")
        x = batch_norm_train(x, params.get(f"batch_norms.{i}.module.weight"),
                             params.get(f"batch_norms.{i}.module.bias"))
        x = F.relu(x)
    return x


def run_get_mlp(params: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """A ``get_mlp`` Sequential without BatchNorm (gnn_models.py:137-178 with
    ``batch_norm=False``): Linear, then (ReLU, Linear) repeated."""
    return run_sequential(params, prefix, x)


def det_net_forward(params: Params, x, edge_index, edge_attr, *, n_layers: int,
                    conv_type: str = "MPNNConv", aggr: str = "max", use_edge_encoder: bool = False,
                    node_embedding: bool = False, edge_embedding: bool = False,
                    dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """gnn_models.py:104-134 for ``batch_norm_in_mlps=False``."""
    x, edge_attr = x.to(dtype), edge_attr.to(dtype)
    if node_embedding:
        x = run_get_mlp(params, "node_emb_mlp", x)
    if edge_embedding:
        edge_attr = run_get_mlp(params, "edge_emb_mlp", edge_attr)
    x = conv_stack_forward(params, x, edge_index, edge_attr, n_layers, conv_type, aggr,
                           use_edge_encoder, dtype)
    return run_get_mlp(params, "classification_head", x), run_get_mlp(params, "regression_head", x)


def relative_error(actual: torch.Tensor, expected: torch.Tensor) -> float:
    """The parity metric for embeddings: max |a - e| / max(|e|) (relative to the
    tensor's scale; north star asks <= 1e-4 in fp32)."""
    a, e = actual.double(), expected.double()
    denom = e.abs().max().clamp(min=1e-30)
    return float((a - e).abs().max() / denom)


def rowwise_relative_error(actual: torch.Tensor, expected: torch.Tensor) -> float:
    """Element-wise companion of ``relative_error``: max over all elements of
    |a - e| / max(|e_row|), every node's row measured against ITS OWN scale (a node whose
    embedding is small must be as accurate, relatively, as the largest one)."""
    a, e = actual.double(), expected.double()
    scale = e.abs().amax(dim=1, keepdim=True).clamp(min=1e-30)
    return float(((a - e).abs() / scale).max())


def parity_report(actual: torch.Tensor, expected: torch.Tensor) -> Dict[str, float]:
    """Both metrics plus the fraction of elements off by more than 1e-4 of their row scale."""
    a, e = actual.double(), expected.double()
    scale = e.abs().amax(dim=1, keepdim=True).clamp(min=1e-30)
    rel = (a - e).abs() / scale
    return {"max_norm": relative_error(actual, expected), "rowwise": float(rel.max()),
            "frac_above_1e-4": float((rel > 1e-4).double().mean())}
