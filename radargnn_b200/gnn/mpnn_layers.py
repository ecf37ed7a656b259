"""Mirror of reference gnn/mpnn_layers.py: ``MPNNConv`` and ``RadarPointGNNConv`` with the same
constructor arguments, attributes and parameter names; ``forward`` runs the CUDA layer
(rgnn_conv_forward) instead of PyG's propagate -> torch_scatter -> addmm sequence."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import ReLU, Sequential

from .. import ops
from ._message_passing import Linear, MessagePassing, reset


def _linears(seq: Sequential):
    return [(m.weight, m.bias) for m in seq if hasattr(m, "weight")]


def _cached_params(module, conv_type, c_in, c_out, edge_dim, aggr, pre, post, enc) -> ops.ConvParams:
    """One ``ops.ConvParams`` per module, kept while the very same weight tensors are installed, so that
    its packed tensor-core images survive from call to call (they are re-packed when a tensor's address or
    version counter changes; after an edit through ``.data`` call ``module.invalidate_packed_weights()``)."""
    tensors = [t for pair in pre + post for t in pair] + (list(enc) if enc is not None else [])
    key = (conv_type, c_in, c_out, edge_dim, aggr, tuple(id(t) for t in tensors))
    hit = module.__dict__.get("_conv_params_cache")
    if hit is not None and hit[0] == key:
        return hit[1]
    params = ops.ConvParams(conv_type, c_in, c_out, edge_dim, aggr, pre, post, enc)
    module.__dict__["_conv_params_cache"] = (key, params)
    return params


def _invalidate(module) -> None:
    module.__dict__.pop("_conv_params_cache", None)


def _run_sequential(seq: Sequential, x: Tensor) -> Tensor:
    """Linear [, ReLU, Linear]* on the CUDA linear kernel, each ReLU fused into the next Linear."""
    relu_pending = False
    for m in seq:
        if isinstance(m, ReLU):
            relu_pending = True
        else:
            x = ops.linear(x, m.weight, m.bias, relu_input=relu_pending)
            relu_pending = False
    if relu_pending:
        x = torch.relu(x)
    return x


class MPNNConv(MessagePassing):
    """General MPNN layer with edge features (reference mpnn_layers.py:11-101).

    ``m_e = pre_mlp([x_i ; x_j ; e])`` with ``x_j = x[edge_index[0]]``, ``x_i = x[edge_index[1]]``,
    aggregated at ``edge_index[1]``; ``h = post_mlp([x ; m])``.
    """

    def __init__(self, in_channels: int, out_channels: int, edge_dim: int, aggr: str = "max",
                 pre_layers: int = 1, post_layers: int = 1, use_edge_encoder: bool = False):
        super().__init__(aggr=aggr)

        self.in_channels = in_channels
        self.out_channels = out_channels
        self.edge_dim = edge_dim
        self.use_edge_encoder = use_edge_encoder

        if use_edge_encoder:
            self.edge_encoder = Linear(edge_dim, self.in_channels)
            pre_mlp_dim = 3 * in_channels
        else:
            pre_mlp_dim = 2 * in_channels + edge_dim

        modules = [Linear(pre_mlp_dim, pre_mlp_dim)]
        for _ in range(pre_layers - 1):
            modules += [ReLU()]
            modules += [Linear(pre_mlp_dim, pre_mlp_dim)]
        self.pre_mlp = Sequential(*modules)

        modules = [Linear(pre_mlp_dim + in_channels, out_channels)]
        for _ in range(post_layers - 1):
            modules += [ReLU()]
            modules += [Linear(out_channels, out_channels)]
        self.post_mlp = Sequential(*modules)

        self.reset_parameters()

    def reset_parameters(self):
        if self.use_edge_encoder:
            self.edge_encoder.reset_parameters()
        for nn in self.pre_mlp:
            reset(nn)
        for nn in self.post_mlp:
            reset(nn)

    def conv_params(self) -> ops.ConvParams:
        """Parameters as the kernels take them, read at call time (the reference's tests swap
        ``layer.weight`` after construction, test/test_gnn.py:13-16)."""
        enc = (self.edge_encoder.weight, self.edge_encoder.bias) if self.use_edge_encoder else None
        post = _linears(self.post_mlp)
        return _cached_params(self, "MPNNConv", self.in_channels, post[-1][0].shape[0], self.edge_dim, self.aggr_name,
                              _linears(self.pre_mlp), post, enc)

    def invalidate_packed_weights(self) -> None:
        """Drop the cached tensor-core weight images (needed only after edits that bypass the tensors'
        version counters, e.g. ``w.data.normal_()``)."""
        _invalidate(self)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> Tensor:
        self._check_inputs(x, edge_index, edge_attr)
        csc = self.csc(edge_index, x.shape[0])
        params = self.conv_params()
        from ._autograd import ConvFunction, wants_grad
        if len(params.pre) == 1 and len(params.post) == 1 and params.edge_encoder is None and \
                wants_grad(x, edge_attr, *[t for pair in params.pre + params.post for t in pair]):
            (w_pre, b_pre), (w_post, b_post) = params.pre[0], params.post[0]
            return ConvFunction.apply(x, edge_attr, w_pre, b_pre, w_post, b_post, params, csc)
        return ops.conv_forward(params, x, csc, edge_attr)

    def message(self, x_i: Tensor, x_j: Tensor, edge_attr: Tensor) -> Tensor:
        """Per-edge messages ``pre_mlp([x_i ; x_j ; e])`` (kept for API parity; ``forward`` does
        not materialise them)."""
        if self.use_edge_encoder:
            edge_attr = self.edge_encoder(edge_attr)
        m = torch.cat([x_i, x_j, edge_attr], dim=-1)
        return _run_sequential(self.pre_mlp, m)


class RadarPointGNNConv(MessagePassing):
    """Adapted Radar-PointGNN convolution with edge features (reference mpnn_layers.py:104-184):
    ``m_e = pre_mlp([x_j ; e])``, ``h = post_mlp([x ; m]) + x``."""

    def __init__(self, init_node_dim: int, init_edge_dim: int, aggr: str = "max",
                 pre_layers: int = 1, post_layers: int = 1):
        super().__init__(aggr=aggr)

        # output dim. = input dim. -> no increase in embedding dimension possible with this layer
        self.in_channels = init_node_dim
        self.out_channels = init_node_dim

        self.init_node_dim = init_node_dim
        self.init_edge_dim = init_edge_dim

        pre_mlp_dim = init_node_dim + init_edge_dim

        modules = [Linear(pre_mlp_dim, pre_mlp_dim)]
        for _ in range(pre_layers - 1):
            modules += [ReLU()]
            modules += [Linear(pre_mlp_dim, pre_mlp_dim)]
        self.pre_mlp = Sequential(*modules)

        modules = [Linear(pre_mlp_dim + init_node_dim, init_node_dim)]
        for _ in range(post_layers - 1):
            modules += [ReLU()]
            modules += [Linear(init_node_dim, init_node_dim)]
        self.post_mlp = Sequential(*modules)

        self.reset_parameters()

    def reset_parameters(self):
        for nn in self.pre_mlp:
            reset(nn)
        for nn in self.post_mlp:
            reset(nn)

    def conv_params(self) -> ops.ConvParams:
        return _cached_params(self, "RadarPointGNNConv", self.init_node_dim, self.init_node_dim, self.init_edge_dim,
                              self.aggr_name, _linears(self.pre_mlp), _linears(self.post_mlp), None)

    def invalidate_packed_weights(self) -> None:
        _invalidate(self)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> Tensor:
        self._check_inputs(x, edge_index, edge_attr)
        csc = self.csc(edge_index, x.shape[0])
        params = self.conv_params()
        from ._autograd import ConvFunction, wants_grad
        if len(params.pre) == 1 and len(params.post) == 1 and params.edge_encoder is None and \
                wants_grad(x, edge_attr, *[t for pair in params.pre + params.post for t in pair]):
            (w_pre, b_pre), (w_post, b_post) = params.pre[0], params.post[0]
            return ConvFunction.apply(x, edge_attr, w_pre, b_pre, w_post, b_post, params, csc)
        return ops.conv_forward(params, x, csc, edge_attr)

    def message(self, x_i: Tensor, x_j: Tensor, edge_attr: Tensor) -> Tensor:
        m = torch.cat([x_j, edge_attr], dim=-1)
        return _run_sequential(self.pre_mlp, m)
