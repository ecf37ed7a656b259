"""Autograd through the CUDA layers (reference gnn/trainer.py:228-231 runs ``loss.backward()`` through
``DetNetBasic``): ``torch.autograd.Function``s around the forward kernels.

* graph convolution (MPNNConv / RadarPointGNNConv with one Linear in ``pre_mlp`` and ``post_mlp``, no edge
  encoder -- the shapes the fused forward factors): the graph-specific half of the backward -- recomputing the
  aggregated messages, arg-max routing of their gradient to target / source nodes, edge weights and edge
  attributes -- is ``rgnn_conv_backward_route`` (csrc/conv_backward.cu); the dense contractions around it are
  library GEMMs (torch.matmul).  Deeper message / update MLPs raise NotImplementedError in backward.
* BatchNorm (training statistics) + ReLU and Linear (+ ReLU on the input): forward on the kernels, backward by the
  closed formulas in torch arithmetic."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib, ops


class ConvFunction(torch.autograd.Function):
    """y = conv(x, edge_attr; W_pre, b_pre, W_post, b_post) on the CSC view ``csc`` of the graph."""

    @staticmethod
    def forward(ctx, x, edge_attr, w_pre, b_pre, w_post, b_post, params, csc):
        ctx.params, ctx.csc = params, csc
        ctx.save_for_backward(x, edge_attr, w_pre, b_pre, w_post, b_post)
        return ops.conv_forward(params, x, csc, edge_attr)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, edge_attr, w_pre, b_pre, w_post, b_post = ctx.saved_tensors
        params, csc = ctx.params, ctx.csc
        if len(params.pre) != 1 or len(params.post) != 1 or params.edge_encoder is not None:
            raise NotImplementedError("backward is implemented for one Linear in pre_mlp / post_mlp and no edge encoder")
        lib = _lib.load()
        mpnn = params.conv_type == "MPNNConv"
        c, de = params.in_channels, params.edge_dim
        p = (2 * c if mpnn else c) + de
        pp = (p + 3) // 4 * 4
        n, e = csc.n_nodes, csc.n_edges
        dev = x.device
        x = x.detach().to(torch.float32).contiguous()
        dy = dy.to(torch.float32).contiguous()
        w_pre_d, w_post_d = w_pre.detach().float().contiguous(), w_post.detach().float().contiguous()
        w_t = w_pre_d[:, :c] if mpnn else None
        w_s = w_pre_d[:, c:2 * c] if mpnn else w_pre_d[:, :c]
        e_off = 2 * c if mpnn else c

        def padded(t):   # [N, p] -> [N, pp], 16-byte rows
            out = torch.zeros((n, pp), dtype=torch.float32, device=dev)
            out[:, :p] = t
            return out

        a = padded(x @ w_t.t()) if mpnn else None
        b = padded(x @ w_s.t())
        g = dy @ w_post_d                       # [N, C + P]: gradient of cat([x, M])
        gm = g[:, c:].contiguous()
        eid = csc.eid.long()
        ea_csc = edge_attr.detach().to(torch.float32)[eid].contiguous() if e > 0 else edge_attr.detach().float()
        m = torch.empty((n, pp), dtype=torch.float32, device=dev)
        ga = torch.empty((n, pp), dtype=torch.float32, device=dev)
        gb = torch.empty((n, pp), dtype=torch.float32, device=dev)
        dwe = torch.zeros((p, max(de, 1)), dtype=torch.float32, device=dev)[:, :de].contiguous()
        dea_csc = torch.empty((e, de), dtype=torch.float32, device=dev)
        bias = b_pre.detach().float().contiguous()
        w_e = w_pre_d[:, e_off:]                # [P, De] view with row stride P
        with torch.cuda.device(dev):
            _lib.check(lib.rgnn_conv_backward_route(
                _lib.AGGR[params.aggr], _lib.ptr(a), b.data_ptr(), p, bias.data_ptr(), w_e.data_ptr() if de > 0 else None,
                w_pre_d.shape[1], de, ea_csc.data_ptr() if e > 0 and de > 0 else None, csc.ptr.data_ptr(),
                csc.src.data_ptr() if e > 0 else None, n, e, gm.data_ptr(), p, m.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                dwe.data_ptr() if de > 0 else None, dea_csc.data_ptr() if e > 0 and de > 0 else None, _lib.stream_ptr()))
        ga_p, gb_p, m_p = ga[:, :p], gb[:, :p], m[:, :p]
        d_w_post = dy.t() @ torch.cat([x, m_p], dim=1)
        d_b_post = dy.sum(dim=0)
        blocks = ([ga_p.t() @ x] if mpnn else []) + [gb_p.t() @ x] + ([dwe] if de > 0 else [])
        d_w_pre = torch.cat(blocks, dim=1)
        d_b_pre = ga_p.sum(dim=0)
        dx = g[:, :c] + gb_p @ w_s
        if mpnn:
            dx = dx + ga_p @ w_t
        else:
            dx = dx + dy                        # residual h = post_mlp([x ; m]) + x
        d_ea = None
        if ctx.needs_input_grad[1]:
            d_ea = torch.zeros((e, de), dtype=torch.float32, device=dev)
            if e > 0 and de > 0:
                d_ea[eid] = dea_csc
        return dx, d_ea, d_w_pre, d_b_pre, d_w_post, d_b_post, None, None


class BatchNormReluFunction(torch.autograd.Function):
    """Training-mode BatchNorm1d (+ ReLU) on ``rgnn_batchnorm_relu_forward``."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, momentum, running_mean, running_var, relu):
        out = ops.batchnorm_relu(x, weight, bias, eps, momentum, running_mean, running_var, relu=relu)
        ctx.eps, ctx.relu = eps, relu
        ctx.save_for_backward(x, weight, out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, weight, out = ctx.saved_tensors
        x = x.detach().float()
        dy = dy.float()
        if ctx.relu:
            dy = dy * (out > 0)
        n = x.shape[0]
        mean = x.mean(dim=0)
        invstd = torch.rsqrt(x.var(dim=0, unbiased=False) + ctx.eps)
        xhat = (x - mean) * invstd
        d_bias = dy.sum(dim=0)
        d_weight = (dy * xhat).sum(dim=0)
        gamma = weight.detach().float() if weight is not None else torch.ones_like(mean)
        dx = (gamma * invstd / n) * (n * dy - d_bias - xhat * d_weight)
        return dx, (d_weight if weight is not None else None), (d_bias if weight is not None else None), None, None, None, None, None


class LinearFunction(torch.autograd.Function):
    """y = act(x) W^T + b on ``rgnn_linear_forward`` (act = ReLU when ``relu_input``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu_input):
        ctx.relu_input = relu_input
        ctx.save_for_backward(x, weight)
        return ops.linear(x, weight, bias, relu_input=relu_input)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        x, w, dy = x.detach().float(), weight.detach().float(), dy.float()
        xin = torch.relu(x) if ctx.relu_input else x
        dx = dy @ w
        if ctx.relu_input:
            dx = dx * (x > 0)
        return dx, dy.t() @ xin, dy.sum(dim=0), None


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


class DetectionLossFunction(torch.autograd.Function):
    """loss = alpha * CrossEntropyLoss(weight)(cls, y[:, 0]) + beta * mean over foreground nodes of HuberLoss(bb, y[:, 1:])
    (gnn/trainer.py:184-231) on ``rgnn_detection_loss``; returns a float32 scalar like the reference's ``loss``."""

    @staticmethod
    def forward(ctx, cls, bb, y, class_weight, bg_index, alpha, beta, delta, nan_to_zero):
        out = ops.detection_loss(cls, bb, y, class_weight, bg_index, alpha, beta, delta, nan_to_zero)
        ctx.save_for_backward(cls, bb, y, class_weight if class_weight is not None else torch.empty(0, device=cls.device), out)
        ctx.cfg = (bg_index, alpha, beta, delta, class_weight is not None, nan_to_zero)
        return out[0].to(torch.float32)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dl):
        cls, bb, y, w, out = ctx.saved_tensors
        bg_index, alpha, beta, delta, has_w, nan_to_zero = ctx.cfg
        cls, bb, y = cls.detach().float(), bb.detach().float(), y.detach().float()
        label = y[:, 0].long()
        valid = (label >= 0) & (label < cls.shape[1])
        lab = label.clamp(0, cls.shape[1] - 1)
        wy = (w.float()[lab] if has_w else torch.ones_like(lab, dtype=torch.float32)) * valid
        p = torch.softmax(cls, dim=1)
        p[torch.arange(cls.shape[0], device=cls.device), lab] -= 1.0
        d_cls = (alpha * dl) * p * (wy / wy.sum()).unsqueeze(1)
        fg = valid & (label != bg_index)
        num_bb = fg.sum()
        diff = bb - y[:, 1:1 + bb.shape[1]]
        hub = torch.where(diff.abs() < delta, diff, delta * torch.sign(diff))
        scale = torch.where(num_bb > 0, (beta * dl) / (num_bb.clamp(min=1) * bb.shape[1]), torch.zeros((), device=cls.device))
        d_bb = hub * fg.unsqueeze(1) * scale
        if nan_to_zero and bool(torch.isnan(d_bb).any()):
            d_bb = torch.zeros_like(d_bb)   # trainer.py:206-216: a NaN box loss is dropped from the step as a whole
        return d_cls, d_bb, None, None, None, None, None, None, None


def detection_loss(cls, bb, y, class_weight, bg_index, cls_loss_weight=1.0, bb_loss_weight=1.0, huber_delta=1.0,
                   nan_to_zero=True):
    """Differentiable training loss of the detection heads (reference gnn/trainer.py:184-231)."""
    return DetectionLossFunction.apply(cls, bb, y, class_weight, bg_index, float(cls_loss_weight), float(bb_loss_weight),
                                       float(huber_delta), bool(nan_to_zero))
