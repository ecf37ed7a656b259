"""Gradients through the CUDA layers (reference gnn/trainer.py:228-231 calls
loss.backward() through DetNetBasic): every gradient against torch autograd on the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import detection_oracle as do
from oracle import mpnn_oracle as mo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.fixture(scope="module")
def gnn():
    from radargnn_b200 import gnn as _gnn
    return _gnn


def _rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("kind,aggr,c,cout,de", [
    ("MPNNConv", "max", 64, 64, 2), ("MPNNConv", "min", 16, 24, 4), ("MPNNConv", "add", 8, 8, 3), ("MPNNConv", "mean", 12, 20, 2),
    ("MPNNConv", "max", 5, 7, 1), ("RadarPointGNNConv", "max", 32, 32, 2), ("RadarPointGNNConv", "mean", 12, 12, 4),
    ("MPNNConv", "max", 128, 128, 2),
])
def test_conv_gradients_match_oracle_autograd(gnn, kind, aggr, c, cout, de):
    g = torch.Generator().manual_seed(c * 3 + de)
    n, e = 300, 2500
    x = torch.randn(n, c, generator=g)
    ei = torch.randint(0, n - 7, (2, e), generator=g)      # the last nodes have no incoming (and no outgoing) edge
    ea = torch.randn(e, de, generator=g)
    layer = (gnn.MPNNConv(c, cout, de, aggr=aggr) if kind == "MPNNConv" else gnn.RadarPointGNNConv(c, de, aggr=aggr)).to(DEV)
    params = {k: v.detach().cpu().double().requires_grad_() for k, v in layer.state_dict().items()}
    r = torch.randn(n, layer.out_channels, generator=g)
    # oracle: fp64 autograd
    xo, eo = x.double().requires_grad_(), ea.double().requires_grad_()
    fwd = mo.mpnn_conv_forward if kind == "MPNNConv" else mo.radar_point_gnn_conv_forward
    yo = fwd(params, xo, ei, eo, aggr, dtype=torch.float64)
    (yo * r.double()).sum().backward()
    # CUDA layer
    xg, eg = x.to(DEV).requires_grad_(), ea.to(DEV).requires_grad_()
    y = layer(xg, ei.to(DEV), eg)
    assert _rel(y.detach(), yo.detach()) <= 2e-5
    (y * r.to(DEV)).sum().backward()
    assert _rel(xg.grad, xo.grad) <= 1e-4
    assert _rel(eg.grad, eo.grad) <= 1e-4
    for name, p in layer.named_parameters():
        assert _rel(p.grad, params[name].grad) <= 1e-4, name


def test_batchnorm_and_linear_gradients(gnn):
    from radargnn_b200.gnn._message_passing import BatchNorm, Linear
    g = torch.Generator().manual_seed(0)
    x = torch.randn(500, 24, generator=g)
    r = torch.randn(500, 24, generator=g)
    bn = BatchNorm(24).to(DEV)
    with torch.no_grad():
        bn.module.weight.copy_(torch.rand(24, generator=g) + 0.5)
        bn.module.bias.copy_(torch.randn(24, generator=g))
    ref = torch.nn.BatchNorm1d(24).double()
    ref.load_state_dict({k: v.cpu().double() if v.is_floating_point() else v.cpu() for k, v in bn.module.state_dict().items()})
    xo = x.double().requires_grad_()
    (torch.relu(ref(xo)) * r.double()).sum().backward()
    xg = x.to(DEV).requires_grad_()
    (bn(xg, relu=True) * r.to(DEV)).sum().backward()
    assert _rel(xg.grad, xo.grad) <= 1e-4
    assert _rel(bn.module.weight.grad, ref.weight.grad) <= 1e-4 and _rel(bn.module.bias.grad, ref.bias.grad) <= 1e-4
    lin = Linear(24, 10).to(DEV)
    lo = torch.nn.Linear(24, 10).double()
    lo.load_state_dict({k: v.cpu().double() for k, v in lin.state_dict().items()})
    xo = x.double().requires_grad_()
    r2 = torch.randn(500, 10, generator=g)
    (lo(torch.relu(xo)) * r2.double()).sum().backward()
    xg = x.to(DEV).requires_grad_()
    (lin(xg, relu_input=True) * r2.to(DEV)).sum().backward()
    assert _rel(xg.grad, xo.grad) <= 1e-5 and _rel(lin.weight.grad, lo.weight.grad) <= 1e-5 and _rel(lin.bias.grad, lo.bias.grad) <= 1e-5


def test_training_step_gradients_of_the_whole_model(gnn):
    """One step of reference gnn/trainer.py:176-231 on the CUDA model: forward, weighted CE + Huber loss,
    loss.backward(); every parameter gradient against fp64 autograd through the oracle model."""
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(3)
    cfg = gnn.GNNArchitectureConfig(6, 2, [16, 16], [12, 5], [12, 5], initial_node_feature_embedding=True,
                                    node_feature_embedding_layer_dimensions=[8, 16], batch_norm_in_mlps=False)
    model = gnn.DetNetBasic(cfg).to(DEV)
    n, e = 400, 3600
    x = torch.randn(n, 6, generator=g)
    ei = torch.randint(0, n, (2, e), generator=g)
    ea = torch.randn(e, 2, generator=g)
    y = torch.cat([torch.randint(0, 5, (n, 1), generator=g).float(), torch.randn(n, 5, generator=g)], dim=1)
    w = torch.rand(5, generator=g) + 0.5
    cls, bb = model(x.to(DEV), ei.to(DEV), ea.to(DEV))
    loss = gnn.detection_loss(cls, bb, y.to(DEV), w.to(DEV), 4, 0.8, 1.2)
    loss.backward()
    params = {k: v.detach().cpu().double().requires_grad_() if v.is_floating_point() else v.cpu() for k, v in model.state_dict().items()}
    co, bo = mo.det_net_forward(params, x, ei, ea, n_layers=2, node_embedding=True, dtype=torch.float64)
    lab = y[:, 0].long()
    fgm = lab != 4
    lo = 0.8 * torch.nn.functional.cross_entropy(co, lab, weight=w.double()) + \
        1.2 * torch.nn.functional.huber_loss(bo[fgm], y[fgm, 1:].double(), reduction="none").mean(dim=1).sum() / fgm.sum()
    lo.backward()
    assert float(loss.detach()) == pytest.approx(float(lo.detach()), rel=1e-4)
    want = do.detection_loss(co.detach().float(), bo.detach().float(), y, w, 4, 0.8, 1.2)
    assert float(loss.detach()) == pytest.approx(want[0], rel=1e-4)
    checked = 0
    scale = max(float(params[name].grad.abs().max()) for name, _ in model.named_parameters())
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        # against the parameter's own gradient scale, floored at 1e-3 of the largest gradient in the model (a bias
        # in front of a BatchNorm has an exactly zero gradient: fp32 leaves rounding noise there)
        err = float((p.grad.double().cpu() - params[name].grad).abs().max())
        assert err <= 2e-3 * max(float(params[name].grad.abs().max()), 1e-3 * scale), name
        checked += 1
    assert checked >= 16
    # an optimizer step on these gradients changes the output (the model is trainable end to end)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    opt.step()
    for conv in model.convs:
        conv.invalidate_packed_weights()
    cls2, _ = model(x.to(DEV), ei.to(DEV), ea.to(DEV))
    assert float((cls2 - cls).abs().max()) > 0
