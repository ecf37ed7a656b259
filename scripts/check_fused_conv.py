"""Debug: one MPNNConv(64 -> 64) through the fused kernel and the two-kernel path on a random graph (rows in
kernel order).  usage: python scripts/check_fused_conv.py N [de] [aggr]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radargnn_b200 import ops  # noqa: E402

n = int(sys.argv[1])
de = int(sys.argv[2]) if len(sys.argv) > 2 else 2
aggr = sys.argv[3] if len(sys.argv) > 3 else "max"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
c, e = 64, 16 * n
x = torch.randn(n, c, generator=g)
ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)])
ea = torch.randn(e, de, generator=g)
p = 2 * c + de
pre = (torch.randn(p, p, generator=g) / p ** 0.5, torch.randn(p, generator=g) * 0.1)
post = (torch.randn(c, p + c, generator=g) / (p + c) ** 0.5, torch.randn(c, generator=g) * 0.1)
cp = ops.ConvParams("MPNNConv", c, c, de, aggr, [tuple(t.to(dev) for t in pre)], [tuple(t.to(dev) for t in post)])
csc = ops.csc_build(ei.to(dev), n)
out = {}
for mode in ("1", "0"):
    os.environ["RGNN_DISABLE_FUSED_LAYER"] = mode
    out[mode] = ops.conv_forward(cp, x.to(dev), csc, ea.to(dev)).cpu()
err = (out["1"] - out["0"]).abs().max(dim=1).values
bad = torch.nonzero(err > 1e-3).flatten()
print("n", n, "max diff", float(err.max()), "bad rows", bad.numel(), bad[:40].tolist(), "..", bad[-8:].tolist())
