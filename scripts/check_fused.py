"""Debug: fused aggregate+update kernel (fused_layer.cu) against the two-kernel path, layer by layer, on a
BASELINE workload.  usage: python scripts/check_fused.py [config4|headline|config5] [n_layers]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from radargnn_b200 import ops, synthetic  # noqa: E402
from test_gpu_ops import _dev, _pipeline_cfg, _stack_params  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "config4"
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if which == "config4":
    frames = [synthetic.nuscenes_frame(2000, seed=s) for s in range(64)]
    X, V, ptr = synthetic.frame_batch(frames)
    de, k, feats, seed = 4, 20, ["point_pair_features"], 6
    if len(sys.argv) > 3 and sys.argv[3] == "rp":
        de, feats = 2, ["relative_position"]
else:
    n = 100_000 if which == "headline" else 125_000
    fr = synthetic.uniform_square(n, seed=7)
    X, V, ptr = fr.X_cc, fr.V_cc_compensated, None
    de, k, feats, seed = 2, 16, ["relative_position"], 0
    if len(sys.argv) > 3 and sys.argv[3] == "ppf":
        de, feats = 4, ["point_pair_features"]
n = X.shape[0]
x0 = synthetic.node_embeddings(n, 64, seed=4)
params = _stack_params(layers, 64, 64, de, "MPNNConv", seed=seed)
cfg = _pipeline_cfg(ops, params, layers, "MPNNConv", "max", algorithm="knn", k=k, edge_features=feats)
out = {}
for mode in ("1", "0"):
    os.environ["RGNN_DISABLE_FUSED_LAYER"] = mode
    ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
    torch.cuda.synchronize()
    out[mode] = h.cpu()
ref, got = out["1"], out["0"]
err = (ref - got).abs()
rowerr = err.max(dim=1).values
bad = torch.nonzero(rowerr > 1e-3 * ref.abs().max()).flatten()
print("max abs diff", float(err.max()), "scale", float(ref.abs().max()), "bad rows", bad.numel(), "of", n)
indeg = torch.bincount(ei[1].cpu(), minlength=n)
if bad.numel():
    print("bad rows (first 20):", bad[:20].tolist())
    print("their in-degree:", indeg[bad[:20]].tolist())
    print("in-degree max", int(indeg.max()), "bad in-degree min/max", int(indeg[bad].min()), int(indeg[bad].max()))
    print("bad columns of first bad row:", torch.nonzero(err[bad[0]] > 1e-3).flatten().tolist()[:64])
top = torch.argsort(rowerr, descending=True)[:12]
print("worst rows", top.tolist(), "err", [round(float(v), 3) for v in rowerr[top]], "in-degree", indeg[top].tolist())
print("in-degree histogram tail:", torch.bincount(indeg)[-8:].tolist(), "max", int(indeg.max()))
