"""Scratch timing of the fused path at a given size (GPU box): per-stage device times."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from radargnn_b200 import ops, synthetic, _lib

def make_cfg(n_layers=4, c=64, de=2, k=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    layers, bn = [], []
    for l in range(n_layers):
        p = 2 * c + de
        def lin(o, i):
            b = 1 / i ** 0.5
            return ((torch.rand(o, i, generator=g) * 2 - 1) * b).cuda(), ((torch.rand(o, generator=g) * 2 - 1) * b).cuda()
        layers.append(ops.ConvParams("MPNNConv", c, c, de, "max", [lin(p, p)], [lin(c, p + c)]))
        bn.append((torch.ones(c).cuda(), torch.zeros(c).cuda()))
    return ops.PipelineConfig(layers=layers, bn=bn, algorithm="knn", k=k)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    fr = synthetic.uniform_square(n, seed=0)
    pos = torch.from_numpy(fr.X_cc).float().cuda(); vel = torch.from_numpy(fr.V_cc_compensated).float().cuda()
    x0 = torch.from_numpy(synthetic.node_embeddings(n, 64)).cuda()
    cfg = make_cfg()
    for _ in range(3):
        out = ops.pipeline_forward(cfg, pos, vel, x0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for _ in range(10):
        ev[0].record(); out = ops.pipeline_forward(cfg, pos, vel, x0, check_errors=False); ev[1].record()
        torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    print("n", n, "E", out[0].shape[1], "ms/step", np.median(ts), "edges/s", out[0].shape[1] / np.median(ts) * 1e3)
    _lib.profile_reset(); _lib.profile_enable(True)
    for _ in range(5):
        ops.pipeline_forward(cfg, pos, vel, x0, check_errors=False)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    tot = _lib.profile_totals()
    for k, (ms, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:20s} {ms / 5:9.3f} ms/step  {cnt // 5} launches/step")
if __name__ == "__main__":
    main()
