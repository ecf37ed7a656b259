"""Wall time of the boundary function over a list of frames: one build_geometric_graph call per frame (the
reference's loop, dataset_creation.py:651-660) against one build_geometric_graphs call for the list."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from radargnn_b200 import synthetic
from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud

def clouds(n_frames, n):
    out = []
    for s in range(n_frames):
        fr = synthetic.radar_frame(n, seed=s)
        pc = RadarPointCloud()
        pc.X_cc, pc.V_cc_compensated = fr.X_cc, fr.V_cc_compensated
        rng = np.random.default_rng(s)
        pc.rcs = rng.normal(size=(n, 1)); pc.timestamp = rng.integers(0, 4, size=(n, 1)).astype(np.float64)
        out.append(pc)
    return out

for n_frames, n, k, feats in ((64, 300, 20, ["relative_position"]), (64, 2000, 20, ["point_pair_features"])):
    cfg = GraphConstructionConfiguration("knn", {"k": k, "r": 1}, ["rcs", "time_index", "degree"], feats, "directed", "X")
    cl = clouds(n_frames, n)
    GraphConstructor.build_geometric_graphs(cfg, cl[:4]); [GraphConstructor.build_geometric_graph(cfg, c) for c in cl[:4]]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = [GraphConstructor.build_geometric_graph(cfg, c) for c in cl]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b = GraphConstructor.build_geometric_graphs(cfg, cl)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    assert all(np.array_equal(x.E, y.E) and np.array_equal(x.E_feat, y.E_feat) for x, y in zip(a, b))
    print(f"{n_frames} frames x {n} points, k = {k}, {feats[0]}: per-frame calls {1e3 * (t1 - t0):.1f} ms, one batched call {1e3 * (t2 - t1):.1f} ms")
