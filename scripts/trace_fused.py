"""Debug: timeline (clock64 cycles) of one CTA of the fused aggregate + update kernel inside the pipeline."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from radargnn_b200 import ops, synthetic, _lib
from scripts.time_pipeline import make_cfg
lib = _lib.load()
fn = lib.rgnn_debug_trace_fused_layer
fn.argtypes = [C.c_void_p, C.c_int]; fn.restype = None
n = 100_000
fr = synthetic.uniform_square(n, seed=0)
pos = torch.from_numpy(fr.X_cc).float().cuda(); vel = torch.from_numpy(fr.V_cc_compensated).float().cuda()
x0 = torch.from_numpy(synthetic.node_embeddings(n, 64)).cuda()
cfg = make_cfg()
for _ in range(3):
    ops.pipeline_forward(cfg, pos, vel, x0)
for cta in (int(a) for a in (sys.argv[1:] or ["70"])):
    buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
    fn(buf.data_ptr(), cta)
    ops.pipeline_forward(cfg, pos, vel, x0)
    torch.cuda.synchronize()
    b = buf.cpu().numpy()
    t0 = b[1000]
    print(f"CTA {cta}: kernel {b[1001] - t0} cycles")
    names = ["start", "x done", "tail done", "full ok", "M' in TMEM", "MMAs issued", "acc ok", "epilogue done"]
    for t in range(8):
        row = b[t * 8: t * 8 + 8]
        if row[0] == 0: break
        print(f"  tile {t}: " + "  ".join(f"{nm} {int(v - t0)}" for nm, v in zip(names, row)))
    for w in range(12):
        a = b[64 + w * 48: 64 + w * 48 + 48].reshape(24, 2)
        a = a[a[:, 0] > 0]
        starts = (a[:, 0] - t0).tolist()
        print(f"  agg warp {w}: pass starts {starts[:16]}  ring waits {a[:, 1].tolist()[:16]}")
