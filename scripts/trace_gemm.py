"""Debug: timeline (clock64 cycles) of CTA 0 of the tcgen05 node GEMM inside the fused path."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from radargnn_b200 import ops, synthetic, _lib
from scripts.time_pipeline import make_cfg
lib = _lib.load()
fn = lib.rgnn_debug_trace_node_gemm
fn.argtypes = [C.c_void_p, C.c_char_p]; fn.restype = None
n = 100_000
fr = synthetic.uniform_square(n, seed=0)
pos = torch.from_numpy(fr.X_cc).float().cuda(); vel = torch.from_numpy(fr.V_cc_compensated).float().cuda()
x0 = torch.from_numpy(synthetic.node_embeddings(n, 64)).cuda()
cfg = make_cfg()
for _ in range(3):
    ops.pipeline_forward(cfg, pos, vel, x0)
for tag in (b"node_gemm_pre", b"node_gemm_post", b"node_gemm_pre@2", b"node_gemm_post@2"):
    buf = torch.zeros(4 * 64 + 2 * 192, dtype=torch.int64, device="cuda")
    fn(buf.data_ptr(), tag)
    ops.pipeline_forward(cfg, pos, vel, x0)
    torch.cuda.synchronize()
    life = buf.cpu().numpy()[256:].reshape(192, 2)
    life = life[life[:, 0] > 0]
    t = buf.cpu().numpy()[:256].reshape(4, 32, 2)
    t0 = t[3, 0, 0]
    print(tag.decode(), "kernel cycles (CTA 0):", t[3, 0, 1] - t0, " setup part 1 cycles:", t0 - buf.cpu().numpy()[3 * 64 + 63])
    tag = tag.split(b"@")[0]
    st, en = life[:, 0] - life[:, 0].min(), life[:, 1] - life[:, 0].min()
    print(f"  CTAs {len(life)}: start spread {int(st.max())} ns, durations min/median/max {int((en-st).min())}/{int(np.median(en-st))}/{int((en-st).max())} ns, last end {int(en.max())} ns")
    for tl in range(8):
        if t[0, tl, 0] == 0: break
        row = [int(v - t0) for v in (t[0, tl, 0], t[0, tl, 1], t[1, tl, 0], t[1, tl, 1], t[2, tl, 0], t[2, tl, 1])]
        print(f"  tile {tl}: loader {row[0]:7d}..{row[1]:7d}  mma {row[2]:7d}..{row[3]:7d}  epilogue {row[4]:7d}..{row[5]:7d}")
    f = buf.cpu().numpy()[3 * 64 + 2: 3 * 64 + 62].reshape(15, 4)
    for g in range(15):
        if f[g, 0] == 0: break
        print(f"  epilogue warp 0, tile 2, block {g}: start {int(f[g,0]-t0):7d}  tmem-ld+sts {int(f[g,1]-f[g,0]):6d}  stores {int(f[g,2]-f[g,1]):6d}")
