#!/usr/bin/env python
"""bench.py -- edges/s of the fused hot path (graph build + L-layer MPNN forward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic radar frames: neighbour search
(-> edge_index, edge_attr) + L x (graph convolution, BatchNorm(train), ReLU) + the loss partial.
Frames are independent, so the N ranks of a box each process their own frames (weak scaling) and
only all-reduce the loss scalar (NCCL).  One JSON line is printed by rank 0:

  value        whole-job edges/s with the inputs already resident in HBM (CUDA-event timed, L2 flushed
               between steps, max over ranks);
  e2e          the same metric through the host-buffer C-ABI entry point (rgnn_pipeline_forward_host):
               H2D of pos / vel / x0 from pinned memory and D2H of the node embeddings inside the timed
               region;
  roofline     dominant kernel: algorithmic bytes per launch / measured device time per launch, against
               MEASURED_PEAKS.json:hbm_gbs;  path_roofline: the whole step against SURVEY.md 8(d)'s B_alg;
  cpu_baseline the CPU port of the reference path (oracle/) on a bounded sample, rank 0 / N = 1 only.

--impl reference times that CPU port alone (the reference is Python on sklearn / PyG; it cannot be
installed here: torch_geometric is absent and /root/reference does not travel to the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: frames per GPU, points per frame, k, layers, channels, generator
    "headline_100k_k16_4x64": dict(frames=1, points=100_000, k=16, layers=4, channels=64, gen="uniform"),
    "config2_10k_k16_4x64": dict(frames=1, points=10_000, k=16, layers=4, channels=64, gen="uniform"),
    "config3_64x300_k20_8x128": dict(frames=64, points=300, k=20, layers=8, channels=128, gen="radar"),
    "config5_125k_k16_4x64": dict(frames=1, points=125_000, k=16, layers=4, channels=64, gen="uniform"),
}
DEFAULT_WORKLOAD = "headline_100k_k16_4x64"
EDGE_FEATURES = ["relative_position"]   # translation-invariant edge_attr, De = 2
CPU_SAMPLE_POINTS = 10_000               # CPU baseline runs the reference path on this many points


# ---------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------
def make_frames(wl: dict, rank: int):
    from radargnn_b200 import synthetic
    frames = []
    for f in range(wl["frames"]):
        seed = rank * 1000 + f
        if wl["gen"] == "uniform":
            frames.append(synthetic.uniform_square(wl["points"], seed=seed))
        else:
            frames.append(synthetic.radar_frame(wl["points"], seed=seed))
    X, V, ptr = synthetic.frame_batch(frames)
    x0 = synthetic.node_embeddings(X.shape[0], wl["channels"], seed=rank)
    return X.astype(np.float32), V.astype(np.float32), ptr, x0


def make_params(wl: dict, seed: int = 0):
    """PyG-default initialised MPNNConv stack (uniform(+-1/sqrt(fan_in))), reference key names."""
    import torch
    g = torch.Generator().manual_seed(seed)
    c, de = wl["channels"], 2
    p = 2 * c + de
    params = {}

    def lin(key, o, i):
        b = 1.0 / i ** 0.5
        params[key + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b
        params[key + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    for l in range(wl["layers"]):
        lin(f"convs.{l}.pre_mlp.0", p, p)
        lin(f"convs.{l}.post_mlp.0", c, p + c)
        params[f"batch_norms.{l}.module.weight"] = torch.ones(c)
        params[f"batch_norms.{l}.module.bias"] = torch.zeros(c)
    return params


def algorithmic_bytes(n: int, e: int, wl: dict) -> int:
    """B_alg = B_build + sum_l B_layer, every tensor at its API dtype."""
    c, de, layers = wl["channels"], 2, wl["layers"]
    build = 16 * n + 16 * e + 4 * de * e
    layer = 4 * n * c + 16 * e + 4 * de * e + 4 * n * c
    return build + layers * layer


def kernel_algorithmic_bytes(name: str, n: int, e: int, wl: dict):
    """Compulsory bytes of ONE launch of a kernel family ("Kernels"): every input and
    output touched once at unique-row granularity, weights ignored."""
    c, de = wl["channels"], 2
    p = 2 * c + de
    pp = (p + 3) // 4 * 4
    table = {
        # B rows (unique) + slot sources + slot edge attributes + row pointers + M rows
        "edge_aggregate": 4 * n * pp + 4 * e + 4 * de * e + 4 * (n + 1) + 4 * n * pp,
        "node_gemm_pre": 4 * n * c + 4 * n * pp,             # read x, write B = x W_s^T
        "node_gemm_post": 4 * n * c + 4 * n * pp + 4 * n * c,  # read x and M, write h
        "linear_pre_node": 4 * n * c + 4 * n * pp,
        "linear_post": 4 * n * c + 4 * n * pp + 4 * n * c,
        "knn_query": 8 * n + 16 * n + 16 * e + 4 * n,        # sorted points + ids/cells + edge_index + in-degree
        "bn_statistics": 4 * n * c,
        "edge_features": 16 * e + 16 * n + 4 * de * e,
    }
    return table.get(name)


# ---------------------------------------------------------------------------------------------
# CPU port of the reference path (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(sample, params, wl):
    """The reference's CPU path on one frame: sklearn k-NN (graph.py:57-63, 1 thread), the per-edge
    Python feature loop (graph.py:172-223), then the PyG-equivalent MPNN forward on all cores."""
    import torch
    from oracle import mpnn_oracle, reference_loop
    X, V, x0 = sample
    E, _ = reference_loop.build_edges_like_reference(X, "knn", wl["k"], 0.0, dense=False)
    ef = reference_loop.edge_feature_loop(X, V, E, EDGE_FEATURES, "directed")
    with torch.no_grad():
        h = mpnn_oracle.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.astype(np.int64)),
                                           torch.from_numpy(ef.astype(np.float32)), wl["layers"], "MPNNConv", "max")
    return E.shape[0], float(h.mean())


def cpu_sample(wl):
    from radargnn_b200 import synthetic
    fr = synthetic.uniform_square(CPU_SAMPLE_POINTS, seed=12345)
    return fr.X_cc, fr.V_cc_compensated, synthetic.node_embeddings(CPU_SAMPLE_POINTS, wl["channels"], seed=3)


def run_cpu_baseline(wl, steps: int, warmup: int):
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host core it can
    try:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    except RuntimeError:
        pass
    params = make_params(wl)
    sample = cpu_sample(wl)
    for _ in range(warmup):
        cpu_reference_step(sample, params, wl)
    t0 = time.perf_counter()
    edges = 0
    for _ in range(steps):
        e, _ = cpu_reference_step(sample, params, wl)
        edges += e
    dt = time.perf_counter() - t0
    return dict(value=edges / dt, unit="edges/s", cores=int(torch.get_num_threads()), kind="port",
                sample=(f"{CPU_SAMPLE_POINTS} of the workload's points (one frame, same density, k={wl['k']}, "
                        f"{wl['layers']}x{wl['channels']}): sklearn k-NN 1 thread + per-edge Python feature loop "
                        f"+ torch-CPU MPNN forward on all cores; {steps} steps"),
                ms_per_step=dt / steps * 1e3, host_cpus=os.cpu_count())


def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    warmup = max(0, min(args.warmup, 3))
    base = run_cpu_baseline(wl, steps, warmup)
    line = {
        "impl": "reference", "metric": "edges/s: graph-build + 4-layer MPNN fwd", "value": base["value"],
        "unit": "edges/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (graph build in f64)", "data": "synthetic",
        "config": {"workload": args.workload, "cpu_sample_points": CPU_SAMPLE_POINTS},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, reasons, mx = [], set(), None
        for ln in self.file.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            # median over the upper half: samples taken while the GPU was busy
            busy = sm[len(sm) // 2:]
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def main_gpu(args, wl):
    import torch
    import torch.distributed as dist
    from radargnn_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: radargnn_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- inputs: pinned host copies (e2e leg) and resident device copies (value leg) ----------
    X, V, ptr, x0 = make_frames(wl, rank)
    n = X.shape[0]
    pos_h, vel_h, x0_h = (torch.from_numpy(a).pin_memory() for a in (X, V, x0))
    pos_d, vel_d, x0_d = pos_h.to(dev), vel_h.to(dev), x0_h.to(dev)
    params = make_params(wl)
    layers, bn = [], []
    for l in range(wl["layers"]):
        pre = [(params[f"convs.{l}.pre_mlp.0.weight"].to(dev), params[f"convs.{l}.pre_mlp.0.bias"].to(dev))]
        post = [(params[f"convs.{l}.post_mlp.0.weight"].to(dev), params[f"convs.{l}.post_mlp.0.bias"].to(dev))]
        layers.append(ops.ConvParams("MPNNConv", wl["channels"], wl["channels"], 2, "max", pre, post))
        bn.append((params[f"batch_norms.{l}.module.weight"].to(dev), params[f"batch_norms.{l}.module.bias"].to(dev)))
    cfg = ops.PipelineConfig(layers=layers, bn=bn, algorithm="knn", k=wl["k"], edge_features=EDGE_FEATURES)
    handle = ops._PipelineHandle(cfg)
    n_frames = len(ptr) - 1
    n_edges = ops.knn_edge_count(ptr, wl["k"])
    c_last = wl["channels"]

    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
    edge_attr = torch.empty((n_edges, handle.edge_dim), dtype=torch.float32, device=dev)
    h = torch.empty((n, c_last), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.rgnn_pipeline_workspace_bytes(C.byref(handle.desc), n, n_frames, n_edges), dev)
    sum_ws = _lib.workspace(lib.rgnn_sum_workspace_bytes(), dev)
    loss = torch.zeros(2, dtype=torch.float64, device=dev)   # [sum of h, element count]
    loss[1] = float(n * c_last)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step_device():
        _lib.check(lib.rgnn_pipeline_forward(C.byref(handle.desc), pos_d.data_ptr(), vel_d.data_ptr(), x0_d.data_ptr(),
                                             ptr.ctypes.data, n_frames, edge_index.data_ptr(), n_edges,
                                             edge_attr.data_ptr(), h.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                                             ws.numel(), sp))
        _lib.check(lib.rgnn_sum_f32(h.data_ptr(), h.numel(), loss.data_ptr(), sum_ws.data_ptr(), sum_ws.numel(), sp))
        if world > 1:
            dist.all_reduce(loss[:1])   # the path's only collective: the loss scalar over NVLink

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(fn, steps):
        """Device time of `steps` calls, L2 flushed (untimed) before each; returns seconds."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) * 1e-3

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(seconds: float) -> float:
        if world == 1:
            return seconds
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: resident inputs ------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    count0 = _lib.launch_count()
    step_device()
    launches_per_step = _lib.launch_count() - count0   # kernels of librgnn_b200.so in one step
    # The ~40 launches of a step are captured once into a CUDA graph and replayed: the first kernels of a
    # step (14 short cell-list launches) are otherwise bound by host launch latency, not by the GPU.
    step_eager = step_device
    graph = None
    if not args.no_graph:
        def step_kernels():
            _lib.check(lib.rgnn_pipeline_forward(C.byref(handle.desc), pos_d.data_ptr(), vel_d.data_ptr(), x0_d.data_ptr(),
                                                 ptr.ctypes.data, n_frames, edge_index.data_ptr(), n_edges,
                                                 edge_attr.data_ptr(), h.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), torch.cuda.current_stream().cuda_stream))
            _lib.check(lib.rgnn_sum_f32(h.data_ptr(), h.numel(), loss.data_ptr(), sum_ws.data_ptr(), sum_ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_kernels()

        def step_device():
            graph.replay()
            if world > 1:
                dist.all_reduce(loss[:1])
        for _ in range(3):
            step_device()
        barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_dev = timed(step_device, args.steps)
    launches = launches_per_step * args.steps   # a graph replay launches the same kernels as the captured eager step
    barrier()
    t_dev = max_over_ranks(t_dev)
    if int(flag.item()) != 0:
        _lib.check(int(flag.item()))
    loss_value = float(loss[0].item() / (loss[1].item() * world)) if world > 1 else float(loss[0].item() / loss[1].item())

    # ---- e2e: host buffers through the C-ABI host entry point --------------------------------------
    h_host = torch.empty((n, c_last), dtype=torch.float32).pin_memory()
    host_ws = _lib.workspace(lib.rgnn_pipeline_host_workspace_bytes(C.byref(handle.desc), n, n_frames, n_edges,
                                                                      wl["channels"]), dev)

    def step_host():
        _lib.check(lib.rgnn_pipeline_forward_host(
            C.byref(handle.desc), pos_h.data_ptr(), vel_h.data_ptr(), x0_h.data_ptr(), wl["channels"],
            ptr.ctypes.data, n_frames, None, n_edges, None, h_host.data_ptr(), host_ws.data_ptr(), host_ws.numel(), sp))
        if world > 1:
            part = torch.tensor([float(h_host[0, 0])], dtype=torch.float64, device=dev)
            dist.all_reduce(part)

    for _ in range(3):
        step_host()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t_e2e = timed(step_host, e2e_steps)
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    h2d = int(pos_h.numel() * 4 + vel_h.numel() * 4 + x0_h.numel() * 4)
    d2h = int(h_host.numel() * 4)

    # ---- roofline of the dominant kernel: per-kernel CUDA events over the same steps ----------------
    _lib.profile_reset()
    _lib.profile_enable(True)
    timed(step_eager, args.steps)   # per-kernel events need eager launches
    _lib.profile_enable(False)
    totals = _lib.profile_totals()
    # keep the GPUs under the same load until nvidia-smi has had time to take a few samples (every
    # rank runs the same number of steps: step_device contains the loss all-reduce)
    for _ in range(30):
        for _ in range(10):
            step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    if rank == 0:
        top = max(totals.items(), key=lambda kv: kv[1][0]) if totals else (None, (0.0, 0))
        kname, (kms, kcount) = top
        kbytes = kernel_algorithmic_bytes(kname, n, n_edges, wl) if kname else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(kname)
        except (OSError, ValueError):
            pass
        roofline = None
        if kname and kbytes and kcount:
            achieved = kbytes / (kms / kcount * 1e-3) / 1e9
            roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                        "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_kind,
                        "us_per_launch": kms / kcount * 1e3, "launches_per_step": kcount / args.steps,
                        "algorithmic_bytes_per_launch": kbytes}
        total_edges = n_edges * world
        b_alg = algorithmic_bytes(n, n_edges, wl)
        step_s = t_dev / args.steps
        line = {
            "metric": "edges/s: graph-build + 4-layer MPNN fwd", "value": total_edges * args.steps / t_dev,
            "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (graph build in f64)", "data": "synthetic",
            "config": {"workload": args.workload, "points_per_gpu": n, "edges_per_gpu": n_edges,
                       "frames_per_gpu": n_frames, "k": wl["k"], "layers": wl["layers"], "channels": wl["channels"],
                       "edge_attr": "relative_position (De=2)", "aggr": "max", "parallelism": f"dp{world} (frames)",
                       "l2": "256 MiB memset between steps (untimed)", "launch": "eager" if graph is None else "cuda-graph replay of the step's kernels", "collective": "loss all-reduce (NCCL)" if world > 1 else "none"},
            "e2e": {"value": total_edges * e2e_steps / t_e2e, "unit": "edges/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / e2e_steps * 1e3,
                    "api": "rgnn_pipeline_forward_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "path_roofline": {"algorithmic_bytes_per_step": b_alg, "achieved": b_alg / step_s / 1e9, "peak": peak_gbs,
                              "unit": "GB/s", "frac": b_alg / step_s / 1e9 / peak_gbs},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(totals.items(), key=lambda kv: -kv[1][0])},
            "clocks": clocks, "loss": loss_value,
        }
        if world == 1 and not args.no_cpu_baseline:
            base = run_cpu_baseline(wl, steps=6, warmup=1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, wl)
    return main_gpu(args, wl)


if __name__ == "__main__":
    sys.exit(main())
