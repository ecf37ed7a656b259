#!/usr/bin/env python
"""bench.py -- edges/s of the fused hot path (graph build + L-layer MPNN forward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic radar frames: neighbour search
(-> edge_index, edge_attr) + L x (graph convolution, BatchNorm(train), ReLU) + the loss partial.
Frames are independent, so the N ranks of a box each process their own frames (weak scaling) and
only all-reduce the loss scalar (NCCL).  One JSON line is printed by rank 0:

  value        whole-job edges/s with the inputs already resident in HBM (CUDA-event timed, L2 flushed
               between steps, max over ranks);
  e2e          the same metric through the host-buffer C-ABI entry point as submit + wait with two batches in
               flight (rgnn_pipeline_submit_host / rgnn_pipeline_wait_host): H2D of pos / vel / x0 from pinned
               memory and D2H of EVERY output of the path (edge_index, edge_attr and the node embeddings) of every
               batch inside the timed region;  e2e_sync: one synchronous call per batch
               (rgnn_pipeline_forward_host), the latency figure;  e2e_embeddings_only: the synchronous call
               downloading only the node embeddings (the graph stays on the device);
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

# BASELINE.json configs as SURVEY.md section 8(d) defines them.  frames / points are PER GPU (weak scaling).
WORKLOADS = {
    "headline_100k_k16_4x64": dict(frames=1, points=100_000, search="knn", k=16, layers=4, c0=64, channels=64,
                                   gen="uniform", edge_features=["relative_position"]),
    "config1_300_r3_1x64": dict(frames=1, points=300, search="radius", r=3.0, k=6, layers=1, c0=5, channels=64,
                                gen="radar", edge_features=["relative_position"]),
    "config2_10k_k16_4x64": dict(frames=1, points=10_000, search="knn", k=16, layers=4, c0=64, channels=64,
                                 gen="uniform", edge_features=["relative_position"]),
    "config3_64x300_k20_8x128": dict(frames=64, points=300, search="knn", k=20, layers=8, c0=128, channels=128,
                                     gen="radar", edge_features=["relative_position"]),
    "config4_64x2000_k20_ppf_4x64": dict(frames=64, points=2000, search="knn", k=20, layers=4, c0=64, channels=64,
                                         gen="nuscenes", edge_features=["point_pair_features"]),
    "config5_125k_k16_4x64": dict(frames=1, points=125_000, search="knn", k=16, layers=4, c0=64, channels=64,
                                  gen="uniform", edge_features=["relative_position"]),
}
DEFAULT_WORKLOAD = "headline_100k_k16_4x64"
# The CPU arm runs the reference's path on ONE frame of at most this many points: the per-edge Python loop
# (graph.py:172-223) costs ~8 us (relative_position) to ~200 us (point-pair features) per edge, i.e. 17 s to
# minutes per step at the full 100 k points, and the driver times --steps 20 --warmup 5 of it.  Its cost
# per edge does not depend on the frame size, so edges/s of the sample is the rate of the full workload.
CPU_SAMPLE_POINTS = 25_000
CPU_SAMPLE_POINTS_PPF = 2_000


def edge_dim(wl: dict) -> int:
    return sum(4 if f == "point_pair_features" else 2 if f in ("relative_position", "relative_velocity") else 1
               for f in wl["edge_features"])


def metric_name(wl: dict) -> str:
    return f"edges/s: graph-build + {wl['layers']}-layer MPNN fwd"


def workload_config(name: str, wl: dict, world: int) -> dict:
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm."""
    de = edge_dim(wl)
    return {"workload": name, "frames_per_gpu": wl["frames"], "points_per_frame": wl["points"],
            "points_per_gpu": wl["frames"] * wl["points"], "search": wl["search"],
            "k": wl["k"] if wl["search"] == "knn" else None, "r": wl.get("r") if wl["search"] == "radius" else None,
            "layers": wl["layers"], "in_channels": wl["c0"], "channels": wl["channels"],
            "edge_attr": "+".join(wl["edge_features"]) + f" (De={de})", "aggr": "max",
            "parallelism": f"dp{world} (frames)"}


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
        elif wl["gen"] == "nuscenes":
            frames.append(synthetic.nuscenes_frame(wl["points"], seed=seed))
        else:
            frames.append(synthetic.radar_frame(wl["points"], seed=seed))
    X, V, ptr = synthetic.frame_batch(frames)
    x0 = synthetic.node_embeddings(X.shape[0], wl["c0"], seed=rank)
    return X.astype(np.float32), V.astype(np.float32), ptr, x0


def make_params(wl: dict, seed: int = 0):
    """PyG-default initialised MPNNConv stack (uniform(+-1/sqrt(fan_in))), reference key names."""
    import torch
    g = torch.Generator().manual_seed(seed)
    c, de = wl["channels"], edge_dim(wl)
    params = {}

    def lin(key, o, i):
        b = 1.0 / i ** 0.5
        params[key + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * b
        params[key + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    cin = wl["c0"]
    for l in range(wl["layers"]):
        p = 2 * cin + de
        lin(f"convs.{l}.pre_mlp.0", p, p)
        lin(f"convs.{l}.post_mlp.0", c, p + cin)
        params[f"batch_norms.{l}.module.weight"] = torch.ones(c)
        params[f"batch_norms.{l}.module.bias"] = torch.zeros(c)
        cin = c
    return params


def algorithmic_bytes(n: int, e: int, wl: dict) -> int:
    """B_alg = B_build + sum_l B_layer, every tensor at its API dtype."""
    c, de, layers = wl["channels"], edge_dim(wl), wl["layers"]
    build = 16 * n + 16 * e + 4 * de * e
    total, cin = build, wl["c0"]
    for _ in range(layers):
        total += 4 * n * cin + 16 * e + 4 * de * e + 4 * n * c
        cin = c
    return total


def kernel_algorithmic_bytes(name: str, n: int, e: int, wl: dict):
    """Compulsory bytes of ONE launch of a kernel family ("Kernels"): every input and
    output touched once at unique-row granularity, weights ignored (channels of the stack's inner layers)."""
    c, de = wl["channels"], edge_dim(wl)
    p = 2 * c + de
    pp = (p + 3) // 4 * 4
    table = {
        # B rows (unique) + slot sources + slot edge attributes + row pointers + M rows
        "edge_aggregate": 4 * n * pp + 4 * e + 4 * de * e + 4 * (n + 1) + 4 * n * pp,
        # fused aggregate + node update: B main rows (unique) + slot sources / attributes / row pointers + reduced
        # tail channels + x + h  (M' never exists in HBM)
        "edge_update_fused": 4 * n * 2 * c + 4 * e + 4 * de * e + 4 * (n + 1) + 16 * n + 4 * n * c + 4 * n * c,
        # tail channels of the messages: B tail rows + slot sources / attributes / row pointers + reduced rows
        "edge_tail_reduce": 16 * n + 4 * e + 4 * de * e + 4 * (n + 1) + 16 * n,
        "node_gemm_pre": 4 * n * c + 4 * n * pp,             # read x, write B = x W_s^T
        "node_gemm_post": 4 * n * c + 4 * n * pp + 4 * n * c,  # read x and M, write h
        "linear_pre_node": 4 * n * c + 4 * n * pp,
        "linear_post": 4 * n * c + 4 * n * pp + 4 * n * c,
        "knn_query": 8 * n + 16 * n + 16 * e + 4 * n,        # sorted points + ids/cells + edge_index + in-degree
        "bn_statistics": 4 * n * c,
        "edge_features": 16 * e + 16 * n + 4 * de * e,
        "csc_build_edge_attr": 16 * e + 8 * e + 2 * 4 * de * e + 16 * n,
    }
    return table.get(name)


# ---------------------------------------------------------------------------------------------
# CPU port of the reference path (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(sample, params, wl):
    """The reference's CPU path on one frame: sklearn k-NN / radius (graph.py:57-63 / 73-79, 1 thread), the
    per-edge Python feature loop (graph.py:172-223), then the PyG-equivalent MPNN forward on all cores."""
    import torch
    from oracle import mpnn_oracle, reference_loop
    X, V, x0 = sample
    E, _ = reference_loop.build_edges_like_reference(X, wl["search"], wl["k"], wl.get("r", 0.0), dense=False)
    ef = reference_loop.edge_feature_loop(X, V, E, wl["edge_features"], "directed")
    with torch.no_grad():
        h = mpnn_oracle.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.astype(np.int64)),
                                           torch.from_numpy(ef.astype(np.float32)), wl["layers"], "MPNNConv", "max")
    return E.shape[0], float(h.mean())


def cpu_sample_points(wl) -> int:
    cap = CPU_SAMPLE_POINTS_PPF if "point_pair_features" in wl["edge_features"] else CPU_SAMPLE_POINTS
    return min(wl["points"], cap)


def cpu_sample(wl):
    """One frame of the workload's generator (same density, same k / r, same layer stack)."""
    from radargnn_b200 import synthetic
    n = cpu_sample_points(wl)
    if wl["gen"] == "uniform":
        fr = synthetic.uniform_square(n, seed=12345)
    elif wl["gen"] == "nuscenes":
        fr = synthetic.nuscenes_frame(n, seed=12345)
    else:
        fr = synthetic.radar_frame(n, seed=12345)
    return fr.X_cc, fr.V_cc_compensated, synthetic.node_embeddings(n, wl["c0"], seed=3)


def run_cpu_baseline(wl, steps: int, warmup: int):
    import torch
    # torchrun exports OMP_NUM_THREADS=1; the reference arm may use every host core it can
    try:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    except RuntimeError:
        pass
    params = make_params(wl)
    sample = cpu_sample(wl)
    n = cpu_sample_points(wl)
    for _ in range(warmup):
        cpu_reference_step(sample, params, wl)
    t0 = time.perf_counter()
    edges = 0
    for _ in range(steps):
        e, _ = cpu_reference_step(sample, params, wl)
        edges += e
    dt = time.perf_counter() - t0
    full = wl["frames"] * wl["points"]
    why = ("the whole per-GPU workload" if n == full else
           f"one frame of {n} of the workload's {full} points per GPU (same generator, density, search and layer "
           f"stack): the reference's per-edge Python loop makes a full-size step take tens of seconds to minutes, "
           f"and its cost per edge does not depend on the frame size")
    return dict(value=edges / dt, unit="edges/s", cores=int(torch.get_num_threads()), kind="port",
                sample=(f"{why}; sklearn {wl['search']} search 1 thread + per-edge Python feature loop + torch-CPU "
                        f"MPNN forward ({wl['layers']}x{wl['channels']}) on all cores; {steps} steps"),
                ms_per_step=dt / steps * 1e3, host_cpus=os.cpu_count(), sample_points=n)


def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    base = run_cpu_baseline(wl, steps, warmup)
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": base["value"],
        "unit": "edges/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (graph build in f64)", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "cpu_sample_points": base["sample_points"],
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


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU's PCIe root hangs off BEFORE any pinned host
    buffer is allocated (first touch places the pages): the e2e leg's copies then stay node-local.
    Returns (numa node or None, pci bus id or None)."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus[len("00000000:"):]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None, bus
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node, bus
    except Exception:
        return None, None


def main_gpu(args, wl):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa_node, pci_bus = bind_to_gpu_numa_node(local_rank)

    import torch
    import torch.distributed as dist
    from radargnn_b200 import _lib, ops

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
    de = edge_dim(wl)
    layers, bn = [], []
    cin = wl["c0"]
    for l in range(wl["layers"]):
        pre = [(params[f"convs.{l}.pre_mlp.0.weight"].to(dev), params[f"convs.{l}.pre_mlp.0.bias"].to(dev))]
        post = [(params[f"convs.{l}.post_mlp.0.weight"].to(dev), params[f"convs.{l}.post_mlp.0.bias"].to(dev))]
        layers.append(ops.ConvParams("MPNNConv", cin, wl["channels"], de, "max", pre, post))
        bn.append((params[f"batch_norms.{l}.module.weight"].to(dev), params[f"batch_norms.{l}.module.bias"].to(dev)))
        cin = wl["channels"]
    cfg = ops.PipelineConfig(layers=layers, bn=bn, algorithm=wl["search"], k=wl["k"], r=wl.get("r", 1.0),
                             edge_features=wl["edge_features"])
    handle = ops._PipelineHandle(cfg)
    n_frames = len(ptr) - 1
    knn = wl["search"] == "knn"
    n_edges = ops.knn_edge_count(ptr, wl["k"]) if knn else ops._pipeline_edge_count(cfg, pos_d, vel_d, ptr)
    c_last = wl["channels"]

    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
    edge_attr = torch.empty((n_edges, handle.edge_dim), dtype=torch.float32, device=dev)
    h = torch.empty((n, c_last), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.rgnn_pipeline_workspace_bytes(C.byref(handle.desc), n, n_frames, n_edges), dev)
    sum_ws = _lib.workspace(lib.rgnn_sum_workspace_bytes(), dev)
    loss = torch.zeros(2, dtype=torch.float64, device=dev)   # [sum of h, element count]
    loss[1] = float(n * c_last)
    loss_global = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step_kernels():
        cur = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.rgnn_pipeline_forward(C.byref(handle.desc), pos_d.data_ptr(), vel_d.data_ptr(), x0_d.data_ptr(),
                                             ptr.ctypes.data, n_frames, edge_index.data_ptr(), n_edges,
                                             edge_attr.data_ptr(), h.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                                             ws.numel(), cur))
        _lib.check(lib.rgnn_sum_f32(h.data_ptr(), h.numel(), loss.data_ptr(), sum_ws.data_ptr(), sum_ws.numel(), cur))

    def all_reduce_loss():
        # the path's only collective: the loss scalar over NVLink (NCCL).  The reduced copy is a separate
        # buffer so that the graph-replayed producer never races with the in-place all-reduce.
        loss_global.copy_(loss[:1])
        dist.all_reduce(loss_global)

    def step_device():
        step_kernels()
        if world > 1:
            all_reduce_loss()

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
    step_kernels()
    launches_per_step = _lib.launch_count() - count0   # kernels of librgnn_b200.so in one step
    # The launches of a step are captured once into a CUDA graph and replayed: the first kernels of a step
    # (short cell-list launches) are otherwise bound by host launch latency, not by the GPU.  With N > 1 the
    # NCCL all-reduce of the loss is captured into the same graph (no eager launch gap after the replay).
    # A radius search hands its edge count to the host inside the call and cannot be captured.
    step_eager = step_device
    graph = None
    graph_has_collective = False
    collective_mode = "eager after the step" if world > 1 else "none"
    if not args.no_graph and knn:
        if world > 1 and not args.no_graph_collective:
            # The 8-byte all-reduce costs ~50 us of latency on the critical path when it follows the step (N = 2:
            # 0.714 -> 0.769 ms).  Nothing in the next step depends on the reduced loss, so it is software-pipelined by
            # one step: the graph of step k carries, as a parallel branch next to the graph build, the all-reduce of the
            # loss step k - 1 left behind.  Every timed step still contains exactly one all-reduce.
            try:
                side = torch.cuda.Stream(device=dev)
                copy_done = torch.cuda.Event()
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    cur = torch.cuda.current_stream()
                    side.wait_stream(cur)                      # fork
                    with torch.cuda.stream(side):
                        loss_global.copy_(loss[:1])            # the previous step's partial
                        copy_done.record(side)
                        dist.all_reduce(loss_global)
                    _lib.check(lib.rgnn_pipeline_forward(C.byref(handle.desc), pos_d.data_ptr(), vel_d.data_ptr(), x0_d.data_ptr(),
                                                         ptr.ctypes.data, n_frames, edge_index.data_ptr(), n_edges,
                                                         edge_attr.data_ptr(), h.data_ptr(), flag.data_ptr(), ws.data_ptr(),
                                                         ws.numel(), cur.cuda_stream))
                    cur.wait_event(copy_done)                  # the partial is overwritten only after it has been copied
                    _lib.check(lib.rgnn_sum_f32(h.data_ptr(), h.numel(), loss.data_ptr(), sum_ws.data_ptr(), sum_ws.numel(),
                                                cur.cuda_stream))
                    cur.wait_stream(side)                      # join
                graph, graph_has_collective = g2, True
                collective_mode = "in the step's CUDA graph, pipelined by one step (overlaps the graph build)"
            except Exception:
                torch.cuda.synchronize()
                graph = None
        if graph is None:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step_kernels()

        def step_device():
            graph.replay()
            if world > 1 and not graph_has_collective:
                all_reduce_loss()
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
    if world > 1 and graph_has_collective:
        all_reduce_loss()   # the pipelined graph reduced the partial of the step before the last one
    loss_value = float(loss_global[0].item() / (loss[1].item() * world)) if world > 1 else float(loss[0].item() / loss[1].item())

    # ---- e2e: host buffers through the C-ABI host entry point --------------------------------------
    h_host = torch.empty((n, c_last), dtype=torch.float32).pin_memory()
    ei_host = torch.empty((2, n_edges), dtype=torch.int64).pin_memory()
    ea_host = torch.empty((n_edges, handle.edge_dim), dtype=torch.float32).pin_memory()
    host_ws = _lib.workspace(lib.rgnn_pipeline_host_workspace_bytes(C.byref(handle.desc), n, n_frames, n_edges,
                                                                      wl["c0"]), dev)

    def make_step_host(full: bool):
        ei_p = ei_host.data_ptr() if full else None
        ea_p = ea_host.data_ptr() if full else None

        def step_host():
            _lib.check(lib.rgnn_pipeline_forward_host(
                C.byref(handle.desc), pos_h.data_ptr(), vel_h.data_ptr(), x0_h.data_ptr(), wl["c0"],
                ptr.ctypes.data, n_frames, ei_p, n_edges, ea_p, h_host.data_ptr(), host_ws.data_ptr(), host_ws.numel(), sp))
            if world > 1:
                all_reduce_loss()   # persistent device scalars: no allocation, no host sync per step
        return step_host

    # the same entry point as submit + wait with two batches in flight (rgnn_pipeline_submit_host): batch i + 1
    # uploads and batch i - 1 downloads while batch i computes.  Each slot has its own pinned outputs and workspace;
    # the L2 flush sits on the caller's stream in front of every submit, INSIDE the timed region.
    depth = max(1, min(int(os.environ.get("RGNN_BENCH_E2E_DEPTH", "2")), _lib.HOST_SLOTS))
    slots = []
    for _ in range(depth):
        slots.append({"h": torch.empty((n, c_last), dtype=torch.float32).pin_memory(),
                      "ei": torch.empty((2, n_edges), dtype=torch.int64).pin_memory(),
                      "ea": torch.empty((n_edges, handle.edge_dim), dtype=torch.float32).pin_memory(),
                      "ws": _lib.workspace(host_ws.numel(), dev)})

    def run_pipelined(steps: int, full: bool = True) -> float:
        """Device-clock time of `steps` batches through submit / wait, every batch's copies and the drain included."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for i in range(steps):
            sl = slots[i % depth]
            if i >= depth:
                _lib.check(lib.rgnn_pipeline_wait_host(i % depth))
            flush.zero_()
            _lib.check(lib.rgnn_pipeline_submit_host(
                i % depth, C.byref(handle.desc), pos_h.data_ptr(), vel_h.data_ptr(), x0_h.data_ptr(), wl["c0"],
                ptr.ctypes.data, n_frames, sl["ei"].data_ptr() if full else None, n_edges,
                sl["ea"].data_ptr() if full else None, sl["h"].data_ptr(),
                sl["ws"].data_ptr(), sl["ws"].numel(), sp))
            if world > 1:
                all_reduce_loss()
        for i in range(max(0, steps - depth), steps):
            _lib.check(lib.rgnn_pipeline_wait_host(i % depth))
        b.record(stream)       # the host has every batch's outputs by now: the event closes the region on the device clock
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3

    e2e_steps = max(3, min(args.steps, 20))
    e2e = {}
    for key, full in (("e2e_sync", True), ("e2e_embeddings_only", False)):
        fn = make_step_host(full)
        for _ in range(3):
            fn()
        barrier()
        t = max_over_ranks(timed(fn, e2e_steps))
        barrier()
        d2h = int(h_host.numel() * 4 + (ei_host.numel() * 8 + ea_host.numel() * 4 if full else 0))
        if full:
            h_host_full = h_host.clone()
        e2e[key] = {"value": n_edges * world * e2e_steps / t, "unit": "edges/s",
                    "h2d_bytes_per_step": int(pos_h.numel() * 4 + vel_h.numel() * 4 + x0_h.numel() * 4),
                    "d2h_bytes_per_step": d2h, "ms_per_step": t / e2e_steps * 1e3,
                    "api": "rgnn_pipeline_forward_host (pinned host buffers)",
                    "outputs": "edge_index + edge_attr + node embeddings" if full else "node embeddings"}

    run_pipelined(4)
    barrier()
    t = max_over_ranks(run_pipelined(e2e_steps))
    barrier()
    for sl in slots:     # the pipelined batches carry the bytes of the synchronous call
        assert torch.equal(sl["h"], h_host_full) and torch.equal(sl["ei"], ei_host), "pipelined host path differs"
    e2e["e2e"] = dict(e2e["e2e_sync"], value=n_edges * world * e2e_steps / t, ms_per_step=t / e2e_steps * 1e3,
                      api="rgnn_pipeline_submit_host / rgnn_pipeline_wait_host (pinned host buffers), "
                          f"{depth} batches in flight, L2 flush inside the timed region",
                      batches_in_flight=depth, sync_call_ms_per_step=e2e["e2e_sync"]["ms_per_step"])
    run_pipelined(4, full=False)
    barrier()
    t = max_over_ranks(run_pipelined(e2e_steps, full=False))
    barrier()
    e2e["e2e_embeddings_only"] = dict(e2e["e2e_embeddings_only"], value=n_edges * world * e2e_steps / t,
                                      ms_per_step=t / e2e_steps * 1e3, api=e2e["e2e"]["api"], batches_in_flight=depth,
                                      sync_call_ms_per_step=e2e["e2e_embeddings_only"]["ms_per_step"])

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
        traffic, traffic_source = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
                tj = json.load(fh)
            traffic = tj.get(args.workload, {}).get(kname)
            if traffic is not None:
                traffic_source = tj.get("_source", "profiles/ncu_traffic.json (static: one ncu --set full capture, not this run)")
        except (OSError, ValueError):
            pass
        roofline = None
        if kname and kbytes and kcount:
            achieved = kbytes / (kms / kcount * 1e-3) / 1e9
            roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                        "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_source,
                        "peak_source": peak_kind,
                        "us_per_launch": kms / kcount * 1e3, "launches_per_step": kcount / args.steps,
                        "algorithmic_bytes_per_launch": kbytes}
        total_edges = n_edges * world
        b_alg = algorithmic_bytes(n, n_edges, wl)
        step_s = t_dev / args.steps
        config = workload_config(args.workload, wl, world)
        config.update({
            "edges_per_gpu": n_edges,
            "l2": "256 MiB memset between steps (untimed)",
            "launch": "eager" if graph is None else ("cuda-graph replay of the step's kernels" +
                                                      (" + the NCCL all-reduce" if graph_has_collective else "")),
            "collective": ("loss all-reduce (NCCL), " + collective_mode) if world > 1 else "none",
            "numa_node": numa_node, "pci_bus": pci_bus})
        line = {
            "metric": metric_name(wl), "value": total_edges * args.steps / t_dev,
            "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (graph build in f64)", "data": "synthetic",
            "config": config,
            "e2e": e2e["e2e"], "e2e_sync": e2e["e2e_sync"], "e2e_embeddings_only": e2e["e2e_embeddings_only"],
            "gpu_launches": int(launches),
            "roofline": roofline,
            "path_roofline": {"algorithmic_bytes_per_step": b_alg, "achieved": b_alg / step_s / 1e9, "peak": peak_gbs,
                              "unit": "GB/s", "frac": b_alg / step_s / 1e9 / peak_gbs},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(totals.items(), key=lambda kv: -kv[1][0])},
            "clocks": clocks, "loss": loss_value,
        }
        if world == 1 and not args.no_cpu_baseline:
            base = run_cpu_baseline(wl, steps=4, warmup=1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown must not be able to hang the launcher (the 2-GPU run of this round printed its line and then sat
        # in the communicator's destruction while the CUDA graph holding the captured NCCL all-reduce was still
        # alive): drop the graph first, synchronise, and keep a watchdog that ends the process if NCCL still blocks.
        import gc
        import threading
        torch.cuda.synchronize()
        step_device = step_eager = None   # noqa: F841 (closures over the graph)
        graph = None                      # noqa: F841
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        watchdog.cancel()
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
    ap.add_argument("--no-graph-collective", action="store_true", help="keep the NCCL all-reduce out of the CUDA graph (N > 1)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, wl)
    return main_gpu(args, wl)


if __name__ == "__main__":
    sys.exit(main())
