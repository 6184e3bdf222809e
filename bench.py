#!/usr/bin/env python
"""Benchmark of the GKGNet graph hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (both clauses of the metric, one line)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores
    python bench.py --workload layer|train|infer ...         # one of the parts alone

BASELINE.json's metric has two clauses and the default run measures both, in one JSON line:

  * "GKGNet-576 images/sec (fwd+bwd, 1/2/4/8 B200)" -- the headline `value` / `ms_per_step` / `e2e`: one training
    step of GKGNet-576 (BASELINE configs[3]: pvig_s backbone + label-query head, random init, bf16 autocast,
    fwd + loss + bwd + grad clip + AdamW, 16 images per GPU).  N GPUs shard by image: the only collective is the NCCL
    gradient all-reduce (mean over ranks, what DDP computes), inside the timed region; value = images of all ranks /
    max-over-ranks device time.  The step is captured once in a CUDA graph and replayed (parallel.GraphedTrainStep;
    --no-graph issues it eagerly through DistributedDataParallel); `e2e` uploads every batch from pinned host memory one
    batch ahead on a copy stream and reads every step's loss back.
  * "Grapher kNN+agg us/layer, % roofline" -- keys `layer`, `phase_ms`, `roofline*`: the stage-1 Grapher layer hot
    path (BASELINE configs[1]: B=32 images per GPU, C=80, N=144x144 patches, M=1296 pooled keys, G=2, k=9): kNN-graph
    construction (normalise + distance + top-k), max-relative aggregation forward and backward through the C ABI,
    plus `roofline_knn_d200`, the kNN of the stage-3 shape (B=64, N=M=1296, D=200, k*d=18 and 27).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "GKGNet-576 images/sec (fwd+bwd, 1/2/4/8 B200); Grapher kNN+agg us/layer, % roofline"
WORKLOAD = dict(B=32, C=80, side=144, r=4, G=2, k=9, dilation=1)


def make_inputs(B, device, dtype, seed):
    """Synthetic stage-1 activations (unit-variance, like post-BN features), pooled keys and the
    analytic relative-position bias of the reference's Grapher(80, n=20736, r=4)."""
    from gkgnet_b200.pos_embed import relative_pos_table
    w = WORKLOAD
    N = w["side"] ** 2
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, N, w["C"], generator=g, dtype=torch.float32)
    x4 = x.view(B, w["side"], w["side"], w["C"]).permute(0, 3, 1, 2)
    y = torch.nn.functional.avg_pool2d(x4, w["r"], w["r"]).permute(0, 2, 3, 1).reshape(B, -1, w["C"])
    rel = relative_pos_table(w["C"], N, w["r"])[0]
    return x.to(dtype).contiguous(), y.to(dtype).contiguous(), rel.contiguous()


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
_REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = f"/tmp/gkg_clocks_{os.getpid()}.csv"

    def start(self):
        q = "index,clocks.sm,clocks.max.sm," + ",".join("clocks_event_reasons." + r for r in _REASONS)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 3 + len(_REASONS):
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(_REASONS, f[3:]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
class HotPath:
    """Device-resident state + one step of the hot path through the C ABI."""

    def __init__(self, x, y, rel, algo, dense_bias=False):
        from gkgnet_b200 import _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.x, self.y, self.rel = x, y, rel
        w = WORKLOAD
        self.B, self.N, self.C = x.shape
        self.M = y.shape[1]
        self.G, self.k, self.d = w["G"], w["k"], w["dilation"]
        self.D = self.C // self.G
        self.algo = algo
        dev = x.device
        self.dt = _lib.GKG_BF16 if x.dtype == torch.bfloat16 else _lib.GKG_F32
        self.ws_bytes = self.lib.gkg_knn_workspace_bytes(self.B, self.G, self.N, self.M, self.D, self.k,
                                                         self.d, 0, self.dt, algo)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.idx = torch.empty(self.B * self.G, self.N, self.k, dtype=torch.int32, device=dev)
        self.out = torch.empty(self.B, self.N, 2 * self.C, dtype=x.dtype, device=dev)
        self.amax = torch.empty(self.B, self.N, self.C, dtype=torch.uint8, device=dev)
        self.gout = torch.randn(self.B, self.N, 2 * self.C, device=dev).to(x.dtype)
        self.gx = torch.empty_like(x)
        self.gy = torch.zeros(self.B, self.M, self.C, dtype=torch.float32, device=dev)
        from gkgnet_b200 import ops
        fit = None if dense_bias else ops.fit_separable_bias(rel)
        self._sep_keep = fit
        self.sep = (None, None, 0, 0) if fit is None else (fit[0].data_ptr(), fit[1].data_ptr(), fit[2], fit[3])
        self.stream = torch.cuda.current_stream(dev)
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step(self, record=False):
        lib, s = self.lib, self.stream.cuda_stream
        x, y = self.x, self.y
        a = (self.B, self.G, self.N, self.M, self.D, self.k, self.d)
        if record:
            self.ev[0].record(self.stream)
        self._lib.check(lib.gkg_knn_prepare(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0),
                                            y.stride(1), *a, self.dt, self.algo, self.ws.data_ptr(),
                                            self.ws_bytes, s), "knn_prepare")
        if record:
            self.ev[1].record(self.stream)
        sa, sb, sw, skw = self.sep
        self._lib.check(lib.gkg_knn_select(x.data_ptr(), x.stride(0), x.stride(1), self.rel.data_ptr(), sa, sb, sw, skw,
                                           self.idx.data_ptr(), *a, 0, self.dt, self.algo, self.ws.data_ptr(),
                                           self.ws_bytes, s), "knn_select")
        if record:
            self.ev[2].record(self.stream)
        self._lib.check(lib.gkg_mr_aggregate_fwd(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(),
                                                 y.stride(0), y.stride(1), self.idx.data_ptr(),
                                                 self.out.data_ptr(), self.amax.data_ptr(), self.B, self.G,
                                                 self.N, self.M, self.D, self.k, self.dt, s), "agg_fwd")
        if record:
            self.ev[3].record(self.stream)
        self.gy.zero_()
        self._lib.check(lib.gkg_mr_aggregate_bwd(self.gout.data_ptr(), self.idx.data_ptr(), self.amax.data_ptr(),
                                                 self.gx.data_ptr(), self.gy.data_ptr(), self.B, self.G, self.N,
                                                 self.M, self.D, self.k, self.dt, s), "agg_bwd")
        if record:
            self.ev[4].record(self.stream)

    def phase_ms(self):
        return [self.ev[i].elapsed_time(self.ev[i + 1]) for i in range(4)]


def dist_setup(n_gpus):
    from gkgnet_b200 import parallel as P
    return P.init_distributed()


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(val, world, device):
    if world == 1:
        return val
    import torch.distributed as dist
    t = torch.tensor([val], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cpu_reference_step(x, y, rel, grad_out):
    """One step of the same workload through the CPU oracle (the reference's algorithm)."""
    from oracle import gkg_oracle as O
    w = WORKLOAD
    B, N, C = x.shape
    G, D = w["G"], C // w["G"]

    def ref_layout(t):
        n = t.shape[1]
        return t.reshape(B, n, G, D).permute(0, 2, 3, 1).reshape(B * G, D, n, 1)

    xr = ref_layout(x).contiguous().requires_grad_(True)
    yr = ref_layout(y).contiguous().requires_grad_(True)
    ei = O.dense_dilated_knn_graph(xr, yr, w["k"], w["dilation"], rel.unsqueeze(0))
    out = O.mr_aggregate(xr, ei, yr, in_channels=C)
    out.backward(grad_out)
    return out


def use_all_host_threads():
    """The CPU arm uses every host core this process may run on.  torchrun exports OMP_NUM_THREADS=1 when it
    starts more than one rank, which would silently make the reference single-threaded."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def time_cpu_reference(n_img, reps, seed=0):
    use_all_host_threads()
    x, y, rel = make_inputs(n_img, "cpu", torch.float32, seed)
    grad_out = torch.randn(n_img, 2 * WORKLOAD["C"], x.shape[1], 1)
    cpu_reference_step(x[:1], y[:1], rel, grad_out[:1])     # warm-up (thread pool, allocator)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(x, y, rel, grad_out)
        best = min(best, time.perf_counter() - t0)
    return n_img / best, best


def run_reference(args):
    """--impl reference: the reference's algorithm (oracle port, PyTorch CPU ops, every host core) for the same
    workloads: GKGNet-576 training step (headline; --workload layer: the stage-1 layer hot path)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    n_img = args.ref_images
    x, y, rel = make_inputs(n_img, "cpu", torch.float32, 0)
    grad_out = torch.randn(n_img, 2 * WORKLOAD["C"], x.shape[1], 1)

    def layer_time(steps, warm):
        for _ in range(warm):
            cpu_reference_step(x, y, rel, grad_out)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(x, y, rel, grad_out)
        return (time.perf_counter() - t0) / steps

    if args.workload == "layer":
        steps = max(1, min(args.steps, args.ref_max_steps))
        dt = layer_time(steps, max(1, min(args.warmup, 2)))
        val = n_img / dt
        sample = (f"{n_img} images per step (of the 32-image batch), fp32, stage-1 Grapher hot path "
                  f"(kNN graph + MR aggregate fwd+bwd), {steps} timed steps")
        cfg = config_dict(args.gpus, "reference algorithm (oracle port, PyTorch CPU ops) on host cores")
        extra = {}
    else:
        steps = max(1, min(args.steps, args.ref_train_steps))
        sd, params = _oracle_state("cpu")
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05)
        g = torch.Generator().manual_seed(5)
        nt = args.ref_train_images
        img = torch.randn(nt, 3, 576, 576, generator=g)
        tgt = (torch.rand(nt, 80, generator=g) < 0.04).float()
        _oracle_train_step(sd, params, img[:1], tgt[:1], opt)
        t0 = time.perf_counter()
        for _ in range(steps):
            _oracle_train_step(sd, params, img, tgt, opt)
        dt = (time.perf_counter() - t0) / steps
        val = nt / dt
        sample = (f"{nt} images per step (of the 16-image batch), fp32, GKGNet-576 fwd + head losses + bwd + grad clip + "
                  f"AdamW through oracle/gkg_oracle.py, {steps} timed steps")
        cfg = train_config(args.gpus, False)
        cfg["note"] = "reference algorithm (oracle port, PyTorch CPU ops) on host cores"
        ldt = layer_time(2, 1)
        extra = {"layer": {"value": n_img / ldt, "unit": "images/s", "ms_per_step": ldt * 1e3,
                           "sample": f"{n_img} images per step, stage-1 Grapher hot path (BASELINE configs[1]), 2 timed steps"}}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def train_config(world, syncbn, B=16):
    return {"workload": "BASELINE configs[3]: GKGNet-576 training fwd+bwd+AdamW, bf16 autocast, DDP",
            "images_per_gpu": B, "global_batch": B * world,
            "parallelism": f"dp{world} (NCCL gradient all-reduce)",
            "norm": "SyncBN" if syncbn else "per-GPU BN",
            "l2": "activations per step exceed the 126 MB L2; no explicit flush"}


def config_dict(n_gpus, note):
    w = WORKLOAD
    return {
        "workload": ("BASELINE configs[1]: stage-1 Grapher layer hot path of GKGNet-576 -- kNN graph "
                     "(normalise+distance+top-k) + max-relative aggregate fwd + bwd; "
                     f"B={w['B']}/GPU, C={w['C']}, N={w['side']}x{w['side']}, M={(w['side'] // w['r']) ** 2}, "
                     f"G={w['G']}, k={w['k']}, dilation={w['dilation']}, r={w['r']}"),
        "images_per_gpu": w["B"], "global_batch": w["B"] * n_gpus, "parallelism": f"dp{n_gpus} (no collective)",
        "l2": "inputs+outputs per step (>=370 MB) exceed the 126 MB L2; no explicit flush",
        "note": note,
    }


def run_layer(args):
    """BASELINE configs[1]: the stage-1 Grapher layer hot path.  Returns the JSON line (rank 0) or None."""
    world, rank, local = dist_setup(args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from gkgnet_b200 import _lib
    algo = {"auto": _lib.KNN_AUTO, "exact": _lib.KNN_EXACT_FP32, "tc": _lib.KNN_TCGEN05}[args.algo]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = WORKLOAD["B"]
    xh, yh, relh = make_inputs(B, "cpu", dtype, seed=rank)
    x, y, rel = xh.to(dev), yh.to(dev), relh.to(dev)
    hp = HotPath(x, y, rel, algo, dense_bias=args.dense_bias)

    for _ in range(max(args.warmup, 3)):
        hp.step()
    barrier(world)

    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    phases = [0.0] * 4
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    start.record(hp.stream)
    for _ in range(args.steps):
        hp.step()
    stop.record(hp.stream)
    barrier(world)
    total_ms = start.elapsed_time(stop)
    launches = _lib.launch_count() - l0 + args.steps     # + the gy.zero_() memset per step
    clocks = sampler.stop()

    # per-phase durations: separate short loop (events inside the timed loop are overwritten)
    reps = min(args.steps, 10)
    for _ in range(reps):
        hp.step(record=True)
        torch.cuda.synchronize()
        for i, v in enumerate(hp.phase_ms()):
            phases[i] += v / reps

    ms_per_step = max_over_ranks(total_ms / args.steps, world, dev)
    value = world * B / (ms_per_step * 1e-3)

    # ---- the grouped 1x1 FC that follows the aggregate, timed beside the step: the
    # step itself keeps the definition of BASELINE configs[1] (kNN + aggregate fwd + bwd)
    fc_ms = None
    if dtype == torch.bfloat16:
        from gkgnet_b200 import ops
        c2 = 2 * hp.C
        if ops.grouped_fc_supported(c2):
            gfc = torch.Generator(device="cpu").manual_seed(1)
            w_op = ops.grouped_fc_weights((torch.randn(c2, c2 // 4, 1, 1, generator=gfc) * (2.0 / (c2 // 4)) ** 0.5).to(dev))
            sh = torch.zeros(c2, device=dev)
            for _ in range(3):
                ops.grouped_fc(hp.out, w_op, sh, "gelu")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(hp.stream)
            for _ in range(reps):
                ops.grouped_fc(hp.out, w_op, sh, "gelu")
            e1.record(hp.stream)
            torch.cuda.synchronize()
            fc_ms = e0.elapsed_time(e1) / reps

    # ---- the aggregate in its inference form (no arg-max plane written), timed beside the step: exactly the
    # traffic of SURVEY 8(d)'s BYTES_agg; inside the step the kernel also writes the uint8 arg-max (B*C*N bytes)
    agg_inf_ms = None
    if True:
        s_ = hp.stream.cuda_stream
        def agg_inf():
            _lib.check(hp.lib.gkg_mr_aggregate_fwd(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0),
                                                   y.stride(1), hp.idx.data_ptr(), hp.out.data_ptr(), None, hp.B, hp.G,
                                                   hp.N, hp.M, hp.D, hp.k, hp.dt, s_), "agg_fwd")
        for _ in range(3):
            agg_inf()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(hp.stream)
        for _ in range(reps):
            agg_inf()
        e1.record(hp.stream)
        torch.cuda.synchronize()
        agg_inf_ms = e0.elapsed_time(e1) / reps

    # ---- the key pooling in front of the kNN (y = avg_pool2d(x, r, r)), timed beside the step
    pool_ms = None
    if True:
        from gkgnet_b200 import ops
        side = int(round(hp.N ** 0.5))
        rr = WORKLOAD.get("r", 4)
        with torch.no_grad():
            for _ in range(3):
                ops.pool_keys(x, side, side, rr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(hp.stream)
            for _ in range(reps):
                ops.pool_keys(x, side, side, rr)
            e1.record(hp.stream)
            torch.cuda.synchronize()
        pool_ms = e0.elapsed_time(e1) / reps

    # ---- e2e: host buffers -> device -> hot path -> host, every step ------------------
    # Every step copies its inputs from pinned host memory and its result back; the three legs run on
    # three streams over two sets of device buffers, so step i+1's upload and step i-1's download overlap
    # step i's kernels (all of it inside the timed region).
    xp, yp = xh.pin_memory(), yh.pin_memory()
    out_host = torch.empty(hp.out.shape, dtype=hp.out.dtype).pin_memory()
    e2e_steps = max(4, min(args.steps, 10))
    hps = [hp, HotPath(torch.empty_like(x), torch.empty_like(y), rel, algo, dense_bias=args.dense_bias)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmp = hp.stream
    hps[1].stream = s_cmp
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        for i in range(n):
            h, j = hps[i % 2], i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[j])            # the kernels that last read these input buffers
                h.x.copy_(xp, non_blocking=True)
                h.y.copy_(yp, non_blocking=True)
                ev_in[j].record(s_in)
            s_cmp.wait_event(ev_in[j])
            s_cmp.wait_event(ev_out[j])               # the download that last read this output buffer
            h.step()
            ev_cmp[j].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[j])
                out_host.copy_(h.out, non_blocking=True)
                ev_out[j].record(s_out)

    e2e_loop(2)
    torch.cuda.synchronize()
    barrier(world)
    s2, t2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record(s_in)
    e2e_loop(e2e_steps)
    s_out.wait_stream(s_cmp)
    s_out.wait_stream(s_in)
    t2.record(s_out)
    torch.cuda.synchronize()
    barrier(world)
    e2e_ms = max_over_ranks(s2.elapsed_time(t2) / e2e_steps, world, dev)
    h2d = xp.numel() * xp.element_size() + yp.numel() * yp.element_size()
    d2h = out_host.numel() * out_host.element_size()

    d200 = knn_d200_roofline(dev) if dtype == torch.bfloat16 else None
    if rank != 0:
        return None
    peaks, peak_src = load_peaks()
    w = WORKLOAD
    N, M, C, G, k = hp.N, hp.M, hp.C, hp.G, hp.k
    es = 2 if dtype == torch.bfloat16 else 4
    flop_knn = 2.0 * B * N * M * C
    bytes_agg = es * B * C * N + es * B * C * M + 4 * B * G * N * k + es * B * 2 * C * N
    bytes_bwd = es * B * 2 * C * N + B * C * N + 4 * B * G * N * k + es * B * C * N + 4 * B * C * M
    knn_ms = phases[1]
    tf = flop_knn / (knn_ms * 1e-3) / 1e12
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    # DRAM bytes of the select kernels per launch: from the ncu --set full capture of this shape (profiles/, see the
    # file for the command); bench.py cannot run under a profiler, so the figure is carried, not measured here
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "knn_select_traffic.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    roofline = {"kernel": "gkg_knn_select (tcgen05 distance + fused top-k kernel, finalize, re-rank and fix-up "
                          "kernels; the first is ~90 % of the time)", "bound": "tensor", "achieved": tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                "measured_in": "layer microbench (BASELINE configs[1]), CUDA events around gkg_knn_select inside the step",
                "frac_of_burst_peak": tf / peaks.get("bf16_tflops", peak_tf),
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside the step)",
                "algorithmic_flop": flop_knn, "ms": knn_ms}
    agg_gbs = bytes_agg / (phases[2] * 1e-3) / 1e9
    bwd_gbs = bytes_bwd / (phases[3] * 1e-3) / 1e9
    extra = {
        "phase_ms": {"knn_prepare": phases[0], "knn_select": phases[1], "agg_fwd": phases[2], "agg_bwd": phases[3]},
        "roofline_agg_fwd": {"bound": "hbm", "achieved": agg_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": agg_gbs / peaks["hbm_gbs"], "algorithmic_bytes": bytes_agg},
        "roofline_agg_bwd": {"bound": "hbm", "achieved": bwd_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": bwd_gbs / peaks["hbm_gbs"], "algorithmic_bytes": bytes_bwd},
    }
    extra["roofline_agg_fwd"]["note"] = ("training form: the launch also writes the uint8 arg-max plane (B*C*N bytes) the "
                                         "backward reads; SURVEY 8(d)'s BYTES_agg does not count it")
    extra["roofline_agg_fwd"]["achieved_with_argmax"] = (bytes_agg + B * C * N) / (phases[2] * 1e-3) / 1e9
    extra["roofline_agg_fwd"]["frac_with_argmax"] = extra["roofline_agg_fwd"]["achieved_with_argmax"] / peaks["hbm_gbs"]
    if agg_inf_ms is not None:
        extra["phase_ms"]["agg_fwd inference form (outside the step)"] = agg_inf_ms
        extra["roofline_agg_fwd_infer"] = {"bound": "hbm", "achieved": bytes_agg / (agg_inf_ms * 1e-3) / 1e9,
                                           "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                           "frac": bytes_agg / (agg_inf_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                           "algorithmic_bytes": bytes_agg}
    if pool_ms is not None:
        bytes_pool = es * B * C * N + es * B * C * M
        extra["phase_ms"]["pool_keys (outside the step)"] = pool_ms
        extra["roofline_pool_keys"] = {"bound": "hbm", "achieved": bytes_pool / (pool_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                       "unit": "GB/s", "frac": bytes_pool / (pool_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                       "algorithmic_bytes": bytes_pool}
    if fc_ms is not None:
        bytes_fc = 2 * es * B * 2 * C * N
        extra["phase_ms"]["fc_fwd (outside the step)"] = fc_ms
        extra["roofline_fc_fwd"] = {"bound": "hbm", "achieved": bytes_fc / (fc_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                    "unit": "GB/s", "frac": bytes_fc / (fc_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "algorithmic_bytes": bytes_fc}
    cpu_val, cpu_s = time_cpu_reference(args.cpu_images, 2) if world == 1 and not args.no_cpu else (None, None)
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(world, f"knn algo={args.algo}"),
        "clocks": clocks,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": None if cpu_val is None else {
            "value": cpu_val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{args.cpu_images} images, fp32, same hot path through oracle/gkg_oracle.py "
                      f"(PyTorch CPU ops), best of 2: {cpu_s:.2f} s"},
    }
    line.update(extra)
    if d200 is not None:
        peak_b = peaks.get("bf16_tflops", peak_tf)
        for e in d200:
            e.update(peak=peak_b, unit="TFLOP/s", frac=e["achieved"] / peak_b, bound="tensor",
                     peak_source=f"{peak_src} bf16_tflops (burst: kernel timed alone)")
        line["roofline_knn_d200"] = d200
    # ---- the same restatement of the reference run with PyTorch's own CUDA kernels on this GPU (fp32 eager, full
    # batch): what the reference's code path costs on a B200 today ("the real bar to beat").  A reported
    # baseline like cpu_baseline -- the checker's code, never the product path.
    if world == 1 and not args.no_cpu:
        try:
            xg, yg, relg = x.float(), y.float(), rel.float()
            gog = hp.gout.float().permute(0, 2, 1).unsqueeze(-1).contiguous()
            for _ in range(2):
                cpu_reference_step(xg, yg, relg, gog)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = 3
            e0.record()
            for _ in range(nrep):
                cpu_reference_step(xg, yg, relg, gog)
            e1.record()
            torch.cuda.synchronize()
            eager_ms = e0.elapsed_time(e1) / nrep
            line["eager_gpu_baseline"] = {
                "value": B / (eager_ms * 1e-3), "unit": "images/s", "ms_per_step": eager_ms, "kind": "port",
                "sample": "the whole 32-image step, fp32, oracle/gkg_oracle.py (the reference's algorithm as plain PyTorch "
                          "ops) on the same B200 with ATen/cuBLAS kernels, device-resident, 3 timed steps"}
        except Exception as e:        # e.g. out of memory next to the bench buffers: the line stands without it
            line["eager_gpu_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    return line


def knn_d200_roofline(dev):
    """kNN graph (gkg_knn_select phase) of the stage-3 shape of GKGNet-576 -- B=64 images, N=M=1296 (self keys),
    G=2 groups of D=200, k=9 with dilation 2 and 3 (k*d = 18, 27), bf16, analytic position bias -- where the
    distance GEMM has enough K for the tensor cores to matter.  Algorithmic FLOP = 2*B*N*M*C."""
    from gkgnet_b200 import _lib, ops
    from gkgnet_b200.pos_embed import relative_pos_table
    lib = _lib.load()
    B, C, n, G = 64, 400, 1296, 2
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(B, n, C, generator=g).to(torch.bfloat16).to(dev)
    rel = relative_pos_table(C, n, 1)[0].contiguous().to(dev)
    fit = ops.fit_separable_bias(rel)
    sep = (None, None, 0, 0) if fit is None else (fit[0].data_ptr(), fit[1].data_ptr(), fit[2], fit[3])
    out = []
    stream = torch.cuda.current_stream(dev)
    for d in (2, 3):
        a = (B, G, n, n, C // G, 9, d)
        wsb = lib.gkg_knn_workspace_bytes(*a, 1, _lib.GKG_BF16, _lib.KNN_AUTO)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        idx = torch.empty(B * G, n, 9, dtype=torch.int32, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        sel_ms = prep_ms = 0.0
        reps = 6
        for it in range(reps + 2):
            ev[0].record(stream)
            _lib.check(lib.gkg_knn_prepare(x.data_ptr(), x.stride(0), x.stride(1), None, 0, 0, *a, _lib.GKG_BF16,
                                           _lib.KNN_AUTO, ws.data_ptr(), wsb, stream.cuda_stream), "knn_prepare")
            ev[1].record(stream)
            _lib.check(lib.gkg_knn_select(x.data_ptr(), x.stride(0), x.stride(1), rel.data_ptr(), *sep, idx.data_ptr(),
                                          *a, 1, _lib.GKG_BF16, _lib.KNN_AUTO, ws.data_ptr(), wsb, stream.cuda_stream),
                       "knn_select")
            ev[2].record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                prep_ms += ev[0].elapsed_time(ev[1]) / reps
                sel_ms += ev[1].elapsed_time(ev[2]) / reps
        flop = 2.0 * B * n * n * C
        out.append({"kernel": "gkg_knn_select", "shape": f"B={B} N=M={n} D={C // G} G={G} k=9 dilation={d} (k*d={9 * d}), bf16",
                    "algorithmic_flop": flop, "ms": sel_ms, "prepare_ms": prep_ms,
                    "achieved": flop / (sel_ms * 1e-3) / 1e12})
    return out



# ----------------------------------------------------------------------------------------
# whole-model workloads (BASELINE configs[2] inference, configs[3] training) -- secondary lines
# ----------------------------------------------------------------------------------------
def run_model(args):
    """GKGNet-576 (pvig_s, 80 labels, random init) + LabelQueryHead under bf16 autocast.
    train: fwd + loss + bwd + grad clip + AdamW step, 16 images / GPU, DDP (NCCL gradient all-reduce
    only: per-GPU BatchNorm unless --syncbn).  infer: eval forward + sigmoid scores, 64 images / GPU,
    replicas only.  value = images of all ranks / max-over-ranks device time."""
    from gkgnet_b200 import parallel as P
    world, rank, local = P.init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import gkgnet_b200 as G
    from gkgnet_b200 import _lib
    train = args.workload == "train"
    B = args.batch or (16 if train else 64)
    G.set_norm_type("SyncBN" if args.syncbn else "BN")
    torch.manual_seed(0)
    net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.1 if train else 0.0)
    head = G.LabelQueryHead(80, 640)

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone, self.head = net, head

        def forward(self, img, tgt=None):
            feats = self.backbone(img)
            if tgt is None:
                return torch.sigmoid(self.head.get_score(feats))
            return sum(self.head.forward_train(feats, tgt).values())

    model = Model().to(dev)
    model.train(train)
    params = [p for p in model.parameters() if p.requires_grad]
    graphed = train and not args.no_graph
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True, capturable=graphed) if train else None
    ddp = P.data_parallel(model, dev) if train and not graphed else model
    g = torch.Generator().manual_seed(100 + rank)
    img_h = torch.randn(B, 3, 576, 576, generator=g).pin_memory()
    tgt_h = (torch.rand(B, 80, generator=g) < 0.04).float().pin_memory()
    img, tgt = img_h.to(dev), tgt_h.to(dev)
    stream = torch.cuda.current_stream(dev)
    # the whole step (fwd, loss, bwd, gradient all-reduce over NCCL, clip, AdamW) as ONE CUDA graph: issued from Python
    # the ~3000 launches of a 16-image step take longer than the kernels run (tools/host_bound.py)
    gstep, graph_note = None, None
    if graphed:
        try:
            gstep = P.GraphedTrainStep(model, opt, params, img, tgt, clip_norm=5.0, warmup=3)
        except Exception as exc:      # capture refused (driver / allocator state): measure the eager path, and say so
            graph_note = f"{type(exc).__name__}: {str(exc)[:160]}"
            graphed, gstep = False, None
            torch.cuda.synchronize(dev)
            for p_ in params:
                p_.grad = None
            opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)
            ddp = P.data_parallel(model, dev)

    igraph = None
    if not train and not args.no_graph:
        # inference: the eval forward captured once and replayed on a static input (device-bound either way: -3 %)
        try:
            static_img = img.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(stream)
            with torch.cuda.stream(side), torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                for _ in range(2):
                    model(static_img)
            stream.wait_stream(side)
            torch.cuda.synchronize(dev)
            igraph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(igraph), torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                static_out = model(static_img)
            igraph_launches = _lib.launch_count() - n0
        except Exception:
            igraph = None
            torch.cuda.synchronize(dev)

    def step(from_host=False):
        if gstep is not None:
            return gstep(img_h, tgt_h) if from_host else gstep()
        if igraph is not None:
            if from_host:
                static_img.copy_(img_h, non_blocking=True)
            igraph.replay()
            return static_out
        x, t = (img_h.to(dev, non_blocking=True), tgt_h.to(dev, non_blocking=True)) if from_host else (img, tgt)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if train:
                loss = ddp(x, t)
            else:
                with torch.no_grad():
                    return ddp(x)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(n, from_host):
        P.barrier(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        if from_host and gstep is not None:
            # end to end through the captured step: every batch is uploaded from pinned host memory inside the timed
            # region -- on a copy stream, one batch ahead (an input pipeline), so that PCIe time hides behind the step --
            # and every step's loss is read back
            gstep.prefetch(img_h, tgt_h)
            for i in range(n):
                out = gstep.step_prefetched()
                if i + 1 < n:
                    gstep.prefetch(img_h, tgt_h)
                out.item()
        else:
            for _ in range(n):
                out = step(from_host)
                if from_host:
                    out.float().sum().item() if not train else out.item()     # device -> host read of the result
        b.record(stream)
        P.barrier(dev)
        return P.max_over_ranks(a.elapsed_time(b) / n, dev)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ms = timed(args.steps, False)
    launches = _lib.launch_count() - l0
    if gstep is not None:
        launches = args.steps * gstep.launches_per_replay       # replayed from the graph: counted at capture
    elif igraph is not None:
        launches = args.steps * igraph_launches
    clocks = sampler.stop()
    # end to end: as many steps as the device-timed run when the input pipeline is one batch deep (its first upload is
    # not hidden; over 5 steps it would weigh 20 % of a copy per step), 3-5 steps otherwise
    timed(2, True)          # untimed warm-up of the host -> device path (first touch of the pinned batch, copy stream)
    e2e_ms = timed(max(3, args.steps if gstep is not None else min(args.steps, 5)), True)
    coll = None
    if train and world > 1:
        # the one collective of the step, timed alone: an all-reduce of the gradient bytes in DDP-sized buckets (25 MB),
        # back to back on the NCCL stream -- what the step would pay if none of it overlapped the backward pass
        import torch.distributed as dist
        nbytes = sum(p.numel() * 4 for p in params)
        # the captured step reduces the flat gradient buffer in one message, the eager DDP path in 25 MB buckets
        bucket = torch.zeros((nbytes if graphed else 25 * 1024 * 1024) // 4, device=dev)
        nb = max(1, (nbytes + bucket.numel() * 4 - 1) // (bucket.numel() * 4))
        for _ in range(2):
            for _ in range(nb):
                dist.all_reduce(bucket)
        torch.cuda.synchronize()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            for _ in range(nb):
                dist.all_reduce(bucket)
        c.record()
        torch.cuda.synchronize()
        coll_ms = P.max_over_ranks(a.elapsed_time(c) / 3, dev)
        coll = {"name": "NCCL all_reduce of the fp32 gradients (" + ("one message over the flat gradient buffer, as inside "
                        "the captured step" if graphed else "DDP buckets of 25 MB") + ")", "bytes_per_step": nbytes,
                "buckets": int(nb), "ms_alone": coll_ms, "share_of_step_if_exposed": coll_ms / ms,
                "bus_gbs": 2.0 * (world - 1) / world * nbytes / (coll_ms * 1e-3) / 1e9}
        del bucket
    eager = None
    if train and world == 1 and not args.no_cpu:
        eager = eager_train_baseline(model, img, tgt, dev)
    if rank != 0:
        return None
    out_bytes = 4 if train else B * 80 * 4
    line = {
        "metric": METRIC, "value": world * B / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(train_config(world, args.syncbn, B),
                       launch=("one CUDA graph per step (fwd + loss + bwd + NCCL gradient all-reduce + clip + AdamW)"
                               if graphed else "eager launches through DistributedDataParallel"
                               + (f" (graph capture failed: {graph_note})" if graph_note else ""))) if train else {
            "workload": "BASELINE configs[2]: GKGNet-576 inference, bf16 autocast, replicas only",
            "images_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world} (replicas)",
            "launch": "one CUDA graph per forward" if igraph is not None else "eager launches",
            "norm": "SyncBN" if args.syncbn else "per-GPU BN",
            "l2": "activations per step exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": img_h.numel() * 4 + tgt_h.numel() * 4, "d2h_bytes_per_step": out_bytes},
        "gpu_launches": int(launches),     # launches of libgkg_b200 kernels only (torch ops not counted)
        "roofline": None, "cpu_baseline": None,
    }
    if coll is not None:
        line["collective"] = coll
    if eager is not None:
        line["eager_gpu_baseline"] = eager
    if train and world == 1 and not args.no_cpu:
        v, cores, sample = time_cpu_train(args.ref_train_images, 1)
        line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    return line


def _oracle_train_step(sd, params, img, tgt, opt):
    from oracle import gkg_oracle as O
    labels, gap, _ = O.gkgnet_forward(sd, img, choice="s", training=True)
    losses = O.head_losses(sd, labels, gap, tgt, prefix="head.")
    loss = sum(losses.values())
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


def _oracle_state(device, dtype=torch.float32):
    """Random-init GKGNet-576 + head weights as the flat state dict the oracle's functional model takes."""
    import gkgnet_b200 as G
    G.set_norm_type("BN")
    torch.manual_seed(0)
    net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.0)
    head = G.LabelQueryHead(80, 640)
    sd = {k: v.detach().to(device=device, dtype=dtype if v.is_floating_point() else v.dtype).clone()
          for k, v in net.state_dict().items()}
    sd.update({"head." + k: v.detach().to(device=device, dtype=dtype).clone() for k, v in head.state_dict().items()})
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and "relative_pos" not in k and "running_" not in k:
            v.requires_grad_(True)
            params.append(v)
    return sd, params


def time_cpu_train(n_img, reps):
    """The reference's algorithm for the headline workload -- GKGNet-576 training step (fwd + head losses + bwd +
    grad clip + AdamW, fp32) through oracle/gkg_oracle.py (plain PyTorch CPU ops) on the host cores."""
    cores = use_all_host_threads()
    sd, params = _oracle_state("cpu")
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(n_img, 3, 576, 576, generator=g)
    tgt = (torch.rand(n_img, 80, generator=g) < 0.04).float()
    _oracle_train_step(sd, params, img[:1], tgt[:1], opt)          # warm-up (thread pool, allocator)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        _oracle_train_step(sd, params, img, tgt, opt)
        best = min(best, time.perf_counter() - t0)
    sample = (f"{n_img} images per step, fp32, GKGNet-576 fwd + losses + bwd + AdamW through oracle/gkg_oracle.py "
              f"(PyTorch CPU ops), best of {reps}: {best:.2f} s")
    return n_img / best, cores, sample


def eager_train_baseline(model, img, tgt, dev):
    """The same restatement of the reference (oracle functional model: ATen / cuBLAS / cuDNN kernels, the N x M distance
    matrices materialised) for one training step on this B200, fp32 and under bf16 autocast: what the reference's code
    path costs on the device today.  A reported baseline -- the checker's code, never the product path."""
    out = {}
    try:
        sd, params = _oracle_state(dev)
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)     # same optimizer kernel as our arm
        for name, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
            def one():
                if ctx is None:
                    return _oracle_train_step(sd, params, img, tgt, opt)
                with torch.autocast("cuda", dtype=ctx):
                    return _oracle_train_step(sd, params, img, tgt, opt)
            one()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            out[name] = {"value": img.shape[0] / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms}
        out["kind"] = "port"
        out["sample"] = (f"{img.shape[0]} images, GKGNet-576 training step through oracle/gkg_oracle.py on the same B200 "
                         "(eager PyTorch kernels, device-resident), 2 timed steps each")
        del sd, params, opt
        torch.cuda.empty_cache()
    except Exception as e:
        out = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
        torch.cuda.empty_cache()
    return out


def run_full(args):
    """Default: both clauses of the metric in one line (see the module docstring)."""
    world, rank, local = dist_setup(args.gpus)
    targs = argparse.Namespace(**vars(args))
    targs.workload = "train"
    tline = run_model(targs)
    torch.cuda.empty_cache()
    lline = run_layer(args)
    if rank != 0:
        return None
    line = dict(tline)
    line["config"] = dict(tline["config"])
    line["config"]["workload"] = (tline["config"]["workload"] + " [headline value / ms_per_step / e2e]; plus " +
                                  lline["config"]["workload"] + " [keys layer, phase_ms, roofline*]")
    line["layer"] = {k: lline[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "clocks",
                                           "cpu_baseline", "eager_gpu_baseline") if k in lline}
    line["layer"]["config"] = lline["config"]
    for k, v in lline.items():
        if k == "phase_ms" or k.startswith("roofline"):
            line[k] = v
    return line



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "exact", "tc"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-images", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dense-bias", action="store_true", help="do not use the separable bias fast path")
    ap.add_argument("--ref-images", type=int, default=4)
    ap.add_argument("--ref-max-steps", type=int, default=20)
    ap.add_argument("--ref-train-images", type=int, default=2)
    ap.add_argument("--ref-train-steps", type=int, default=6)
    ap.add_argument("--workload", default="full", choices=["full", "layer", "train", "infer"],
                    help="full = training step (headline) + stage-1 layer microbench in one line; or one part alone")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU for --workload train / infer")
    ap.add_argument("--syncbn", action="store_true", help="keep the reference's SyncBN (default: per-GPU BN)")
    ap.add_argument("--no-graph", action="store_true",
                    help="training: issue the step from Python through DistributedDataParallel instead of replaying "
                         "the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        line = {"full": run_full, "layer": run_layer}.get(args.workload, run_model)(args)
        if line is not None:
            print(json.dumps(line), flush=True)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
