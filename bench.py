#!/usr/bin/env python
"""Benchmark of the Factorizer hot path on B200 (see BASELINE.json: metric / configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 2 -- ft.SWMatricize(head_dim 8, patch 8) + ReLU +
ft.NMF(rank 1, 5 HALS sweeps) + inverse, forward + backward on one (1, 32, 128^3) fp32 volume per GPU,
i.e. the fused FactMixer core that FactorizerBlock runs between its two 1x1 projections.  A "step" is
one forward + one backward through the C ABI (fz_swnmf_forward / fz_swnmf_backward).  The same line
also carries the whole FactorizerBlock (BASELINE config 3: fused glue kernels around the fused core) in
`block`, and the whole Swin Factorizer (configs 4-5: inference pass and single-rank training step) in `model`.

value      voxels/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks); the K steps are
           timed twice, as plain stream launches (which also gives the fwd / bwd split) and replayed from one CUDA graph
           per step, and the shorter total is reported (config.launch says which)
e2e        same metric with HOST buffers: pinned-host x and dY copied in, y and dX copied out, every step
           (steps pipelined over copy-in / kernel / copy-out streams with double-buffered device tensors)
roofline   dominant kernel (phase_bwd_apply = pass 3 of the backward: reads X and dY, writes dX, i.e. exactly
           the backward's compulsory traffic): algorithmic bytes / event-timed launch duration vs the
           measured HBM peak in MEASURED_PEAKS.json.  The kernel is isolated with the C ABI's measurement
           hook fz_set_pass_mask() after complete calls have filled the intermediate buffers; `passes_us`
           carries all six kernels of a step timed the same way.
cpu_baseline / --impl reference
           the oracle's C/OpenMP port of the reference path (oracle/nmf_oracle.c) on the host cores,
           on a bounded sample of the same workload (the reference itself is pure PyTorch and does not
           exist on the GPU box)
Multi-GPU: one process per GPU (torchrun), batch-sharded (one volume per rank, no data-path
collective), weak scaling.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, N, HEAD_DIM, PATCH, T_ITERS = 32, 128, 8, 8, 5
METRIC = "FactorizerBlock voxels/s fwd+bwd @128^3 (fused SWMatricize+NMF core), % of HBM roofline"
UNIT = "voxels/s"
WORKLOAD = ("SWMatricize(head_dim=8,patch=8,shifts=[None,4])+ReLU+NMF(rank=1,iters=5,hals)+inverse "
            "fwd+bwd on (1,32,128,128,128) fp32 per GPU")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_rate(sample_n: int, reps: int):
    """C/OpenMP oracle port of the reference path on a (1,32,sample_n^3) sample; voxels/s."""
    import numpy as np

    from oracle import c_oracle

    c_oracle.use_all_cores()
    rng = np.random.default_rng(0)
    x = rng.random((1, C, sample_n, sample_n, sample_n), dtype=np.float32)
    gy = rng.standard_normal(x.shape, dtype=np.float32)
    v0 = rng.random(512, dtype=np.float32)
    shifts = [(0, 0, 0), (PATCH // 2,) * 3]
    times = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        c_oracle.swnmf_forward(x, v0, HEAD_DIM, (PATCH,) * 3, shifts, relu=True, num_iters=T_ITERS)
        c_oracle.swnmf_backward(x, gy, v0, HEAD_DIM, (PATCH,) * 3, shifts, relu=True, num_iters=T_ITERS)
        times.append(time.perf_counter() - t0)
    times = times[1:]  # first rep warms the page cache / thread pool
    return sample_n ** 3 / statistics.median(times), times, c_oracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_n = 64
    # each "step" is one fwd+bwd of the port on the bounded sample
    rate, times, threads = cpu_reference_rate(sample_n, max(args.steps, 1) + args.warmup - 1)
    times = times[-max(args.steps, 1):]
    ms = 1e3 * statistics.mean(times)
    sample = f"(1,{C},{sample_n}^3) = 1/8 of the workload volume, same geometry/solver, fwd+bwd"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "oracle/nmf_oracle.c (C/OpenMP port of the reference's "
                   "PyTorch path; the reference is pure Python and absent on the GPU box)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_cpus(index: int):
    """Run this rank on the CPU cores NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end leg are first-touched on the NUMA node next to the GPU (with one rank per GPU every rank otherwise
    allocates wherever the launcher happened to start it, and the copies of several ranks share one socket's
    memory controllers and inter-socket links).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(masks) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = cpus & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} GPU-local cores"
        return "all cores are GPU-local"
    except Exception as e:       # no NVML / not permitted: keep the inherited affinity
        return f"not bound ({type(e).__name__})"


class ClockSampler:
    def __init__(self, index: int):
        self.path = f"/tmp/fz_clocks_{os.getpid()}.csv"
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def model_leg(dev, n=128, steps=3):
    """Configs 4 and 5: the README Swin Factorizer on this package's kernels, one 128^3 volume per GPU --
    an inference pass and a training step (forward, sigmoid-BCE + soft-Dice, backward, AdamW), fp32, cuDNN / cuBLAS for
    the convolutions and the wide stages' GEMMs.  Side measurement; never raises."""
    import torch
    import factorizer_b200 as ft
    from torch import nn
    keep = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.benchmark = True       # the heuristic choice for the 4->32 stem's fp32 wgrad is 10x slower
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.manual_seed(1234)
        net = ft.Factorizer(in_channels=4, out_channels=3, spatial_size=(n, n, n), encoder_depth=(1, 1, 1, 1, 1),
                            encoder_width=(32, 64, 128, 256, 512), strides=(1, 2, 2, 2, 2), decoder_depth=(1, 1, 1, 1),
                            norm=ft.LayerNorm, reshape=(ft.SWMatricize, {"head_dim": HEAD_DIM, "patch_size": PATCH}),
                            act=nn.ReLU, factorize=ft.NMF, rank=1, num_iters=T_ITERS, init="uniform", solver="hals",
                            mlp_ratio=2, dropout=0.1).to(dev)
        x = torch.rand(1, 4, n, n, n, device=dev)
        target = torch.randint(0, 2, (1, 3, n, n, n), device=dev).float()
        ev = lambda: torch.cuda.Event(enable_timing=True)

        def timed(fn):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            a, b = ev(), ev()
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / steps

        net.eval()
        with torch.no_grad():
            infer_ms = timed(lambda: net(x))
        net.train()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-5)

        def train_step():
            opt.zero_grad(set_to_none=True)
            logits = net(x)
            p = torch.sigmoid(logits)
            dice = 1 - (2 * (p * target).sum((2, 3, 4)) + 1e-5) / (p.sum((2, 3, 4)) + target.sum((2, 3, 4)) + 1e-5)
            loss = nn.functional.binary_cross_entropy_with_logits(logits, target) + dice.mean()
            loss.backward()
            opt.step()

        train_ms = timed(train_step)
        out = {"workload": f"Swin Factorizer (README 78-96) 4->3 ch, {n}^3, widths (32,64,128,256,512), HALS r1, B=1/GPU, fp32, "
                           "cudnn.benchmark; wide-stage GEMMs and the patch (down / up / head) convolutions' forward are cuBLAS, the stem's "
                           "forward and every weight gradient of those layers are csrc/fz_linear.cu",
               "params": sum(q.numel() for q in net.parameters()),
               "infer_ms": infer_ms, "infer_voxels_per_s_per_gpu": n ** 3 / (infer_ms * 1e-3),
               "train_step_ms": train_ms, "train_voxels_per_s_per_gpu": n ** 3 / (train_ms * 1e-3),
               "train_step": "forward, sigmoid-BCE + soft-Dice, backward, AdamW; no gradient all-reduce (single-rank step)"}
        del net, opt, x, target
        torch.cuda.empty_cache()
        return out
    except Exception as e:                              # a side measurement must not cost the bench line
        try:
            torch.cuda.synchronize(dev)
        except Exception:
            pass
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = keep


def run_ours(args):
    import torch
    import torch.distributed as dist

    import factorizer_b200 as ft
    from factorizer_b200 import _lib, _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local)       # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    torch.manual_seed(1234 + rank)
    sw = ft.SWMatricize((None, C, N, N, N), head_dim=HEAD_DIM, patch_size=PATCH)
    nmf = ft.NMF(sw.output_size[2:], rank=1, num_iters=T_ITERS, init="uniform", solver="hals").to(dev)
    geom, spec = sw._geom, nmf.solver_spec()
    g, s = geom.c_geom(1), spec.c_solver()
    x = torch.rand(1, C, N, N, N, device=dev)
    gy = torch.randn(1, C, N, N, N, device=dev)
    y = torch.empty_like(x)
    gx = torch.empty_like(x)
    saved = torch.empty(lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
    ws = torch.zeros(lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
    u0, v0 = nmf.init.u0, nmf.init.v0
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    launches = [0]

    def fwd(on=None):
        _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), saved.data_ptr(),
                                        ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp if on is None else on))
        launches[0] += lib.fz_last_launches()

    def bwd(on=None):
        _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(),
                                         gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp if on is None else on))
        launches[0] += lib.fz_last_launches()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- resident-in-HBM timing ----------------
    for _ in range(max(args.warmup, 3)):
        fwd(); bwd()
    fast_path = lib.fz_last_path()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches[0] = 0
    barrier()
    for k in range(args.steps):
        ev[k][0].record(stream); fwd(); ev[k][1].record(stream); bwd(); ev[k][2].record(stream)
    barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    fwd_us = 1e3 * statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    bwd_us = 1e3 * statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    timed_launches = launches[0]

    # The same K steps replayed from ONE CUDA graph (the C ABI is capturable: no host synchronisation, no allocation):
    # identical kernels and arguments, without the per-launch host work and the event records between them.  `value`
    # is taken from whichever of the two timings is shorter; `config.launch` says which.
    launch_mode = "stream launches"
    try:
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            n_before = launches[0]
            with torch.cuda.graph(graph, stream=cap):
                fwd(cap.cuda_stream); bwd(cap.cuda_stream)
            per_step_launches = launches[0] - n_before
        stream.wait_stream(cap)
        for _ in range(3):
            graph.replay()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            graph.replay()
        g1.record(stream)
        barrier()
        graph_ms = g0.elapsed_time(g1)
        if graph_ms < total_ms:
            total_ms, launch_mode = graph_ms, "one CUDA graph per step (fwd+bwd, %d kernel nodes)" % per_step_launches
            timed_launches = per_step_launches * args.steps
    except Exception as e:            # capture not available: keep the stream-launch timing
        launch_mode = f"stream launches (graph capture failed: {type(e).__name__})"
        torch.cuda.synchronize(dev)

    # ---------------- the kernels of one step, one at a time (octant path only) ----------------
    passes_us = None
    if fast_path == 2:
        passes_us = {}
        names = {("fwd", 1): "phase_fwd_gram", ("fwd", 2): "phase_fwd_solve", ("fwd", 4): "phase_fwd_apply",
                 ("bwd", 1): "phase_bwd_reduce", ("bwd", 2): "phase_bwd_solve", ("bwd", 4): "phase_bwd_apply"}
        reps = max(5, min(args.steps, 20))
        for (direction, mask), name in names.items():
            fn = fwd if direction == "fwd" else bwd
            lib.fz_set_pass_mask(mask)
            fn(); fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize(dev)
            passes_us[name] = 1e3 * a.elapsed_time(b) / reps
        lib.fz_set_pass_mask(7)
        fwd(); bwd()      # leave every buffer consistent again
        torch.cuda.synchronize(dev)

    # ---------------- end-to-end with host buffers ----------------
    # Every step copies its own x and dY in from pinned host memory and its y and dX back out; consecutive steps
    # are pipelined over three streams (copy-in | kernels | copy-out) with two sets of device buffers, so the
    # host->device copy of step k+1 and the device->host copy of step k-1 run while step k computes (PCIe is
    # full duplex).
    hx = torch.rand(1, C, N, N, N).pin_memory()
    hgy = torch.randn(1, C, N, N, N).pin_memory()
    hy = torch.empty(1, C, N, N, N).pin_memory()
    hgx = torch.empty(1, C, N, N, N).pin_memory()
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    bufs = [dict(x=x, gy=gy, y=y, gx=gx),
            dict(x=torch.empty_like(x), gy=torch.empty_like(x), y=torch.empty_like(x), gx=torch.empty_like(x))]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    comp_done = [torch.cuda.Event() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]

    def e2e_step(k):
        b = bufs[k % 2]
        s_in.wait_event(comp_done[k % 2])          # the kernels of step k-2 have consumed these input buffers
        with torch.cuda.stream(s_in):
            b["x"].copy_(hx, non_blocking=True)
            b["gy"].copy_(hgy, non_blocking=True)
            in_ready[k % 2].record(s_in)
        stream.wait_event(in_ready[k % 2])
        stream.wait_event(out_done[k % 2])          # step k-2's results have left these output buffers
        _lib.check(lib.fz_swnmf_forward(b["x"].data_ptr(), u0.data_ptr(), v0.data_ptr(), b["y"].data_ptr(), saved.data_ptr(),
                                        ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
        _lib.check(lib.fz_swnmf_backward(b["x"].data_ptr(), b["gy"].data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(),
                                         b["gx"].data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
        comp_done[k % 2].record(stream)
        s_out.wait_event(comp_done[k % 2])
        with torch.cuda.stream(s_out):
            hy.copy_(b["y"], non_blocking=True)
            hgx.copy_(b["gx"], non_blocking=True)
            out_done[k % 2].record(s_out)

    e2e_steps = max(4, min(args.steps, 10))
    for k in range(2):
        e2e_step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s_in.wait_event(e0)
    for k in range(e2e_steps):
        e2e_step(k)
    stream.wait_stream(s_out)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    del bufs

    # ---------------- whole FactorizerBlock (config 3): fused glue kernels around the fused core ----------------
    block = None
    if not args.no_block:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        del hx, hgy, hy, hgx
        blk = ft.FactorizerBlock(channels=C, spatial_size=(N, N, N), norm=ft.LayerNorm,
                                 reshape=(ft.SWMatricize, {"head_dim": HEAD_DIM, "patch_size": PATCH}),
                                 act=torch.nn.ReLU, factorize=ft.NMF, rank=1, num_iters=T_ITERS, init="uniform",
                                 solver="hals", mlp_ratio=2, dropout=0.0).to(dev)
        xb = torch.rand(1, C, N, N, N, device=dev, requires_grad=True)
        block_fused = blk._fused_args(xb) is not None
        for _ in range(2):
            blk(xb).backward(gy)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = 5
        b0.record(stream)
        for _ in range(nb):
            blk(xb).backward(gy)
        b1.record(stream)
        barrier()
        bms = b0.elapsed_time(b1) / nb
        block_launch = "stream launches from autograd"
        bgraph = None
        try:        # the same step replayed from one CUDA graph (forward, autograd backward, gradient accumulation)
            for p_ in blk.parameters():
                p_.grad = None
            xb.grad = None
            cap = torch.cuda.Stream(dev)
            cap.wait_stream(stream)
            with torch.cuda.stream(cap):
                for _ in range(2):
                    blk(xb).backward(gy)
                for p_ in blk.parameters():
                    p_.grad = None
                xb.grad = None
                bgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(bgraph, stream=cap):
                    blk(xb).backward(gy)
            stream.wait_stream(cap)
            for _ in range(2):
                bgraph.replay()
            barrier()
            b0.record(stream)
            for _ in range(nb):
                bgraph.replay()
            b1.record(stream)
            barrier()
            gms = b0.elapsed_time(b1) / nb
            if gms < bms:
                bms, block_launch = gms, "one CUDA graph per step"
        except Exception as e:
            block_launch = f"stream launches from autograd (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize(dev)
        block = {"workload": "FactorizerBlock(32,128^3,LayerNorm,SWMatricize,HALS r1,mlp_ratio=2,dropout=0) fwd+bwd incl. "
                             "parameter gradients, B=1/GPU, fp32",
                 "path": "hand-written glue kernels (fz_block_glue.cu on the FP32 pipe; forward out_proj+norm2+MLP on tcgen05/TMEM, 3xTF32, fz_block_glue_tc.cu) + fused core: 3 launches fwd, 4 bwd" if block_fused
                         else "layer by layer (library GEMMs) around the fused core",
                 "launch": block_launch, "ms_per_step": bms, "voxels_per_s_per_gpu": N ** 3 / (bms * 1e-3)}

    # ---------------- whole model (configs 4 and 5), every rank its own replica ----------------
    model = None
    if not args.no_model:
        if block is not None:
            blk = xb = bgraph = None                # release the block leg's tensors and graph pool
        torch.cuda.empty_cache()
        model = model_leg(dev)
        barrier()

    # ---------------- reduce over ranks ----------------
    dom_us = passes_us["phase_bwd_apply"] if passes_us else bwd_us
    vals = torch.tensor([total_ms, e2e_ms, fwd_us, bwd_us, block["ms_per_step"] if block else 0.0, dom_us],
                        device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_us, bwd_us, block_ms, dom_us = vals.tolist()

    if rank == 0:
        peak, peak_src = load_peaks()
        n_el = C * N ** 3
        voxels = N ** 3
        ms_per_step = total_ms / args.steps
        bwd_bytes, fwd_bytes = 3 * n_el * 4, 2 * n_el * 4
        dom_kernel = "phase_bwd_apply" if passes_us else "swnmf_bwd (whole backward call)"
        achieved = bwd_bytes / (dom_us * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("phase_bwd_apply_dram_bytes_per_launch" if passes_us else "swnmf_bwd_fast_dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": world * voxels / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "volumes_per_gpu": 1, "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "inputs (3 x 256 MiB per step) are larger than the 126 MB L2; no explicit flush",
                       "host_affinity": numa,
                       "path": {0: "generic", 1: "window-at-a-time TMA/register kernels", 2: "three-pass octant kernels"}[fast_path],
                       "launch": launch_mode, "fwd_us": fwd_us, "bwd_us": bwd_us,
                       "fused_op_hbm_frac": (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / peak},
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bwd_bytes, "kernel_us": dom_us,
                         "note": "pass 3 of the backward reads X and dY and writes dX once each = the backward's algorithmic "
                                 "bytes (12*C B/voxel); the two passes before it re-read X and dY, which is why the whole "
                                 "op sits lower (fused_op_hbm_frac)",
                         "passes_us": passes_us,
                         "bwd_op": {"achieved": bwd_bytes / (bwd_us * 1e-6) / 1e9, "frac": bwd_bytes / (bwd_us * 1e-6) / 1e9 / peak,
                                    "algorithmic_bytes": bwd_bytes, "us": bwd_us},
                         "fwd_op": {"achieved": fwd_bytes / (fwd_us * 1e-6) / 1e9, "frac": fwd_bytes / (fwd_us * 1e-6) / 1e9 / peak,
                                    "algorithmic_bytes": fwd_bytes, "us": fwd_us}},
            "e2e": {"value": world * voxels / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * n_el * 4, "d2h_bytes_per_step": 2 * n_el * 4},
            "gpu_launches": timed_launches,
            "clocks": clocks,
        }
        if block:
            block["ms_per_step"] = block_ms
            block["voxels_per_s"] = world * voxels / (block_ms * 1e-3)
            # 20*C bytes per voxel is the absolute floor for the whole block too: x, dOut in; out, dX out
            block["hbm_frac_of_absolute_floor"] = 5 * n_el * 4 / (block_ms * 1e-3) / 1e9 / peak
            block.pop("voxels_per_s_per_gpu", None)
            line["block"] = block
        if model:
            line["model"] = model
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)       # the CPU baseline gets every core again
            rate, times, threads = cpu_reference_rate(64, 3)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"(1,{C},64^3) = 1/8 of the workload volume, same geometry/solver, fwd+bwd, "
                                              f"median of 3 ({1e3*statistics.median(times):.1f} ms)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-block", action="store_true", help="skip the FactorizerBlock (config 3) side measurement")
    ap.add_argument("--no-model", action="store_true", help="skip the whole-model (configs 4-5) side measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
