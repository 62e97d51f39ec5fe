"""Swin Factorizer (README.md:78-96 architecture, random init, synthetic BraTS-shaped volumes) on this package's
kernels: BASELINE config 4 (inference) and config 5 (training step: forward, sigmoid-BCE + soft-Dice loss, backward,
AdamW; DDP gradient all-reduce over NCCL when launched with torchrun).  Side measurement, not the bench line.

    python bench_probes/model_step.py [--mode infer|train] [--steps K] [--size 128]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_probes/model_step.py --mode train
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch import nn

import factorizer_b200 as ft


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--tf32", action="store_true", help="let the cuDNN / cuBLAS glue use TF32 (off: true fp32)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true", help="leave cuDNN's heuristic algorithm choice on")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark   # heuristics pick a 20 ms fp32 wgrad for the stem
    torch.backends.cudnn.allow_tf32 = args.tf32
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.manual_seed(1234)          # identical weights and NMF buffers on every rank
    n = args.size
    net = ft.Factorizer(in_channels=4, out_channels=3, spatial_size=(n, n, n), encoder_depth=(1, 1, 1, 1, 1),
                        encoder_width=(32, 64, 128, 256, 512), strides=(1, 2, 2, 2, 2), decoder_depth=(1, 1, 1, 1),
                        norm=ft.LayerNorm, reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU,
                        factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2,
                        dropout=0.1).to(dev)
    nparams = sum(p.numel() for p in net.parameters())
    torch.manual_seed(1234 + rank)
    x = torch.rand(1, 4, n, n, n, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    if args.mode == "infer":
        net.eval()
        with torch.no_grad():
            for _ in range(2):
                y = net(x)
            torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            for _ in range(args.steps):
                y = net(x)
            b.record()
            torch.cuda.synchronize()
        assert tuple(y.shape) == (1, 3, n, n, n) and bool(torch.isfinite(y).all())
    else:
        model = nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        target = torch.randint(0, 2, (1, 3, n, n, n), device=dev).float()

        def step():
            opt.zero_grad(set_to_none=True)
            logits = model(x)
            p = torch.sigmoid(logits)
            dice = 1 - (2 * (p * target).sum((2, 3, 4)) + 1e-5) / (p.sum((2, 3, 4)) + target.sum((2, 3, 4)) + 1e-5)
            loss = nn.functional.binary_cross_entropy_with_logits(logits, target) + dice.mean()
            loss.backward()
            opt.step()
            return loss

        for _ in range(2):
            loss = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(args.steps):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(loss))
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"Swin Factorizer 4->3 ch, {n}^3, widths (32,64,128,256,512), HALS rank 1, {args.mode}",
                          "params": nparams, "n_gpus": world, "ms_per_step": ms.item(),
                          "voxels_per_s": world * n ** 3 / (ms.item() * 1e-3), "tf32_glue": args.tf32, "cudnn_benchmark": not args.no_cudnn_benchmark,
                          "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
