"""Diagnostic: FactorizerBlock at n^3 against the fp64 torch + C-oracle composition, output by output."""
import sys, os
import numpy as np
import torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import factorizer_b200 as ft
from oracle import c_oracle as CO
import test_gpu_fullsize as T

dev = torch.device("cuda:0")
CO.use_all_cores()


def ratio(a, b, scale=1.0):
    a = a.detach().double().cpu() / scale; b = b.detach().double().cpu() / scale
    r = (a - b).abs() / (1e-5 + 1e-4 * b.abs())
    i = int(r.argmax())
    return float(r.max()), np.unravel_index(i, tuple(a.shape))


for n in (64, 96, 128):
    C = 32
    torch.manual_seed(11)
    blk = ft.FactorizerBlock(channels=C, spatial_size=(n, n, n), norm=ft.LayerNorm,
                             reshape=(ft.SWMatricize, {"head_dim": 8, "patch_size": 8}), act=nn.ReLU,
                             factorize=ft.NMF, rank=1, num_iters=5, init="uniform", solver="hals", mlp_ratio=2,
                             dropout=0.0).to(dev)
    x = torch.randn(1, C, n, n, n, device=dev).requires_grad_(True)
    gy = torch.randn(1, C, n, n, n, device=dev)
    y = blk(x)
    params = dict(blk.named_parameters())
    grads = torch.autograd.grad((y * gy).sum(), [x] + list(params.values()))
    torch.cuda.synchronize()
    y_ref, gx_ref, gp_ref = T._block_reference(blk, x, gy)
    print(n, "y", ratio(y, y_ref), "gx", ratio(grads[0], gx_ref), flush=True)
    d = (grads[0].double() - gx_ref).abs()
    bad = (d > 1e-3).nonzero()
    print("  bad count", bad.shape[0], "of", d.numel(), "first", bad[:5].tolist(), "last", bad[-5:].tolist())
    if bad.shape[0]:
        for ax in range(1, 5):
            vals = torch.unique(bad[:, ax])
            print("   axis", ax, "bad idx:", vals[:20].tolist(), "... n=", vals.numel())
    for (k, _), gp in zip(params.items(), grads[1:]):
        ref = gp_ref[k].reshape(gp.shape)
        print("  ", k, ratio(gp, ref, max(1.0, float(ref.abs().max())))[0])
