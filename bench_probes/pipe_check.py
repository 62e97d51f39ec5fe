"""Pipelined vs three-launch octant kernels: agreement and timing at config 2."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _ops, _lib
dev = torch.device('cuda:0')
shapes = [(1, 8, 16, 16, 16), (2, 8, 24, 16, 32), (1, 32, 64, 64, 64), (1, 32, 128, 128, 128)]
for shape in shapes:
    C = shape[1]
    sw = ft.SWMatricize((None, *shape[1:]), head_dim=8, patch_size=8)
    torch.manual_seed(0)
    nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
    x = torch.randn(*shape, device=dev, requires_grad=True)
    gy = torch.randn(*shape, device=dev)
    res = {}
    for name, path in (("3launch", _lib.FZ_PATH_OCTANT_3LAUNCH), ("pipe", _lib.FZ_PATH_OCTANT_PIPELINE)):
        sw._geom.path = path
        for _ in range(2):
            y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
            (gx,) = torch.autograd.grad((y * gy).sum(), x)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0; NIT = 5
        for _ in range(NIT):
            e[0].record()
            y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
            e[1].record()
            (gx,) = torch.autograd.grad((y * gy).sum(), x)
            e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        res[name] = (y.detach().clone(), gx.clone())
        print(f"{shape} {name} path={_lib.lib().fz_last_path()} fwd {tf/NIT*1e3:.1f} us bwd {tb/NIT*1e3:.1f} us", flush=True)
    dy = (res["pipe"][0] - res["3launch"][0]).abs().max().item()
    dg = (res["pipe"][1] - res["3launch"][1]).abs().max().item()
    print(f"   maxdiff y {dy:.3g} gx {dg:.3g}  (|y| max {res['3launch'][0].abs().max().item():.3g})", flush=True)
