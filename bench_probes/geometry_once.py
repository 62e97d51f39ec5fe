"""One forward + backward of the fused core at a reduced isles22 geometry (8x64 windows, S=4) and a reduced brats23
geometry (8x512 windows, shifts [0,2,4,6]): the command ncu wraps for the sub-warp and paired-octant paths."""
import sys
sys.path.insert(0, '.')
import torch
import factorizer_b200 as ft
from factorizer_b200 import _ops

dev = torch.device('cuda:0')
for C, n, kw in ((32, 64, dict(head_dim=8, patch_size=4, shifts=[None, 1, 2, 3])),
                 (32, 128, dict(head_dim=8, patch_size=8, shifts=[None, 2, 4, 6]))):
    sw = ft.SWMatricize((None, C, n, n, n), **kw)
    nmf = ft.NMF(sw.output_size[2:], rank=1, num_iters=5, init="uniform", solver="hals").to(dev)
    x = torch.rand(2, C, n, n, n, device=dev, requires_grad=True)
    for _ in range(2):
        y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
