import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _ops, _lib
dev = torch.device('cuda:0')
shape = tuple(int(v) for v in sys.argv[1].split(","))
sw = ft.SWMatricize((None, *shape[1:]), head_dim=8, patch_size=8)
sw._geom.path = _lib.FZ_PATH_OCTANT_PIPELINE
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.randn(*shape, device=dev, requires_grad=True)
gy = torch.randn(*shape, device=dev)
y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
torch.cuda.synchronize(); print("fwd ok")
(gx,) = torch.autograd.grad((y * gy).sum(), x)
torch.cuda.synchronize(); print("bwd ok")
