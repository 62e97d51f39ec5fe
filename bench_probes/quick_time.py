import torch, time, sys
sys.path.insert(0, '.')
import factorizer_b200 as ft
from factorizer_b200 import _ops, _lib
dev = torch.device('cuda:0')
NIT = int(sys.argv[1]) if len(sys.argv) > 1 else 10
C, n = 32, 128
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.rand(1, C, n, n, n, device=dev, requires_grad=True)
gy = torch.randn(1, C, n, n, n, device=dev)
for path in (_lib.FZ_PATH_AUTO, _lib.FZ_PATH_OCTANT_3LAUNCH):
    sw._geom.path = path
    for _ in range(3):
        y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
        y.backward(gy)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0
    for _ in range(NIT):
        e[0].record()
        y = _ops.SWNMF.apply(x, nmf.init.u0, nmf.init.v0, sw._geom, nmf.solver_spec(), True)
        e[1].record()
        y.backward(gy)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f"path={path} last_path={_lib.lib().fz_last_path()} fwd {tf/NIT*1e3:.1f} us  bwd {tb/NIT*1e3:.1f} us  total {(tf+tb)/NIT*1e3:.1f} us  -> {1342.177/((tf+tb)/NIT):.0f} GB/s algorithmic")
    if path == _lib.FZ_PATH_AUTO: yf, gf = y.detach().clone(), x.grad.clone()
    else:
        print("pipelined vs three-launch: y maxdiff", (yf - y).abs().max().item(), "gx maxdiff", (gf - x.grad).abs().max().item(), "gx max", x.grad.abs().max().item())
    x.grad = None
