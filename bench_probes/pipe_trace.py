"""Timeline of the pipelined forward/backward as CTA 0's watcher saw it (library built with FZ_TUNING=1)."""
import sys, os, ctypes, re, io
os.environ["FZ_PIPE_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _lib
n, C = 128, 32
dev = torch.device('cuda:0')
lib = _lib.lib()
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
sw._geom.path = _lib.FZ_PATH_OCTANT_PIPELINE
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.randn(1, C, n, n, n, device=dev); gy = torch.randn(1, C, n, n, n, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
u0, v0 = nmf.init.u0, nmf.init.v0
st = torch.cuda.current_stream().cuda_stream
g, s = sw._geom.c_geom(1), nmf.solver_spec().c_solver()
saved = torch.empty(lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
nws = lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s))
ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
NP = int(os.environ.get("NP", "64"))
ntr = 3 * (NP + 1)
off = int(os.environ.get('OFF', '23528448'))
for which in ("fwd", "bwd"):
    for _ in range(3):
        _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
        if which == "bwd":
            _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(), gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
    torch.cuda.synchronize()
    tr = ws[off:off + ntr * 8].view(torch.int64).cpu().numpy().reshape(3, NP + 1)
    t0 = tr[tr > 0].min()
    print(which, "step: A-complete  B-complete  C-complete (us since first event)")
    for k in range(0, NP + 1, 4):
        print(f"  {k:3d}: {(tr[0,k]-t0)/1e3:8.1f} {(tr[1,k]-t0)/1e3:8.1f} {(tr[2,k]-t0)/1e3:8.1f}")
    lg = ws[off + ntr * 8: off + ntr * 8 + 100 * 5 * 8].view(torch.int64).cpu().numpy().reshape(100, 5)
    for base, name in ((0, "solver"), (50, "consumer0")):
        print(name, "items: start | wait | compute | release (us), tag")
        for k in range(50):
            a, b, c, d, tag = lg[base + k]
            if a == 0: break
            print(f"  {k:2d}: start {(a-t0)/1e3:8.1f} wait {(b-a)/1e3:6.2f} compute {(c-b)/1e3:6.2f} release {(d-c)/1e3:6.2f} tag {tag}")
    ws[off:].zero_()
