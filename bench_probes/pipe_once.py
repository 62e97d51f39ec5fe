"""One forward + one backward of the pipelined octant kernels at config 2 (for ncu)."""
import sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _lib
n, C = 128, 32
dev = torch.device('cuda:0')
lib = _lib.lib()
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
sw._geom.path = int(sys.argv[1]) if len(sys.argv) > 1 else _lib.FZ_PATH_AUTO
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.randn(1, C, n, n, n, device=dev); gy = torch.randn(1, C, n, n, n, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
u0, v0 = nmf.init.u0, nmf.init.v0
st = torch.cuda.current_stream().cuda_stream
g, s = sw._geom.c_geom(1), nmf.solver_spec().c_solver()
saved = torch.empty(lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
ws = torch.zeros(lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
for _ in range(2):
    _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
    _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(), gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
torch.cuda.synchronize()
print("path", lib.fz_last_path())
