"""Fused core at config 2 through the C ABI, CUDA events: three-launch vs pipelined octant kernels.
usage: core_time.py [n=128] [C=32] [reps=20]"""
import sys, os, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device('cuda:0')
lib = _lib.lib()
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.randn(1, C, n, n, n, device=dev); gy = torch.randn(1, C, n, n, n, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
u0, v0 = nmf.init.u0, nmf.init.v0
st = torch.cuda.current_stream().cuda_stream
outs = {}
for name, path in (("3launch", _lib.FZ_PATH_OCTANT_3LAUNCH), ("pipe", _lib.FZ_PATH_OCTANT_PIPELINE)):
    sw._geom.path = path
    g, s = sw._geom.c_geom(1), nmf.solver_spec().c_solver()
    saved = torch.empty(lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
    ws = torch.zeros(lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
    fwd = lambda: _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
    bwd = lambda: _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(), gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, st))
    for _ in range(3):
        fwd(); bwd()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    for k in range(reps):
        ev[k][0].record(); fwd(); ev[k][1].record(); bwd(); ev[k][2].record()
    torch.cuda.synchronize()
    tf = sum(e[0].elapsed_time(e[1]) for e in ev) / reps * 1e3
    tb = sum(e[1].elapsed_time(e[2]) for e in ev) / reps * 1e3
    nel = C * n ** 3
    print(f"n={n} C={C} {name}: path {lib.fz_last_path()} fwd {tf:.1f} us ({2*nel*4/tf/1e3:.0f} GB/s) bwd {tb:.1f} us ({3*nel*4/tb/1e3:.0f} GB/s) total {tf+tb:.1f} us -> frac of 6541.5 = {5*nel*4/(tf+tb)/1e3/6541.5:.3f}", flush=True)
    outs[name] = (y.clone(), gx.clone())
print("equal:", torch.equal(outs["pipe"][0], outs["3launch"][0]), torch.equal(outs["pipe"][1], outs["3launch"][1]))
