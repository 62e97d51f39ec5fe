"""Time fwd / bwd of the fused core at config 2 for several values of an environment knob.
    python bench_probes/sweep.py FZ_LAG_FWD 20,40,76 FZ_LAG_BWD 20,39,60
Launches are queued back to back (no host sync inside the timed loop)."""
import ctypes, os, sys
sys.path.insert(0, '.')
import torch
import factorizer_b200 as ft
from factorizer_b200 import _lib

dev = torch.device('cuda:0')
C, n = 32, 128
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
lib = _lib.lib()
g, s = sw._geom.c_geom(1), nmf.solver_spec().c_solver()
x = torch.rand(1, C, n, n, n, device=dev); gy = torch.randn(1, C, n, n, n, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
saved = torch.empty(lib.fz_swnmf_saved_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
ws = torch.zeros(lib.fz_swnmf_workspace_bytes(ctypes.byref(g), ctypes.byref(s)), dtype=torch.uint8, device=dev)
u0, v0 = nmf.init.u0, nmf.init.v0
sp = torch.cuda.current_stream(dev).cuda_stream
def fwd():
    _lib.check(lib.fz_swnmf_forward(x.data_ptr(), u0.data_ptr(), v0.data_ptr(), y.data_ptr(), saved.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
def bwd():
    _lib.check(lib.fz_swnmf_backward(x.data_ptr(), gy.data_ptr(), u0.data_ptr(), v0.data_ptr(), saved.data_ptr(), gx.data_ptr(), ws.data_ptr(), ctypes.byref(g), ctypes.byref(s), 1, sp))
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
args = sys.argv[1:]
if args and args[0] == "once":      # for ncu: a few fwd/bwd pairs and nothing else
    for _ in range(3):
        fwd(); bwd()
    torch.cuda.synchronize()
    sys.exit(0)
fwd(); bwd(); torch.cuda.synchronize()
y0, gx0 = y.clone(), gx.clone()
def check():
    gx.fill_(float("nan")); y.fill_(float("nan"))
    fwd(); bwd(); torch.cuda.synchronize()
    ye = torch.equal(y, y0)
    gd = (gx - gx0).abs().max().item()
    return f"y_equal={ye} gx_maxdiff={gd:.3g}"
print(f"default: fwd {timeit(fwd):.1f} us  bwd {timeit(bwd):.1f} us  {check()}")
for k in range(0, len(args), 2):
    name, vals = args[k], args[k + 1].split(',')
    for v in vals:
        os.environ[name] = v
        print(f"{name}={v}: fwd {timeit(fwd):.1f} us  bwd {timeit(bwd):.1f} us  {check()}")
    os.environ.pop(name, None)
