"""Sweep of the pipelined octant kernels' plan parameters (needs a library built with FZ_TUNING=1)."""
import sys, os, ctypes, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import factorizer_b200 as ft
from factorizer_b200 import _lib
n, C, reps = 128, 32, 10
dev = torch.device('cuda:0')
lib = _lib.lib()
sw = ft.SWMatricize((None, C, n, n, n), head_dim=8, patch_size=8)
nmf = ft.NMF((8, 512), rank=1, num_iters=5, init='uniform', solver='hals').to(dev)
x = torch.randn(1, C, n, n, n, device=dev); gy = torch.randn(1, C, n, n, n, device=dev)
y = torch.empty_like(x); gx = torch.empty_like(x)
u0, v0 = nmf.init.u0, nmf.init.v0
st = torch.cuda.current_stream().cuda_stream
ref = None


def run(path, label):
    global ref
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
    ok = ""
    if ref is None:
        ref = (y.clone(), gx.clone())
    else:
        ok = f"equal={torch.equal(y, ref[0]) and torch.equal(gx, ref[1])}"
    print(f"{label}: fwd {tf:.1f} bwd {tb:.1f} total {tf+tb:.1f} us frac {5*C*n**3*4/(tf+tb)/1e3/6541.5:.3f} {ok}", flush=True)


run(_lib.FZ_PATH_OCTANT_3LAUNCH, "3launch")
groups = [int(v) for v in os.environ.get("SWEEP_GROUPS", "256,512,1024").split(",")]
targets = [int(v) for v in os.environ.get("SWEEP_LEADS", "1024,2048,3072,4096").split(",")]
flagss = [int(v) for v in os.environ.get("SWEEP_FLAGS", "7,0").split(",")]
actas = os.environ.get("SWEEP_ROLES", "10,1,4").split(";")
for grp, tgt, fl, ac in itertools.product(groups, targets, flagss, actas):
    os.environ["FZ_PIPE_GROUP"], os.environ["FZ_PIPE_LEAD"], os.environ["FZ_PIPE_FLAGS"], os.environ["FZ_PIPE_ROLES"] = str(grp), str(tgt), str(fl), str(ac)
    run(_lib.FZ_PATH_OCTANT_PIPELINE, f"pipe group={grp} lead={tgt} flags={fl} roles={ac}")
